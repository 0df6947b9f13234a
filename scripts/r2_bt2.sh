VERBOSE=0 TD=12 python scripts/block_trace.py
python scripts/mmvq_probe.py time types=16,17,22,18,21,19,29 static=1 B=1,8 reps=2 shapes=14336x4096
