python scripts/mmvq_trace.py N=57344 K=8192 B=1 static=1
python scripts/mmvq_trace.py N=8192 K=28672 B=1 static=1
