VERBOSE=1 TD=12,12,14,12,12,14 python scripts/block_trace.py
