TD=12 python scripts/block_trace.py
TD=12 PLAIN=1 python scripts/block_trace.py
