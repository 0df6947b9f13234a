#!/bin/bash
python scripts/mmvq_probe.py time types=12,14 static=1 B=1 reps=4 shapes=4096x14336
python scripts/mmvq_probe.py time types=12,14 static=1 B=1 reps=4 rg=513 shapes=4096x14336
