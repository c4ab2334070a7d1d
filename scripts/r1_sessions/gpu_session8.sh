#!/bin/bash
# Session 8: BASELINE config 4 shape on one GPU (300K regions as a stepping stone, then 1M regions x 24 = ~1e9 edges).
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 400 python scripts/config4_1m.py --regions 300000 --walks 16000000 --sgns-walks 4000000 --out gpurun_out/config4_300k.json 2>&1 | tail -12
timeout 900 python scripts/config4_1m.py --regions 1000000 --walks 32000000 --sgns-walks 8000000 --out gpurun_out/config4_1m.json 2>&1 | tail -12
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
