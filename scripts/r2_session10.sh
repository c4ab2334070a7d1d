#!/bin/bash
# Round 2, GPU session 10 (1 GPU): kernel G with the sentence's negatives drawn before the rounds.
mkdir -p gpurun_out
echo "== pytest sgns"; timeout 900 python -m pytest tests/test_sgns_gpu.py -m gpu -q --tb=short 2>&1 | tail -6
echo "== kernel G at full size"
timeout 1200 python scripts/fullsize_staleness.py 0,148 0 2>&1 | tail -4
echo "== CA 1M"
timeout 600 python scripts/sgns_ab.py ca 1000000 --variants default:0 --conc 0 --tag r2s10_ca 2>&1 | tail -1
