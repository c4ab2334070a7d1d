#!/bin/bash
# Session 14: unit-range trimming + warp-merged syn0 reductions: parity, then timing against session 12/13 numbers.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sgns_gpu.py tests/test_pipeline_gpu.py tests/test_comm_gpu.py -m gpu -q 2>&1 | tail -6
echo "== tract24 4M walks"; timeout 300 python scripts/prof_path.py tract24 4000000 2>&1 | tail -1
echo "== tract24 4M walks, no reductions"; DGE_SGNS_DEBUG=1 timeout 300 python scripts/prof_path.py tract24 4000000 2>&1 | tail -1
echo "== tract8 4M walks"; timeout 300 python scripts/prof_path.py tract 4000000 2>&1 | tail -1
echo "== synth D=128 1M walks"; timeout 300 python scripts/prof_path.py synth 100000 1000000 2>&1 | tail -1
echo "== synth D=16 1M walks"; timeout 300 python scripts/prof_path.py synth 100000 1000000 16 2>&1 | tail -1
echo "== uniform V=19224 L=24 D=20"; timeout 300 python scripts/uniform_corpus.py 19224 24 2000000 20 2>&1 | tail -1
timeout 600 python scripts/sgns_sweep.py --walks 1000000 --dims 16,128 --negatives 5,10 --windows 5,10 --out gpurun_out/sgns_sweep14.json 2>&1 | grep "D="
