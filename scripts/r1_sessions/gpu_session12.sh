#!/bin/bash
# Session 12: narrow-row item kernel (4-lane groups, two slots per lane): parity, then A/B against the 8-lane build.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sgns_gpu.py tests/test_pipeline_gpu.py -m gpu -q 2>&1 | tail -15
for dbg in 2 0 3 1; do
  echo "== tract24 4M walks DGE_SGNS_DEBUG=$dbg"
  DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py tract24 4000000 2>&1 | tail -1
done
for dbg in 2 0; do
  echo "== tract8 4M walks DGE_SGNS_DEBUG=$dbg"
  DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py tract 4000000 2>&1 | tail -1
  echo "== synth D=32 1M walks DGE_SGNS_DEBUG=$dbg"
  DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py synth 100000 1000000 32 2>&1 | tail -1
  echo "== synth D=16 1M walks DGE_SGNS_DEBUG=$dbg"
  DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py synth 100000 1000000 16 2>&1 | tail -1
done
