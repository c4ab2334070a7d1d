#!/bin/bash
# Session 6: new tests (eval, exports, sgns), A/B of item-kernel builds, cost of the reductions.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu6.log
grep -E "passed|failed|rc=|Error|assert |FAILED" gpurun_out/pytest_gpu6.log | head -40
for dbg in 0 1 4 16 5 17; do
  echo "== tract24 4M walks DGE_SGNS_DEBUG=$dbg"
  DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py tract24 4000000 2>&1 | tail -1
done
for dbg in 0 1 4 16; do
  echo "== synth100k 1M walks D=128 DGE_SGNS_DEBUG=$dbg"
  DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py synth 100000 1000000 2>&1 | tail -1
done
