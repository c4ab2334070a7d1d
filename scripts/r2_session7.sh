#!/bin/bash
# Round 2, GPU session 7 (1 GPU): kernel G with the wavefront schedule -- arithmetic test, agreement / throughput sweep.
mkdir -p gpurun_out
echo "== pytest sgns"; timeout 900 python -m pytest tests/test_sgns_gpu.py -m gpu -q --tb=short 2>&1 | tail -8
echo "== kernel G (wavefront) at full size"
timeout 1500 python scripts/fullsize_staleness.py 0,444,370,296,222,148 2052 2>&1 | tail -8
