#!/bin/bash
# Round 2, GPU session 11 (1 GPU): kernel H (pipelined wavefront) -- arithmetic, agreement / throughput at full size.
mkdir -p gpurun_out
echo "== pytest sgns"; timeout 600 python -m pytest tests/test_sgns_gpu.py -m gpu -q --tb=short 2>&1 | tail -8
echo "== kernel H at full size (flags 131072)"
timeout 900 python scripts/fullsize_staleness.py 0,148,74 131072 2>&1 | tail -5
