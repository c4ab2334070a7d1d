#!/bin/bash
# Round 2, GPU session 26 (1 GPU): the automatic schedule for narrow rows = kernel F, sentence counter, write-through words,
# a full GPU of sentences in flight (<= V): whole GPU suite, full-size runs (tract x 24 twice, CA), bench at N = 1.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -s 2>&1 | grep -v Warning | tail -12
echo "== full size, automatic schedule: tract x 24 (x2), CA"
timeout 900 python scripts/fullsize_staleness.py 0,0 0 r2s26_auto 2>&1 | grep -v Warning | tail -2
timeout 900 python scripts/fullsize_ca.py 0 0 r2s26_auto 2>&1 | grep -v Warning | tail -1
echo "== bench N=1"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s26_bench_n1.json 2> gpurun_out/r2s26_bench_n1.err; cut -c1-400 gpurun_out/r2s26_bench_n1.json; tail -3 gpurun_out/r2s26_bench_n1.err
