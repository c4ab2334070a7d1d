#!/bin/bash
# Round 2, GPU session 21 (1 GPU): the automatic schedule now picks kernel F with write-through words where it pays
# (tract x 24: 12 warps per SM): whole GPU suite, full-size parity three times (run-to-run spread), bench at N = 1.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -s 2>&1 | grep -v Warning | tail -12
echo "== full size, automatic schedule x 3"
timeout 900 python scripts/fullsize_staleness.py 0,0,0 0 r2s21_auto 2>&1 | grep -v Warning | tail -3
echo "== bench N=1"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s21_bench_n1.json 2> gpurun_out/r2s21_bench_n1.err; cut -c1-400 gpurun_out/r2s21_bench_n1.json; tail -3 gpurun_out/r2s21_bench_n1.err
