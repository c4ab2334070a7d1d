#!/bin/bash
# Round 2, GPU session 8 (1 GPU): sentence-resident kernels as the default -- whole suite, full-size agreement sweep of
# kernel G with row prefetch, CA quality, bench N = 1.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -s 2>&1 | grep -v Warning | tail -15
echo "== kernel G (wavefront + prefetch) at full size"
timeout 1200 python scripts/fullsize_staleness.py 0,222,148 0 2>&1 | tail -5
echo "== CA 1M walks: default (kernel G) vs item kernels"
timeout 900 python scripts/sgns_ab.py ca 1000000 --quality --variants default:0,items:65536 --conc 0 --tag r2s8_ca 2>&1 | tail -4
echo "== synth D=128 2M walks: default (kernel F) vs item kernel"
timeout 600 python scripts/sgns_ab.py synth 100000 2000000 --dim 128 --variants default:0,items:65536 --tag r2s8_synth128 2>&1 | tail -3
echo "== bench N=1"
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2s8_bench_n1.json 2> gpurun_out/r2s8_bench_n1.err; cut -c1-400 gpurun_out/r2s8_bench_n1.json; tail -3 gpurun_out/r2s8_bench_n1.err
