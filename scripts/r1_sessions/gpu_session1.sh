#!/bin/bash
# one GPU session: tests, bench (both arms), ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; lscpu | grep "Model name" >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 120 python scripts/prof_path.py tract 2000000 > gpurun_out/prof_tract.log 2>&1
timeout 300 python scripts/prof_path.py synth 20000 2000000 > gpurun_out/prof_synth.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_tract24.json 2> gpurun_out/bench_tract24.err; echo "bench rc=$?" >> gpurun_out/bench_tract24.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/prof_tract.log gpurun_out/prof_synth.log; cat gpurun_out/bench_tract24.json; tail -3 gpurun_out/bench_tract24.err; cat gpurun_out/bench_ref.json
