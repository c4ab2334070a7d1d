#!/bin/bash
# Session 17: kernel E with the table lookups two pair steps ahead: parity, timing, the CA bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sgns_gpu.py tests/test_pipeline_gpu.py -m gpu -q 2>&1 | tail -4
echo "== ca 2M walks D=8, target-parallel (auto)"; timeout 300 python scripts/prof_path.py ca 2000000 2>&1 | tail -1
echo "== ca 2M walks D=2, target-parallel (auto)"; timeout 300 python scripts/prof_path.py ca 2000000 0 2 2>&1 | tail -1
echo "== ca 2M walks D=16, target-parallel (auto)"; timeout 300 python scripts/prof_path.py ca 2000000 0 16 2>&1 | tail -1
timeout 600 python bench.py --workload ca --steps 3 --warmup 3 > gpurun_out/bench17_ca.json 2> gpurun_out/bench17_ca.err; tail -c 400 gpurun_out/bench17_ca.err; cut -c1-300 gpurun_out/bench17_ca.json
