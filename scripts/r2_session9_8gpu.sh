#!/bin/bash
# Round 2, GPU session 9 (gpurun --gpus 8): the data-parallel skip-gram across the 8 B200 of one NVSwitch box, as the
# driver's scaling run will launch it -- the data-parallel object alone (default rounds, peer-memory transport), the same
# with NCCL for the A/B, then the full bench line at N = 8.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== nvidia-smi"; nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== dp-only N=8 default"
timeout 900 $TR --master-port 29631 bench.py --gpus 8 --dp-only 2> gpurun_out/r2s9_dp8_default.err | tail -1 > gpurun_out/r2s9_dp8_default.json; cut -c1-2500 gpurun_out/r2s9_dp8_default.json; tail -3 gpurun_out/r2s9_dp8_default.err | cut -c1-300
echo "== dp-only N=8 NCCL transport"
timeout 900 $TR --master-port 29632 bench.py --gpus 8 --dp-only --transport 2 2> gpurun_out/r2s9_dp8_nccl.err | tail -1 > gpurun_out/r2s9_dp8_nccl.json; cut -c1-1500 gpurun_out/r2s9_dp8_nccl.json
echo "== full bench N=8"
timeout 1200 $TR --master-port 29633 bench.py --gpus 8 --steps 3 --warmup 2 2> gpurun_out/r2s9_bench_n8.err | tail -1 > gpurun_out/r2s9_bench_n8.json; cut -c1-1200 gpurun_out/r2s9_bench_n8.json; tail -3 gpurun_out/r2s9_bench_n8.err | cut -c1-300
