#!/bin/bash
# Session 10 (gpurun --gpus 8): BASELINE config 4 -- 1M regions x 24 slices, walks sharded by walk id over 8 B200,
# data-parallel skip-gram with NCCL delta all-reduce over NVLink.
mkdir -p gpurun_out
nvidia-smi -L | wc -l; free -g | head -2; nproc
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
NCCL_DEBUG=WARN timeout 500 $TR --master-port 29701 scripts/config4_1m.py --regions 1000000 --walks 32000000 --sgns-walks 8000000 --out gpurun_out/config4_1m_n8.json 2>&1 | grep -v Warning | tail -14
