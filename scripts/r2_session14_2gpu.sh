#!/bin/bash
# Round 2, GPU session 14 (gpurun --gpus 2): the 2-rank parity tests with the hub-concurrency bound on the sentences in flight.
mkdir -p gpurun_out
echo "== pytest comm"; timeout 900 python -m pytest tests/test_comm_gpu.py -m gpu -q --tb=short -s 2>&1 | grep -v Warning | tail -12
