#!/bin/bash
# Session 18 (gpurun --gpus 2): the multi-GPU row with the final library: NCCL data-parallel tests, both bench arms at N=2
# exactly as the driver launches them.
mkdir -p gpurun_out
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_comm_gpu.py -m gpu -q 2>&1 | tail -3
timeout 300 $TR --master-port 29611 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench18_reference_n2.json 2> gpurun_out/bench18_reference_n2.err; cut -c1-200 gpurun_out/bench18_reference_n2.json
timeout 600 $TR --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench18_tract24_n2.json 2> gpurun_out/bench18_tract24_n2.err; tail -c 300 gpurun_out/bench18_tract24_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench18_tract24_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
print("n_gpus", d["n_gpus"], "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], d["clocks"], d["gpu_launches"])
for k in ("walk","sgns"):
    st=d["stages"][k]; print(k, "value %.4g"%st["value"], "e2e %.4g"%st["e2e"]["value"], "kernel_ms %.3f"%st["kernel_ms"], st.get("kernel"))
PY
