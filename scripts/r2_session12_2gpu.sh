#!/bin/bash
# Round 2, GPU session 12 (gpurun --gpus 2): the 2-rank parity tests with the sentence-resident kernels as the default, and
# the data-parallel object with the replica arena / cached peer mappings (exchange set-up paid once per communicator).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== pytest comm"; timeout 900 python -m pytest tests/test_comm_gpu.py -m gpu -q --tb=short -s 2>&1 | grep -v Warning | tail -12
echo "== bench N=2 data-parallel only"
timeout 900 $TR --master-port 29641 bench.py --gpus 2 --dp-only 2> gpurun_out/r2s12_dp2.err | tail -1 > gpurun_out/r2s12_dp2.json; tail -2 gpurun_out/r2s12_dp2.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s12_dp2.json")); s = d["sgns"]
print("agg pairs/s %.4g (kernel-only %.4g) call_ms %.0f phases %s sync_ms %.1f rounds %s %s" % (s["value"], s["kernel_pairs_per_s"], s["call_ms"], s["call_phases_ms"], s["sync_ms"], s["sync_rounds"], s["transport"]))
print("single", d["single_gpu_reference"]); print("agreement", d["agreement"]); print("stats", d["model_stats"])
PY
