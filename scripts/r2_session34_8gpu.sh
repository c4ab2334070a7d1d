#!/bin/bash
# Round 2, GPU session 34 (gpurun --gpus 8): the N = 8 bench line, as the driver's scaling run launches it, with the round's
# final kernels (counter + write-through schedule on the ranks of the data-parallel skip-gram).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29663 bench.py --gpus 8 --steps 3 --warmup 3 2> gpurun_out/r2s34_bench_n8.err | tail -1 > gpurun_out/r2s34_bench_n8.json; tail -2 gpurun_out/r2s34_bench_n8.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s34_bench_n8.json"))
print("value %.4g steps/s, ms_per_step %.1f, e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
dp = d.get("data_parallel") or {}
s = dp.get("sgns", {})
print("dp: agg pairs/s %.4g (kernel-only %.4g) call_ms %.0f sync_ms %.1f rounds %s %s" % (s.get("value", 0), s.get("kernel_pairs_per_s", 0), s.get("call_ms", 0), s.get("sync_ms", 0), s.get("sync_rounds"), s.get("transport")))
print("single", dp.get("single_gpu_reference")); print("agreement", dp.get("agreement")); print("stats", dp.get("model_stats"))
PY
