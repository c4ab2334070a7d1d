#!/bin/bash
# Round 2, GPU session 30 (gpurun --gpus 2): the 2-rank parity tests and the N = 2 bench line with the round's final kernels
# (kernel F's sentence loop was restructured for the counter hand-out; the data-parallel path uses its strided form).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== pytest comm"; timeout 900 python -m pytest tests/test_comm_gpu.py -m gpu -q --tb=short -s 2>&1 | grep -v Warning | tail -12
echo "== bench N=2"
timeout 900 $TR --master-port 29653 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/r2s30_bench_n2.err | tail -1 > gpurun_out/r2s30_bench_n2.json; tail -2 gpurun_out/r2s30_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s30_bench_n2.json"))
print("value %.4g steps/s, ms_per_step %.1f, e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
dp = d.get("data_parallel") or {}
s = dp.get("sgns", {})
print("dp: agg pairs/s %.4g (kernel-only %.4g) call_ms %.0f sync_ms %.1f rounds %s %s" % (s.get("value", 0), s.get("kernel_pairs_per_s", 0), s.get("call_ms", 0), s.get("sync_ms", 0), s.get("sync_rounds"), s.get("transport")))
print("single", dp.get("single_gpu_reference")); print("agreement", dp.get("agreement")); print("stats", dp.get("model_stats"))
PY
