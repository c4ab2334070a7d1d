#!/bin/bash
# Round 2, GPU session 2 (gpurun --gpus 2): the data-parallel skip-gram on two GPUs -- the 2-rank parity test with both
# transports (peer-memory kernel over NVLink; NCCL), then bench.py at N = 2 (tract replicas + data-parallel synthetic run
# with neighbourhood agreement against one GPU), then a sweep of rounds / transport / combine rule on the synthetic run.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== nvidia-smi"; nvidia-smi --query-gpu=index,name --format=csv,noheader; nvidia-smi topo -m 2>/dev/null | head -6
echo "== pytest comm"; timeout 900 python -m pytest tests/test_comm_gpu.py -m gpu -q --tb=short -s 2>&1 | grep -v Warning | tail -15
echo "== bench N=2 data-parallel only, default (auto rounds, auto transport, aligned rule)"
timeout 900 $TR --master-port 29621 bench.py --gpus 2 --dp-only 2> gpurun_out/r2s2_dp_default.err | tail -1 > gpurun_out/r2s2_dp_default.json; cut -c1-1800 gpurun_out/r2s2_dp_default.json; tail -3 gpurun_out/r2s2_dp_default.err
for cfg in "8 1 0" "32 1 0" "16 2 0" "16 1 5"; do
  set -- $cfg
  echo "== dp sweep rounds=$1 transport=$2 combine=$3"
  timeout 600 $TR --master-port 29622 bench.py --gpus 2 --dp-only --sync-rounds $1 --transport $2 --combine $3 2> gpurun_out/r2s2_dp_$1_$2_$3.err | tail -1 > gpurun_out/r2s2_dp_$1_$2_$3.json
  python - "$1" "$2" "$3" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/r2s2_dp_%s_%s_%s.json" % tuple(sys.argv[1:4])))
    s = d["sgns"]
    print("pairs/s %.4g  kernel_ms %.1f  sync_ms %.1f  rounds %s  transport %s  single %s  agreement %s  stats %s" % (
        s["value"], s["kernel_ms"], s["sync_ms"], s["sync_rounds"], s["transport"], d.get("single_gpu_reference"), d.get("agreement"), d.get("model_stats")))
except Exception as e:
    print("failed:", e)
PY
done
echo "== full bench N=2"
timeout 1200 $TR --master-port 29623 bench.py --gpus 2 --steps 3 --warmup 2 2> gpurun_out/r2s2_bench_n2.err | tail -1 > gpurun_out/r2s2_bench_n2.json; cut -c1-3000 gpurun_out/r2s2_bench_n2.json; tail -3 gpurun_out/r2s2_bench_n2.err
