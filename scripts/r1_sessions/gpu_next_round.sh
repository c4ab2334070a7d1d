#!/bin/bash
# First GPU session of the next round (gpurun --gpus 2): what round 1 could not verify once its GPU budget ended.
#  1. the 2-rank data-parallel skip-gram test with the per-row contributor average (tests/test_comm_gpu.py),
#  2. the exchange diagnosis under every combine rule (DGE_SGNS_COMBINE = 0 sum, 1 mean, 2 contributors, 3 sqrt),
#  3. bench.py at N=2 on the synthetic workload (data-parallel path) with table health from dge_model_stats.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_comm_gpu.py -m gpu -q --tb=short 2>&1 | tail -8
for c in 2 3 0; do
  echo "== dp_diagnose DGE_SGNS_COMBINE=$c"
  DGE_SGNS_COMBINE=$c timeout 300 python scripts/dp_diagnose.py 2>&1 | grep -v Warning | tail -12
  cp gpurun_out/dp_diagnose.json gpurun_out/dp_diagnose_combine$c.json
done
timeout 900 $TR --master-port 29621 scripts/config4_1m.py --regions 100000 --walks 4000000 --sgns-walks 2000000 --out gpurun_out/config4_100k_n2.json 2>&1 | grep -v Warning | tail -8
