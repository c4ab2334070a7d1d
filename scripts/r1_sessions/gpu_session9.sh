#!/bin/bash
# Session 9 (gpurun --gpus 2): multi-GPU row -- NCCL tests, bench at N=2 (tract24: walk ids sharded, SGNS replicas;
# synth100k: data-parallel SGNS with delta all-reduce), config 4 at N=2, one sweep cell at N=2.
mkdir -p gpurun_out
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_comm_gpu.py tests/test_walk_gpu.py -m gpu -q 2>&1 | tail -4
timeout 600 $TR --master-port 29601 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench9_tract24_n2.json 2> gpurun_out/bench9_tract24_n2.err
timeout 600 $TR --master-port 29602 bench.py --gpus 2 --workload synth100k --steps 2 --warmup 3 > gpurun_out/bench9_synth_n2.json 2> gpurun_out/bench9_synth_n2.err
timeout 900 $TR --master-port 29603 scripts/config4_1m.py --regions 1000000 --walks 32000000 --sgns-walks 8000000 --out gpurun_out/config4_1m_n2.json 2>&1 | grep -v Warning | tail -12
timeout 600 $TR --master-port 29604 scripts/sgns_sweep.py --walks 2000000 --dims 32,128 --negatives 5 --windows 10 --out gpurun_out/sgns_sweep_n2.json 2>&1 | grep "D=" 
python - <<'PY'
import json
for f in ("bench9_tract24_n2","bench9_synth_n2"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
        for k in ("walk","sgns"):
            st=d["stages"][k]; print(f, k, "value %.4g"%st["value"], "e2e %.4g"%st["e2e"]["value"], "kernel_ms %.3f"%st["kernel_ms"], {x:st.get(x) for x in ("sync_rounds","sync_ms")})
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
PY
