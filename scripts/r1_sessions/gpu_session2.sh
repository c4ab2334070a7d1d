#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench2_tract24.json 2> gpurun_out/bench2_tract24.err
timeout 900 python bench.py --workload synth100k --steps 2 --warmup 3 > gpurun_out/bench2_synth_n1.json 2> gpurun_out/bench2_synth_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload synth100k --steps 2 --warmup 3 > gpurun_out/bench2_synth_n2.json 2> gpurun_out/bench2_synth_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench2_tract24_n2.json 2> gpurun_out/bench2_tract24_n2.err
tail -5 gpurun_out/pytest_gpu2.log
for f in bench2_tract24 bench2_synth_n1 bench2_synth_n2 bench2_tract24_n2; do echo "== $f"; tail -c 600 gpurun_out/$f.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    for k in ("walk","sgns"):
        st=d["stages"][k]; print(k, "value %.4g"%st["value"], "e2e %.4g"%st["e2e"]["value"], "kernel_ms %.3f"%st["kernel_ms"], "ms_per_step %.2f"%st["ms_per_step"], "frac %.3f"%st["roofline"]["frac"], {x:st.get(x) for x in ("sync_rounds","sync_ms","groups_in_flight")})
    print("n_gpus", d["n_gpus"], "launches", d["gpu_launches"], d["clocks"])
except Exception as e: print("ERR", e)
PY
done
