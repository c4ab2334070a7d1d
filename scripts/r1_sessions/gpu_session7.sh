#!/bin/bash
# Session 7: L2 reduction / load microbenchmark, config-5 SGNS sweep at N=1, bench lines with the default item-kernel build.
mkdir -p gpurun_out
timeout 300 scripts/bin/red_microbench > gpurun_out/red_microbench.txt 2>&1; cat gpurun_out/red_microbench.txt
timeout 900 python scripts/sgns_sweep.py --walks 2000000 --out gpurun_out/sgns_sweep_n1.json 2>&1 | tail -32
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench7_tract24.json 2> gpurun_out/bench7_tract24.err
tail -c 300 gpurun_out/bench7_tract24.err
python - <<'PY'
import json
for f in ("bench7_tract24",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        for k in ("walk","sgns"):
            st=d["stages"][k]; print(f, k, "value %.4g"%st["value"], "e2e %.4g"%st["e2e"]["value"], "kernel_ms %.3f"%st["kernel_ms"], "frac %.3f"%st["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
