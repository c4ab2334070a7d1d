#!/bin/bash
# Session 5: parity (full suite), ncu --set full of the rebuilt item kernel at reduced corpus sizes, bench lines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu5.log
grep -E "passed|failed|rc=|Error|assert " gpurun_out/pytest_gpu5.log | head -40
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:k_sgns_items -s 1 -c 1 -f -o gpurun_out/sgns5_tract24 \
    python scripts/prof_path.py tract24 2000000 > gpurun_out/ncu5_sgns_tract24.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_sgns_items -s 1 -c 1 -f -o gpurun_out/sgns5_synth \
    python scripts/prof_path.py synth 100000 300000 > gpurun_out/ncu5_sgns_synth.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench5_tract24.json 2> gpurun_out/bench5_tract24.err
timeout 600 python bench.py --workload synth100k --steps 2 --warmup 3 > gpurun_out/bench5_synth_n1.json 2> gpurun_out/bench5_synth_n1.err
tail -n 3 gpurun_out/ncu5_*.log gpurun_out/bench5_*.err
python - <<'PY'
import json
for f in ("bench5_tract24","bench5_synth_n1"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        for k in ("walk","sgns"):
            st=d["stages"][k]; print(f, k, "value %.4g"%st["value"], "e2e %.4g"%st["e2e"]["value"], "kernel_ms %.3f"%st["kernel_ms"], "frac %.3f"%st["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
