#!/bin/bash
# Round 2, GPU session 37 (1 GPU): final tree (table placement probe with 24 candidates): whole GPU suite, smoke, short bench.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v Warning | tail -6
echo "== smoke"; timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1
echo "== bench (short: no synthetic stage, no CPU baseline)"
timeout 200 python bench.py --steps 5 --warmup 3 --no-synth --no-cpu-baseline > gpurun_out/r2s37_bench_n1_short.json 2> gpurun_out/r2s37_bench.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s37_bench_n1_short.json"))
print("value %.4g steps/s ms_per_step %.1f | e2e %.4g (%.1f ms) | sgns %.4g pairs/s kernel_ms %.1f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["stages"]["sgns"]["value"], d["stages"]["sgns"]["kernel_ms"], d["roofline"]["frac"]))
PY
