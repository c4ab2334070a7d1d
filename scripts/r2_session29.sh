#!/bin/bash
# Round 2, GPU session 29 (1 GPU): the automatic schedule on wide rows (synthetic 100K regions, D = 128): agreement with a
# near-sequential run against the hub-bounded schedule it replaces, and the synthetic stage pair of the bench.
mkdir -p gpurun_out
timeout 1200 python scripts/synth_schedule_check.py 1000000 148 2>&1 | grep -v Warning | tail -6
echo "== bench --dp-only (N = 1: the synthetic stage pair)"
timeout 900 python bench.py --dp-only 2> gpurun_out/r2s29_synth.err | tail -1 > gpurun_out/r2s29_synth.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s29_synth.json")); s = d["sgns"]
print("walk %.4g steps/s; sgns %.4g pairs/s kernel %s kernel_ms %.1f call_ms %.1f frac %.3f" % (d["walk"]["value"], s["value"], s["kernel"], s["kernel_ms"], s["call_ms"], s["roofline"]["frac"]))
PY
