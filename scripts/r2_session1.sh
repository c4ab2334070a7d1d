#!/bin/bash
# Round 2, GPU session 1 (1 GPU): the suite after the refactor (every gated test un-gated), the two micro-benchmarks
# (random-sector walk ceiling; L2 reductions incl. TMA bulk reductions), skip-gram kernel A/B on tract x 24 with the
# downstream metric, the CA staleness sweep, and the e2e > device anomaly of round 1 (clock sampler on / off).
mkdir -p gpurun_out
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15
echo "== red_microbench"; timeout 120 scripts/bin/red_microbench 2>&1 | tee gpurun_out/r2s1_red_microbench.txt
echo "== walk_microbench"; timeout 180 scripts/bin/walk_microbench 24 2>&1 | tee gpurun_out/r2s1_walk_microbench.txt
echo "== sgns A/B tract24 2M walks"
timeout 900 python scripts/sgns_ab.py tract24 2000000 --quality --variants v2:0,staged:256,staged6:24832,staged7:28928,plain:512,noupd:1,staged_noupd:257 --tag r2s1_tract24 2>&1 | tail -12
echo "== sgns A/B tract24 15M walks (bench size), throughput only"
timeout 600 python scripts/sgns_ab.py tract24 15000000 --variants v2:0,staged:256,staged6:24832,staged7:28928 --tag r2s1_tract24_full 2>&1 | tail -5
echo "== sgns A/B synth 100K D=128 2M walks"
timeout 600 python scripts/sgns_ab.py synth 100000 2000000 --dim 128 --variants v2:0,staged:256 --tag r2s1_synth128 2>&1 | tail -3
echo "== CA staleness sweep, 1M walks"
timeout 900 python scripts/sgns_ab.py ca 1000000 --quality --conc 0,206,412,824,1648,3296 --variants tp:0 --tag r2s1_ca 2>&1 | tail -10
echo "== bench: clock sampler on / off"
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-synth > gpurun_out/r2s1_bench_sampler_on.json 2> gpurun_out/r2s1_bench_on.err; python - <<'PY'
import json
for f in ("on",):
    d = json.load(open("gpurun_out/r2s1_bench_sampler_%s.json" % f))
    s = d["stages"]["sgns"]
    print(f, "device", s["value"], "kernel_ms", s["kernel_ms"], "e2e", s["e2e"]["value"])
PY
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-synth --no-clock-sampler > gpurun_out/r2s1_bench_sampler_off.json 2> gpurun_out/r2s1_bench_off.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s1_bench_sampler_off.json"))
s = d["stages"]["sgns"]
print("off", "device", s["value"], "kernel_ms", s["kernel_ms"], "e2e", s["e2e"]["value"])
PY
