#!/bin/bash
# Session 15 (one GPU): full GPU suite incl. the full-size property tests, smoke, both bench arms, CA workload,
# e2e breakdown, then the ncu launch list of the bench command and --set full captures of the two hot kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu15.log; tail -14 gpurun_out/pytest_gpu15.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench15_reference.json 2> gpurun_out/bench15_reference.err; cut -c1-300 gpurun_out/bench15_reference.json
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench15_tract24.json 2> gpurun_out/bench15_tract24.err; tail -c 400 gpurun_out/bench15_tract24.err; cut -c1-400 gpurun_out/bench15_tract24.json
timeout 600 python bench.py --workload ca --steps 3 --warmup 3 > gpurun_out/bench15_ca.json 2> gpurun_out/bench15_ca.err; tail -c 400 gpurun_out/bench15_ca.err; cut -c1-400 gpurun_out/bench15_ca.json
timeout 300 python scripts/e2e_breakdown.py tract24 4 > gpurun_out/e2e_breakdown15_tract24.json 2>&1; tail -40 gpurun_out/e2e_breakdown15_tract24.json
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches15_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench15_under_ncu.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:k_sgns_items -s 1 -c 1 -f -o gpurun_out/sgns15_tract24 \
    python scripts/prof_path.py tract24 500000 > gpurun_out/ncu15_sgns.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:k_walk_alias -s 1 -c 1 -f -o gpurun_out/walk15_tract24 \
    python scripts/prof_path.py tract24 4000000 > gpurun_out/ncu15_walk.log 2>&1
ls -la gpurun_out | tail -15
