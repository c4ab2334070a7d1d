#!/bin/bash
# Profiling session (one GPU): launch list of the bench command + ncu --set full of the hot kernels.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# tract (L2-resident) : walk + sgns kernels, second repetition (-s 1)
$NCU --set full --import-source on -k regex:k_walk_alias -s 1 -c 1 -f -o gpurun_out/walk_tract \
    python scripts/prof_path.py tract 2000000 > gpurun_out/ncu_walk_tract.log 2>&1
$NCU --set full --import-source on -k regex:k_sgns -s 1 -c 1 -f -o gpurun_out/sgns_tract \
    python scripts/prof_path.py tract 500000 > gpurun_out/ncu_sgns_tract.log 2>&1
# synthetic power-law graph (HBM-resident): 100K regions
$NCU --set full --import-source on -k regex:k_walk_alias -s 1 -c 1 -f -o gpurun_out/walk_synth \
    python scripts/prof_path.py synth 100000 4000000 > gpurun_out/ncu_walk_synth.log 2>&1
$NCU --set full --import-source on -k regex:k_sgns -s 1 -c 1 -f -o gpurun_out/sgns_synth \
    python scripts/prof_path.py synth 100000 500000 > gpurun_out/ncu_sgns_synth.log 2>&1
# un-profiled timing of the synthetic config
python scripts/prof_path.py synth 100000 4000000 > gpurun_out/prof_synth100k.log 2>&1
tail -2 gpurun_out/*.log
