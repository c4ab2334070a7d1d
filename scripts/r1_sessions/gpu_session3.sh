#!/bin/bash
# Session 3: full GPU test suite, bench lines, launch list of the bench command, ncu --set full at bench sizes.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi3.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench3_tract24.json 2> gpurun_out/bench3_tract24.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench3_reference.json 2> gpurun_out/bench3_reference.err
timeout 900 python bench.py --workload synth100k --steps 2 --warmup 3 > gpurun_out/bench3_synth_n1.json 2> gpurun_out/bench3_synth_n1.err
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches3_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_under_ncu.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_sgns -s 1 -c 1 -f -o gpurun_out/sgns3_tract24 \
    python scripts/prof_path.py tract24 15000000 > gpurun_out/ncu3_sgns_tract24.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_walk_alias -s 1 -c 1 -f -o gpurun_out/walk3_tract24 \
    python scripts/prof_path.py tract24 15000000 > gpurun_out/ncu3_walk_tract24.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_sgns -s 1 -c 1 -f -o gpurun_out/sgns3_synth \
    python scripts/prof_path.py synth 100000 4000000 > gpurun_out/ncu3_sgns_synth.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_walk_alias -s 1 -c 1 -f -o gpurun_out/walk3_synth \
    python scripts/prof_path.py synth 100000 4000000 > gpurun_out/ncu3_walk_synth.log 2>&1
tail -3 gpurun_out/pytest_gpu3.log gpurun_out/*3*.err gpurun_out/ncu3*.log
