#!/bin/bash
# Round 2, GPU session 31 (1 GPU): final verification of the shipped build -- whole suite, smoke, bench at N = 1 (both
# arms), ncu evidence of the timed build (launch list, DRAM traffic at the bench sizes, --set full of kernel G and F).
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -s 2>&1 | grep -v Warning | tail -20
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench N=1 (b200 arm)"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s31_bench_n1.json 2> gpurun_out/r2s31_bench_n1.err; cut -c1-500 gpurun_out/r2s31_bench_n1.json; tail -3 gpurun_out/r2s31_bench_n1.err
echo "== bench N=1 (reference arm)"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s31_bench_ref.json 2> gpurun_out/r2s31_bench_ref.err; cut -c1-300 gpurun_out/r2s31_bench_ref.json
NCU="ncu --clock-control none"
echo "== ncu launch list of the bench command"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2s31_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-synth > gpurun_out/r2s31_bench_under_ncu.log 2>&1; tail -1 gpurun_out/r2s31_bench_under_ncu.log | cut -c1-200
echo "== ncu DRAM traffic at the bench sizes"
timeout 900 $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:'k_sgns|k_walk_alias' -c 4 --csv --log-file gpurun_out/r2s31_traffic_tract24.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-synth --no-e2e > /dev/null 2>&1; tail -4 gpurun_out/r2s31_traffic_tract24.csv | cut -c1-260
timeout 900 $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:'k_sgns|k_walk_alias' -c 2 --csv --log-file gpurun_out/r2s31_traffic_synth100k.csv python bench.py --workload synth100k --steps 1 --warmup 0 --no-cpu-baseline --no-synth --no-e2e > /dev/null 2>&1; tail -3 gpurun_out/r2s31_traffic_synth100k.csv | cut -c1-260
echo "== ncu --set full: kernel F on synth100k (1M walks, wide rows); the narrow-row capture is session 27's"
timeout 600 $NCU --set full --import-source on -k regex:k_sgns_sent -s 1 -c 1 -f -o gpurun_out/r2s31_sgns_sent_synth100k python scripts/prof_path.py synth 100000 1000000 > gpurun_out/r2s31_ncu_f.log 2>&1; tail -1 gpurun_out/r2s31_ncu_f.log | cut -c1-200
ls -la gpurun_out/r2s31*.ncu-rep
