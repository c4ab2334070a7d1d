#!/bin/bash
# Session 11: quality sweep (stage-2 downstream parity vs the oracle under several schedules), DRAM traffic of the
# HBM-resident 1M-region walk, full GPU suite + smoke + bench with the final library.
mkdir -p gpurun_out
timeout 600 python scripts/quality_sweep.py tract 1000000 > gpurun_out/quality_tract.log 2>&1; tail -3 gpurun_out/quality_tract.log | cut -c1-400
timeout 600 python scripts/quality_sweep.py CA 100000 > gpurun_out/quality_CA.log 2>&1; tail -3 gpurun_out/quality_CA.log | cut -c1-400
NCU="ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct"
timeout 600 $NCU -k regex:k_walk_alias -s 1 -c 1 --csv --log-file gpurun_out/traffic11_walk_1m.csv python scripts/config4_1m.py --regions 1000000 --walks 32000000 --sgns-walks 1000 --out gpurun_out/config4_1m_ncu.json > gpurun_out/traffic11_walk_1m.log 2>&1
tail -n 6 gpurun_out/traffic11_walk_1m.csv | cut -d, -f5,13-15
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu11.log; tail -3 gpurun_out/pytest_gpu11.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench11_tract24.json 2> gpurun_out/bench11_tract24.err; tail -c 400 gpurun_out/bench11_tract24.err; cut -c1-600 gpurun_out/bench11_tract24.json
