#!/bin/bash
# Session 4: parity of the rebuilt item kernel + A/B timing (previous kernel / v2 at 4 and 3 blocks per SM) + dram traffic.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
tail -15 gpurun_out/pytest_gpu4.log
for dbg in 2 0 4; do
  echo "== tract24 4M walks DGE_SGNS_DEBUG=$dbg"
  DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py tract24 4000000 2>&1 | tail -1
done
for dbg in 2 0 4; do
  echo "== tract8 4M walks DGE_SGNS_DEBUG=$dbg"
  DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py tract 4000000 2>&1 | tail -1
done
for dbg in 2 0 4; do
  echo "== synth100k 1M walks D=128 DGE_SGNS_DEBUG=$dbg"
  DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py synth 100000 1000000 2>&1 | tail -1
done
for dbg in 2 0; do
  echo "== synth100k 1M walks D=64 DGE_SGNS_DEBUG=$dbg"
  DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py synth 100000 1000000 64 2>&1 | tail -1
done
NCU="ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors.sum,lts__t_sectors_op_red.sum"
timeout 600 $NCU -k regex:k_sgns_items -s 1 -c 1 --csv --log-file gpurun_out/traffic4_sgns_tract24.csv python scripts/prof_path.py tract24 15000000 > gpurun_out/traffic4_sgns_tract24.log 2>&1
timeout 600 $NCU -k regex:k_walk_alias -s 1 -c 1 --csv --log-file gpurun_out/traffic4_walk_tract24.csv python scripts/prof_path.py tract24 15000000 > gpurun_out/traffic4_walk_tract24.log 2>&1
timeout 600 $NCU -k regex:k_sgns_items -s 1 -c 1 --csv --log-file gpurun_out/traffic4_sgns_synth.csv python scripts/prof_path.py synth 100000 4000000 > gpurun_out/traffic4_sgns_synth.log 2>&1
tail -2 gpurun_out/traffic4_*.csv
