#!/bin/bash
# Session 16: target-parallel skip-gram kernel for small vocabularies with narrow rows (CA: V = 1 848, D = 8):
# parity, A/B timing against the 8-lane kernel (DGE_SGNS_DEBUG=128), downstream quality, one ncu capture.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sgns_gpu.py tests/test_pipeline_gpu.py -m gpu -q 2>&1 | tail -6
echo "== ca 2M walks D=8, target-parallel (auto)"; timeout 300 python scripts/prof_path.py ca 2000000 2>&1 | tail -1
echo "== ca 2M walks D=8, 8-lane kernel";         DGE_SGNS_DEBUG=128 timeout 300 python scripts/prof_path.py ca 2000000 2>&1 | tail -1
echo "== ca 2M walks D=8, target-parallel, no reductions"; DGE_SGNS_DEBUG=1 timeout 300 python scripts/prof_path.py ca 2000000 2>&1 | tail -1
echo "== ca 2M walks D=2, target-parallel (auto)"; timeout 300 python scripts/prof_path.py ca 2000000 0 2 2>&1 | tail -1
echo "== ca 2M walks D=2, 8-lane kernel";         DGE_SGNS_DEBUG=128 timeout 300 python scripts/prof_path.py ca 2000000 0 2 2>&1 | tail -1
echo "== ca 2M walks D=16, target-parallel (auto)"; timeout 300 python scripts/prof_path.py ca 2000000 0 16 2>&1 | tail -1
echo "== ca 2M walks D=16, 8-lane kernel";         DGE_SGNS_DEBUG=128 timeout 300 python scripts/prof_path.py ca 2000000 0 16 2>&1 | tail -1
timeout 600 python scripts/quality_sweep.py CA 100000 > gpurun_out/quality_CA16.log 2>&1; grep -o '"name": "[a-z_0-9]*", "seconds": [0-9.e-]*, "pairs": [0-9]*, "mpairs_per_s": [0-9.]*, "metric": {[^}]*}' gpurun_out/quality_CA16.log | cut -c1-260
cp gpurun_out/quality_CA.json gpurun_out/quality_CA16.json
timeout 400 ncu --clock-control none --set full --import-source on -k regex:k_sgns_items -s 1 -c 1 -f -o gpurun_out/sgns16_ca \
    python scripts/prof_path.py ca 500000 > gpurun_out/ncu16_sgns_ca.log 2>&1; tail -2 gpurun_out/ncu16_sgns_ca.log
