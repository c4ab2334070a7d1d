#!/bin/bash
# Round 2, GPU session 5 (1 GPU): kernel F (sentence-resident) -- arithmetic parity, then agreement with the oracle and
# throughput at the full bench size against the number of sentences in flight, 640-thread and 128-thread blocks; D=128 A/B.
mkdir -p gpurun_out
echo "== pytest sgns"; timeout 900 python -m pytest tests/test_sgns_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -8
echo "== kernel F, 640-thread blocks"
timeout 1500 python scripts/fullsize_staleness.py 0,1480,740,256,64,20 2048 2>&1 | tail -8
echo "== kernel F, 128-thread blocks"
timeout 900 python scripts/fullsize_staleness.py 0,740 2064 2>&1 | tail -3
echo "== synth D=128 2M walks: item kernel vs kernel F"
timeout 600 python scripts/sgns_ab.py synth 100000 2000000 --dim 128 --variants v2:0,sent:2048 --tag r2s5_synth128 2>&1 | tail -3
echo "== CA 1M walks: tp kernel vs kernel F (quality)"
timeout 900 python scripts/sgns_ab.py ca 1000000 --quality --variants tp:0,sent:2048 --tag r2s5_ca 2>&1 | tail -5
