#!/bin/bash
# Round 2, GPU session 6 (1 GPU): kernel G (a block per sentence) -- arithmetic test, then agreement with the oracle and
# throughput at the full bench size against the number of sentences in flight; CA with kernels F / G.
mkdir -p gpurun_out
echo "== pytest sgns"; timeout 900 python -m pytest tests/test_sgns_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -8
echo "== kernel G at full size"
timeout 1500 python scripts/fullsize_staleness.py 0,444,296,222,148,74 2052 2>&1 | tail -8
echo "== CA 1M walks: tp kernel vs kernel F vs kernel G (quality)"
timeout 900 python scripts/sgns_ab.py ca 1000000 --quality --variants tp:0,sentF:2176,blockG:2180 --conc 0 --tag r2s6_ca 2>&1 | tail -5
timeout 600 python scripts/sgns_ab.py ca 1000000 --quality --variants blockG:2180 --conc 37,74,148,296 --tag r2s6_ca_conc 2>&1 | tail -4
