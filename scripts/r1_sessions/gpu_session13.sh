#!/bin/bash
# Session 13: parity of all item kernels in strict item order; does row skew (hot rows) or the reductions' sheer count
# bound the tract x 24 skip-gram?  (uniform corpus of the same shape vs the walk corpus)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sgns_gpu.py -m gpu -q 2>&1 | tail -8
echo "== uniform corpus V=19224 L=24 D=20, 2M sentences"
timeout 300 python scripts/uniform_corpus.py 19224 24 2000000 20 2>&1 | tail -1
echo "== uniform corpus, no reductions"
DGE_SGNS_DEBUG=1 timeout 300 python scripts/uniform_corpus.py 19224 24 2000000 20 2>&1 | tail -1
echo "== uniform corpus V=192240 (10x rows)"
timeout 300 python scripts/uniform_corpus.py 192240 24 2000000 20 2>&1 | tail -1
echo "== CA shape: V=1848 L=24 D=8"
timeout 300 python scripts/uniform_corpus.py 1848 24 1000000 8 2>&1 | tail -1
DGE_SGNS_DEBUG=32 timeout 300 python scripts/uniform_corpus.py 1848 24 1000000 8 2>&1 | tail -1
echo "== D=16 synth default"
timeout 300 python scripts/prof_path.py synth 100000 1000000 16 2>&1 | tail -1
