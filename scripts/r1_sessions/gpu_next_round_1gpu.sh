#!/bin/bash
# Next round, one GPU: A/B of the experimental skip-gram kernel C' (rows of a unit staged in shared memory by cp.async,
# DGE_SGNS_DEBUG=256) against kernel C -- parity first, then timing on the bench workload and the synthetic one.
mkdir -p gpurun_out
DGE_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_sgns_gpu.py tests/test_pipeline_gpu.py tests/test_walk_gpu.py -m gpu -q -k "item_kernel or deepwalk_main or tokens_u16" --tb=short 2>&1 | tail -6
for dbg in 0 256; do
  echo "== tract24 4M walks DGE_SGNS_DEBUG=$dbg"; DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py tract24 4000000 2>&1 | tail -1
  echo "== tract8 4M walks DGE_SGNS_DEBUG=$dbg";  DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py tract 4000000 2>&1 | tail -1
  echo "== synth D=128 1M walks DGE_SGNS_DEBUG=$dbg"; DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py synth 100000 1000000 2>&1 | tail -1
  echo "== synth D=32 1M walks DGE_SGNS_DEBUG=$dbg"; DGE_SGNS_DEBUG=$dbg timeout 300 python scripts/prof_path.py synth 100000 1000000 32 2>&1 | tail -1
done
# 16-bit token download in the host-buffer walk number (dge_corpus_tokens_u16), against the int32 download
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --tokens16 > gpurun_out/bench_tokens16.json 2> gpurun_out/bench_tokens16.err; cut -c1-700 gpurun_out/bench_tokens16.json
