#!/bin/bash
# Round 2, GPU session 35 (1 GPU): the whole GPU suite and the smoke test on the round's final tree (sgns.cu split into
# headers, arena sizing, the automatic-schedule test).
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v Warning | tail -8
echo "== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
