"""SGNS throughput on a uniform random corpus (no hot rows):  python scripts/uniform_corpus.py V L n dim [conc]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embedding_b200 import abi
V, L, n, dim = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
conc = int(sys.argv[5]) if len(sys.argv) > 5 else 0
ctx = abi.Context(0)
rng = np.random.default_rng(0)
tok = rng.integers(0, V, size=(n, L)).astype(np.int32)
c = abi.Corpus.from_tokens(ctx, tok, V)
for rep in range(2):
    m = abi.Model.train(ctx, [c], abi.sgns_params(dim=dim, window=L, seed=1, concurrency=conc))
print("uniform V=%d L=%d dim=%d sgns ms %.3f pairs %d groups %d Mpairs/s %.1f" % (V, L, dim, ctx.phase_ms("sgns"), m.pairs, ctx.phase_ms("sgns_groups"), m.pairs / ctx.phase_ms("sgns") / 1e3))
