"""ctypes binding of the CPU ORACLE (oracle/libdge_oracle.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package `embedding_b200` never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdge_oracle.so")

SAMPLER_ALIAS, SAMPLER_CDF = 0, 1
RNG_PHILOX, RNG_JAVA_LCG = 0, 1
ALIAS_LITERAL, ALIAS_FAST = 0, 1


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("dge_oracle.c", "sgns_oracle.c", "dge_oracle.h", "sgns_oracle.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libdge_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


class SgnsParams(C.Structure):
    _fields_ = [("dim", C.c_int32), ("window", C.c_int32), ("negative", C.c_int32), ("min_count", C.c_int32),
                ("epochs", C.c_int32), ("threads", C.c_int32), ("use_hs", C.c_int32),
                ("neg_table_size", C.c_int32), ("exp_table_size", C.c_int32),
                ("lr", C.c_float), ("min_lr", C.c_float), ("seed", C.c_uint64)]


def sgns_params(dim=20, window=8, negative=5, min_count=2, epochs=1, threads=1, use_hs=0,
                neg_table_size=100000, exp_table_size=1000, lr=0.025, min_lr=1e-4, seed=1):
    return SgnsParams(dim, window, negative, min_count, epochs, threads, use_hs, neg_table_size,
                      exp_table_size, lr, min_lr, seed)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp, i32, i64, f64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_uint64
    pi32, pi64, pf64, pf32 = C.POINTER(i32), C.POINTER(i64), C.POINTER(f64), C.POINTER(C.c_float)
    L.ora_alias_literal.argtypes = [i32, pf64, f64, pf64, pi32]
    L.ora_alias_fast.argtypes = [i32, pf64, f64, pf64, pi32]
    L.ora_graph_build.restype = vp
    L.ora_graph_build.argtypes = [i32, i64, pi32, pi32, pf64, i32, pi32, pf64, pf64, C.c_int]
    L.ora_graph_free.argtypes = [vp]
    L.ora_graph_num_edges.restype = i64
    L.ora_graph_num_edges.argtypes = [vp]
    L.ora_graph_num_vertices.argtypes = [vp]
    L.ora_graph_num_sources.argtypes = [vp]
    L.ora_graph_tables.argtypes = [vp, pi64, pi32, pf64, pf64, pi32, pf64, pf64, pi32, pf64]
    L.ora_sample_next.argtypes = [vp, i32, f64]
    L.ora_sample_next_ov.argtypes = [vp, i32, f64]
    L.ora_sample_source.argtypes = [vp, f64, C.c_int]
    L.ora_walk.argtypes = [vp, i64, i64, i32, u64, C.c_int, C.c_int, pi32]
    L.ora_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
    L.ora_philox_uniform.restype = f64
    L.ora_philox_uniform.argtypes = [u64, u64, C.c_uint32]
    L.ora_java_random_doubles.argtypes = [i64, i32, pf64]
    L.ora_keep_nearest_k.argtypes = [i32, pf64, i32, pi32, pf64, pf64]
    L.ora_java8_stream_sum.restype = f64
    L.ora_java8_stream_sum.argtypes = [pf64, i64]
    L.ora_flow_ca.argtypes = [pi32, i32, i32, i32, i32, i32]
    L.ora_flow_tract.argtypes = [pi32, i32, i32, i32, i32, i32]
    L.ora_crosstime_edges.restype = i64
    L.ora_crosstime_edges.argtypes = [pi32, i32, pi32, i32, C.c_int, pi32, pi32, pi32, pf64, pi32, pi32, pi32,
                                      pi32, pi32]
    L.ora_write_seq.restype = i64
    L.ora_write_seq.argtypes = [pi32, i64, i32, pi32, pi32, C.c_char_p]
    # stage 2
    pp = C.POINTER(SgnsParams)
    L.ora_vocab_build.restype = vp
    L.ora_vocab_build.argtypes = [pi32, i64, i32, i32]
    L.ora_vocab_free.argtypes = [vp]
    L.ora_vocab_size.argtypes = [vp]
    L.ora_vocab_total_words.restype = i64
    L.ora_vocab_total_words.argtypes = [vp]
    L.ora_vocab_tables.argtypes = [vp, pi32, pi32, pi64]
    L.ora_neg_table.argtypes = [vp, i32, pi32]
    L.ora_init_syn0.argtypes = [i32, i32, u64, pf32]
    L.ora_sgns_train.restype = vp
    L.ora_sgns_train.argtypes = [pi32, i64, i32, i32, pp, pi64]
    L.ora_sgns_train_dp.restype = vp
    L.ora_sgns_train_dp.argtypes = [pi32, i64, i32, i32, pp, i32, i32, i32, pi64]
    L.ora_sgns_train_dp_tiered.restype = vp
    L.ora_sgns_train_dp_tiered.argtypes = [pi32, i64, i32, i32, pp, i32, i32, i32, i32, i32, pi64]
    L.ora_model_free.argtypes = [vp]
    L.ora_model_vocab_size.argtypes = [vp]
    L.ora_model_get.argtypes = [vp, pf32, pf32, pi32]
    L.ora_sgns_count_pairs.restype = i64
    L.ora_sgns_count_pairs.argtypes = [pi32, i64, i32, i32, pp]
    L.ora_sentence_rng.restype = u64
    L.ora_sentence_rng.argtypes = [u64, i32, i64]
    L.ora_alpha.restype = C.c_float
    L.ora_alpha.argtypes = [pp, i32, i64, i64]
    _lib = L
    return L


def alias_table(w, out_degree=None, mode=ALIAS_LITERAL):
    """Vertex.initiateAliasTable (LayeredGraph.java:54-82) on one weight list."""
    w = np.ascontiguousarray(w, dtype=np.float64)
    k = len(w)
    if out_degree is None:
        out_degree = 0.0
        for x in w:  # left-to-right, as Vertex.addOutEdge :46-49
            out_degree += float(x)
    prob = np.empty(k, np.float64)
    alias = np.empty(k, np.int32)
    fn = lib().ora_alias_literal if mode == ALIAS_LITERAL else lib().ora_alias_fast
    fn(k, _p(w, C.c_double), float(out_degree), _p(prob, C.c_double), _p(alias, C.c_int32))
    return prob, alias


class Graph:
    """LayeredGraph (LayeredGraph.java:142-281) over integer vertex ids."""

    def __init__(self, n_vertices, src, dst, w, sources, out_degree=None, source_weight_sum=None,
                 alias_mode=ALIAS_LITERAL):
        src = np.ascontiguousarray(src, np.int32)
        dst = np.ascontiguousarray(dst, np.int32)
        w = np.ascontiguousarray(w, np.float64)
        sources = np.ascontiguousarray(sources, np.int32)
        od = None if out_degree is None else np.ascontiguousarray(out_degree, np.float64)
        sws = None if source_weight_sum is None else np.array([source_weight_sum], np.float64)
        self._h = lib().ora_graph_build(n_vertices, len(src), _p(src, C.c_int32), _p(dst, C.c_int32),
                                        _p(w, C.c_double), len(sources), _p(sources, C.c_int32),
                                        _p(od, C.c_double), _p(sws, C.c_double), alias_mode)
        self.nv, self.ne, self.ns = n_vertices, len(src), len(sources)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_graph_free(self._h)
            self._h = None

    def tables(self):
        t = dict(row_ptr=np.empty(self.nv + 1, np.int64), col=np.empty(self.ne, np.int32),
                 w=np.empty(self.ne, np.float64), prob=np.empty(self.ne, np.float64),
                 alias=np.empty(self.ne, np.int32), out_degree=np.empty(self.nv, np.float64),
                 src_prob=np.empty(self.ns, np.float64), src_alias=np.empty(self.ns, np.int32))
        sws = np.empty(1, np.float64)
        lib().ora_graph_tables(self._h, _p(t["row_ptr"], C.c_int64), _p(t["col"], C.c_int32),
                               _p(t["w"], C.c_double), _p(t["prob"], C.c_double), _p(t["alias"], C.c_int32),
                               _p(t["out_degree"], C.c_double), _p(t["src_prob"], C.c_double),
                               _p(t["src_alias"], C.c_int32), _p(sws, C.c_double))
        t["source_weight_sum"] = float(sws[0])
        return t

    def sample_next(self, v, x):
        return lib().ora_sample_next(self._h, v, x)

    def sample_next_ov(self, v, x):
        return lib().ora_sample_next_ov(self._h, v, x)

    def sample_source(self, x, sampler=SAMPLER_ALIAS):
        return lib().ora_sample_source(self._h, x, sampler)

    def walk(self, n_walks, L, seed, sampler=SAMPLER_ALIAS, rng=RNG_PHILOX, first_walk_id=0):
        out = np.empty((n_walks, L), np.int32)
        lib().ora_walk(self._h, n_walks, first_walk_id, L, seed, sampler, rng, _p(out, C.c_int32))
        return out


def write_seq(tokens, layer, region, path):
    """String.join(" ", seq) + "\\n" per walk, one thread (CrossTimeGraph.java:134-141); returns the bytes written."""
    tokens = np.ascontiguousarray(tokens, np.int32)
    layer, region = np.ascontiguousarray(layer, np.int32), np.ascontiguousarray(region, np.int32)
    n = lib().ora_write_seq(_p(tokens, C.c_int32), tokens.shape[0], tokens.shape[1], _p(layer, C.c_int32), _p(region, C.c_int32),
                            os.fsencode(path))
    if n < 0:
        raise IOError("ora_write_seq failed: %s" % path)
    return n


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().ora_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def philox_uniform(seed, walk_id, draw):
    return lib().ora_philox_uniform(seed, walk_id, draw)


def java_random_doubles(seed, n):
    out = np.empty(n, np.float64)
    lib().ora_java_random_doubles(seed, n, _p(out, C.c_double))
    return out


def keep_nearest_k(w, k):
    w = np.ascontiguousarray(w, np.float64)
    n = w.shape[0]
    col = np.empty((n, k), np.int32)
    wk = np.empty((n, k), np.float64)
    od = np.empty(n, np.float64)
    lib().ora_keep_nearest_k(n, _p(w, C.c_double), k, _p(col, C.c_int32), _p(wk, C.c_double), _p(od, C.c_double))
    return col, wk, od


def java8_stream_sum(v):
    v = np.ascontiguousarray(v, np.float64)
    return lib().ora_java8_stream_sum(_p(v, C.c_double), len(v))


def flow_ca(F, src, dst, lo, hi):
    """CommunityArea.getFlowTo(dst, lo, hi) CommunityAreas.java:240-245 over the dense F[n, 24, n]."""
    F = np.ascontiguousarray(F, np.int32)
    return int(lib().ora_flow_ca(_p(F, C.c_int32), F.shape[0], int(src), int(dst), int(lo), int(hi)))


def flow_tract(F, src, dst, lo, hi):
    """Tract.getFlowTo(dst, lo, hi) Tracts.java:477-482."""
    F = np.ascontiguousarray(F, np.int32)
    return int(lib().ora_flow_tract(_p(F, C.c_int32), F.shape[0], int(src), int(dst), int(lo), int(hi)))


def crosstime_edges(F, order, L, mode, intervals=None):
    """CrossTimeGraph.constructGraph_CA(int[]) (mode 0) / constructGraph_tract() (mode 1)."""
    F = np.ascontiguousarray(F, np.int32)
    n = F.shape[0]
    assert F.shape == (n, 24, n)
    order = np.ascontiguousarray(order, np.int32)
    iv = None if intervals is None else np.ascontiguousarray(intervals, np.int32)
    cap_e = int(L) * n * n
    src = np.empty(cap_e, np.int32)
    dst = np.empty(cap_e, np.int32)
    w = np.empty(cap_e, np.float64)
    vl = np.empty(L * n, np.int32)
    vr = np.empty(L * n, np.int32)
    sources = np.empty(n, np.int32)
    nv = C.c_int32()
    ns = C.c_int32()
    ne = lib().ora_crosstime_edges(_p(F, C.c_int32), n, _p(order, C.c_int32), L, mode, _p(iv, C.c_int32),
                                   _p(src, C.c_int32), _p(dst, C.c_int32), _p(w, C.c_double),
                                   _p(vl, C.c_int32), _p(vr, C.c_int32), C.byref(nv), _p(sources, C.c_int32),
                                   C.byref(ns))
    return dict(src=src[:ne].copy(), dst=dst[:ne].copy(), w=w[:ne].copy(), v_layer=vl[:nv.value].copy(),
                v_region=vr[:nv.value].copy(), sources=sources[:ns.value].copy(), n_vertices=nv.value)


# ------------------------------------------------------------------ stage 2

def vocab(tokens, n_ids, min_count):
    tokens = np.ascontiguousarray(tokens, np.int32)
    h = lib().ora_vocab_build(_p(tokens, C.c_int32), tokens.size, n_ids, min_count)
    V = lib().ora_vocab_size(h)
    word_of_id = np.empty(n_ids, np.int32)
    id_of_word = np.empty(V, np.int32)
    count = np.empty(V, np.int64)
    lib().ora_vocab_tables(h, _p(word_of_id, C.c_int32), _p(id_of_word, C.c_int32), _p(count, C.c_int64))
    total = lib().ora_vocab_total_words(h)
    return h, dict(V=V, word_of_id=word_of_id, id_of_word=id_of_word, count=count, total=total)


def neg_table(vocab_handle, size):
    t = np.empty(size, np.int32)
    lib().ora_neg_table(vocab_handle, size, _p(t, C.c_int32))
    return t


def init_syn0(V, dim, seed):
    a = np.empty((V, dim), np.float32)
    lib().ora_init_syn0(V, dim, seed, _p(a, C.c_float))
    return a


def sgns_train(tokens, n_ids, params):
    tokens = np.ascontiguousarray(tokens, np.int32)
    n_sent, L = tokens.shape
    pairs = C.c_int64()
    h = lib().ora_sgns_train(_p(tokens, C.c_int32), n_sent, L, n_ids, C.byref(params), C.byref(pairs))
    V = lib().ora_model_vocab_size(h)
    syn0 = np.empty((V, params.dim), np.float32)
    syn1 = np.empty((V, params.dim), np.float32)
    ids = np.empty(V, np.int32)
    lib().ora_model_get(h, _p(syn0, C.c_float), _p(syn1, C.c_float), _p(ids, C.c_int32))
    lib().ora_model_free(h)
    return dict(syn0=syn0, syn1neg=syn1, id_of_word=ids, pairs=pairs.value)


COMBINE_SUM, COMBINE_MEAN, COMBINE_CONTRIBUTORS, COMBINE_SQRT, COMBINE_ALIGNED, COMBINE_DELAYED = 0, 1, 2, 3, 4, 16


def sgns_train_dp(tokens, n_ids, params, world, rounds, combine, full_every=1, hot_rows=0):
    """Data-parallel emulation of stage 2 (ora_sgns_train_dp[_tiered]): contiguous sentence shards, per-round delta
    exchange; with full_every > 1 only the hot prefix rows [0, hot_rows) take part in most exchanges."""
    tokens = np.ascontiguousarray(tokens, np.int32)
    n_sent, L = tokens.shape
    pairs = C.c_int64()
    h = lib().ora_sgns_train_dp_tiered(_p(tokens, C.c_int32), n_sent, L, n_ids, C.byref(params), world, rounds, combine,
                                       full_every, hot_rows, C.byref(pairs))
    V = lib().ora_model_vocab_size(h)
    syn0 = np.empty((V, params.dim), np.float32)
    syn1 = np.empty((V, params.dim), np.float32)
    ids = np.empty(V, np.int32)
    lib().ora_model_get(h, _p(syn0, C.c_float), _p(syn1, C.c_float), _p(ids, C.c_int32))
    lib().ora_model_free(h)
    return dict(syn0=syn0, syn1neg=syn1, id_of_word=ids, pairs=pairs.value)


def sgns_count_pairs(tokens, n_ids, params):
    tokens = np.ascontiguousarray(tokens, np.int32)
    n_sent, L = tokens.shape
    return lib().ora_sgns_count_pairs(_p(tokens, C.c_int32), n_sent, L, n_ids, C.byref(params))
