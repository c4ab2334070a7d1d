"""ctypes binding of libdge.so (include/dge.h).  Plumbing only: numpy arrays in, numpy arrays out.

The CUDA library is the product; there is no CPU fallback.  Importing this module never touches the
GPU; `lib()` fails loudly if libdge.so has not been built, `Context()` fails loudly without a B200.
"""
import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdge.so")

SAMPLER_ALIAS, SAMPLER_CDF = 0, 1
SCHEDULE_ITEMS, SCHEDULE_SENTENCE = 0, 1
COMBINE_DEFAULT, COMBINE_MEAN, COMBINE_CONTRIBUTORS, COMBINE_SQRT, COMBINE_ALIGNED, COMBINE_SUM = 0, 1, 2, 3, 4, 5
TRANSPORT_AUTO, TRANSPORT_PEER, TRANSPORT_NCCL = 0, 1, 2
# dge_sgns_params.flags (include/dge.h DGE_SGNS_F_*): kernel-selection / measurement hooks for tests and A/B runs
F_NO_UPDATES, F_NO_NARROW, F_ONE_WARP, F_NARROW, F_TARGET_PARALLEL, F_NO_TARGET_PARALLEL, F_STAGED_ROWS, F_PLAIN_STORES, F_SMEM_NEG_TABLE = \
    1, 2, 8, 32, 64, 128, 256, 512, 1024
F_BLOCK_PER_SENTENCE, F_SMALL_BLOCKS, F_SENTENCE_RESIDENT, F_ITEM_KERNELS, F_PIPELINED, F_PAIR_WARPS, F_HELPER_WARPS, F_ROW_PREFETCH, F_DYNAMIC, F_ROW_PREFETCH_SMEM = 4, 16, 2048, 65536, 131072, 262144, 524288, 1 << 24, 1 << 25, 1 << 26
COMM_ID_BYTES = 128

class DgeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libdge error %d: %s" % (code, msg))
        self.code = code


class SgnsParams(C.Structure):
    _fields_ = [("dim", C.c_int32), ("window", C.c_int32), ("negative", C.c_int32), ("min_count", C.c_int32),
                ("epochs", C.c_int32), ("neg_table_size", C.c_int32), ("exp_table_size", C.c_int32),
                ("concurrency", C.c_int32), ("schedule", C.c_int32), ("sync_rounds", C.c_int32), ("combine", C.c_int32),
                ("transport", C.c_int32), ("flags", C.c_uint32), ("lr", C.c_float), ("min_lr", C.c_float), ("seed", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -m embedding_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    P = C.POINTER
    pi32, pi64, pf64, pf32 = P(i32), P(i64), P(f64), P(C.c_float)
    L.dge_version.restype = C.c_int
    L.dge_create.argtypes = [C.c_int, P(vp)]
    L.dge_destroy.argtypes = [vp]
    L.dge_destroy.restype = None
    L.dge_last_error.argtypes = [vp]
    L.dge_last_error.restype = C.c_char_p
    L.dge_host_alloc.argtypes = [C.c_size_t]
    L.dge_host_alloc.restype = vp
    L.dge_host_free.argtypes = [vp]
    L.dge_host_free.restype = None
    L.dge_phase_ms.argtypes = [vp, C.c_char_p, P(C.c_float)]
    L.dge_kernel_launches.argtypes = [vp]
    L.dge_kernel_launches.restype = i64
    L.dge_graph_build.argtypes = [vp, i32, i64, pi32, pi32, pf64, i32, pi32, pf64, pf64, P(vp)]
    L.dge_graph_sizes.argtypes = [vp, pi32, pi64, pi32]
    L.dge_graph_tables.argtypes = [vp, pi64, pi32, pf64, pf64, pi32, pf64, pf64, pi32, pf64]
    L.dge_graph_sample_next.argtypes = [vp, i64, pi32, pf64, C.c_int, pi32]
    L.dge_graph_free.argtypes = [vp]
    L.dge_graph_free.restype = None
    L.dge_walk.argtypes = [vp, i64, i64, i32, u64, C.c_int, P(vp)]
    L.dge_corpus_from_tokens.argtypes = [vp, pi32, i64, i32, i32, P(vp)]
    L.dge_corpus_shape.argtypes = [vp, pi64, pi32, pi32]
    L.dge_corpus_tokens.argtypes = [vp, pi32]
    L.dge_corpus_tokens_u16.argtypes = [vp, P(C.c_uint16)]
    L.dge_corpus_count_tokens.argtypes = [vp, pi64]
    L.dge_corpus_relabel.argtypes = [vp, pi32, i32, i32]
    L.dge_corpus_write_seq.argtypes = [vp, pi32, pi32, C.c_int, C.c_char_p, C.c_int]
    L.dge_corpus_read_seq.argtypes = [vp, C.c_char_p, pi32, pi32, i32, C.c_int, P(vp)]
    L.dge_corpus_free.argtypes = [vp]
    L.dge_corpus_free.restype = None
    L.dge_sgns_default_params.argtypes = [P(SgnsParams)]
    L.dge_sgns_default_params.restype = None
    L.dge_sgns_train.argtypes = [vp, P(vp), i32, P(SgnsParams), P(vp)]
    L.dge_model_shape.argtypes = [vp, pi32, pi32, pi64]
    L.dge_model_vectors.argtypes = [vp, pf32, pf32, pi32]
    L.dge_model_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), pi64]
    L.dge_model_write_vec.argtypes = [vp, pi32, pi32, C.c_char_p]
    L.dge_model_free.argtypes = [vp]
    L.dge_model_free.restype = None
    L.dge_flows_create.argtypes = [vp, i32, pi32, P(vp)]
    L.dge_flows_add_trips.argtypes = [vp, i64, pi32, pi32, pi32]
    L.dge_flows_tensor.argtypes = [vp, pi32]
    L.dge_flows_slot_weights.argtypes = [vp, C.c_int, i32, i32, pi32]
    L.dge_flows_write_matrix.argtypes = [vp, C.c_int, i32, i32, pi32, i32, pi32, i32, C.c_char, C.c_char_p]
    L.dge_flows_write_od.argtypes = [vp, C.c_int, i32, i32, pi32, i32, pi32, i32, pi32, C.c_int, i32, C.c_char_p]
    L.dge_flows_free.argtypes = [vp]
    L.dge_flows_free.restype = None
    L.dge_crosstime_graph_build.argtypes = [vp, pi32, i32, C.c_int, pi32, P(vp)]
    L.dge_graph_labels.argtypes = [vp, pi32, pi32, pi32]
    L.dge_eval_knn.argtypes = [vp, pf32, i32, i32, i32, pi32, pf64]
    L.dge_eval_ndcg.argtypes = [vp, pf32, i32, i32, pi32, pf64, i32, i32, pf64, P(C.c_double)]
    L.dge_timer_start.argtypes = [vp]
    L.dge_timer_stop.argtypes = [vp, P(C.c_float)]
    L.dge_comm_unique_id.argtypes = [vp, C.c_size_t]
    L.dge_comm_init.argtypes = [vp, C.c_int, C.c_int, vp, C.c_size_t]
    L.dge_comm_shape.argtypes = [vp, P(C.c_int), P(C.c_int)]
    L.dge_comm_destroy.argtypes = [vp]
    L.dge_comm_destroy.restype = None
    L.dge_comm_nccl_version.restype = C.c_int
    _lib = L
    return L


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _check(rc, ctx_handle=None):
    if rc != 0:
        msg = lib().dge_last_error(ctx_handle)
        raise DgeError(rc, msg.decode() if msg else "?")


class PinnedArray:
    """numpy view over cudaMallocHost memory (dge_host_alloc)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = lib().dge_host_alloc(max(nbytes, 1))
        if not self._p:
            raise DgeError(-3, "dge_host_alloc failed")
        buf = (C.c_char * max(nbytes, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._p:
            self.array = None
            lib().dge_host_free(self._p)
            self._p = None

    def __del__(self):
        self.free()


class Context:
    """dge_ctx: one per GPU / process."""

    def __init__(self, device=None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        rc = lib().dge_create(int(device), C.byref(h))
        if rc != 0:
            _check(rc, None)
        self._h = h
        self.device = int(device)
        self._children = weakref.WeakSet()   # graphs / corpora / models: released before the ctx (dge.h: handles
                                             # must be freed before dge_destroy)

    def _adopt(self, child):
        self._children.add(child)

    def close(self):
        if getattr(self, "_h", None):
            for ch in list(self._children):
                ch.free()
            lib().dge_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def phase_ms(self, name):
        ms = C.c_float()
        rc = lib().dge_phase_ms(self._h, name.encode(), C.byref(ms))
        return None if rc != 0 else float(ms.value)

    def kernel_launches(self):
        return int(lib().dge_kernel_launches(self._h))

    def eval_knn(self, X, topk, want_dist=False):
        """dge_eval_knn: topk nearest other rows by cosine distance (fp64), ties by row index."""
        X = np.ascontiguousarray(X, np.float32)
        m, dim = X.shape
        nbr = np.empty((m, topk), np.int32)
        dist = np.empty((m, topk), np.float64) if want_dist else None
        _check(lib().dge_eval_knn(self._h, _ptr(X, C.c_float), m, dim, int(topk), _ptr(nbr, C.c_int32), _ptr(dist, C.c_double)),
               self._h)
        return (nbr, dist) if want_dist else nbr

    def eval_ndcg(self, X, gt_index, gt_dist, topk):
        """dge_eval_ndcg: (per-row nDCG@topk, mean) of one embedding layer against a ground-truth distance matrix."""
        X = np.ascontiguousarray(X, np.float32)
        gi = np.ascontiguousarray(gt_index, np.int32)
        gt = np.ascontiguousarray(gt_dist, np.float64)
        m, dim = X.shape
        n = gt.shape[0]
        if gt.shape != (n, n) or len(gi) != m:
            raise ValueError("gt_dist must be [n, n] and gt_index must have one entry per row of X")
        nd = np.empty(m, np.float64)
        mean = C.c_double()
        _check(lib().dge_eval_ndcg(self._h, _ptr(X, C.c_float), m, dim, _ptr(gi, C.c_int32), _ptr(gt, C.c_double), n, int(topk),
                                   _ptr(nd, C.c_double), C.byref(mean)), self._h)
        return nd, float(mean.value)

    def timer_start(self):
        _check(lib().dge_timer_start(self._h), self._h)

    def timer_stop(self):
        """ms between timer_start() and now, from CUDA events on the ctx stream."""
        ms = C.c_float()
        _check(lib().dge_timer_stop(self._h, C.byref(ms)), self._h)
        return float(ms.value)

    # ---- multi-GPU (one process per GPU): the host carries the opaque NCCL id from rank 0 to the others
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(COMM_ID_BYTES)
        _check(lib().dge_comm_unique_id(C.cast(buf, C.c_void_p), COMM_ID_BYTES), None)
        return buf.raw

    def comm_init(self, rank, world, unique_id):
        if len(unique_id) != COMM_ID_BYTES:
            raise ValueError("unique_id must be %d bytes" % COMM_ID_BYTES)
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        _check(lib().dge_comm_init(self._h, int(rank), int(world), C.cast(buf, C.c_void_p), COMM_ID_BYTES), self._h)
        self.rank, self.world = int(rank), int(world)

    def comm_shape(self):
        r, w = C.c_int(), C.c_int()
        _check(lib().dge_comm_shape(self._h, C.byref(r), C.byref(w)), self._h)
        return r.value, w.value


class Flows:
    """dge_flows: the hourly flow counts F[src][hour][dst] on the device."""

    def __init__(self, ctx, n_regions, F=None):
        self.ctx, self.n = ctx, int(n_regions)
        if F is not None:
            F = np.ascontiguousarray(F, np.int32)
            if F.shape != (self.n, 24, self.n):
                raise ValueError("F must be [n_regions, 24, n_regions]")
        h = C.c_void_p()
        _check(lib().dge_flows_create(ctx._h, self.n, _ptr(F, C.c_int32), C.byref(h)), ctx._h)
        self._h = h
        ctx._adopt(self)

    def add_trips(self, src_region, dst_region, start_hour):
        s = np.ascontiguousarray(src_region, np.int32)
        d = np.ascontiguousarray(dst_region, np.int32)
        h = np.ascontiguousarray(start_hour, np.int32)
        if not (len(s) == len(d) == len(h)):
            raise ValueError("trip arrays must have equal length")
        _check(lib().dge_flows_add_trips(self._h, len(s), _ptr(s, C.c_int32), _ptr(d, C.c_int32), _ptr(h, C.c_int32)),
               self.ctx._h)

    def tensor(self):
        F = np.empty((self.n, 24, self.n), np.int32)
        _check(lib().dge_flows_tensor(self._h, _ptr(F, C.c_int32)), self.ctx._h)
        return F

    def slot_weights(self, mode, lo, hi):
        """W[src, dst] = getFlowTo(dst, lo, hi): mode 0 CA (circular, half-open), mode 1 tract (inclusive)."""
        W = np.empty((self.n, self.n), np.int32)
        _check(lib().dge_flows_slot_weights(self._h, int(mode), int(lo), int(hi), _ptr(W, C.c_int32)), self.ctx._h)
        return W

    def write_matrix(self, mode, lo, hi, rows, cols, sep, path):
        r, c = np.ascontiguousarray(rows, np.int32), np.ascontiguousarray(cols, np.int32)
        _check(lib().dge_flows_write_matrix(self._h, int(mode), int(lo), int(hi), _ptr(r, C.c_int32), len(r),
                                            _ptr(c, C.c_int32), len(c), sep.encode(), os.fsencode(path)), self.ctx._h)

    def write_od(self, mode, lo, hi, rows, cols, region_ids, path, keep_zero=False, presence_hour=-1):
        r, c = np.ascontiguousarray(rows, np.int32), np.ascontiguousarray(cols, np.int32)
        ids = np.ascontiguousarray(region_ids, np.int32)
        if len(ids) != self.n:
            raise ValueError("region_ids must have n_regions entries")
        _check(lib().dge_flows_write_od(self._h, int(mode), int(lo), int(hi), _ptr(r, C.c_int32), len(r), _ptr(c, C.c_int32),
                                        len(c), _ptr(ids, C.c_int32), int(bool(keep_zero)), int(presence_hour),
                                        os.fsencode(path)), self.ctx._h)

    def crosstime_graph(self, order, num_layer, mode, intervals=None):
        """mode 0 = constructGraph_CA(intervals), mode 1 = constructGraph_tract(); returns a Graph with labels."""
        o = np.ascontiguousarray(order, np.int32)
        if len(o) != self.n:
            raise ValueError("order must have n_regions entries")
        iv = None if intervals is None else np.ascontiguousarray(intervals, np.int32)
        if iv is not None and len(iv) != num_layer + 1:
            raise ValueError("intervals must have num_layer + 1 entries")
        h = C.c_void_p()
        _check(lib().dge_crosstime_graph_build(self._h, _ptr(o, C.c_int32), int(num_layer), int(mode), _ptr(iv, C.c_int32),
                                               C.byref(h)), self.ctx._h)
        return Graph._from_handle(self.ctx, h)

    def free(self):
        if getattr(self, "_h", None):
            lib().dge_flows_free(self._h)
            self._h = None

    def __del__(self):
        self.free()


class Graph:
    """dge_graph: CSR + alias tables + packed walk records on the device."""

    @classmethod
    def _from_handle(cls, ctx, h):
        g = cls.__new__(cls)
        g.ctx, g._h = ctx, h
        ctx._adopt(g)
        nv, ne, ns = C.c_int32(), C.c_int64(), C.c_int32()
        _check(lib().dge_graph_sizes(h, C.byref(nv), C.byref(ne), C.byref(ns)), ctx._h)
        g.nv, g.ne, g.ns = nv.value, ne.value, ns.value
        return g

    def labels(self):
        """(v_layer, v_region_index, sources) of a graph built from flows on the device."""
        vl, vr, so = np.empty(self.nv, np.int32), np.empty(self.nv, np.int32), np.empty(self.ns, np.int32)
        _check(lib().dge_graph_labels(self._h, _ptr(vl, C.c_int32), _ptr(vr, C.c_int32), _ptr(so, C.c_int32)), self.ctx._h)
        return vl, vr, so

    def __init__(self, ctx, n_vertices, src, dst, w, sources, out_degree=None, source_weight_sum=None):
        self.ctx = ctx
        src = np.ascontiguousarray(src, np.int32)
        dst = np.ascontiguousarray(dst, np.int32)
        w = np.ascontiguousarray(w, np.float64)
        sources = np.ascontiguousarray(sources, np.int32)
        if not (len(src) == len(dst) == len(w)):
            raise ValueError("src, dst, w must have equal length")
        od = None if out_degree is None else np.ascontiguousarray(out_degree, np.float64)
        if od is not None and len(od) != n_vertices:
            raise ValueError("out_degree must have n_vertices entries")
        sws = None if source_weight_sum is None else np.array([source_weight_sum], np.float64)
        h = C.c_void_p()
        _check(lib().dge_graph_build(ctx._h, int(n_vertices), len(src), _ptr(src, C.c_int32), _ptr(dst, C.c_int32),
                                     _ptr(w, C.c_double), len(sources), _ptr(sources, C.c_int32),
                                     _ptr(od, C.c_double), _ptr(sws, C.c_double), C.byref(h)), ctx._h)
        self._h = h
        ctx._adopt(self)
        self.nv, self.ne, self.ns = int(n_vertices), len(src), len(sources)

    def free(self):
        if getattr(self, "_h", None):
            lib().dge_graph_free(self._h)
            self._h = None

    def __del__(self):
        self.free()

    def tables(self):
        t = dict(row_ptr=np.empty(self.nv + 1, np.int64), col=np.empty(self.ne, np.int32),
                 w=np.empty(self.ne, np.float64), prob=np.empty(self.ne, np.float64),
                 alias=np.empty(self.ne, np.int32), out_degree=np.empty(self.nv, np.float64),
                 src_prob=np.empty(self.ns, np.float64), src_alias=np.empty(self.ns, np.int32))
        sws = np.empty(1, np.float64)
        _check(lib().dge_graph_tables(self._h, _ptr(t["row_ptr"], C.c_int64), _ptr(t["col"], C.c_int32),
                                      _ptr(t["w"], C.c_double), _ptr(t["prob"], C.c_double),
                                      _ptr(t["alias"], C.c_int32), _ptr(t["out_degree"], C.c_double),
                                      _ptr(t["src_prob"], C.c_double), _ptr(t["src_alias"], C.c_int32),
                                      _ptr(sws, C.c_double)), self.ctx._h)
        t["source_weight_sum"] = float(sws[0])
        return t

    def sample_next(self, v, x, sampler=SAMPLER_ALIAS):
        v = np.ascontiguousarray(v, np.int32)
        x = np.ascontiguousarray(x, np.float64)
        out = np.empty(len(v), np.int32)
        _check(lib().dge_graph_sample_next(self._h, len(v), _ptr(v, C.c_int32), _ptr(x, C.c_double), sampler,
                                           _ptr(out, C.c_int32)), self.ctx._h)
        return out

    def walk(self, n_walks, num_layer, seed, sampler=SAMPLER_ALIAS, first_walk_id=0):
        h = C.c_void_p()
        _check(lib().dge_walk(self._h, int(n_walks), int(first_walk_id), int(num_layer), int(seed), sampler,
                              C.byref(h)), self.ctx._h)
        return Corpus(self.ctx, h)


class Corpus:
    """dge_corpus: walk tokens resident on the device."""

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self._h = handle
        ctx._adopt(self)
        n, L, ids = C.c_int64(), C.c_int32(), C.c_int32()
        _check(lib().dge_corpus_shape(handle, C.byref(n), C.byref(L), C.byref(ids)), ctx._h)
        self.n_walks, self.L, self.n_ids = n.value, L.value, ids.value

    @classmethod
    def from_tokens(cls, ctx, tokens, n_ids):
        tokens = np.ascontiguousarray(tokens, np.int32)
        if tokens.ndim != 2:
            raise ValueError("tokens must be [n_walks, L]")
        h = C.c_void_p()
        _check(lib().dge_corpus_from_tokens(ctx._h, _ptr(tokens, C.c_int32), tokens.shape[0], tokens.shape[1],
                                            int(n_ids), C.byref(h)), ctx._h)
        return cls(ctx, h)

    @classmethod
    def read_seq(cls, ctx, path, label_region, label_layer=None, position_prefix=False):
        lr = np.ascontiguousarray(label_region, np.int32)
        ll = None if label_layer is None else np.ascontiguousarray(label_layer, np.int32)
        h = C.c_void_p()
        _check(lib().dge_corpus_read_seq(ctx._h, os.fsencode(path), _ptr(ll, C.c_int32), _ptr(lr, C.c_int32), len(lr),
                                         1 if position_prefix else 0, C.byref(h)), ctx._h)
        return cls(ctx, h)

    def free(self):
        if getattr(self, "_h", None):
            lib().dge_corpus_free(self._h)
            self._h = None

    def __del__(self):
        self.free()

    def tokens(self, out=None):
        if out is None:
            out = np.empty((self.n_walks, self.L), np.int32)
        assert out.dtype == np.int32 and out.size == self.n_walks * self.L and out.flags.c_contiguous
        _check(lib().dge_corpus_tokens(self._h, _ptr(out, C.c_int32)), self.ctx._h)
        return out

    def tokens_u16(self, out=None):
        """Walk-major tokens as uint16 (0xFFFF = padding); only for id spaces below 65 535."""
        if out is None:
            out = np.empty((self.n_walks, self.L), np.uint16)
        assert out.dtype == np.uint16 and out.size == self.n_walks * self.L and out.flags.c_contiguous
        _check(lib().dge_corpus_tokens_u16(self._h, _ptr(out, C.c_uint16)), self.ctx._h)
        return out

    def relabel(self, id_map, new_n_ids, position_stride=0):
        m = np.ascontiguousarray(id_map, np.int32)
        if len(m) != self.n_ids:
            raise ValueError("id_map must have n_ids entries")
        _check(lib().dge_corpus_relabel(self._h, _ptr(m, C.c_int32), int(new_n_ids), int(position_stride)),
               self.ctx._h)
        self.n_ids = int(new_n_ids)

    def count_tokens(self):
        n = C.c_int64()
        _check(lib().dge_corpus_count_tokens(self._h, C.byref(n)), self.ctx._h)
        return n.value

    def write_seq(self, path, label_region, label_layer=None, position_prefix=False, append=False):
        lr = np.ascontiguousarray(label_region, np.int32)
        ll = None if label_layer is None else np.ascontiguousarray(label_layer, np.int32)
        _check(lib().dge_corpus_write_seq(self._h, _ptr(ll, C.c_int32), _ptr(lr, C.c_int32),
                                          1 if position_prefix else 0, os.fsencode(path), 1 if append else 0),
               self.ctx._h)


def sgns_params(**kw):
    """dge_sgns_default_params + overrides.  A/B scripts may set DGE_SGNS_FLAGS in the environment: it is read HERE, on
    the host side, and travels in the explicit `flags` field -- the library itself never reads the environment."""
    p = SgnsParams()
    lib().dge_sgns_default_params(C.byref(p))
    if "flags" not in kw and os.environ.get("DGE_SGNS_FLAGS"):
        kw = dict(kw, flags=int(os.environ["DGE_SGNS_FLAGS"]))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError("unknown SGNS parameter %r" % k)
        setattr(p, k, v)
    return p


class Model:
    """dge_model: syn0 / syn1neg resident on the device."""

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self._h = handle
        ctx._adopt(self)
        V, d, pairs = C.c_int32(), C.c_int32(), C.c_int64()
        _check(lib().dge_model_shape(handle, C.byref(V), C.byref(d), C.byref(pairs)), ctx._h)
        self.V, self.dim, self.pairs = V.value, d.value, pairs.value

    @classmethod
    def train(cls, ctx, corpora, params):
        arr = (C.c_void_p * len(corpora))(*[c._h for c in corpora])
        h = C.c_void_p()
        _check(lib().dge_sgns_train(ctx._h, arr, len(corpora), C.byref(params), C.byref(h)), ctx._h)
        return cls(ctx, h)

    def free(self):
        if getattr(self, "_h", None):
            lib().dge_model_free(self._h)
            self._h = None

    def __del__(self):
        self.free()

    def vectors(self, want_syn1neg=False):
        syn0 = np.empty((self.V, self.dim), np.float32)
        syn1 = np.empty((self.V, self.dim), np.float32) if want_syn1neg else None
        ids = np.empty(self.V, np.int32)
        _check(lib().dge_model_vectors(self._h, _ptr(syn0, C.c_float), _ptr(syn1, C.c_float), _ptr(ids, C.c_int32)),
               self.ctx._h)
        return (syn0, syn1, ids) if want_syn1neg else (syn0, ids)

    def stats(self):
        """(mean |syn0 row|, max |element| of both tables, non-finite elements) computed on the device."""
        norm, mx, bad = C.c_double(), C.c_double(), C.c_int64()
        _check(lib().dge_model_stats(self._h, C.byref(norm), C.byref(mx), C.byref(bad)), self.ctx._h)
        return dict(mean_row_norm=norm.value, max_abs=mx.value, nonfinite=bad.value)

    def write_vec(self, path, label_layer, label_region):
        ll = np.ascontiguousarray(label_layer, np.int32)
        lr = np.ascontiguousarray(label_region, np.int32)
        _check(lib().dge_model_write_vec(self._h, _ptr(ll, C.c_int32), _ptr(lr, C.c_int32), os.fsencode(path)),
               self.ctx._h)
