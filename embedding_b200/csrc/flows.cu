// flows.cu -- the step in front of stage 1 (8(f) N1): trip records -> hourly flow
// counts -> slot weights -> the CrossTimeGraph edge list, vertex ids and sources, all on device, handed straight to
// the CSR / alias builder of graph.cu.
//
// Reference behaviour reproduced (embedding/src/main/java/embedding/):
//   CommunityAreas.mapTripsIntoCommunities :55-103, Tracts.mapTripsIntoTracts :71-102
//        per trip: taxiFlows[src].get(hour).put(dst, count + 1)    (point-in-polygon lookup stays on the host)
//   CommunityArea.getFlowTo(dst, lo, hi) :240-245   circular half-open  for (h = lo; h != hi; h = (h+1) % 24)
//   Tract.getFlowTo(dst, lo, hi)  Tracts.java:477-482   inclusive       for (h = lo; h <= hi; h++)
//   CrossTimeGraph.constructGraph_CA(int[]) :68-95 / constructGraph_tract() :25-52
//        for h, for src in regions, for dst in regions: w = getFlowTo(...); if (w > 0) addEdge("h-src", "h+1-dst", w)
//        then for each region: if "0-id" is a vertex: addSourceVertex
//   LayeredGraph.addEdge :157-174   vertex id = first appearance in the interleaved mention order s0,d0,s1,d1,...
// The storage is the dense tensor F[src][hour][dst] int32 (the reference keeps 24 HashMaps per region; dense is the
// GPU-friendly equivalent at CA / tract scale: 77 -> 0.6 MB, 801 -> 62 MB).
#include "dge_internal.cuh"

#define FLOWS_MAX_REGIONS 8192

__global__ void k_add_trips(const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                            const int32_t *__restrict__ hour, int64_t n_trips, int32_t n, int32_t *F, int *bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n_trips; i += stride) {
        int32_t s = src[i], d = dst[i], h = hour[i];
        if (s < 0 || d < 0) continue; // trip outside every region: the Java loops never find s / e, nothing is counted
        if (s >= n || d >= n || h < 0 || h > 23) { *bad = 1; continue; }
        atomicAdd(&F[((int64_t)s * 24 + h) * n + d], 1);
    }
}

// W[h][a][b] = getFlowTo(order[b], slot h) of region order[a]; flag = (W > 0).   One thread per (h, a, b), b fastest.
// mode 0 (CA): hours lo = iv[h], hi = iv[h+1], circular half-open.  mode 1 (tract): [h, h + time_step - 1] inclusive.
__global__ void k_slot_weights(const int32_t *__restrict__ F, int32_t n, const int32_t *__restrict__ order, int32_t L,
                               int mode, const int32_t *__restrict__ iv, int32_t time_step, int32_t *__restrict__ W,
                               int32_t *__restrict__ flag) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = (int64_t)L * n * n;
    if (t >= total) return;
    int32_t b = (int32_t)(t % n);
    int32_t a = (int32_t)((t / n) % n);
    int32_t h = (int32_t)(t / ((int64_t)n * n));
    const int32_t *row = F + (int64_t)order[a] * 24 * n + order[b];
    int32_t cnt = 0;
    if (mode == 0) {
        int32_t lo = iv[h], hi = iv[h + 1];
        for (int32_t x = lo; x != hi; x = (x + 1) % 24) cnt += row[(int64_t)x * n];
    } else {
        int32_t lo = h, hi = h + time_step - 1;
        for (int32_t x = lo; x <= hi; x++) cnt += row[(int64_t)x * n];
    }
    W[t] = cnt;
    flag[t] = cnt > 0;
}

// edge e = pos[t] for every (h, a, b) with W > 0: vertex keys (layer * n + region index), weight, first mentions
__global__ void k_emit_edges(const int32_t *__restrict__ W, const int64_t *__restrict__ pos, int32_t n,
                             const int32_t *__restrict__ order, int32_t L, int32_t *__restrict__ skey,
                             int32_t *__restrict__ dkey, double *__restrict__ w, unsigned long long *__restrict__ first) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = (int64_t)L * n * n;
    if (t >= total) return;
    int32_t c = W[t];
    if (c <= 0) return;
    int32_t b = (int32_t)(t % n);
    int32_t a = (int32_t)((t / n) % n);
    int32_t h = (int32_t)(t / ((int64_t)n * n));
    int64_t e = pos[t];
    int32_t sk = h * n + order[a], dk = ((h + 1) % L) * n + order[b];
    skey[e] = sk; dkey[e] = dk; w[e] = (double)c; // Edge.weight is a double (LayeredGraph.java:17-27)
    atomicMin(&first[sk], (unsigned long long)(2 * e));
    atomicMin(&first[dk], (unsigned long long)(2 * e + 1));
}

__global__ void k_mention_flags(const int32_t *__restrict__ skey, const int32_t *__restrict__ dkey, int64_t ne,
                                const unsigned long long *__restrict__ first, int32_t *__restrict__ mflag) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    mflag[2 * e] = first[skey[e]] == (unsigned long long)(2 * e);
    mflag[2 * e + 1] = first[dkey[e]] == (unsigned long long)(2 * e + 1);
}

__global__ void k_assign_ids(const unsigned long long *__restrict__ first, int32_t n_keys, int32_t n,
                             const int64_t *__restrict__ mpos, int32_t *__restrict__ vid, int32_t *__restrict__ v_layer,
                             int32_t *__restrict__ v_region) {
    int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_keys) return;
    unsigned long long f = first[k];
    if (f == ~0ULL) { vid[k] = -1; return; }
    int32_t id = (int32_t)mpos[f];
    vid[k] = id;
    v_layer[id] = k / n;
    v_region[id] = k % n;
}

__global__ void k_keys_to_ids(int32_t *__restrict__ skey, int32_t *__restrict__ dkey, int64_t ne,
                              const int32_t *__restrict__ vid) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    skey[e] = vid[skey[e]];
    dkey[e] = vid[dkey[e]];
}

// W[a][b] = getFlowTo(b, lo, hi) of region a for all pairs (static exports); b fastest => coalesced.
__global__ void k_slot_matrix(const int32_t *__restrict__ F, int32_t n, int mode, int32_t lo, int32_t hi,
                              int32_t *__restrict__ W) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * n) return;
    int32_t b = (int32_t)(t % n), a = (int32_t)(t / n);
    const int32_t *row = F + (int64_t)a * 24 * n + b;
    int32_t cnt = 0;
    if (mode == 0) for (int32_t x = lo; x != hi; x = (x + 1) % 24) cnt += row[(int64_t)x * n];
    else for (int32_t x = lo; x <= hi; x++) cnt += row[(int64_t)x * n];
    W[t] = cnt;
}

struct flows_tmp {
    dge_ctx *ctx;
    void *p = nullptr;
    explicit flows_tmp(dge_ctx *c) : ctx(c) {}
    ~flows_tmp() { dge_free(ctx, p); }
};

extern "C" {

int dge_flows_create(dge_ctx *ctx, int32_t n_regions, const int32_t *F, dge_flows **out) {
    if (!ctx) return dge_fail(nullptr, DGE_E_INVALID, "dge_flows_create: ctx is NULL");
    if (!out) return dge_fail(ctx, DGE_E_INVALID, "dge_flows_create: out is NULL");
    *out = nullptr;
    if (n_regions < 0) return dge_fail(ctx, DGE_E_INVALID, "dge_flows_create: negative size");
    if (n_regions > FLOWS_MAX_REGIONS)
        return dge_fail(ctx, DGE_E_LIMIT, "dge_flows_create: the dense flow tensor supports at most 8192 regions");
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    dge_flows *f = dge_new_handle<dge_flows>(ctx);
    f->ctx = ctx; f->n = n_regions;
    size_t total = (size_t)n_regions * 24 * (size_t)n_regions;
    cudaError_t e = dge_malloc(ctx, &f->F, total);
    if (e == cudaSuccess) {
        if (F) e = cudaMemcpyAsync(f->F, F, total * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
        else e = cudaMemsetAsync(f->F, 0, (total ? total : 1) * sizeof(int32_t), ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        dge_free(ctx, f->F);
        dge_delete_handle(f);
        return dge_fail(ctx, DGE_E_CUDA, std::string("dge_flows_create: ") + cudaGetErrorString(e));
    }
    *out = f;
    return DGE_OK;
}

int dge_flows_add_trips(dge_flows *f, int64_t n_trips, const int32_t *src_region, const int32_t *dst_region,
                        const int32_t *start_hour) {
    if (!f) return dge_fail(nullptr, DGE_E_INVALID, "dge_flows_add_trips: flows is NULL");
    dge_ctx *ctx = f->ctx;
    if (n_trips < 0 || (n_trips > 0 && (!src_region || !dst_region || !start_hour)))
        return dge_fail(ctx, DGE_E_INVALID, "dge_flows_add_trips: bad arguments");
    if (n_trips == 0) return DGE_OK;
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    flows_tmp ts(ctx), td(ctx), th(ctx), tb(ctx);
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&ts.p, (size_t)n_trips));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&td.p, (size_t)n_trips));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&th.p, (size_t)n_trips));
    DGE_CUDA(ctx, dge_malloc(ctx, (int **)&tb.p, 1));
    DGE_CUDA(ctx, cudaMemsetAsync(tb.p, 0, sizeof(int), st));
    dge_phase_timer t(ctx, "trips");
    DGE_CUDA(ctx, cudaMemcpyAsync(ts.p, src_region, sizeof(int32_t) * (size_t)n_trips, cudaMemcpyHostToDevice, st));
    DGE_CUDA(ctx, cudaMemcpyAsync(td.p, dst_region, sizeof(int32_t) * (size_t)n_trips, cudaMemcpyHostToDevice, st));
    DGE_CUDA(ctx, cudaMemcpyAsync(th.p, start_hour, sizeof(int32_t) * (size_t)n_trips, cudaMemcpyHostToDevice, st));
    k_add_trips<<<ctx->sm_count * 8, 256, 0, st>>>((const int32_t *)ts.p, (const int32_t *)td.p, (const int32_t *)th.p, n_trips,
                                                  f->n, f->F, (int *)tb.p);
    DGE_LAUNCH_CHECK(ctx);
    t.stop();
    int h_bad = 0;
    DGE_CUDA(ctx, cudaMemcpyAsync(&h_bad, tb.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    DGE_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_bad) return dge_fail(ctx, DGE_E_INVALID, "dge_flows_add_trips: region index >= n_regions or hour outside 0..23 (those records were skipped)");
    f->trips += n_trips;
    return DGE_OK;
}

int dge_flows_tensor(const dge_flows *f, int32_t *F) {
    if (!f || !F) return dge_fail(f ? f->ctx : nullptr, DGE_E_INVALID, "dge_flows_tensor: NULL argument");
    dge_ctx *ctx = f->ctx;
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t total = (size_t)f->n * 24 * (size_t)f->n;
    if (total) DGE_CUDA(ctx, cudaMemcpyAsync(F, f->F, total * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    DGE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DGE_OK;
}

int dge_flows_slot_weights(const dge_flows *f, int mode, int32_t lo, int32_t hi, int32_t *W) {
    if (!f) return dge_fail(nullptr, DGE_E_INVALID, "dge_flows_slot_weights: flows is NULL");
    dge_ctx *ctx = f->ctx;
    if (!W && f->n) return dge_fail(ctx, DGE_E_INVALID, "dge_flows_slot_weights: W is NULL");
    if ((mode != 0 && mode != 1) || lo < 0 || lo > 23 || hi < 0 || hi > 23)
        return dge_fail(ctx, DGE_E_INVALID, "dge_flows_slot_weights: mode must be 0 or 1 and the hours must lie in 0..23");
    const int64_t cells = (int64_t)f->n * f->n;
    if (!cells) return DGE_OK;
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    flows_tmp t_W(ctx);
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_W.p, (size_t)cells));
    dge_phase_timer t(ctx, "slot_weights");
    k_slot_matrix<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(f->F, f->n, mode, lo, hi, (int32_t *)t_W.p);
    DGE_LAUNCH_CHECK(ctx);
    t.stop();
    DGE_CUDA(ctx, cudaMemcpyAsync(W, t_W.p, sizeof(int32_t) * (size_t)cells, cudaMemcpyDeviceToHost, ctx->stream));
    DGE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DGE_OK;
}

void dge_flows_free(dge_flows *f) {
    if (!f) return;
    cudaSetDevice(f->ctx->device);
    dge_free(f->ctx, f->F);
    dge_delete_handle(f);
}

int dge_crosstime_graph_build(const dge_flows *f, const int32_t *order, int32_t num_layer, int mode,
                              const int32_t *intervals, dge_graph **out) {
    if (!f) return dge_fail(nullptr, DGE_E_INVALID, "dge_crosstime_graph_build: flows is NULL");
    dge_ctx *ctx = f->ctx;
    if (!out) return dge_fail(ctx, DGE_E_INVALID, "dge_crosstime_graph_build: out is NULL");
    *out = nullptr;
    const int32_t n = f->n, L = num_layer;
    if (L < 1 || L > 24 || (mode != 0 && mode != 1) || (n > 0 && !order) || (mode == 0 && !intervals))
        return dge_fail(ctx, DGE_E_INVALID, "dge_crosstime_graph_build: bad arguments (1 <= num_layer <= 24, mode 0 = CA needs intervals)");
    const int32_t time_step = 24 / L; // CrossTimeGraph.java:30 (integer division)
    std::vector<char> seen((size_t)n, 0);
    for (int32_t i = 0; i < n; i++) {
        if (order[i] < 0 || order[i] >= n || seen[order[i]])
            return dge_fail(ctx, DGE_E_INVALID, "dge_crosstime_graph_build: order must be a permutation of the region indices");
        seen[order[i]] = 1;
    }
    if (mode == 0)
        for (int32_t h = 0; h <= L; h++)
            if (intervals[h] < 0 || intervals[h] > 23)
                return dge_fail(ctx, DGE_E_INVALID, "dge_crosstime_graph_build: interval bounds must lie in 0..23 (the Java loop would not terminate)");
    const int64_t cells = (int64_t)L * n * n;
    if (cells >= ((int64_t)1 << 31)) return dge_fail(ctx, DGE_E_LIMIT, "dge_crosstime_graph_build: num_layer * n_regions^2 must be < 2^31");
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int T = 256;
    const int32_t n_keys = L * n;

    dge_phase_timer t_flow(ctx, "crosstime_edges");
    flows_tmp t_order(ctx), t_iv(ctx), t_W(ctx), t_flag(ctx), t_pos(ctx), t_first(ctx);
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_order.p, (size_t)n));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_iv.p, (size_t)L + 1));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_W.p, (size_t)cells));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_flag.p, (size_t)cells));
    DGE_CUDA(ctx, dge_malloc(ctx, (int64_t **)&t_pos.p, (size_t)cells + 1));
    DGE_CUDA(ctx, dge_malloc(ctx, (unsigned long long **)&t_first.p, (size_t)n_keys));
    if (n) DGE_CUDA(ctx, cudaMemcpyAsync(t_order.p, order, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    if (mode == 0) DGE_CUDA(ctx, cudaMemcpyAsync(t_iv.p, intervals, sizeof(int32_t) * ((size_t)L + 1), cudaMemcpyHostToDevice, st));
    DGE_CUDA(ctx, cudaMemsetAsync(t_first.p, 0xFF, sizeof(unsigned long long) * (size_t)(n_keys ? n_keys : 1), st));
    int64_t ne = 0;
    if (cells) {
        k_slot_weights<<<(unsigned)((cells + T - 1) / T), T, 0, st>>>(f->F, n, (const int32_t *)t_order.p, L, mode,
                                                                      (const int32_t *)t_iv.p, time_step, (int32_t *)t_W.p, (int32_t *)t_flag.p);
        DGE_LAUNCH_CHECK(ctx);
        int rc = dge_scan_i32(ctx, (const int32_t *)t_flag.p, (int32_t)cells, (int64_t *)t_pos.p);
        if (rc != DGE_OK) return rc;
        DGE_CUDA(ctx, cudaMemcpyAsync(&ne, (int64_t *)t_pos.p + cells, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        DGE_CUDA(ctx, cudaStreamSynchronize(st));
    }
    // the mention scan below runs over 2 * ne int32 flags: keep it inside int32 (a dense tensor of ~8192 regions could exceed it)
    if (2 * ne >= ((int64_t)1 << 31)) return dge_fail(ctx, DGE_E_LIMIT, "dge_crosstime_graph_build: more than 2^30 edges");
    flows_tmp t_sk(ctx), t_dk(ctx), t_w(ctx), t_mflag(ctx), t_mpos(ctx), t_vid(ctx);
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_sk.p, (size_t)ne));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_dk.p, (size_t)ne));
    DGE_CUDA(ctx, dge_malloc(ctx, (double **)&t_w.p, (size_t)ne));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_mflag.p, (size_t)ne * 2));
    DGE_CUDA(ctx, dge_malloc(ctx, (int64_t **)&t_mpos.p, (size_t)ne * 2 + 1));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_vid.p, (size_t)n_keys));
    int32_t *d_vl = nullptr, *d_vr = nullptr;
    DGE_CUDA(ctx, dge_malloc(ctx, &d_vl, (size_t)n_keys));
    if (dge_malloc(ctx, &d_vr, (size_t)n_keys) != cudaSuccess) { dge_free(ctx, d_vl); return dge_fail(ctx, DGE_E_CUDA, "dge_crosstime_graph_build: cudaMalloc"); }
    auto drop_labels = [&]() { dge_free(ctx, d_vl); dge_free(ctx, d_vr); };
    int64_t nv64 = 0;
    if (ne) {
        k_emit_edges<<<(unsigned)((cells + T - 1) / T), T, 0, st>>>((const int32_t *)t_W.p, (const int64_t *)t_pos.p, n,
                                                                    (const int32_t *)t_order.p, L, (int32_t *)t_sk.p, (int32_t *)t_dk.p,
                                                                    (double *)t_w.p, (unsigned long long *)t_first.p);
        ctx->launches++;
        k_mention_flags<<<(unsigned)((ne + T - 1) / T), T, 0, st>>>((const int32_t *)t_sk.p, (const int32_t *)t_dk.p, ne,
                                                                    (const unsigned long long *)t_first.p, (int32_t *)t_mflag.p);
        ctx->launches++;
        int rc = dge_scan_i32(ctx, (const int32_t *)t_mflag.p, (int32_t)(2 * ne), (int64_t *)t_mpos.p);
        if (rc != DGE_OK) { drop_labels(); return rc; }
        k_assign_ids<<<(n_keys + T - 1) / T, T, 0, st>>>((const unsigned long long *)t_first.p, n_keys, n, (const int64_t *)t_mpos.p,
                                                         (int32_t *)t_vid.p, d_vl, d_vr);
        ctx->launches++;
        k_keys_to_ids<<<(unsigned)((ne + T - 1) / T), T, 0, st>>>((int32_t *)t_sk.p, (int32_t *)t_dk.p, ne, (const int32_t *)t_vid.p);
        ctx->launches++;
        cudaMemcpyAsync(&nv64, (int64_t *)t_mpos.p + 2 * ne, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    }
    // sources: every region in iteration order whose layer-0 vertex exists (CrossTimeGraph.java:43-47, :86-90)
    std::vector<int32_t> vid0((size_t)n, -1), sources;
    if (ne && n) cudaMemcpyAsync(vid0.data(), t_vid.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st);
    cudaError_t ce = cudaStreamSynchronize(st);
    if (ce == cudaSuccess) ce = cudaGetLastError();
    if (ce != cudaSuccess) { drop_labels(); return dge_fail(ctx, DGE_E_CUDA, std::string("dge_crosstime_graph_build: ") + cudaGetErrorString(ce)); }
    for (int32_t a = 0; a < n; a++)
        if (vid0[order[a]] >= 0) sources.push_back(vid0[order[a]]);
    t_flow.stop();
    dge_graph *g = nullptr;
    int rc = dge_graph_build_device(ctx, (int32_t)nv64, ne, (const int32_t *)t_sk.p, (const int32_t *)t_dk.p, (const double *)t_w.p,
                                    (int32_t)sources.size(), sources.data(), nullptr, nullptr, &g);
    if (rc != DGE_OK) { drop_labels(); return rc; }
    g->v_layer = d_vl;
    g->v_region = d_vr;
    *out = g;
    return DGE_OK;
}

int dge_graph_labels(const dge_graph *g, int32_t *v_layer, int32_t *v_region_index, int32_t *sources) {
    if (!g) return dge_fail(nullptr, DGE_E_INVALID, "dge_graph_labels: graph is NULL");
    dge_ctx *ctx = g->ctx;
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((v_layer || v_region_index) && !g->v_layer)
        return dge_fail(ctx, DGE_E_INVALID, "dge_graph_labels: only graphs built by dge_crosstime_graph_build carry labels");
    cudaStream_t st = ctx->stream;
    if (v_layer && g->nv) DGE_CUDA(ctx, cudaMemcpyAsync(v_layer, g->v_layer, sizeof(int32_t) * (size_t)g->nv, cudaMemcpyDeviceToHost, st));
    if (v_region_index && g->nv) DGE_CUDA(ctx, cudaMemcpyAsync(v_region_index, g->v_region, sizeof(int32_t) * (size_t)g->nv, cudaMemcpyDeviceToHost, st));
    if (sources && g->ns) DGE_CUDA(ctx, cudaMemcpyAsync(sources, g->sources, sizeof(int32_t) * (size_t)g->ns, cudaMemcpyDeviceToHost, st));
    DGE_CUDA(ctx, cudaStreamSynchronize(st));
    return DGE_OK;
}

} // extern "C"
