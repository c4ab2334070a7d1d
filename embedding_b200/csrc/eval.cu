// eval.cu -- the reference's downstream pairwise-similarity metric on device.
//
// Reference behaviour reproduced (python/embeddingEvaluation_tract.py of the reference):
//   pairwiseEstimator :169-196   cosine distance of every pair of rows of one layer's embedding
//                                (scipy.spatial.distance.cosine, NaN -> 2 :186-189), the topk nearest OTHER rows
//   dcg_atK / ndcg_atK :249-260  relevance of neighbour j = 1 - ground-truth distance, discount 1 / log2(j + 2),
//                                normalised by the DCG of the ground truth's own ordering
//   generatePairWiseGT :63-103   the ground-truth ordering is the same kNN over the POI count vectors
// All arithmetic is fp64 as in numpy / scipy.  Neighbour order = ascending distance, ties by ascending row index
// (numpy's stable argsort).  O(m^2 dim) per layer: it only matters for the synthetic configs, where it closes the
// quality loop of the SGNS sweep without moving the embeddings to the host.
#include "dge_internal.cuh"

#define EVAL_THREADS 256

__global__ void k_row_norms(const float *__restrict__ X, int32_t m, int32_t dim, double *__restrict__ norm) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double s = 0.0;
    for (int32_t d = 0; d < dim; d++) { double v = (double)X[(int64_t)i * dim + d]; s += v * v; }
    norm[i] = sqrt(s);
}

// topk rounds of a block-wide argmin over D[0..m) with (distance, index) ordering; selected entries become +inf.
// Row i's own entry must already be +inf.  Fewer than topk finite entries: the tail is -1 / +inf.
__device__ void block_select_topk(double *D, int32_t m, int32_t i, int32_t topk, int32_t *__restrict__ nbr,
                                  double *__restrict__ ndist) {
    __shared__ double s_best[EVAL_THREADS / 32];
    __shared__ int32_t s_idx[EVAL_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int32_t r = 0; r < topk; r++) {
        double best = INFINITY;
        int32_t bi = 0x7fffffff;
        for (int32_t j = threadIdx.x; j < m; j += blockDim.x) {
            double v = D[j];
            if (v < best || (v == best && j < bi && v < INFINITY)) { best = v; bi = j; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, o);
            int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { s_best[wid] = best; s_idx[wid] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w2 = 1; w2 < EVAL_THREADS / 32; w2++)
                if (s_best[w2] < best || (s_best[w2] == best && s_idx[w2] < bi)) { best = s_best[w2]; bi = s_idx[w2]; }
            const bool found = bi != 0x7fffffff;
            nbr[(int64_t)i * topk + r] = found ? bi : -1;
            if (ndist) ndist[(int64_t)i * topk + r] = found ? best : INFINITY;
            if (found) D[bi] = INFINITY;
        }
        __syncthreads(); // the cleared entry is visible to every thread before the next round
    }
}

// One block per query row i of the chunk [row0, row0 + rows): distances to every row j into Dm (global scratch),
// then topk rounds of a block-wide argmin with (distance, index) ordering.
__global__ void __launch_bounds__(EVAL_THREADS)
k_knn_cosine(const float *__restrict__ X, const double *__restrict__ norm, int32_t m, int32_t dim, int32_t row0,
             int32_t topk, double *__restrict__ Dm, int32_t *__restrict__ nbr, double *__restrict__ ndist) {
    extern __shared__ unsigned char eval_smem[];
    double *xi = reinterpret_cast<double *>(eval_smem);                 // [dim]
    const int32_t i = row0 + blockIdx.x;
    double *D = Dm + (int64_t)blockIdx.x * m;
    for (int32_t d = threadIdx.x; d < dim; d += blockDim.x) xi[d] = (double)X[(int64_t)i * dim + d];
    __syncthreads();
    const double ni = norm[i];
    for (int32_t j = threadIdx.x; j < m; j += blockDim.x) {
        const float *xj = X + (int64_t)j * dim;
        double dot = 0.0;
        for (int32_t d = 0; d < dim; d++) dot += xi[d] * (double)xj[d];
        double dist = 1.0 - dot / (ni * norm[j]);
        if (!isfinite(dist)) dist = 2.0;                               // zero vector: the reference's NaN -> 2
        D[j] = j == i ? INFINITY : dist;                               // `if k2 == k: continue`
    }
    __syncthreads();
    block_select_topk(D, m, i, topk, nbr, ndist);
}

// dcg of every row against the ground truth: relv_j = 1 - gt[gi[i]][gi[nbr_j]], summed left to right
__global__ void k_dcg(const int32_t *__restrict__ nbr, int32_t m, int32_t topk, const int32_t *__restrict__ gi,
                      const double *__restrict__ gt, int32_t n, double *__restrict__ dcg) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double s = 0.0;
    const int32_t a = gi ? gi[i] : i;
    for (int32_t j = 0; j < topk; j++) {
        int32_t b = nbr[(int64_t)i * topk + j];
        if (b < 0) break;
        if (gi) b = gi[b];
        double relv = 1.0 - gt[(int64_t)a * n + b];
        s += relv / log2((double)(j + 2));
    }
    dcg[i] = s;
}

struct eval_tmp {
    dge_ctx *ctx;
    void *p = nullptr;
    explicit eval_tmp(dge_ctx *c) : ctx(c) {}
    ~eval_tmp() { dge_free(ctx, p); }
};

// kNN of the rows of device matrix d_X into device nbr (and optionally ndist)
static int knn_device(dge_ctx *ctx, const float *d_X, int32_t m, int32_t dim, int32_t topk, int32_t *d_nbr, double *d_ndist) {
    cudaStream_t st = ctx->stream;
    eval_tmp t_norm(ctx), t_D(ctx);
    DGE_CUDA(ctx, dge_malloc(ctx, (double **)&t_norm.p, (size_t)m));
    int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(m, ((int64_t)256 << 20) / ((int64_t)m * 8)));
    DGE_CUDA(ctx, dge_malloc(ctx, (double **)&t_D.p, (size_t)chunk * (size_t)m));
    k_row_norms<<<(m + 255) / 256, 256, 0, st>>>(d_X, m, dim, (double *)t_norm.p);
    DGE_LAUNCH_CHECK(ctx);
    size_t smem = sizeof(double) * (size_t)dim;
    for (int64_t r0 = 0; r0 < m; r0 += chunk) {
        int32_t rows = (int32_t)std::min<int64_t>(chunk, m - r0);
        k_knn_cosine<<<rows, EVAL_THREADS, smem, st>>>(d_X, (const double *)t_norm.p, m, dim, (int32_t)r0, topk, (double *)t_D.p,
                                                     d_nbr, d_ndist);
        DGE_LAUNCH_CHECK(ctx);
    }
    return DGE_OK;
}

// topk smallest entries of every row of a device distance matrix (diagonal excluded), ascending, ties by index
__global__ void __launch_bounds__(EVAL_THREADS)
k_select_rows(const double *__restrict__ Din, int32_t n, int32_t row0, int32_t topk, double *__restrict__ Dm, int32_t *__restrict__ nbr) {
    const int32_t i = row0 + blockIdx.x;
    double *D = Dm + (int64_t)blockIdx.x * n;
    for (int32_t j = threadIdx.x; j < n; j += blockDim.x) D[j] = j == i ? INFINITY : Din[(int64_t)i * n + j];
    __syncthreads();
    block_select_topk(D, n, i, topk, nbr, nullptr);
}

static int dge_eval_select_rows(dge_ctx *ctx, const double *d_D, int32_t n, int32_t topk, int32_t *d_nbr) {
    cudaStream_t st = ctx->stream;
    eval_tmp t_D(ctx);
    int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n, ((int64_t)256 << 20) / ((int64_t)n * 8)));
    DGE_CUDA(ctx, dge_malloc(ctx, (double **)&t_D.p, (size_t)chunk * (size_t)n));
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        int32_t rows = (int32_t)std::min<int64_t>(chunk, n - r0);
        k_select_rows<<<rows, EVAL_THREADS, 0, st>>>(d_D, n, (int32_t)r0, topk, (double *)t_D.p, d_nbr);
        DGE_LAUNCH_CHECK(ctx);
    }
    return DGE_OK;
}

extern "C" {

int dge_eval_knn(dge_ctx *ctx, const float *X, int32_t m, int32_t dim, int32_t topk, int32_t *nbr, double *dist) {
    if (!ctx) return dge_fail(nullptr, DGE_E_INVALID, "dge_eval_knn: ctx is NULL");
    if (m < 0 || dim < 1 || topk < 1 || (m > 0 && (!X || !nbr))) return dge_fail(ctx, DGE_E_INVALID, "dge_eval_knn: bad arguments");
    if (dim > 4096) return dge_fail(ctx, DGE_E_LIMIT, "dge_eval_knn: dim must be <= 4096");
    if (m == 0) return DGE_OK;
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    eval_tmp t_X(ctx), t_nbr(ctx), t_nd(ctx);
    DGE_CUDA(ctx, dge_malloc(ctx, (float **)&t_X.p, (size_t)m * dim));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_nbr.p, (size_t)m * topk));
    if (dist) DGE_CUDA(ctx, dge_malloc(ctx, (double **)&t_nd.p, (size_t)m * topk));
    DGE_CUDA(ctx, cudaMemcpyAsync(t_X.p, X, sizeof(float) * (size_t)m * dim, cudaMemcpyHostToDevice, st));
    dge_phase_timer t(ctx, "eval_knn");
    int rc = knn_device(ctx, (const float *)t_X.p, m, dim, topk, (int32_t *)t_nbr.p, (double *)t_nd.p);
    if (rc != DGE_OK) return rc;
    t.stop();
    DGE_CUDA(ctx, cudaMemcpyAsync(nbr, t_nbr.p, sizeof(int32_t) * (size_t)m * topk, cudaMemcpyDeviceToHost, st));
    if (dist) DGE_CUDA(ctx, cudaMemcpyAsync(dist, t_nd.p, sizeof(double) * (size_t)m * topk, cudaMemcpyDeviceToHost, st));
    DGE_CUDA(ctx, cudaStreamSynchronize(st));
    return DGE_OK;
}

int dge_eval_ndcg(dge_ctx *ctx, const float *X, int32_t m, int32_t dim, const int32_t *gt_index, const double *gt_dist,
                  int32_t n, int32_t topk, double *ndcg, double *mean) {
    if (!ctx) return dge_fail(nullptr, DGE_E_INVALID, "dge_eval_ndcg: ctx is NULL");
    if (m < 0 || n < 0 || dim < 1 || topk < 1 || (m > 0 && (!X || !gt_index || !gt_dist)))
        return dge_fail(ctx, DGE_E_INVALID, "dge_eval_ndcg: bad arguments");
    if (dim > 4096) return dge_fail(ctx, DGE_E_LIMIT, "dge_eval_ndcg: dim must be <= 4096");
    for (int32_t i = 0; i < m; i++)
        if (gt_index[i] < 0 || gt_index[i] >= n) return dge_fail(ctx, DGE_E_INVALID, "dge_eval_ndcg: gt_index out of range");
    if (mean) *mean = NAN;
    if (m <= topk) return DGE_OK; // the reference needs topk other regions in the layer; nothing to report
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    eval_tmp t_X(ctx), t_nbr(ctx), t_gi(ctx), t_gt(ctx), t_gnbr(ctx), t_dcg(ctx), t_max(ctx);
    DGE_CUDA(ctx, dge_malloc(ctx, (float **)&t_X.p, (size_t)m * dim));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_nbr.p, (size_t)m * topk));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_gi.p, (size_t)m));
    DGE_CUDA(ctx, dge_malloc(ctx, (double **)&t_gt.p, (size_t)n * n));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_gnbr.p, (size_t)n * topk));
    DGE_CUDA(ctx, dge_malloc(ctx, (double **)&t_dcg.p, (size_t)m));
    DGE_CUDA(ctx, dge_malloc(ctx, (double **)&t_max.p, (size_t)n));
    DGE_CUDA(ctx, cudaMemcpyAsync(t_X.p, X, sizeof(float) * (size_t)m * dim, cudaMemcpyHostToDevice, st));
    DGE_CUDA(ctx, cudaMemcpyAsync(t_gi.p, gt_index, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, st));
    DGE_CUDA(ctx, cudaMemcpyAsync(t_gt.p, gt_dist, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice, st));
    dge_phase_timer t(ctx, "eval_ndcg");
    int rc = knn_device(ctx, (const float *)t_X.p, m, dim, topk, (int32_t *)t_nbr.p, nullptr);
    if (rc != DGE_OK) return rc;
    k_dcg<<<(m + 255) / 256, 256, 0, st>>>((const int32_t *)t_nbr.p, m, topk, (const int32_t *)t_gi.p, (const double *)t_gt.p, n,
                                          (double *)t_dcg.p);
    DGE_LAUNCH_CHECK(ctx);
    // ideal DCG: the ground truth's own topk ordering (generatePairWiseGT :63-103 + dcg_atK of :302-304)
    {
        // the rows of gt are already distances: select on them
        rc = dge_eval_select_rows(ctx, (const double *)t_gt.p, n, topk, (int32_t *)t_gnbr.p);
        if (rc != DGE_OK) return rc;
    }
    k_dcg<<<(n + 255) / 256, 256, 0, st>>>((const int32_t *)t_gnbr.p, n, topk, nullptr, (const double *)t_gt.p, n, (double *)t_max.p);
    DGE_LAUNCH_CHECK(ctx);
    t.stop();
    std::vector<double> dcg((size_t)m), mx((size_t)n + 1);
    DGE_CUDA(ctx, cudaMemcpyAsync(dcg.data(), t_dcg.p, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, st));
    DGE_CUDA(ctx, cudaMemcpyAsync(mx.data(), t_max.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    DGE_CUDA(ctx, cudaStreamSynchronize(st));
    double s = 0.0;
    for (int32_t i = 0; i < m; i++) {
        double v = dcg[i] / mx[gt_index[i]];
        if (ndcg) ndcg[i] = v;
        s += v;
    }
    if (mean) *mean = s / (double)m;
    return DGE_OK;
}

} // extern "C"

