// graph.cu -- stage 1a on device: COO -> CSR (insertion order kept), out-degrees, per-row and
// source alias tables bit-identical to the reference's Java, and the packed 32 B walk records.
//
// Reference behaviour reproduced (paths under embedding/src/main/java/embedding/):
//   LayeredGraph.addEdge :157-174 / Vertex.addOutEdge :46-49   -> CSR rows in insertion order,
//                                                                  outDegree = left-to-right sum
//   Vertex.initiateAliasTable :54-82                            -> k_alias_small / k_alias_big
//   LayeredGraph.initiateAliasTables :195-226                   -> same kernels on the source list
// Compile this file with --fmad=false; the table arithmetic additionally uses explicit
// round-to-nearest intrinsics so no contraction can change a rounding.
#include "dge_internal.cuh"
#include <algorithm>

#define ALIAS_SMALL_MAX 1024            // rows up to this size keep their S/G bitmaps in registers
#define ALIAS_BIG_MAX (1 << 25)         // documented limit of the hierarchical-bitmap kernel

// ------------------------------------------------------------------ CSR construction

// Validates ids, histograms sources, counts runs of equal consecutive sources and records the
// first edge of each run.  If #runs == #non-empty rows every row is one contiguous run ("grouped"
// input, which is what CrossTimeGraph / SpatialGraph emit) and placement needs no sort.
__global__ void k_count(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int64_t ne, int32_t nv,
                        int32_t *__restrict__ deg, int64_t *__restrict__ row_first, unsigned long long *runs,
                        int *bad) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long my_runs = 0;
    for (; e < ne; e += stride) {
        int32_t s = src[e], d = dst[e];
        if (s < 0 || s >= nv || d < 0 || d >= nv) { *bad = 1; continue; }
        atomicAdd(&deg[s], 1);
        if (e == 0 || src[e - 1] != s) { my_runs++; row_first[s] = e; }
    }
    for (int o = 16; o; o >>= 1) my_runs += __shfl_xor_sync(0xffffffffu, my_runs, o);
    if ((threadIdx.x & 31) == 0 && my_runs) atomicAdd(runs, my_runs);
}

#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__global__ void k_scan_tile_sums(const int32_t *__restrict__ deg, int32_t n, int64_t *tile_sums, int32_t *nonempty) {
    __shared__ int64_t sh[SCAN_THREADS / 32];
    __shared__ int shc[SCAN_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t s = 0;
    int c = 0;
    for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_THREADS) {
        int64_t j = base + i;
        if (j < n) { int32_t d = deg[j]; s += d; c += d > 0; }
    }
    for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = s; shc[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0; int tc = 0;
        for (int i = 0; i < SCAN_THREADS / 32; i++) { t += sh[i]; tc += shc[i]; }
        tile_sums[blockIdx.x] = t;
        if (tc) atomicAdd(nonempty, tc);
    }
}

// single block: exclusive scan of the tile sums in place
__global__ void k_scan_sums(int64_t *tile_sums, int32_t n_tiles) {
    __shared__ int64_t sh[1024];
    int per = (n_tiles + 1023) / 1024;
    int lo = threadIdx.x * per, hi = min(lo + per, n_tiles);
    int64_t s = 0;
    for (int i = lo; i < hi; i++) s += tile_sums[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = 0;
        for (int i = 0; i < 1024; i++) { int64_t t = sh[i]; sh[i] = run; run += t; }
    }
    __syncthreads();
    int64_t run = sh[threadIdx.x];
    for (int i = lo; i < hi; i++) { int64_t t = tile_sums[i]; tile_sums[i] = run; run += t; }
}

__global__ void k_scan_apply(const int32_t *__restrict__ deg, int32_t n, const int64_t *__restrict__ tile_off,
                             int64_t *__restrict__ row_ptr) {
    __shared__ int64_t warp_tot[SCAN_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = (base + i < n) ? deg[base + i] : 0; s += v[i]; }
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int64_t incl = s;
    for (int o = 1; o < 32; o <<= 1) { int64_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int64_t woff = 0;
    for (int i = 0; i < wid; i++) woff += warp_tot[i];
    int64_t run = tile_off[blockIdx.x] + woff + incl - s;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) row_ptr[base + i] = run;
        run += v[i];
    }
    if (base <= (int64_t)n - 1 && base + SCAN_ITEMS > (int64_t)n - 1) row_ptr[n] = run; // thread owning the last row
}

__global__ void k_place_grouped(const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                                const double *__restrict__ w, int64_t ne, const int64_t *__restrict__ row_ptr,
                                const int64_t *__restrict__ row_first, int32_t *__restrict__ col,
                                double *__restrict__ wc) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; e < ne; e += stride) {
        int32_t s = src[e];
        int64_t pos = row_ptr[s] + (e - row_first[s]);
        col[pos] = dst[e];
        wc[pos] = w[e];
    }
}

// generic input: claim a slot in the row (arbitrary order), then sort each row's slots by edge id,
// which restores insertion order deterministically.
__global__ void k_place_atomic(const int32_t *__restrict__ src, int64_t ne, const int64_t *__restrict__ row_ptr,
                               int32_t *__restrict__ cursor, int32_t *__restrict__ perm) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; e < ne; e += stride) {
        int32_t s = src[e];
        int32_t p = atomicAdd(&cursor[s], 1);
        perm[row_ptr[s] + p] = (int32_t)e;
    }
}

// warp per row: in-place normalised bitonic sort (all compare-exchanges ascending, so a virtual
// +inf padding to the next power of two never has to move).
__global__ void k_sort_rows(const int64_t *__restrict__ row_ptr, int32_t nv, int32_t *__restrict__ perm) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < nv; row += nwarps) {
        int64_t b = row_ptr[row];
        int64_t k = row_ptr[row + 1] - b;
        if (k < 2) continue;
        int32_t *a = perm + b;
        int64_t n2 = 1;
        while (n2 < k) n2 <<= 1;
        for (int64_t size = 2; size <= n2; size <<= 1) {
            for (int64_t stride = size >> 1; stride > 0; stride >>= 1) {
                bool first = (stride == (size >> 1));
                for (int64_t t = lane; t < (n2 >> 1); t += 32) {
                    // t enumerates the lower element of each pair
                    int64_t lo = ((t / stride) * (stride << 1)) + (t % stride);
                    int64_t hi = first ? (lo ^ (size - 1)) : (lo + stride);
                    if (first) { // mirror partner inside the size-block
                        int64_t blk = lo & ~(size - 1);
                        hi = blk + (size - 1) - (lo - blk);
                    }
                    if (hi < k && lo < k) {
                        int32_t x = a[lo], y = a[hi];
                        if (x > y) { a[lo] = y; a[hi] = x; }
                    }
                }
                __syncwarp();
            }
        }
    }
}

__global__ void k_gather(const int32_t *__restrict__ perm, const int32_t *__restrict__ dst,
                         const double *__restrict__ w, int64_t ne, int32_t *__restrict__ col,
                         double *__restrict__ wc) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < ne; j += stride) {
        int32_t e = perm[j];
        col[j] = dst[e];
        wc[j] = w[e];
    }
}

// Vertex.addOutEdge :46-49: outDegree += weight, strictly left to right (fp add is not associative).
__global__ void k_out_degree(const int64_t *__restrict__ row_ptr, int32_t nv, const double *__restrict__ wc,
                             double *__restrict__ od) {
    int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; v < nv; v += stride) {
        double s = 0.0;
        for (int64_t j = row_ptr[v]; j < row_ptr[v + 1]; j++) s = __dadd_rn(s, wc[j]);
        od[v] = s;
    }
}

__global__ void k_gather_src_w(const int32_t *__restrict__ sources, int32_t ns, int32_t nv,
                               const double *__restrict__ od, double *__restrict__ sw, int *bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ns) {
        int32_t v = sources[i];
        if (v < 0 || v >= nv) { *bad = 1; sw[i] = 0; } else sw[i] = od[v];
    }
}
// addSourceVertex :188: sourceWeightSum += v.outDegree in list order
__global__ void k_seq_sum(const double *__restrict__ a, int32_t n, double *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int32_t i = 0; i < n; i++) s = __dadd_rn(s, a[i]);
        *out = s;
    }
}

// ------------------------------------------------------------------ alias tables
//
// Exact ordered-set form of the reference's O(k^2) greedy pairing (LayeredGraph.java:65-81).
// An entry is small (prob<1, alias unset), large (prob>1) or inert.  Small entries never become
// large, large entries only shrink, an alias is written once.  The Java inner loop therefore is:
//   large l1 : repeatedly take the MINIMUM small entry l2, alias[l2]=l1, prob[l1]-=1-prob[l2],
//              while prob[l1]>1; if it ends <1 at scan position q it continues from q+1 and stops
//              at the first large l2: alias[l1]=l2, prob[l2]-=1-prob[l1].
//   small l1 : first large l2 from 0: alias[l1]=l2, prob[l2]-=1-prob[l1].
// The double subtractions happen in the same order with the same operands => identical bits.
// (oracle/dge_oracle.c proves the equivalence against the literal double loop.)

// One warp per row, k <= 1024: lane j owns the 32-bit S (small) and G (large) words of entries
// 32j..32j+31.  Control flow is warp-uniform; prob/alias live in global memory in place.
__global__ void __launch_bounds__(256)
k_alias_small(const int64_t *__restrict__ row_ptr, int32_t n_rows, const double *__restrict__ w,
              const double *__restrict__ od, double *prob, int32_t *alias, int32_t *big_rows, int32_t *n_big) {
    const unsigned FULL = 0xffffffffu;
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < n_rows; row += nwarps) {
        int64_t b = row_ptr[row];
        int64_t k64 = row_ptr[row + 1] - b;
        if (k64 == 0) continue;
        int32_t k = (int32_t)k64;
        double odv = od[row];
        double kd = (double)k;
        // :59-62  probTable[i] = k * w / outDegree ; aliasTable = -1
        for (int32_t i = lane; i < k; i += 32) {
            prob[b + i] = __ddiv_rn(__dmul_rn(kd, w[b + i]), odv);
            alias[b + i] = -1;
        }
        if (k > ALIAS_SMALL_MAX) {
            if (lane == 0) big_rows[atomicAdd(n_big, 1)] = (int32_t)row;
            continue;
        }
        __syncwarp();
        uint32_t sw = 0, gw = 0;
        for (int32_t c = 0; c * 32 < k; c++) {
            int32_t i = c * 32 + lane;
            double p = i < k ? prob[b + i] : 1.0;
            uint32_t sm = __ballot_sync(FULL, p < 1.0), gm = __ballot_sync(FULL, p > 1.0);
            if (lane == c) { sw = sm; gw = gm; }
        }
        for (int32_t l1 = 0; l1 < k; l1++) {
            int wl1 = l1 >> 5;
            uint32_t bit1 = 1u << (l1 & 31);
            uint32_t s1 = __shfl_sync(FULL, sw, wl1), g1 = __shfl_sync(FULL, gw, wl1);
            if (g1 & bit1) {
                double p1 = prob[b + l1];
                int32_t pos = -1;
                bool exhausted = false;
                while (p1 > 1.0) {
                    uint32_t nz = __ballot_sync(FULL, sw != 0);
                    if (!nz) { exhausted = true; break; }
                    int wl = __ffs(nz) - 1;
                    uint32_t word = __shfl_sync(FULL, sw, wl);
                    int bt = __ffs(word) - 1;
                    int32_t l2 = wl * 32 + bt;
                    double p2 = prob[b + l2];
                    if (lane == 0) alias[b + l2] = l1;
                    p1 = __dsub_rn(p1, __dsub_rn(1.0, p2));
                    if (lane == wl) sw &= ~(1u << bt);
                    pos = l2;
                }
                if (lane == 0) prob[b + l1] = p1;
                if (exhausted) { __syncwarp(); continue; }
                if (lane == wl1) gw &= ~bit1;
                if (p1 < 1.0) {
                    int32_t q = pos + 1;
                    int qw = q >> 5;
                    uint32_t m = gw;
                    if (lane < qw) m = 0;
                    else if (lane == qw) m &= (FULL << (q & 31));
                    uint32_t nz = __ballot_sync(FULL, m != 0);
                    if (nz) {
                        int wl = __ffs(nz) - 1;
                        uint32_t word = __shfl_sync(FULL, m, wl);
                        int bt = __ffs(word) - 1;
                        int32_t l2 = wl * 32 + bt;
                        double pl2 = __dsub_rn(prob[b + l2], __dsub_rn(1.0, p1));
                        if (lane == 0) { alias[b + l1] = l2; prob[b + l2] = pl2; }
                        if (!(pl2 > 1.0) && lane == wl) {
                            gw &= ~(1u << bt);
                            if (pl2 < 1.0) sw |= (1u << bt);
                        }
                    } else if (lane == wl1) {
                        sw |= bit1; // dangling small entry: a later large entry may still take it
                    }
                }
                __syncwarp();
            } else if (s1 & bit1) {
                uint32_t nz = __ballot_sync(FULL, gw != 0);
                if (nz) {
                    int wl = __ffs(nz) - 1;
                    uint32_t word = __shfl_sync(FULL, gw, wl);
                    int bt = __ffs(word) - 1;
                    int32_t l2 = wl * 32 + bt;
                    double p1 = prob[b + l1];
                    double pl2 = __dsub_rn(prob[b + l2], __dsub_rn(1.0, p1));
                    if (lane == 0) { alias[b + l1] = l2; prob[b + l2] = pl2; }
                    if (lane == wl1) sw &= ~bit1;
                    if (!(pl2 > 1.0) && lane == wl) {
                        gw &= ~(1u << bt);
                        if (pl2 < 1.0) sw |= (1u << bt);
                    }
                    __syncwarp();
                }
            }
        }
    }
}

// Hierarchical 64-bit bitmap in global scratch (ordered set with successor queries).
struct hset_dev {
    unsigned long long *lv[6];
    int64_t nw[6];
    int levels;
};
__host__ __device__ static inline int64_t hset_words(int64_t k) {
    int64_t total = 0, bits = k > 0 ? k : 1;
    do { int64_t wds = (bits + 63) >> 6; total += wds; bits = wds; } while (bits > 1);
    return total;
}
__device__ static inline void hset_bind(hset_dev &h, unsigned long long *base, int64_t k) {
    h.levels = 0;
    int64_t bits = k > 0 ? k : 1;
    do {
        int64_t wds = (bits + 63) >> 6;
        h.lv[h.levels] = base; h.nw[h.levels] = wds;
        base += wds; h.levels++; bits = wds;
    } while (bits > 1);
}
__device__ static inline void hset_add(hset_dev &h, int64_t i) {
    for (int l = 0; l < h.levels; l++) { h.lv[l][i >> 6] |= 1ULL << (i & 63); i >>= 6; }
}
__device__ static inline void hset_del(hset_dev &h, int64_t i) {
    for (int l = 0; l < h.levels; l++) {
        unsigned long long v = h.lv[l][i >> 6] & ~(1ULL << (i & 63));
        h.lv[l][i >> 6] = v;
        if (v) break;
        i >>= 6;
    }
}
__device__ static inline int64_t hset_succ(const hset_dev &h, int64_t p) {
    int64_t pos = p;
    for (int l = 0; l < h.levels; l++) {
        int64_t wi = pos >> 6;
        if (wi >= h.nw[l]) return -1;
        unsigned long long m = h.lv[l][wi] & (~0ULL << (pos & 63));
        if (m) {
            int64_t idx = (wi << 6) + (__ffsll((long long)m) - 1);
            for (int d = l - 1; d >= 0; d--) idx = (idx << 6) + (__ffsll((long long)h.lv[d][idx]) - 1);
            return idx;
        }
        pos = wi + 1;
    }
    return -1;
}

// One warp per big row (k > 1024): lanes classify in parallel, lane 0 runs the sequential pairing
// on the hierarchical bitmaps.  scratch_off[r] is the row's offset (in 64-bit words) into scratch;
// each row owns 2 * hset_words(k) zero-initialised words.
__global__ void __launch_bounds__(32)
k_alias_big(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ big_rows, int32_t n_big,
            const int64_t *__restrict__ scratch_off, unsigned long long *scratch, double *prob, int32_t *alias) {
    int r = blockIdx.x;
    if (r >= n_big) return;
    int lane = threadIdx.x;
    int32_t row = big_rows[r];
    int64_t b = row_ptr[row];
    int64_t k = row_ptr[row + 1] - b;
    hset_dev S, G;
    int64_t words = hset_words(k);
    hset_bind(S, scratch + scratch_off[r], k);
    hset_bind(G, scratch + scratch_off[r] + words, k);
    double *p = prob + b;
    int32_t *al = alias + b;
    for (int64_t i = lane; i < k; i += 32) {
        double v = p[i];
        if (v < 1.0) atomicOr(&S.lv[0][i >> 6], 1ULL << (i & 63));
        else if (v > 1.0) atomicOr(&G.lv[0][i >> 6], 1ULL << (i & 63));
    }
    __syncwarp();
    for (int l = 1; l < S.levels; l++) {
        for (int64_t i = lane; i < S.nw[l - 1]; i += 32) {
            if (S.lv[l - 1][i]) atomicOr(&S.lv[l][i >> 6], 1ULL << (i & 63));
            if (G.lv[l - 1][i]) atomicOr(&G.lv[l][i >> 6], 1ULL << (i & 63));
        }
        __syncwarp();
    }
    if (lane != 0) return;
    for (int64_t l1 = 0; l1 < k; l1++) {
        double p1 = p[l1];
        if (!(p1 != 1.0 && al[l1] == -1)) continue;
        if (p1 > 1.0) {
            int64_t pos = -1;
            bool exhausted = false;
            while (p1 > 1.0) {
                int64_t l2 = hset_succ(S, 0);
                if (l2 < 0) { exhausted = true; break; }
                al[l2] = (int32_t)l1;
                p1 = __dsub_rn(p1, __dsub_rn(1.0, p[l2]));
                hset_del(S, l2);
                pos = l2;
            }
            p[l1] = p1;
            if (exhausted) continue;
            hset_del(G, l1);
            if (p1 < 1.0) {
                int64_t l2 = hset_succ(G, pos + 1);
                if (l2 >= 0) {
                    al[l1] = (int32_t)l2;
                    double pl2 = __dsub_rn(p[l2], __dsub_rn(1.0, p1));
                    p[l2] = pl2;
                    if (!(pl2 > 1.0)) { hset_del(G, l2); if (pl2 < 1.0) hset_add(S, l2); }
                } else {
                    hset_add(S, l1);
                }
            }
        } else if (p1 < 1.0) {
            int64_t l2 = hset_succ(G, 0);
            if (l2 >= 0) {
                al[l1] = (int32_t)l2;
                double pl2 = __dsub_rn(p[l2], __dsub_rn(1.0, p1));
                p[l2] = pl2;
                hset_del(S, l1);
                if (!(pl2 > 1.0)) { hset_del(G, l2); if (pl2 < 1.0) hset_add(S, l2); }
            }
        }
    }
}

// ------------------------------------------------------------------ packed walk records

__global__ void k_pack_rows(const int64_t *__restrict__ row_ptr, int32_t nv, const int32_t *__restrict__ col,
                            const double *__restrict__ prob, const int32_t *__restrict__ alias,
                            dge_edge_rec *__restrict__ rec) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < nv; row += nwarps) {
        int64_t b = row_ptr[row], e = row_ptr[row + 1];
        for (int64_t j = b + lane; j < e; j += 32) {
            int32_t a = alias[j];
            int32_t d0 = col[j];
            int32_t d1 = a < 0 ? d0 : col[b + a];
            dge_edge_rec r;
            r.prob = prob[j];
            r.dst = d0; r.adst = d1;
            r.start0 = (uint32_t)row_ptr[d0]; r.deg0 = (uint32_t)(row_ptr[d0 + 1] - row_ptr[d0]);
            r.start1 = (uint32_t)row_ptr[d1]; r.deg1 = (uint32_t)(row_ptr[d1 + 1] - row_ptr[d1]);
            rec[j] = r;
        }
    }
}

__global__ void k_pack_sources(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ sources, int32_t ns,
                               const double *__restrict__ sprob, const int32_t *__restrict__ salias,
                               dge_edge_rec *__restrict__ srec) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    int32_t a = salias[i];
    int32_t v0 = sources[i];
    int32_t v1 = a < 0 ? v0 : sources[a];
    dge_edge_rec r;
    r.prob = sprob[i];
    r.dst = v0; r.adst = v1;
    r.start0 = (uint32_t)row_ptr[v0]; r.deg0 = (uint32_t)(row_ptr[v0 + 1] - row_ptr[v0]);
    r.start1 = (uint32_t)row_ptr[v1]; r.deg1 = (uint32_t)(row_ptr[v1 + 1] - row_ptr[v1]);
    srec[i] = r;
}

// ------------------------------------------------------------------ batched single draws (test hook)

// LayeredGraph.java:107-115: i=(int)(x*k); y=x*k-i; y<prob[i] ? column i : alias column
__device__ static inline const dge_edge_rec *alias_pick(const dge_edge_rec *row, uint32_t k, double x, bool &first) {
    double xk = __dmul_rn(x, (double)k);
    int32_t i = __double2int_rz(xk);
    double y = __dsub_rn(xk, (double)i);
    const dge_edge_rec *r = row + i;
    first = y < r->prob;
    return r;
}

__global__ void k_sample_next(const dge_edge_rec *__restrict__ rec, const dge_edge_rec *__restrict__ srec, int32_t ns,
                              int32_t nv, const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                              const double *__restrict__ wc, const double *__restrict__ od,
                              const int32_t *__restrict__ sources, const double *__restrict__ sws, int64_t n,
                              const int32_t *__restrict__ v, const double *__restrict__ x, int sampler,
                              int32_t *__restrict__ out) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    int32_t u = v[q];
    double xx = x[q];
    int32_t res = -1;
    if (u >= nv) { out[q] = -2; return; }
    if (sampler == DGE_SAMPLER_ALIAS) {
        bool first;
        if (u < 0) {
            if (ns > 0) { const dge_edge_rec *r = alias_pick(srec, (uint32_t)ns, xx, first); res = first ? r->dst : r->adst; }
        } else {
            int64_t b = row_ptr[u];
            uint32_t k = (uint32_t)(row_ptr[u + 1] - b);
            if (k) { const dge_edge_rec *r = alias_pick(rec + b, k, xx, first); res = first ? r->dst : r->adst; }
        }
    } else { // LayeredGraph.java:89-98 / :261-270
        if (u < 0) {
            double s = __dmul_rn(xx, *sws), cnt = 0.0;
            for (int32_t i = 0; i < ns; i++) {
                cnt = __dadd_rn(cnt, od[sources[i]]);
                if (cnt >= s) { res = sources[i]; break; }
            }
        } else {
            double s = __dmul_rn(xx, od[u]), cnt = 0.0;
            for (int64_t j = row_ptr[u]; j < row_ptr[u + 1]; j++) {
                cnt = __dadd_rn(cnt, wc[j]);
                if (cnt >= s) { res = col[j]; break; }
            }
        }
    }
    out[q] = res;
}

// ------------------------------------------------------------------ host side

__global__ void k_big_sizes(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ big_rows, int32_t n_big,
                            int64_t *__restrict__ k_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_big) k_out[i] = row_ptr[big_rows[i] + 1] - row_ptr[big_rows[i]];
}

struct dev_tmp { // device scratch returned to the ctx pool on scope exit
    dge_ctx *ctx;
    void *p = nullptr;
    explicit dev_tmp(dge_ctx *c) : ctx(c) {}
    ~dev_tmp() { if (p) dge_free(ctx, p); }
};

int dge_scan_i32(dge_ctx *ctx, const int32_t *d_in, int32_t n, int64_t *d_pos) {
    cudaStream_t st = ctx->stream;
    if (n <= 0) {
        DGE_CUDA(ctx, cudaMemsetAsync(d_pos, 0, sizeof(int64_t), st));
        return DGE_OK;
    }
    int32_t n_tiles = (int32_t)(((int64_t)n + SCAN_TILE - 1) / SCAN_TILE);
    dev_tmp t_tiles(ctx), t_cnt(ctx);
    DGE_CUDA(ctx, dge_malloc(ctx, (int64_t **)&t_tiles.p, (size_t)n_tiles));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_cnt.p, 1));
    DGE_CUDA(ctx, cudaMemsetAsync(t_cnt.p, 0, sizeof(int32_t), st));
    k_scan_tile_sums<<<n_tiles, SCAN_THREADS, 0, st>>>(d_in, n, (int64_t *)t_tiles.p, (int32_t *)t_cnt.p);
    DGE_LAUNCH_CHECK(ctx);
    k_scan_sums<<<1, 1024, 0, st>>>((int64_t *)t_tiles.p, n_tiles);
    DGE_LAUNCH_CHECK(ctx);
    k_scan_apply<<<n_tiles, SCAN_THREADS, 0, st>>>(d_in, n, (const int64_t *)t_tiles.p, d_pos);
    DGE_LAUNCH_CHECK(ctx);
    return DGE_OK;
}

static int grid_for(int64_t n, int threads, int sm_count, int per_sm = 8) {
    int64_t g = (n + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// Runs the alias kernels over a CSR-shaped problem (row_ptr/w/od -> prob/alias).
static int run_alias(dge_ctx *ctx, const int64_t *d_row_ptr, int32_t n_rows, int64_t n_entries, const double *d_w,
                     const double *d_od, double *d_prob, int32_t *d_alias) {
    if (n_rows == 0 || n_entries == 0) return DGE_OK;
    int32_t *d_big = nullptr, *d_nbig = nullptr;
    int64_t max_big = n_entries / (ALIAS_SMALL_MAX + 1) + 1;
    DGE_CUDA(ctx, dge_malloc(ctx, &d_big, (size_t)max_big));
    DGE_CUDA(ctx, dge_malloc(ctx, &d_nbig, 1));
    DGE_CUDA(ctx, cudaMemsetAsync(d_nbig, 0, sizeof(int32_t), ctx->stream));
    int threads = 256;
    int64_t warps_needed = n_rows;
    int grid = grid_for(warps_needed * 32, threads, ctx->sm_count, 8);
    k_alias_small<<<grid, threads, 0, ctx->stream>>>(d_row_ptr, n_rows, d_w, d_od, d_prob, d_alias, d_big, d_nbig);
    DGE_LAUNCH_CHECK(ctx);
    int32_t n_big = 0;
    DGE_CUDA(ctx, cudaMemcpyAsync(&n_big, d_nbig, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    DGE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int rc = DGE_OK;
    if (n_big > 0) {
        std::vector<int32_t> big(n_big);
        DGE_CUDA(ctx, cudaMemcpy(big.data(), d_big, sizeof(int32_t) * n_big, cudaMemcpyDeviceToHost));
        std::sort(big.begin(), big.end());
        // sizes of the big rows in one transfer
        dev_tmp t_k(ctx);
        DGE_CUDA(ctx, dge_malloc(ctx, (int64_t **)&t_k.p, (size_t)n_big));
        DGE_CUDA(ctx, cudaMemcpyAsync(d_big, big.data(), sizeof(int32_t) * n_big, cudaMemcpyHostToDevice, ctx->stream));
        k_big_sizes<<<(n_big + 255) / 256, 256, 0, ctx->stream>>>(d_row_ptr, d_big, n_big, (int64_t *)t_k.p);
        DGE_LAUNCH_CHECK(ctx);
        std::vector<int64_t> ks(n_big), off(n_big);
        DGE_CUDA(ctx, cudaMemcpyAsync(ks.data(), t_k.p, sizeof(int64_t) * n_big, cudaMemcpyDeviceToHost, ctx->stream));
        DGE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        int64_t total = 0;
        for (int32_t i = 0; i < n_big && rc == DGE_OK; i++) {
            if (ks[i] > ALIAS_BIG_MAX)
                rc = dge_fail(ctx, DGE_E_LIMIT, "alias table: a row (or the source list) has more than 2^25 entries");
            off[i] = total;
            total += 2 * hset_words(ks[i]);
        }
        if (rc == DGE_OK) {
            unsigned long long *d_scratch = nullptr;
            int64_t *d_off = nullptr;
            DGE_CUDA(ctx, dge_malloc(ctx, &d_scratch, (size_t)total));
            DGE_CUDA(ctx, dge_malloc(ctx, &d_off, (size_t)n_big));
            DGE_CUDA(ctx, cudaMemsetAsync(d_scratch, 0, sizeof(unsigned long long) * (size_t)total, ctx->stream));
            DGE_CUDA(ctx, cudaMemcpyAsync(d_off, off.data(), sizeof(int64_t) * n_big, cudaMemcpyHostToDevice, ctx->stream));
            k_alias_big<<<n_big, 32, 0, ctx->stream>>>(d_row_ptr, d_big, n_big, d_off, d_scratch, d_prob, d_alias);
            DGE_LAUNCH_CHECK(ctx);
            DGE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            dge_free(ctx, d_scratch);
            dge_free(ctx, d_off);
        }
    }
    dge_free(ctx, d_big);
    dge_free(ctx, d_nbig);
    return rc;
}

static void graph_release(dge_graph *g) {
    if (!g) return;
    dge_free(g->ctx, g->row_ptr); dge_free(g->ctx, g->col); dge_free(g->ctx, g->w); dge_free(g->ctx, g->prob); dge_free(g->ctx, g->alias);
    dge_free(g->ctx, g->out_degree); dge_free(g->ctx, g->sources); dge_free(g->ctx, g->src_w); dge_free(g->ctx, g->src_prob);
    dge_free(g->ctx, g->src_alias); dge_free(g->ctx, g->sws); dge_free(g->ctx, g->rec); dge_free(g->ctx, g->srec);
    dge_free(g->ctx, g->v_layer); dge_free(g->ctx, g->v_region);
    dge_delete_handle(g);
}

struct graph_guard { // frees a half-built graph on an early error return
    dge_graph *g;
    ~graph_guard() { if (g) graph_release(g); }
};
// Core of dge_graph_build over a DEVICE COO (d_src / d_dst / d_w stay owned by the caller; sources, out_degree and
// source_weight_sum are host pointers).  Also used by flows.cu, which enumerates the CrossTimeGraph edges on device.
int dge_graph_build_device(dge_ctx *ctx, int32_t nv, int64_t ne, const int32_t *d_src, const int32_t *d_dst,
                           const double *d_w, int32_t ns, const int32_t *sources, const double *out_degree,
                           const double *source_weight_sum, dge_graph **out) {
    *out = nullptr;
    if (ne >= (int64_t)1 << 31) return dge_fail(ctx, DGE_E_LIMIT, "dge_graph_build: n_edges must be < 2^31");
    if (ne > 0 && nv == 0) return dge_fail(ctx, DGE_E_INVALID, "dge_graph_build: edges but no vertices");
    if (ns > ALIAS_BIG_MAX) return dge_fail(ctx, DGE_E_LIMIT, "dge_graph_build: more than 2^25 source vertices");
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    dge_graph *g = dge_new_handle<dge_graph>(ctx);
    graph_guard guard{g};
    g->ctx = ctx; g->nv = nv; g->ne = ne; g->ns = ns;

    dge_phase_timer t_csr(ctx, "csr");
    DGE_CUDA(ctx, dge_malloc(ctx, &g->row_ptr, (size_t)nv + 1));
    DGE_CUDA(ctx, dge_malloc(ctx, &g->col, (size_t)ne));
    DGE_CUDA(ctx, dge_malloc(ctx, &g->w, (size_t)ne));
    DGE_CUDA(ctx, dge_malloc(ctx, &g->prob, (size_t)ne));
    DGE_CUDA(ctx, dge_malloc(ctx, &g->alias, (size_t)ne));
    DGE_CUDA(ctx, dge_malloc(ctx, &g->out_degree, (size_t)nv));
    DGE_CUDA(ctx, dge_malloc(ctx, &g->sources, (size_t)ns));
    DGE_CUDA(ctx, dge_malloc(ctx, &g->src_w, (size_t)ns));
    DGE_CUDA(ctx, dge_malloc(ctx, &g->src_prob, (size_t)ns));
    DGE_CUDA(ctx, dge_malloc(ctx, &g->src_alias, (size_t)ns));
    DGE_CUDA(ctx, dge_malloc(ctx, &g->sws, 1));
    DGE_CUDA(ctx, dge_malloc(ctx, &g->rec, (size_t)ne));
    DGE_CUDA(ctx, dge_malloc(ctx, &g->srec, (size_t)ns));

    dev_tmp t_deg(ctx), t_first(ctx), t_misc(ctx), t_tiles(ctx), t_perm(ctx);
    int32_t *d_deg;
    int64_t *d_first;
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_deg.p, (size_t)nv + 1)); d_deg = (int32_t *)t_deg.p;
    DGE_CUDA(ctx, dge_malloc(ctx, (int64_t **)&t_first.p, (size_t)nv)); d_first = (int64_t *)t_first.p;
    // misc: [0] runs (u64), [1] bad (int), [2] nonempty (int)
    DGE_CUDA(ctx, dge_malloc(ctx, (unsigned long long **)&t_misc.p, 4));
    unsigned long long *d_runs = (unsigned long long *)t_misc.p;
    int *d_bad = (int *)(d_runs + 1);
    int32_t *d_nonempty = (int32_t *)(d_runs + 2);
    DGE_CUDA(ctx, cudaMemsetAsync(t_misc.p, 0, 4 * sizeof(unsigned long long), st));
    DGE_CUDA(ctx, cudaMemsetAsync(d_deg, 0, sizeof(int32_t) * ((size_t)nv + 1), st));
    if (ns) DGE_CUDA(ctx, cudaMemcpyAsync(g->sources, sources, sizeof(int32_t) * (size_t)ns, cudaMemcpyHostToDevice, st));
    const int T = 256;
    if (ne) {
        k_count<<<grid_for(ne, T, ctx->sm_count, 16), T, 0, st>>>(d_src, d_dst, ne, nv, d_deg, d_first, d_runs, d_bad);
        DGE_LAUNCH_CHECK(ctx);
    }
    // exclusive scan deg -> row_ptr
    int32_t n_tiles = (int32_t)(((int64_t)nv + SCAN_TILE - 1) / SCAN_TILE);
    if (n_tiles < 1) n_tiles = 1;
    int64_t *d_tiles;
    DGE_CUDA(ctx, dge_malloc(ctx, (int64_t **)&t_tiles.p, (size_t)n_tiles)); d_tiles = (int64_t *)t_tiles.p;
    if (nv > 0) {
        k_scan_tile_sums<<<n_tiles, SCAN_THREADS, 0, st>>>(d_deg, nv, d_tiles, d_nonempty);
        DGE_LAUNCH_CHECK(ctx);
        k_scan_sums<<<1, 1024, 0, st>>>(d_tiles, n_tiles);
        DGE_LAUNCH_CHECK(ctx);
        k_scan_apply<<<n_tiles, SCAN_THREADS, 0, st>>>(d_deg, nv, d_tiles, g->row_ptr);
        DGE_LAUNCH_CHECK(ctx);
    } else {
        DGE_CUDA(ctx, cudaMemsetAsync(g->row_ptr, 0, sizeof(int64_t), st));
    }
    unsigned long long h_misc[4];
    DGE_CUDA(ctx, cudaMemcpyAsync(h_misc, t_misc.p, sizeof(h_misc), cudaMemcpyDeviceToHost, st));
    DGE_CUDA(ctx, cudaStreamSynchronize(st));
    if (*(int *)&h_misc[1]) return dge_fail(ctx, DGE_E_INVALID, "dge_graph_build: edge endpoint out of [0, n_vertices)");
    bool grouped = h_misc[0] == (unsigned long long)(*(int32_t *)&h_misc[2]);
    if (ne) {
        if (grouped) {
            k_place_grouped<<<grid_for(ne, T, ctx->sm_count, 16), T, 0, st>>>(d_src, d_dst, d_w, ne, g->row_ptr, d_first,
                                                                           g->col, g->w);
            DGE_LAUNCH_CHECK(ctx);
        } else {
            int32_t *d_perm;
            DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_perm.p, (size_t)ne)); d_perm = (int32_t *)t_perm.p;
            DGE_CUDA(ctx, cudaMemsetAsync(d_deg, 0, sizeof(int32_t) * ((size_t)nv + 1), st)); // reuse as cursor
            k_place_atomic<<<grid_for(ne, T, ctx->sm_count, 16), T, 0, st>>>(d_src, ne, g->row_ptr, d_deg, d_perm);
            DGE_LAUNCH_CHECK(ctx);
            k_sort_rows<<<grid_for((int64_t)nv * 32, T, ctx->sm_count, 8), T, 0, st>>>(g->row_ptr, nv, d_perm);
            DGE_LAUNCH_CHECK(ctx);
            k_gather<<<grid_for(ne, T, ctx->sm_count, 16), T, 0, st>>>(d_perm, d_dst, d_w, ne, g->col, g->w);
            DGE_LAUNCH_CHECK(ctx);
        }
    }
    ctx->phase_ms["grouped_input"] = grouped ? 1.f : 0.f;
    // out-degrees
    if (nv) {
        if (out_degree) {
            DGE_CUDA(ctx, cudaMemcpyAsync(g->out_degree, out_degree, sizeof(double) * (size_t)nv, cudaMemcpyHostToDevice, st));
        } else {
            k_out_degree<<<grid_for(nv, T, ctx->sm_count, 16), T, 0, st>>>(g->row_ptr, nv, g->w, g->out_degree);
            DGE_LAUNCH_CHECK(ctx);
        }
    }
    t_csr.stop();

    dge_phase_timer t_alias(ctx, "alias");
    int rc = run_alias(ctx, g->row_ptr, nv, ne, g->w, g->out_degree, g->prob, g->alias);
    if (rc != DGE_OK) return rc;
    // source table: one "row" of ns entries with weights out_degree[source]
    if (ns) {
        k_gather_src_w<<<(ns + T - 1) / T, T, 0, st>>>(g->sources, ns, nv, g->out_degree, g->src_w, d_bad);
        DGE_LAUNCH_CHECK(ctx);
        int h_bad = 0;
        DGE_CUDA(ctx, cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
        DGE_CUDA(ctx, cudaStreamSynchronize(st));
        if (h_bad) return dge_fail(ctx, DGE_E_INVALID, "dge_graph_build: source vertex out of [0, n_vertices)");
        if (source_weight_sum) {
            DGE_CUDA(ctx, cudaMemcpyAsync(g->sws, source_weight_sum, sizeof(double), cudaMemcpyHostToDevice, st));
        } else {
            k_seq_sum<<<1, 32, 0, st>>>(g->src_w, ns, g->sws);
            DGE_LAUNCH_CHECK(ctx);
        }
        DGE_CUDA(ctx, cudaMemcpyAsync(&g->source_weight_sum, g->sws, sizeof(double), cudaMemcpyDeviceToHost, st));
        int64_t h_rp[2] = {0, ns};
        dev_tmp t_rp(ctx);
        DGE_CUDA(ctx, dge_malloc(ctx, (int64_t **)&t_rp.p, 2));
        DGE_CUDA(ctx, cudaMemcpyAsync(t_rp.p, h_rp, sizeof(h_rp), cudaMemcpyHostToDevice, st));
        rc = run_alias(ctx, (const int64_t *)t_rp.p, 1, ns, g->src_w, g->sws, g->src_prob, g->src_alias);
        if (rc != DGE_OK) return rc;
    }
    t_alias.stop();

    dge_phase_timer t_pack(ctx, "pack");
    if (ne) {
        k_pack_rows<<<grid_for((int64_t)nv * 32, T, ctx->sm_count, 8), T, 0, st>>>(g->row_ptr, nv, g->col, g->prob, g->alias, g->rec);
        DGE_LAUNCH_CHECK(ctx);
    }
    if (ns) {
        k_pack_sources<<<(ns + T - 1) / T, T, 0, st>>>(g->row_ptr, g->sources, ns, g->src_prob, g->src_alias, g->srec);
        DGE_LAUNCH_CHECK(ctx);
    }
    t_pack.stop();
    DGE_CUDA(ctx, cudaStreamSynchronize(st));
    DGE_CUDA(ctx, cudaGetLastError());
    guard.g = nullptr;
    *out = g;
    return DGE_OK;
}

extern "C" {

int dge_graph_build(dge_ctx *ctx, int32_t nv, int64_t ne, const int32_t *src, const int32_t *dst, const double *w,
                    int32_t ns, const int32_t *sources, const double *out_degree, const double *source_weight_sum,
                    dge_graph **out) {
    if (!ctx) return dge_fail(nullptr, DGE_E_INVALID, "dge_graph_build: ctx is NULL");
    if (!out) return dge_fail(ctx, DGE_E_INVALID, "dge_graph_build: out is NULL");
    *out = nullptr;
    if (nv < 0 || ne < 0 || ns < 0) return dge_fail(ctx, DGE_E_INVALID, "dge_graph_build: negative size");
    if (ne > 0 && (!src || !dst || !w)) return dge_fail(ctx, DGE_E_INVALID, "dge_graph_build: NULL edge arrays");
    if (ns > 0 && !sources) return dge_fail(ctx, DGE_E_INVALID, "dge_graph_build: NULL source list");
    if (ne >= (int64_t)1 << 31) return dge_fail(ctx, DGE_E_LIMIT, "dge_graph_build: n_edges must be < 2^31");
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    dev_tmp t_src(ctx), t_dst(ctx), t_w(ctx);
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_src.p, (size_t)ne));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&t_dst.p, (size_t)ne));
    DGE_CUDA(ctx, dge_malloc(ctx, (double **)&t_w.p, (size_t)ne));
    if (ne) {
        dge_phase_timer t_h2d(ctx, "coo_h2d");
        DGE_CUDA(ctx, cudaMemcpyAsync(t_src.p, src, sizeof(int32_t) * (size_t)ne, cudaMemcpyHostToDevice, st));
        DGE_CUDA(ctx, cudaMemcpyAsync(t_dst.p, dst, sizeof(int32_t) * (size_t)ne, cudaMemcpyHostToDevice, st));
        DGE_CUDA(ctx, cudaMemcpyAsync(t_w.p, w, sizeof(double) * (size_t)ne, cudaMemcpyHostToDevice, st));
        t_h2d.stop();
    }
    return dge_graph_build_device(ctx, nv, ne, (const int32_t *)t_src.p, (const int32_t *)t_dst.p, (const double *)t_w.p, ns,
                                  sources, out_degree, source_weight_sum, out);
}

int dge_graph_sizes(const dge_graph *g, int32_t *nv, int64_t *ne, int32_t *ns) {
    if (!g) return dge_fail(nullptr, DGE_E_INVALID, "dge_graph_sizes: graph is NULL");
    if (nv) *nv = g->nv;
    if (ne) *ne = g->ne;
    if (ns) *ns = g->ns;
    return DGE_OK;
}

int dge_graph_tables(const dge_graph *g, int64_t *row_ptr, int32_t *col, double *w, double *prob, int32_t *alias,
                     double *out_degree, double *src_prob, int32_t *src_alias, double *source_weight_sum) {
    if (!g) return dge_fail(nullptr, DGE_E_INVALID, "dge_graph_tables: graph is NULL");
    dge_ctx *ctx = g->ctx;
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
#define D2H(dstp, srcp, n) \
    if ((dstp) && (n)) DGE_CUDA(ctx, cudaMemcpyAsync((dstp), (srcp), sizeof(*(dstp)) * (size_t)(n), cudaMemcpyDeviceToHost, st))
    D2H(row_ptr, g->row_ptr, (size_t)g->nv + 1);
    D2H(col, g->col, g->ne);
    D2H(w, g->w, g->ne);
    D2H(prob, g->prob, g->ne);
    D2H(alias, g->alias, g->ne);
    D2H(out_degree, g->out_degree, g->nv);
    D2H(src_prob, g->src_prob, g->ns);
    D2H(src_alias, g->src_alias, g->ns);
#undef D2H
    if (source_weight_sum) *source_weight_sum = g->source_weight_sum;
    DGE_CUDA(ctx, cudaStreamSynchronize(st));
    return DGE_OK;
}

int dge_graph_sample_next(const dge_graph *g, int64_t n, const int32_t *v, const double *x, int sampler, int32_t *out) {
    if (!g) return dge_fail(nullptr, DGE_E_INVALID, "dge_graph_sample_next: graph is NULL");
    dge_ctx *ctx = g->ctx;
    if (n < 0 || (n > 0 && (!v || !x || !out))) return dge_fail(ctx, DGE_E_INVALID, "dge_graph_sample_next: bad arguments");
    if (sampler != DGE_SAMPLER_ALIAS && sampler != DGE_SAMPLER_CDF)
        return dge_fail(ctx, DGE_E_INVALID, "dge_graph_sample_next: unknown sampler");
    if (n == 0) return DGE_OK;
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    dev_tmp tv(ctx), tx(ctx), to(ctx);
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&tv.p, (size_t)n));
    DGE_CUDA(ctx, dge_malloc(ctx, (double **)&tx.p, (size_t)n));
    DGE_CUDA(ctx, dge_malloc(ctx, (int32_t **)&to.p, (size_t)n));
    DGE_CUDA(ctx, cudaMemcpyAsync(tv.p, v, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    DGE_CUDA(ctx, cudaMemcpyAsync(tx.p, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    k_sample_next<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g->rec, g->srec, g->ns, g->nv, g->row_ptr, g->col, g->w,
                                                              g->out_degree, g->sources, g->sws, n, (const int32_t *)tv.p,
                                                              (const double *)tx.p, sampler, (int32_t *)to.p);
    DGE_LAUNCH_CHECK(ctx);
    DGE_CUDA(ctx, cudaMemcpyAsync(out, to.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
    DGE_CUDA(ctx, cudaStreamSynchronize(st));
    for (int64_t i = 0; i < n; i++)
        if (out[i] == -2) return dge_fail(ctx, DGE_E_INVALID, "dge_graph_sample_next: vertex id out of range");
    return DGE_OK;
}

void dge_graph_free(dge_graph *g) {
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    graph_release(g);
}

} // extern "C"
