// walk_microbench.cu -- what the memory system delivers for the access shape of k_walk_alias on an HBM-resident
// graph (BASELINE config 4: 1M regions x 24 slices, ~25 GB of 32-byte walk records): every walk step is ONE
// dependent, uniformly random 32-byte sector read (ld.global.nc.v4.u64, LDG.256) whose address comes from the
// previous read, plus one coalesced 4-byte token store.  This gives the denominator for "how close is the walk
// kernel to the hardware's random-sector rate" (DESIGN.md 3.2 / 3.4).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/bin/walk_microbench scripts/walk_microbench.cu
//   scripts/bin/walk_microbench [GiB of records, default 24]
// Variants: C = independent chains per thread (1, 2, 4: more sectors in flight per warp at the same occupancy),
// threads per block 256, grid sized so that all chains are resident at once or oversubscribed 4x; a second table
// size (64 MB: L2-resident) shows the same loop without DRAM.  Prints G steps/s and sector GB/s (32 B per step).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

struct rec { unsigned long long a, b, c, d; };

__device__ __forceinline__ rec ld_rec(const rec *p) {
    rec r;
    asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p));
    return r;
}

// every record holds two pseudo-random successor indices (like start0/start1 of dge_edge_rec) and a threshold
__global__ void k_fill(rec *t, uint64_t n) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t z = i * 0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
        uint64_t y = z * 0xD6E8FEB86659FD93ULL; y ^= y >> 32;
        rec r; r.a = z; r.b = (z >> 11) % n; r.c = (y >> 11) % n; r.d = y;
        t[i] = r;
    }
}

template <int C>
__global__ void __launch_bounds__(256) k_chase(const rec *__restrict__ t, uint64_t n, int steps, int64_t n_chains,
                                                int32_t *__restrict__ tok) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i * C >= n_chains) return;
    uint64_t cur[C];
    uint32_t coin = (uint32_t)i * 2654435761u;
#pragma unroll
    for (int c = 0; c < C; c++) cur[c] = ((uint64_t)(i * C + c) * 0x9E3779B97F4A7C15ULL >> 20) % n;
    for (int j = 0; j < steps; j++) {
        rec r[C];
#pragma unroll
        for (int c = 0; c < C; c++) r[c] = ld_rec(t + cur[c]);     // C independent sectors in flight per thread
        coin = coin * 1664525u + 1013904223u;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const bool first = ((uint32_t)r[c].a ^ coin) & 0x10000u;  // data-dependent select, like y < prob
            cur[c] = first ? r[c].b : r[c].c;
            tok[(int64_t)j * n_chains + (int64_t)c * (n_chains / C) + i] = (int32_t)r[c].d; // coalesced 4 B store per step
        }
    }
}

template <int C>
static void run(const char *what, const rec *t, uint64_t n, int64_t n_chains, int steps, int32_t *tok) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int64_t threads = n_chains / C;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    k_chase<C><<<grid, 256>>>(t, n, 2, n_chains, tok);
    cudaEventRecord(e0);
    k_chase<C><<<grid, 256>>>(t, n, steps, n_chains, tok);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double st = (double)n_chains * steps;
    printf("%-22s chains/thread=%d chains=%9lld steps=%d : %7.2f G steps/s  %7.1f GB/s of sectors  %7.1f GB/s incl. token stores  (%.2f ms)%s\n",
           what, C, (long long)n_chains, steps, st / ms / 1e6, st * 32 / ms / 1e6, st * 36 / ms / 1e6, ms,
           cudaGetLastError() == cudaSuccess ? "" : "  CUDA ERROR");
}

int main(int argc, char **argv) {
    const double gib = argc > 1 ? atof(argv[1]) : 24.0;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    const uint64_t n_big = (uint64_t)(gib * 1024.0 * 1024.0 * 1024.0 / 32.0), n_small = (64ull << 20) / 32;
    rec *big = nullptr;
    if (cudaMalloc(&big, n_big * sizeof(rec)) != cudaSuccess) { printf("cudaMalloc of %.0f GiB failed\n", gib); return 1; }
    k_fill<<<p.multiProcessorCount * 8, 256>>>(big, n_big);
    const int steps = 24;
    const int64_t resident = (int64_t)p.multiProcessorCount * 2048;          // 64 warps per SM: k_walk_alias at 32 registers
    int32_t *tok = nullptr;
    cudaMalloc(&tok, sizeof(int32_t) * (size_t)resident * 32 * steps);
    cudaDeviceSynchronize();
    for (int over : {1, 4, 16}) {                                            // walks per launch / resident walks
        const int64_t n = resident * over;
        printf("-- %.0f GiB of records (HBM), %lld chains (%dx the resident threads at 1 chain/thread)\n", gib, (long long)n, over);
        run<1>("HBM random sector", big, n_big, n, steps, tok);
        run<2>("HBM random sector", big, n_big, n, steps, tok);
        run<4>("HBM random sector", big, n_big, n, steps, tok);
    }
    printf("-- 64 MB of records (L2-resident)\n");
    k_fill<<<p.multiProcessorCount * 8, 256>>>(big, n_small);
    run<1>("L2 random sector", big, n_small, resident * 4, steps, tok);
    run<2>("L2 random sector", big, n_small, resident * 4, steps, tok);
    run<4>("L2 random sector", big, n_small, resident * 4, steps, tok);
    cudaFree(big); cudaFree(tok);
    return 0;
}
