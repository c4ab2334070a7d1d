// red_microbench.cu -- what the L2 can take: random-row 128-bit reductions (red.global.add.v4.f32) and 128-bit
// L2 loads (ld.global.cg.v4.f32) in the access shapes of the skip-gram item kernel.  It gives the denominator for
// "how close is k_sgns_items_v2 to the reduction / load throughput of the memory system" (DESIGN.md 3.3).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/bin/red_microbench scripts/red_microbench.cu
//   scripts/bin/red_microbench
// Shapes: G lanes per row, `live` of them active, row pitch in bytes, V rows.  Every warp-level instruction
// touches 32/G random rows.  mode 0 = red only, 1 = load only, 2 = load + red of the same row,
// 3 = the row update staged in shared memory and sent as ONE TMA bulk reduction per row by one lane
// (cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32: no per-lane LSU reduction at all), 4 = load + 3.
#include <cstdio>
#include <cstdint>
#include <string>
#include <cuda_runtime.h>

__device__ __forceinline__ void red4(float4 *p, float v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float4 ld4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void bulk_red(float *gptr, uint32_t smem_addr, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gptr), "r"(smem_addr), "r"(bytes) : "memory");
}

template <int G, int MODE>
__global__ void __launch_bounds__(128) k(float *table, uint32_t V, uint32_t pitch, int live, int iters, float *sink) {
    const int lane = threadIdx.x % G;
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    uint32_t s = gid * 2654435761u + 12345u;
    float acc = 0.f;
    char *base = reinterpret_cast<char *>(table) + (lane < live ? lane : 0) * 16;
    __shared__ __align__(128) float4 stage[2][6][128]; // [double buffer][row of the unit][thread]: group g's row at [..][g * G]
    const uint32_t my_stage = (uint32_t)__cvta_generic_to_shared(&stage[0][0][threadIdx.x]);
    for (int it = 0; it < iters; it++) {
        if (MODE >= 3) { // the stage written two units ago must have been read by the TMA engine
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
        }
#pragma unroll
        for (int u = 0; u < 6; u++) { // 6 rows per "pair", like K+1 targets
            s = s * 1664525u + 1013904223u;
            uint32_t row = (uint32_t)(((uint64_t)(s >> 4) * V) >> 28);
            float4 *p = reinterpret_cast<float4 *>(base + (uint64_t)row * pitch);
            if (MODE < 3) {
                if (lane < live) {
                    if (MODE >= 1) { float4 v = ld4(p); acc += v.x + v.y + v.z + v.w; }
                    if (MODE != 1) red4(p, 1e-9f);
                }
            } else {
                float4 v = make_float4(1e-9f, 1e-9f, 1e-9f, 1e-9f);
                if (lane < live && MODE == 4) { float4 x = ld4(p); acc += x.x + x.y + x.z + x.w; }
                const uint32_t dst = my_stage + (uint32_t)(((it & 1) * 6 + u) * 128 * 16);
                if (lane < live) asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) bulk_red(reinterpret_cast<float *>(p), dst, (uint32_t)live * 16u);
            }
        }
        if (MODE >= 3 && lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (MODE >= 3 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 123.456f) *sink = acc;
}

static bool g_json = false;
static double g_last_rows_per_s = 0;

template <int G, int MODE>
static void run(const char *name, float *table, uint32_t V, uint32_t pitch, int live, int sms) {
    float *sink;
    cudaMalloc(&sink, 4);
    const int blocks = sms * 8, iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<G, MODE><<<blocks, 128>>>(table, V, pitch, live, 50, sink);
    cudaEventRecord(e0);
    k<G, MODE><<<blocks, 128>>>(table, V, pitch, live, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double rows = (double)blocks * 128 / G * iters * 6;
    double lane_ops = rows * live * (MODE == 2 || MODE == 4 ? 2 : 1);
    g_last_rows_per_s = rows / ms * 1e3;
    if (!g_json) printf("%-34s G=%2d live=%2d pitch=%4u V=%8u mode=%d : %8.1f M rows/s  %8.1f G lane-ops/s  %7.1f GB/s  (%.2f ms)\n", name, G, live, pitch,
           V, MODE, rows / ms / 1e3, lane_ops / ms / 1e6, lane_ops * 16 / ms / 1e6, ms);
    cudaFree(sink);
}

int main(int argc, char **argv) {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    if (argc > 1 && std::string(argv[1]) == "--json") {
        // bench.py's live ceiling: the tract x 24 access shape (80-byte rows on a 96-byte pitch, 5 of 8 lanes, 19 224 rows,
        // L2-resident) and the community-area shape (32-byte rows, 2 of 8 lanes, 1 848 rows); one JSON line
        g_json = true;
        float *t;
        cudaMalloc(&t, (size_t)19224 * 128);
        cudaMemset(t, 0, (size_t)19224 * 128);
        double r[9];
        run<8, 0>("", t, 19224, 96, 5, sms); r[0] = g_last_rows_per_s;
        run<8, 1>("", t, 19224, 96, 5, sms); r[1] = g_last_rows_per_s;
        run<8, 2>("", t, 19224, 96, 5, sms); r[2] = g_last_rows_per_s;
        run<8, 0>("", t, 1848, 32, 2, sms); r[3] = g_last_rows_per_s;
        run<8, 1>("", t, 1848, 32, 2, sms); r[4] = g_last_rows_per_s;
        run<8, 2>("", t, 1848, 32, 2, sms); r[5] = g_last_rows_per_s;
        // the synthetic shape: 512-byte rows (D = 128), 32 lanes, a hot set of 100 000 rows (what a skewed walk corpus touches most)
        float *big;
        cudaMalloc(&big, (size_t)100000 * 512);
        cudaMemset(big, 0, (size_t)100000 * 512);
        run<32, 0>("", big, 100000, 512, 32, sms); r[6] = g_last_rows_per_s;
        run<32, 1>("", big, 100000, 512, 32, sms); r[7] = g_last_rows_per_s;
        run<32, 2>("", big, 100000, 512, 32, sms); r[8] = g_last_rows_per_s;
        printf("{\"device\": \"%s\", \"tract24_rows_per_s\": {\"red\": %.4g, \"load\": %.4g, \"load_red\": %.4g}, "
               "\"ca_rows_per_s\": {\"red\": %.4g, \"load\": %.4g, \"load_red\": %.4g}, "
               "\"d128_hot100k_rows_per_s\": {\"red\": %.4g, \"load\": %.4g, \"load_red\": %.4g}}\n", p.name, r[0], r[1], r[2], r[3], r[4], r[5],
               r[6], r[7], r[8]);
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
    }
    printf("%s, %d SMs\n", p.name, sms);
    float *small, *big;
    const uint32_t Vs = 19224, Vb = 2400000;
    cudaMalloc(&small, (size_t)Vs * 128);
    cudaMalloc(&big, (size_t)Vb * 512);
    cudaMemset(small, 0, (size_t)Vs * 128);
    cudaMemset(big, 0, (size_t)Vb * 512);
    // tract x 24 shape: 80-byte rows on a 96-byte pitch, 5 of 8 lanes live, 19 224 rows (L2-resident)
    run<8, 0>("tract24 rows: red", small, Vs, 96, 5, sms);
    run<8, 1>("tract24 rows: load", small, Vs, 96, 5, sms);
    run<8, 2>("tract24 rows: load+red", small, Vs, 96, 5, sms);
    run<8, 3>("tract24 rows: TMA bulk red", small, Vs, 96, 5, sms);
    run<8, 4>("tract24 rows: load+bulk red", small, Vs, 96, 5, sms);
    run<8, 0>("128 B rows, 8/8 live: red", small, Vs, 128, 8, sms);
    run<8, 0>("32 B rows (CA D=8), 2/8 live: red", small, Vs, 32, 2, sms);
    run<4, 0>("64 B rows, G=4 4/4 live: red", small, Vs, 64, 4, sms);
    // synthetic shape: 512-byte rows, 32 lanes, 2.4 M rows (1.2 GB: HBM unless the draw is skewed)
    run<32, 0>("512 B rows, uniform 2.4M: red", big, Vb, 512, 32, sms);
    run<32, 1>("512 B rows, uniform 2.4M: load", big, Vb, 512, 32, sms);
    run<32, 2>("512 B rows, uniform 2.4M: load+red", big, Vb, 512, 32, sms);
    run<32, 3>("512 B rows, uniform 2.4M: bulk red", big, Vb, 512, 32, sms);
    run<32, 4>("512 B rows, uniform 2.4M: ld+bulk", big, Vb, 512, 32, sms);
    run<32, 0>("512 B rows, 100K hot rows: red", big, 100000, 512, 32, sms);
    run<32, 3>("512 B rows, 100K hot: bulk red", big, 100000, 512, 32, sms);
    run<32, 4>("512 B rows, 100K hot: ld+bulk red", big, 100000, 512, 32, sms);
    run<32, 2>("512 B rows, 100K hot rows: ld+red", big, 100000, 512, 32, sms);
    return 0;
}
