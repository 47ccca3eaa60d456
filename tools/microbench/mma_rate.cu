// Issue-rate microbenchmark of tcgen05.mma kind::tf32 on one SM (cta_group::1): cycles per
// instruction for the operand forms the graph-convolution engines use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ist-gcn_b200/csrc \
//        tools/microbench/mma_rate.cu -o tools/microbench/mma_rate && tools/microbench/mma_rate
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace istgcn::tc;

// mode 0: TS (A from tensor memory), B K-major SWIZZLE_128B
// mode 1: SS (A K-major SWIZZLE_128B in shared memory), B K-major
// mode 2: TS, B MN-major 32-byte-atom swizzle (the aggregation operand)
// mode 3: TS with a disable-output-lane mask (gcn_tc2's aggregation)
template <int MODE, int N, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int random) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<float*>(smem)[i] = random ? ((int)(h & 0xFFFF) - 32768) * (1.0f / 32768.f) : 1.0f;
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (random) {                                   // random A operand / accumulators in tensor memory
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = reinterpret_cast<float*>(smem)[(threadIdx.x * 37 + j * 101) & 8191];
        for (int c = 0; c < 512; c += 32) tmem_st32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(128, N, false, MODE == 2 || MODE == 3);
        const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem) + 16384;
        long long t0 = 0, t1 = 0;
        unsigned long long g0 = 0, g1 = 0;
        for (int rep = 0; rep < 2; ++rep) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
            t0 = clock64();
            if (elect_one()) {
                for (int i = 0; i < iters; ++i) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t d = tm + 256 + (NACC > 1 ? (i % NACC) * (N < 128 ? N : 128) : 0);
                        if (MODE == 0)
                            tc_mma_tf32_ts(d, tm + ks * 8, make_desc(b_s + ks * 32, 16, 1024), idesc, 1u);
                        else if (MODE == 1)
                            tc_mma_tf32(d, make_desc(a_s + ks * 32, 16, 1024), make_desc(b_s + ks * 32, 16, 1024),
                                        idesc, 1u);
                        else if (MODE == 2)
                            tc_mma_tf32_ts(d, tm + ks * 8, make_desc(b_s + ks * 1024, 4096, 512, 1), idesc, 1u);
                        else
                            tc_mma_tf32_ts_masked(d, tm + ks * 8, make_desc(b_s + ks * 1024, 4096, 512, 1), idesc, 1u,
                                                  0u, ~0u, ~0u, ~0u);
                    }
                }
                tc_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, rep & 1);
            t1 = clock64();
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        }
        if (threadIdx.x == 32 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)(g1 - g0); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int MODE, int N, int NACC>
void run(const char* what, long long* d_out, int grid = 1, int random = 0, int iters = 256) {
    auto k = rate_kernel<MODE, N, NACC>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k<<<grid, 128, 64 * 1024>>>(d_out, iters, random);
    cudaError_t e = cudaDeviceSynchronize();
    long long r[2] = {0, 0};
    cudaMemcpy(r, d_out, sizeof(r), cudaMemcpyDeviceToHost);
    const double per = (double)r[0] / (iters * 4.0);
    const double mhz = r[1] > 0 ? (double)r[0] / (double)r[1] * 1e3 : 0.0;
    printf("%-44s N=%3d grid=%3d %s iters=%6d  %6.1f clk/instr  %5.0f MHz  %6.1f ns/instr  %5.0f TFLOP/s chip%s\n", what, N,
           grid, random ? "random" : "ones  ", iters, per, mhz, (double)r[1] / (iters * 4.0),
           2.0 * 128 * N * 8 * grid / ((double)r[1] / (iters * 4.0)) * 1e-3, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 64);
    run<0, 32, 1>("TS  A=tmem, B K-major SW128", d_out);
    run<0, 64, 1>("TS  A=tmem, B K-major SW128", d_out);
    run<0, 128, 1>("TS  A=tmem, B K-major SW128", d_out);
    run<0, 256, 1>("TS  A=tmem, B K-major SW128", d_out);
    run<0, 128, 2>("TS  A=tmem, B K-major SW128, 2 accumulators alternating", d_out);
    run<1, 32, 1>("SS  A=smem K-major, B K-major", d_out);
    run<1, 64, 1>("SS  A=smem K-major, B K-major", d_out);
    run<1, 128, 1>("SS  A=smem K-major, B K-major", d_out);
    run<1, 256, 1>("SS  A=smem K-major, B K-major", d_out);
    run<2, 32, 1>("TS  A=tmem, B MN-major 32B-atom", d_out);
    run<2, 64, 1>("TS  A=tmem, B MN-major 32B-atom (LBO 4096)", d_out);
    run<3, 32, 1>("TS  masked (one lane quadrant), B MN-major", d_out);
    printf("-- all SMs, long runs (power / clock behaviour)\n");
    run<0, 128, 1>("TS  B K-major", d_out, 148, 0, 40000);
    run<0, 128, 1>("TS  B K-major", d_out, 148, 1, 40000);
    run<0, 128, 1>("TS  B K-major", d_out, 148, 1, 40000);
    run<1, 128, 1>("SS  A,B K-major", d_out, 148, 1, 40000);
    run<0, 64, 1>("TS  B K-major", d_out, 148, 1, 40000);
    run<2, 32, 1>("TS  B MN-major", d_out, 148, 1, 40000);
    run<0, 128, 1>("TS  B K-major", d_out, 74, 1, 40000);
    run<0, 128, 1>("TS  B K-major", d_out, 1, 1, 40000);
    return 0;
}
