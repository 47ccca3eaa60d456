// Blackwell (sm_100a) building blocks for the tensor-core graph-convolution kernels:
// mbarrier pipelines, TMA tensor-map loads, tcgen05.mma with TMEM accumulators, tcgen05.ld.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace istgcn {
namespace tc {

constexpr int kAtomRows = 128;                       // MMA M
constexpr int kAtomK = 32;                           // tf32 elements per 128-byte swizzle row
constexpr int kAtomBytes = kAtomRows * 128;          // [128 rows][32 tf32], SWIZZLE_128B K-major

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of the (converged) warp: the way to issue single-thread instructions (TMA, tcgen05.mma,
// tcgen05.commit) from warp-uniform control flow.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// Bounded wait: a pipeline bug must abort the kernel (trap -> CUDA error), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin > (1u << 26)) {
            printf("istgcn: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x,
                   threadIdx.x, addr, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 2-D tile load: coordinates (c0 = fastest dim, c1), completes `bytes` on `bar`
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
        "%4}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 2-D tile store / reduce-add from shared memory (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const void* smem_src, const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const void* smem_src, const CUtensorMap* map, int c0,
                                                  int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* smem_src, const CUtensorMap* map, int c0, int c1,
                                             int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory source of every committed bulk store has been read (buffer reusable)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// every committed bulk store is complete (writes performed)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 4-D tile load (coordinates fastest first), completes on `bar`
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
        "%4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // one warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// arrive on `bar` when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 inputs, fp32 accumulate; one thread issues
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
        "%4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x 8 consecutive 32-bit columns
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with a disable-output-lane mask (bit i of word q set: lane 32q+i is NOT written)
__device__ __forceinline__ void tc_mma_tf32_ts_masked(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                                      uint32_t idesc, uint32_t accumulate, uint32_t m0,
                                                      uint32_t m1, uint32_t m2, uint32_t m3) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
        : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 32 consecutive columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
        "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])),
        "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])),
        "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])),
        "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading/stride
// byte offsets (all >> 4), version 1 (Blackwell), SWIZZLE_128B.
//   K-major operand  : rows of 128 B (32 tf32 along K); 8-row groups 1024 B apart (SBO = 1024);
//                      LBO unused (1).  K-step of 8 tf32 = start address + 32 B.
//   MN-major operand : the same tile read the other way: 32 tf32 along M/N per 128-B row, the
//                      8 rows of a 1024-B group are 8 consecutive K; SBO = 1024 (next 8 K),
//                      LBO = byte distance between 32-wide M/N groups.
//   MN-major TF32    : only SWIZZLE_128B_BASE32B (layout type 1) is legal (cutlass
//                      sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only
//                      available smem layout"): 128-B rows, 32-B chunks XORed with row % 4,
//                      swizzle atom = 4 K-rows (512 B), so SBO = 512 and a K-step of 8 rows =
//                      start address + 1024 B.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes, uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;                                        // version = 1
    d |= static_cast<uint64_t>(layout_type) << 61;          // 2 = SWIZZLE_128B, 1 = 128B_BASE32B
    return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, TF32 A/B.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
           (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

// float index of element (row, e) inside a [128][32] SWIZZLE_128B atom (16-byte chunk index
// XORed with row % 8 -- the layout both TMA and tcgen05 use for SWIZZLE_128B)
__device__ __forceinline__ int atom_index(int row, int e) {
    return ((row >> 3) << 8) + ((row & 7) << 5) + ((((e >> 2) ^ (row & 7)) << 2) | (e & 3));
}

// the same for the 32-byte-atom variant (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B / UMMA
// SWIZZLE_128B_BASE32B): 32-byte chunk index XORed with row % 4
__device__ __forceinline__ int atom_index32(int row, int e) {
    return (row << 5) + ((((e >> 3) ^ (row & 3)) << 3) | (e & 7));
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Column sums of a 32x32 (lane x register) tile in 31 shuffles: afterwards v[0] of lane l holds
// the sum over all lanes of the original v[l].
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const bool upper = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < step; ++i) {
            const float send = upper ? v[i] : v[i + step];
            const float keep = upper ? v[i + step] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    return v[0];
}

// One destination joint of one partition: acc[f] = sum_j a_j * xs[f][v_j] for FR frames, four
// channels per lane, entries taken two at a time so that 2*FR 128-bit loads are in flight.
template <int FR, bool SW32>
__device__ __forceinline__ void aggregate_joint(float* __restrict__ A, const float* __restrict__ xs,
                                                const int2* __restrict__ s_ent, int beg, int end,
                                                int fstride, int V, int w, int c4) {
    float4 acc[FR];
#pragma unroll
    for (int f = 0; f < FR; ++f) acc[f] = make_float4(0.f, 0.f, 0.f, 0.f);
    int j = beg;
    for (; j + 1 < end; j += 2) {
        const int2 e0 = s_ent[j], e1 = s_ent[j + 1];
        const float a0 = __int_as_float(e0.y), a1 = __int_as_float(e1.y);
        const float* x0 = xs + e0.x;
        const float* x1 = xs + e1.x;
        float4 u0[FR], u1[FR];
#pragma unroll
        for (int f = 0; f < FR; ++f) {
            u0[f] = ld4(x0 + f * fstride);
            u1[f] = ld4(x1 + f * fstride);
        }
#pragma unroll
        for (int f = 0; f < FR; ++f) {
            acc[f].x = fmaf(a1, u1[f].x, fmaf(a0, u0[f].x, acc[f].x));
            acc[f].y = fmaf(a1, u1[f].y, fmaf(a0, u0[f].y, acc[f].y));
            acc[f].z = fmaf(a1, u1[f].z, fmaf(a0, u0[f].z, acc[f].z));
            acc[f].w = fmaf(a1, u1[f].w, fmaf(a0, u0[f].w, acc[f].w));
        }
    }
    if (j < end) {
        const int2 e0 = s_ent[j];
        const float a0 = __int_as_float(e0.y);
        const float* x0 = xs + e0.x;
#pragma unroll
        for (int f = 0; f < FR; ++f) {
            const float4 xv = ld4(x0 + f * fstride);
            acc[f].x = fmaf(a0, xv.x, acc[f].x);
            acc[f].y = fmaf(a0, xv.y, acc[f].y);
            acc[f].z = fmaf(a0, xv.z, acc[f].z);
            acc[f].w = fmaf(a0, xv.w, acc[f].w);
        }
    }
#pragma unroll
    for (int f = 0; f < FR; ++f)
        st4(A + (SW32 ? atom_index32(f * V + w, c4) : atom_index(f * V + w, c4)), acc[f]);
}

// dispatch on the number of frames: 5 (V=25), 7 (V=18), 8 (V<=16) per tile, or the share of one
// aggregator team (a team working on frames [f0, f0+n) passes xs + f0*fstride and w + f0*V)
template <bool SW32 = false>
__device__ __forceinline__ void aggregate_joint_any(int F, float* __restrict__ A,
                                                    const float* __restrict__ xs,
                                                    const int2* __restrict__ s_ent, int beg, int end,
                                                    int fstride, int V, int w, int c4) {
    switch (F) {
        case 5: aggregate_joint<5, SW32>(A, xs, s_ent, beg, end, fstride, V, w, c4); break;
        case 7: aggregate_joint<7, SW32>(A, xs, s_ent, beg, end, fstride, V, w, c4); break;
        case 8: aggregate_joint<8, SW32>(A, xs, s_ent, beg, end, fstride, V, w, c4); break;
        case 6: aggregate_joint<6, SW32>(A, xs, s_ent, beg, end, fstride, V, w, c4); break;
        case 3: aggregate_joint<3, SW32>(A, xs, s_ent, beg, end, fstride, V, w, c4); break;
        case 2: aggregate_joint<2, SW32>(A, xs, s_ent, beg, end, fstride, V, w, c4); break;
        case 1: aggregate_joint<1, SW32>(A, xs, s_ent, beg, end, fstride, V, w, c4); break;
        default: aggregate_joint<4, SW32>(A, xs, s_ent, beg, end, fstride, V, w, c4); break;
    }
}

// Host: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed).
// 2-D fp32 row-major matrix [rows][cols] -> tiles of [box_rows][32 floats], SWIZZLE_128B.
int encode_tile_map(CUtensorMap* map, const float* base, long long rows, long long cols,
                    int box_rows, bool atom32 = false);
// 4-D (C, V, T, NM) view of a channels-last activation: box of 32 channels x V x frames (strided)
int encode_frames_map(CUtensorMap* map, const float* base, int NM, int T, int V, int C, int frames,
                      int t_stride, bool atom32 = false);
// 3-D (C, V, frames) view of a channels-last activation: box = 32 channels x V joints x 1 frame,
// 32-byte-atom 128B swizzle (gcn_tc2.cu)
int encode_frame_slices(CUtensorMap* map, const float* base, long long frames, int V, int C);
// 3-D (C, V, frames) view: box = 32 channels x ONE joint x box_frames consecutive frames, 32-byte-atom
// 128B swizzle: the rows-contracted (MN-major) operand of gcn_pair_tc.cu
int encode_joint_frames_map(CUtensorMap* map, const float* base, long long frames, int V, int C,
                            int box_frames);
// second-generation graph-convolution engine (gcn_tc2.cu)
bool gcn_tc2_eligible(int V, int K, int Cin, int Cout, const float* in, const float* out);
int launch_gcn_tc2(const float* in, const float* w_rows, const float* vals, const int* lptr,
                   const int* lsrc, const int* lid, const float* bias_k, const float* colsum, float* out,
                   int reduce, double* stat_sum, double* stat_sumsq, int frames, int V, int K, int Cin,
                   int Cout, cudaStream_t st);
// third-generation engine (gcn_tc3.cu): partitions on the lanes + block exchange
int launch_gcn_tc3(const float* in, const float* w_rows, const float* vals, const int* lptr,
                   const int* lsrc, const int* lid, const float* bias_k, const float* colsum, float* out,
                   int reduce, double* stat_sum, double* stat_sumsq, int frames, int V, int K, int Cin,
                   int Cout, cudaStream_t st);
// second-generation weight-gradient kernel (gcn_tc_dw2.cu)
bool gcn_tc_dw2_eligible(int V, int K, int Cin, int Cout);
int launch_gcn_tc_dw2(const float* dz, const float* x, const float* vals, const int* lptr,
                      const int* lsrc, const int* lid, float* dWc, int frames, int V, int K, int Cin,
                      int Cout, cudaStream_t st);
// second-generation adjacency-gradient kernel (gcn_tc_da2.cu)
bool gcn_tc_da2_eligible(int V, int K, int Cin, int Cout);
int launch_gcn_tc_da2(const float* dz, const float* x, const float* Wc, const int* lptr, const int* lsrc,
                      const int* lid, int nnz, float* dvals, int frames, int V, int K, int Cin, int Cout,
                      cudaStream_t st);
// out[j] += sum_f in[f][j], j < n (gcn_tc_dw.cu)
int launch_frame_colsum(const float* in, float* out, int frames, int n, cudaStream_t st);

}  // namespace tc
}  // namespace istgcn
