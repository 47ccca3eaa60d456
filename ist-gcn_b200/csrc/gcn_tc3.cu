// Third-generation tcgen05 engine of the fused graph convolution (reference:
// net/utils/tgcn.py:76-89, net/utils/inceptionv2_gcn.py:64-89).  Same two contractions as
// gcn_tc2.cu,
//
//     X'_k[(f,w)][ci] = sum_v A_eff[k][v][w] * IN[(f,v)][ci]            (MMA 1, per frame)
//     OUT[(f,w)][n]   = sum_k sum_ci X'_k[(f,w)][ci] * W[k][n][ci]      (MMA 2)
//
// but the aggregation no longer wastes three quarters of every M = 128 instruction.  gcn_tc2 puts
// the four frames of a tile on the block diagonal (lanes = (frame, joint); each instruction keeps
// one lane quadrant) and pays K * 4 frames * 4 K-steps = 64 instructions per 32-channel slice.
// Here MMA 1 stacks the PARTITIONS on the lanes instead,
//
//     D1[(k,w)][f][32 ci] = ADJ[(k,w)][32 v] * X_f[32 v][32 ci]          M128 N32 K8 x 4 per frame,
//
// all 128 lanes useful, 16 instructions per slice.  MMA 2 wants (frame, joint) on the lanes and
// (partition, channel) on the columns, i.e. the 4 x 4 grid of [32 lanes][32 columns] blocks
// transposed.  tensor-memory lanes are private to one warp of each warpgroup-quarter, so four
// exchange warps carry the 12 off-diagonal blocks through shared memory:
//
//     tcgen05.ld (quadrant k) -> round to TF32 -> st.shared -> bar -> ld.shared -> tcgen05.st (quadrant f)
//
// in place (block (k, f) and block (f, k) swap; the diagonal stays), 48 KB of scratch, while the
// tensor pipe runs MMA 2 of the previous slice.  MMA 2, the weight / input TMA rings and the
// epilogue are those of gcn_tc2.
//
//   warp 0  TMA producer (weights)        warp 2  TMA producer (input frames)
//   warp 1  MMA issuer                    warps 4-7, 12-15  epilogue (two per frame / lane quadrant)
//   warps 8-11  block exchange (one lane quadrant each)
#include <stdlib.h>

#include "tc_common.cuh"

#ifndef ISTGCN_TC3_PROF
#define ISTGCN_TC3_PROF 0      // 1: CTA 0 prints per-role wait cycles (tools/bench_gcn_fwd.py --iters 1)
#endif
#if ISTGCN_TC3_PROF
#include <stdio.h>
#define PROF_T0() const long long _t0 = clock64()
#define PROF_ADD(acc) acc += clock64() - _t0
#else
#define PROF_T0()
#define PROF_ADD(acc)
#endif

namespace istgcn {
namespace tc {

constexpr int kThreads3 = 512;
constexpr int kSlot3 = 32;                             // padded rows of one frame
constexpr int kFr3 = 4;                                // frames per tile
constexpr int kXStage3 = kFr3 * kSlot3 * 128;          // one 32-channel slice of a tile
constexpr int kAdj3 = 0;                               // TMEM: stacked adjacency [0, 32)
                                                       //       exchange buffers [kSB, kSB + 128*NSB)
constexpr int kExch3 = 12 * 4096;                      // off-diagonal blocks in flight

template <int NCOLS>
struct Cfg3 {
    static constexpr int WU = NCOLS > 128 ? 128 : NCOLS;          // weight rows per TMA box
    static constexpr int NU = NCOLS / WU;                         // boxes per partition
    static constexpr int GW = NCOLS == 64 ? 4 : 1;                // weight boxes per stage (one barrier)
    static constexpr int WBYTES = WU * 128;
    static constexpr int WSTAGE = GW * WBYTES;                    // 32 / 16 / 16 KB
    static constexpr int NW = NCOLS == 64 ? 2 : 5;                // weights arrive with L2 latency: deep ring
    static constexpr int NX = NCOLS == 64 ? 5 : (NCOLS == 128 ? 3 : 2);
    static constexpr int NSB = 2;                                 // exchange buffers in tensor memory
    static constexpr int ND2 = NCOLS == 64 ? 2 : 1;
    // Cout >= 128: the accumulators (2 x 128 or 1 x 256 columns) + 2 x 128 exchange columns are all of
    // tensor memory; there the stacked adjacency is a shared-memory (SS-mode) operand instead
    static constexpr bool ADJS = NCOLS == 256;
    static constexpr int kSB = ADJS ? 0 : 32;
    static constexpr int kD2Col = 512 - ND2 * NCOLS;
    static_assert(kSB + NSB * 128 <= kD2Col, "tensor-memory budget exceeded");
    static constexpr int x_off = 0;
    static constexpr int w_off = x_off + NX * kXStage3;
    static constexpr int exch_off = w_off + NW * WSTAGE;
    static constexpr int stage_off = exch_off + kExch3;
    static constexpr int adj_off = stage_off + 8 * 4096;          // one staging tile per epilogue warp
    static constexpr int bias_off = adj_off + (ADJS ? 16384 : 0);
    static constexpr int stat_off = bias_off + 4 * NCOLS * 4;
    static constexpr int bar_off = stat_off + 2 * NCOLS * 8;
    static constexpr int kNumBars = 2 * NX + 2 * NW + 8;
    static constexpr int total = bar_off + kNumBars * 8 + 16;
    static_assert(total <= 232448, "shared-memory budget exceeded");
    static_assert(NX * kXStage3 >= 4 * 32 * 33 * 4, "adjacency scratch lives in the input ring");
};

struct GcnTc3Params {
    const float* vals;
    const int *lptr, *lsrc, *lid;
    const float *bias_k, *colsum;
    double *stat_sum, *stat_sumsq;
    int frames, V, K, Cin, Cout, tiles, reduce;
    int variant;        // 0 except in ISTGCN_TC3_PROF builds: ablation bits for timing experiments (1 exchange
                        // without shared-memory traffic, 2 no MMA 1, 4 no epilogue work, 8 no exchange,
                        // 16 no weight reloads, 32 / 64 / 128 no sums / bias / stores; results are wrong)
};

__device__ __forceinline__ void named_bar(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// four 32-column loads in flight, one wait; the empty asm pins every later use behind the wait
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tmem_ld_pin(uint32_t (&r)[32]) {
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
                   "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]),
                   "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),
                   "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                   "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}
__device__ __forceinline__ void tmem_st32_issue(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// Exchange role of lane quadrant Q (compile-time: the register blocks are indexed statically).
// Before: lanes (k = Q, w), column block f = frame.  After: lanes (f = Q, w), column block k.
// Two blocks in registers at a time (the CTA runs 512 threads, 128 registers each).
__device__ __forceinline__ void exch_put(uint8_t* dst, int sw, const uint32_t (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j)                       // + half a TF32 ulp: the tensor core truncates
        *reinterpret_cast<uint4*>(dst + ((j ^ sw) << 4)) =
            make_uint4(v[4 * j] + 0x1000u, v[4 * j + 1] + 0x1000u, v[4 * j + 2] + 0x1000u, v[4 * j + 3] + 0x1000u);
}
__device__ __forceinline__ void exch_get(const uint8_t* src, int sw, uint32_t (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint4 t = *reinterpret_cast<const uint4*>(src + ((j ^ sw) << 4));
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int Q, int NSB>
__device__ __forceinline__ void exchange_role(uint32_t tmem_sb, uint8_t* exch, uint64_t* d1_full,
                                              uint64_t* a2_ready, uint32_t total, int lane, int variant) {
    const uint32_t lane_base = tmem_sb + (static_cast<uint32_t>(Q * 32) << 16);
    const int sw = lane & 7;
    constexpr int F0 = (Q + 1) & 3, F1 = (Q + 2) & 3, F2 = (Q + 3) & 3;
    auto slot = [&](int k, int f) { return exch + (k * 3 + (f > k ? f - 1 : f)) * 4096 + lane * 128; };
#if ISTGCN_TC3_PROF
    long long w_d1 = 0, w_bar = 0;
    const long long t_begin = clock64();
#endif
    for (uint32_t s = 0; s < total; ++s) {
        const uint32_t b = s % NSB;
        {
            PROF_T0();
            mbar_wait(&d1_full[b], (s / NSB) & 1);
            PROF_ADD(w_d1);
        }
        tc_fence_after();
        if (variant & 8) {
            tc_fence_before();
            if (lane == 0) mbar_arrive(&a2_ready[b]);
            continue;
        }
        const uint32_t base = lane_base + b * 128;
        uint32_t x[32], y[32];
        // my three off-diagonal blocks leave; the diagonal one is rounded in place
        tmem_ld32_issue(base + F0 * 32, x);
        tmem_ld32_issue(base + F1 * 32, y);
        tmem_wait_ld();
        tmem_ld_pin(x);
        tmem_ld_pin(y);
        if (!(variant & 1)) { exch_put(slot(Q, F0), sw, x); exch_put(slot(Q, F1), sw, y); }
        tmem_ld32_issue(base + F2 * 32, x);
        tmem_ld32_issue(base + Q * 32, y);
        tmem_wait_ld();
        tmem_ld_pin(x);
        tmem_ld_pin(y);
        if (!(variant & 1)) exch_put(slot(Q, F2), sw, x);
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] += 0x1000u;
        tmem_st32_issue(base + Q * 32, y);
        {
            PROF_T0();
            named_bar(1, 128);
            PROF_ADD(w_bar);
        }
        // the blocks of the other quadrants arrive
        if (!(variant & 1)) { exch_get(slot(F0, Q), sw, x); exch_get(slot(F1, Q), sw, y); }
        tmem_st32_issue(base + F0 * 32, x);
        tmem_st32_issue(base + F1 * 32, y);
        if (!(variant & 1)) exch_get(slot(F2, Q), sw, x);
        tmem_st32_issue(base + F2 * 32, x);
        tmem_wait_st();
        tc_fence_before();
        named_bar(2, 128);                   // every read of the scratch is done: next slice may write
        if (lane == 0) mbar_arrive(&a2_ready[b]);
    }
#if ISTGCN_TC3_PROF
    if (blockIdx.x == 0 && lane == 0)
        printf("tc3 exch%d: total %lld  wait d1_full %lld  bar1 %lld  (%u slices)\n", Q, clock64() - t_begin, w_d1,
               w_bar, total);
#endif
}

template <int NCOLS>
__global__ void __launch_bounds__(kThreads3, 1)
gcn_tc3_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap wmap,
               const __grid_constant__ CUtensorMap omap, GcnTc3Params p) {
    using L = Cfg3<NCOLS>;
    constexpr int NX = L::NX, NW = L::NW, NSB = L::NSB, ND2 = L::ND2, WU = L::WU, NU = L::NU, GW = L::GW;
    constexpr int kSB = L::kSB;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* Xs = smem + L::x_off;
    uint8_t* Ws = smem + L::w_off;
    float* s_bias = reinterpret_cast<float*>(smem + L::bias_off);
    double* s_sum = reinterpret_cast<double*>(smem + L::stat_off);
    double* s_sq = s_sum + NCOLS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bar_off);
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + NX;
    uint64_t* b_full = x_empty + NX;
    uint64_t* b_empty = b_full + NW;
    uint64_t* t_full = b_empty + NW;
    uint64_t* t_empty = t_full + 2;
    uint64_t* d1_full = t_empty + 2;
    uint64_t* a2_ready = d1_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::kNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = p.V, K = p.K, Cout = p.Cout;
    const int nchunk = p.Cin / 32;
    const int nbox = K * NU;                            // weight boxes per slice (partition, column half)
    const int ngrp = (nbox + GW - 1) / GW;              // weight stages per slice
    // the whole weight set fits the ring exactly (Cin = Cout = 64): load it once, never release it
    const bool w_resident = nchunk * ngrp == NW;
    const int my_tiles = p.tiles > (int)blockIdx.x
                             ? (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    // ---- one-time setup: barriers, TMEM, bias factors; the dense adjacency goes through a
    // scratch table [k][w][33] in the (still idle) input ring into tensor memory
    float* adjT = reinterpret_cast<float*>(Xs);
    for (int i = tid; i < 4 * 32 * 33; i += kThreads3) adjT[i] = 0.f;
    for (int i = tid; i < 2 * NCOLS; i += kThreads3) s_sum[i] = 0.0;
    if (p.bias_k)
        for (int i = tid; i < K * NCOLS; i += kThreads3) {
            const int k = i / NCOLS, c = i % NCOLS;
            s_bias[i] = c < Cout ? p.bias_k[(size_t)k * Cout + c] : 0.f;
        }
    if (tid == 0) {
        for (int i = 0; i < NX; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < NW; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 8);
            mbar_init(&d1_full[i], 1); mbar_init(&a2_ready[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) tma_prefetch_desc(&wmap);
    if (warp == 2 && lane == 0) tma_prefetch_desc(&xmap);
    if (warp == 4 && lane == 0) tma_prefetch_desc(&omap);
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    for (int i = tid; i < K * V; i += kThreads3) {          // one thread per (k, w): no races
        const int k = i / V, w = i - k * V;
        for (int j = p.lptr[i]; j < p.lptr[i + 1]; ++j)
            adjT[(k * 32 + w) * 33 + p.lsrc[j]] += __uint_as_float(to_tf32(p.vals[p.lid[j]]));
    }
    __syncthreads();
    if (L::ADJS) {                                          // rows (k, w), K-major SWIZZLE_128B operand
        float* adjS = reinterpret_cast<float*>(smem + L::adj_off);
        for (int i = tid; i < 128 * 32; i += kThreads3) {
            const int row = i >> 5, e = i & 31;
            adjS[atom_index(row, e)] = (row >> 5) < K ? adjT[row * 33 + e] : 0.f;
        }
        fence_proxy_async();
    } else if (warp >= 4 && warp < 8) {                     // lanes (k, w): quadrant = partition
        const int k = warp - 4;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = k < K ? adjT[(k * 32 + lane) * 33 + j] : 0.f;
        tmem_st32(tmem_base + (static_cast<uint32_t>(k * 32) << 16) + kAdj3, v);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // pad rows of the frame slots must be finite: zero the whole ring once (TMA never writes them)
    for (int i = tid; i < NX * kXStage3 / 16; i += kThreads3)
        reinterpret_cast<float4*>(Xs)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    __syncthreads();

    if (warp == 0) {
        // =========================== TMA producer: weight stages (<= KG partitions, one barrier)
        uint32_t it = 0;
        for (int t = 0; t < my_tiles; ++t)
            for (int ch = 0; ch < nchunk; ++ch)
                for (int g = 0; g < ngrp; ++g, ++it) {
                    const int sb = it % NW;
                    const int bn = min(GW, nbox - g * GW);
                    if (w_resident && t > 0) continue;
                    mbar_wait(&b_empty[sb], ((it / NW) & 1) ^ 1);
                    if ((p.variant & 16) && it >= (uint32_t)NW) {
                        if (elect_one()) mbar_arrive(&b_full[sb]);
                        __syncwarp();
                        continue;
                    }
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&b_full[sb], bn * L::WBYTES);
                        for (int j = 0; j < bn; ++j) {
                            const int bx = g * GW + j, k = bx / NU, u = bx % NU;
                            tma_load_2d(Ws + sb * L::WSTAGE + j * L::WBYTES, &wmap, &b_full[sb], ch * 32,
                                        k * Cout + u * WU);
                        }
                    }
                    __syncwarp();
                }
    } else if (warp == 2) {
        // =========================== TMA producer: the frames of the tile, one slot each
        uint32_t it = 0;
        const uint32_t bytes = kFr3 * V * 128;
        for (int t = 0; t < my_tiles; ++t) {
            const int f0 = (blockIdx.x + t * gridDim.x) * kFr3;
            for (int ch = 0; ch < nchunk; ++ch, ++it) {
                const int xs = it % NX;
                mbar_wait(&x_empty[xs], ((it / NX) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&x_full[xs], bytes);
#pragma unroll
                    for (int f = 0; f < kFr3; ++f)
                        tma_load_3d(Xs + xs * kXStage3 + f * (kSlot3 * 128), &xmap, &x_full[xs],
                                    ch * 32, 0, f0 + f);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer
        if (my_tiles > 0) {
            constexpr uint32_t idesc1 = make_idesc(128, 32, false, true);
            constexpr uint32_t idesc2 = make_idesc(128, WU, false, false);
            const uint32_t total = (uint32_t)my_tiles * nchunk;
            const uint32_t xs0 = smem_u32(Xs), ws0 = smem_u32(Ws);
            const uint32_t adj = tmem_base + kAdj3;
            const uint32_t adj_s = smem_u32(smem + L::adj_off);
#if ISTGCN_TC3_PROF
            long long w_x = 0, w_t = 0, w_a2 = 0, w_b = 0;
            const long long t_begin = clock64();
#endif
            uint32_t xsb = 0, xph = 0;                  // input ring slot and its barrier phase
            const uint64_t adj_desc = make_desc(adj_s, 16, 1024);
            auto issue1 = [&](uint32_t s) {             // called for s = 0, 1, 2, ... in order
                const uint32_t xs = xsb;
                {
                    PROF_T0();
                    mbar_wait(&x_full[xs], xph);
                    PROF_ADD(w_x);
                }
                if (++xsb == NX) { xsb = 0; xph ^= 1; }
                tc_fence_after();
                const uint32_t d1 = tmem_base + kSB + (s % NSB) * 128;
                if (elect_one()) {
                    if (!(p.variant & 2))
#pragma unroll
                    for (int f = 0; f < kFr3; ++f) {
                        const uint64_t xdesc = make_desc(xs0 + xs * kXStage3, 4096, 512, 1);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t bd = xdesc + (uint64_t)((f * (kSlot3 * 128) + ks * 1024) >> 4);
                            if (L::ADJS)
                                tc_mma_tf32(d1 + f * 32, adj_desc + (uint64_t)(ks * 2), bd, idesc1, ks ? 1u : 0u);
                            else
                                tc_mma_tf32_ts(d1 + f * 32, adj + ks * 8, bd, idesc1, ks ? 1u : 0u);
                        }
                    }
                    tc_commit(&x_empty[xs]);
                    tc_commit(&d1_full[s % NSB]);
                }
                __syncwarp();
            };
            uint32_t wsb = 0, wph = 0;                  // weight ring slot and its barrier phase
            for (uint32_t s = 0; s < (uint32_t)NSB && s < total; ++s) issue1(s);
            for (uint32_t s = 0; s < total; ++s) {
                const uint32_t t = s / nchunk, ch = s - t * nchunk;
                const uint32_t buf = t % ND2, use = t / ND2;
                if (ch == 0) {
                    PROF_T0();
                    mbar_wait(&t_empty[buf], (use & 1) ^ 1);
                    PROF_ADD(w_t);
                    tc_fence_after();
                }
                {
                    PROF_T0();
                    mbar_wait(&a2_ready[s % NSB], (s / NSB) & 1);
                    PROF_ADD(w_a2);
                }
                tc_fence_after();
                const uint32_t d2 = tmem_base + L::kD2Col + buf * NCOLS;
                const uint32_t a2 = tmem_base + kSB + (s % NSB) * 128;
                // fully unrolled, descriptors = base + constant: the issuing thread must spend fewer
                // cycles per instruction than the tensor pipe does (16 / 32 / 64 at N = 32 / 64 / 128)
#pragma unroll
                for (int g = 0; g < 4 * NU / GW; ++g) {
                    if (g >= ngrp) break;
                    if (!w_resident || t == 0) {
                        PROF_T0();
                        mbar_wait(&b_full[wsb], wph);
                        PROF_ADD(w_b);
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t bdesc = make_desc(ws0 + wsb * L::WSTAGE, 16, 1024);
#pragma unroll
                        for (int j = 0; j < GW; ++j) {
                            constexpr int kDummy = 0; (void)kDummy;
                            const int bx = g * GW + j, k = bx / NU, u = bx % NU;     // compile-time
                            if (bx < nbox) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    tc_mma_tf32_ts(d2 + u * WU, a2 + k * 32 + ks * 8,
                                                   bdesc + (uint64_t)((j * L::WBYTES + ks * 32) >> 4), idesc2,
                                                   (k | ks) ? 1u : (ch ? 1u : 0u));
                            }
                        }
                        if (!w_resident) tc_commit(&b_empty[wsb]);
                        if (ch == (uint32_t)nchunk - 1 && g == ngrp - 1) tc_commit(&t_full[buf]);
                    }
                    __syncwarp();
                    if (++wsb == NW) { wsb = 0; wph ^= 1; }
                }
                // the next slice for this exchange buffer: executes after MMA 2 above (issue order)
                if (s + NSB < total) issue1(s + NSB);
            }
#if ISTGCN_TC3_PROF
            if (blockIdx.x == 0 && lane == 0)
                printf("tc3 issuer: total %lld  wait x_full %lld  t_empty %lld  a2_ready %lld  b_full %lld  (%u slices)\n",
                       clock64() - t_begin, w_x, w_t, w_a2, w_b, total);
#endif
        }
    } else if (warp >= 8 && warp < 12) {
        // =========================== block exchange: D1 (partition on lanes) -> A2 (frame on lanes)
        const uint32_t total = (uint32_t)my_tiles * nchunk;
        uint8_t* exch = smem + L::exch_off;
        switch (warp - 8) {
            case 0: exchange_role<0, NSB>(tmem_base + kSB, exch, d1_full, a2_ready, total, lane, p.variant); break;
            case 1: exchange_role<1, NSB>(tmem_base + kSB, exch, d1_full, a2_ready, total, lane, p.variant); break;
            case 2: exchange_role<2, NSB>(tmem_base + kSB, exch, d1_full, a2_ready, total, lane, p.variant); break;
            default: exchange_role<3, NSB>(tmem_base + kSB, exch, d1_full, a2_ready, total, lane, p.variant); break;
        }
    } else if (warp >= 4) {
        // =========================== epilogue: two warps per lane quadrant (warps 4-7 take the even
        // 32-column blocks of the accumulator, warps 12-15 the odd ones); quadrant ew = frame ew
        const int ew = warp & 3, half = warp >= 12 ? 1 : 0;
        const int w = lane;
        uint8_t* stage = smem + L::stage_off + (half * 4 + ew) * 4096;    // [32 rows][128 B], SWIZZLE_128B
        constexpr int NB = NCOLS / 64;                        // column blocks of this warp
        double acc_s[NB], acc_q[NB];                          // their column sums (lane = column)
#pragma unroll
        for (int i = 0; i < NB; ++i) acc_s[i] = acc_q[i] = 0.0;
        float cs[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.bias_k && w < V) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < K) cs[k] = p.colsum[k * V + w];
        }
        // one column block per warp (Cout <= 64): its bias term sum_k colsum[k][w] * bias_k[n] stays in
        // registers for the whole kernel; otherwise it is rebuilt per block from shared memory
        float bt[NB == 1 ? 32 : 1];
        if (NB == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (p.bias_k && k < K) a = fmaf(cs[k], s_bias[k * NCOLS + half * 32 + j], a);
                bt[NB == 1 ? j : 0] = a;
            }
        }
        // column `lane` of staging row r sits at srow[r & 7] + (r >> 3) * 1024
        const uint8_t* srow[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) srow[i] = stage + i * 128 + (((lane >> 2) ^ i) << 4) + (lane & 3) * 4;
#if ISTGCN_TC3_PROF
        long long w_tf = 0, w_bulk = 0;
        const long long t_begin = clock64();
#endif
        for (int t = 0; t < my_tiles; ++t) {
            const uint32_t buf = t % ND2, use = t / ND2;
            const int frame = (blockIdx.x + t * gridDim.x) * kFr3 + ew;
            const bool fok = frame < p.frames;
            {
                PROF_T0();
                mbar_wait(&t_full[buf], use & 1);
                PROF_ADD(w_tf);
            }
            tc_fence_after();
            // one 32-column block: bias term, staging tile, TMA store / reduce-add, BatchNorm sums
            auto finish_block = [&](float (&v)[32], int c0, int ib) {
                if (p.variant & 4) return;
                if (NB == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += bt[NB == 1 ? j : 0];
                } else if (p.bias_k && !(p.variant & 64)) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k >= K) break;
                        const float ck = cs[k];
                        const float* b = s_bias + k * NCOLS + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bv = *reinterpret_cast<const float4*>(b + j);
                            v[j] = fmaf(ck, bv.x, v[j]); v[j + 1] = fmaf(ck, bv.y, v[j + 1]);
                            v[j + 2] = fmaf(ck, bv.z, v[j + 2]); v[j + 3] = fmaf(ck, bv.w, v[j + 3]);
                        }
                    }
                }
                {
                    PROF_T0();
                    if (lane == 0) bulk_wait_read();          // the previous store has read the tile
                    __syncwarp();
                    PROF_ADD(w_bulk);
                }
                if (!(p.variant & 128))
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0 && fok && !(p.variant & 128)) {
                    if (p.reduce) tma_reduce_add_2d(stage, &omap, c0, frame * V);
                    else tma_store_2d(stage, &omap, c0, frame * V);
                    bulk_commit();
                }
                if (p.stat_sum && fok && !(p.variant & 32)) {
                    // column sums straight from the staging tile: lane c adds column c of all 32 rows
                    // (rows >= V are exact zeros: zero adjacency rows, no bias term), one conflict-free
                    // wavefront per row, addresses = 8 bases + immediates; the TMA store reads the tile
                    // concurrently.  Then the warp's double accumulators.
                    float a[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float x = *reinterpret_cast<const float*>(srow[r & 7] + (r >> 3) * 1024);
                        a[r & 3] += x;
                        q[r & 3] = fmaf(x, x, q[r & 3]);
                    }
                    acc_s[ib] += (double)((a[0] + a[1]) + (a[2] + a[3]));
                    acc_q[ib] += (double)((q[0] + q[1]) + (q[2] + q[3]));
                }
            };
            const uint32_t d2 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + L::kD2Col + buf * NCOLS;
            // two column blocks leave tensor memory at a time, so that the accumulator goes back to the
            // MMA issuer before the slow part (single-buffered accumulators at Cout >= 128)
#pragma unroll
            for (int ib = 0; ib < NB; ib += 2) {
                const int c0 = (2 * ib + half) * 32, c1 = c0 + 64;
                const bool live0 = c0 < Cout, live1 = NB > 1 && c1 < Cout;     // warp-uniform
                const bool last = ib + 2 >= NB || c1 + 64 >= Cout;
                uint32_t r0[32], r1[32];
                if (live0) tmem_ld32_issue(d2 + c0, r0);
                if (live1) tmem_ld32_issue(d2 + c1, r1);
                tmem_wait_ld();
                if (live0) tmem_ld_pin(r0);
                if (live1) tmem_ld_pin(r1);
                if (last) {                                   // this warp's last read of the accumulator
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[buf]);
                }
                float v[32];
                if (live0) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]);
                    finish_block(v, c0, ib);
                }
                if (NB > 1 && live1) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r1[i]);
                    finish_block(v, c1, ib + 1);
                }
                if (last) break;
            }
        }
#if ISTGCN_TC3_PROF
        if (blockIdx.x == 0 && lane == 0)
            printf("tc3 epi%d.%d: total %lld  wait t_full %lld  bulk_wait_read %lld  (%d tiles)\n", ew, half,
                   clock64() - t_begin, w_tf, w_bulk, my_tiles);
#endif
        if (lane == 0) bulk_wait_all();
        if (p.stat_sum) {
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                atomicAdd(&s_sum[(2 * i + half) * 32 + lane], acc_s[i]);
                atomicAdd(&s_sq[(2 * i + half) * 32 + lane], acc_q[i]);
            }
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (p.stat_sum) {
        for (int c = tid; c < NCOLS; c += kThreads3) {
            if (c < Cout) {
                atomicAdd(&p.stat_sum[c], s_sum[c]);
                atomicAdd(&p.stat_sumsq[c], s_sq[c]);
            }
        }
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int NCOLS>
static int launch_tc3_n(const CUtensorMap& xmap, const CUtensorMap& wmap, const CUtensorMap& omap,
                        const GcnTc3Params& p, cudaStream_t s) {
    using L = Cfg3<NCOLS>;
    auto kern = gcn_tc3_kernel<NCOLS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
    int nx = num_sms();
    if (nx > p.tiles) nx = p.tiles;
    kern<<<nx, kThreads3, L::total, s>>>(xmap, wmap, omap, p);
    return finish_launch("gcn_tc3");
}

// Same shapes as the second-generation engine (gcn_tc2_eligible decides).
int launch_gcn_tc3(const float* in, const float* w_rows, const float* vals, const int* lptr,
                   const int* lsrc, const int* lid, const float* bias_k, const float* colsum, float* out,
                   int reduce, double* stat_sum, double* stat_sumsq, int frames, int V, int K, int Cin,
                   int Cout, cudaStream_t st) {
    GcnTc3Params p{vals, lptr, lsrc, lid, bias_k, colsum, stat_sum, stat_sumsq, frames, V, K, Cin, Cout,
                   (frames + kFr3 - 1) / kFr3, reduce, 0};
#if ISTGCN_TC3_PROF
    static const char* var_env = getenv("ISTGCN_TC3_VARIANT");
    if (var_env) p.variant = atoi(var_env);
#endif
    const int ncols = Cout > 128 ? 256 : (Cout > 64 ? 128 : 64);
    CUtensorMap xmap, wmap, omap;
    if (int e = encode_frame_slices(&xmap, in, frames, V, Cin)) return e;
    if (int e = encode_tile_map(&wmap, w_rows, (long long)K * Cout, Cin, ncols > 128 ? 128 : ncols)) return e;
    if (int e = encode_tile_map(&omap, out, (long long)frames * V, Cout, V)) return e;
    if (ncols == 256) return launch_tc3_n<256>(xmap, wmap, omap, p, st);
    if (ncols == 128) return launch_tc3_n<128>(xmap, wmap, omap, p, st);
    return launch_tc3_n<64>(xmap, wmap, omap, p, st);
}

}  // namespace tc
}  // namespace istgcn
