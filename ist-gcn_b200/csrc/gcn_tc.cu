// tcgen05 engine of the fused graph convolution (reference: net/utils/tgcn.py:76-89,
// net/utils/inceptionv2_gcn.py:64-89).  Same math as gcn.cu:
//
//     OUT[(f,w)][n] = sum_{k,ci} ( sum_v A_eff[k][v][w] * IN[(f,v)][ci] ) * W[k][n][ci]  (+ epilogue)
//
// but the channel-mix GEMM runs on the 5th-generation tensor cores: a persistent CTA per SM owns
// frame tiles (128 rows = floor(128/V) frames), warp-specialised:
//
//   warp 0        TMA producer: weight atoms W[k][n0..][ci0..ci0+32) -> smem (SWIZZLE_128B), ring
//   warp 1        MMA issuer: tcgen05.mma kind::tf32, M=128, N=NCOLS, K=8 per instruction,
//                 accumulators in TMEM (2 x NCOLS columns, double buffered across tiles)
//   warps 2-3,16-17  loaders: 32-channel slices of the input rows -> smem (optionally through
//                 the BatchNorm-backward transform, so the same kernel computes the input gradient)
//   warps 4-7     epilogue: tcgen05.ld (one TMEM lane = one row per thread), + bias term /
//                 residual, BatchNorm statistics via shuffle column sums, global stores
//   warps 8-15    aggregators: sparse A_eff aggregation of the slice on CUDA cores, written
//                 straight into the K-major SWIZZLE_128B operand atom the tensor core reads
//
// Forward:          IN = x,  lists grouped by destination (k,w), W = conv weight (K*Cout, Cin)
// Input gradient:   IN = dz (formed on load), lists grouped by (k,v) (transposed adjacency),
//                   W = Wc (K*Cin, Cout), epilogue adds the residual gradient.
// Every mbarrier wait is bounded (trap on timeout) so a pipeline bug cannot hang the device.
#include <stdlib.h>

#include "tc_common.cuh"

namespace istgcn {
namespace tc {

#ifndef ISTGCN_TC_PROF
#define ISTGCN_TC_PROF 0      // 1: per-role clock64 accounting of CTA 0 (tools/dbg_tc_time.py)
#endif
#if ISTGCN_TC_PROF
__device__ unsigned long long g_prof[32];
#define PROF_DECL long long prof[32] = {0}
#define PROF_T0() long long _t0 = clock64()
#define PROF_ADD(i) do { long long _t1 = clock64(); if (blockIdx.x == 0 && blockIdx.y == 0) prof[i] += _t1 - _t0; _t0 = _t1; } while (0)
#else
#define PROF_T0()
#define PROF_ADD(i)
#define PROF_DECL
#endif

constexpr int kThreadsTC = 576;     // 18 warps, see the role table above
constexpr int kAggThreads = 256;    // warps 8-15
// Ring geometry per output-tile width (smem budget 227 KB).  The aggregated operand ring is made
// of NH half-slots of two atoms each; the aggregators hand a whole 32-channel slice (all K
// partitions = up to two half-slots) to the MMA issuer with ONE barrier round trip, because a
// round trip costs ~600 cycles and used to dominate the slice time when paid per atom.  The
// weight ring must cover the TMA round trip (~1.5 us) or the MMA issuer starves on b_full.
template <int NCOLS> struct Rings { static constexpr int NH = 3, NB = 2, NX = 2; };
template <> struct Rings<64> { static constexpr int NH = 3, NB = 6, NX = 3; };
template <> struct Rings<128> { static constexpr int NH = 3, NB = 4, NX = 2; };
constexpr int kNC = 4;            // slice-ready barriers (>= slices in flight)

struct BnBackTC {
    const float *p, *m1, *cc, *mu;
};

struct GcnTcParams {
    const float* in;              // [rows][Cin]
    const float* in2;             // z for the BatchNorm-backward transform (or NULL)
    BnBackTC bn;
    const float* vals;
    const int *lptr, *lsrc, *lid; // lists grouped by (k, destination joint)
    const float* bias_k;          // [K][Cout] conv bias and colsum[K][V] = column sums of A_eff[k]:
    const float* colsum;          //   bias term = sum_k colsum[k][w] * bias_k[k][n]  (forward) or NULL
    const float* add_rows;        // [rows][Cout] added per row             (backward) or NULL
    float* out;                   // [rows][Cout]
    float* in_out;                // optional copy of the (transformed) input, [rows][Cin]
    double *stat_sum, *stat_sumsq;
    int frames, V, K, Cin, CinPad, Cout, nnz, tiles;
    FrameMap in_map, out_map;     // temporal stride of the residual branch (t_out == 0: none)
    int tma_out;                  // 0: direct stores, 1: TMA tile store, 2: TMA reduce-add (add_rows == out)
};

template <int NCOLS>
struct SmemLayout {
    static constexpr int kNH = Rings<NCOLS>::NH, kNA = 2 * kNH, kNB = Rings<NCOLS>::NB;
    static constexpr int kBAtomBytes = NCOLS * 128;
    static constexpr int a_off = 0;
    static constexpr int b_off = a_off + kNA * kAtomBytes;
    static constexpr int kNX = Rings<NCOLS>::NX;
    static constexpr int x_off = b_off + kNB * kBAtomBytes;
    static constexpr int stage_off = x_off + kNX * kAtomRows * 32 * 4;           // 4 x [32 rows][128 B]
    static constexpr int list_off = stage_off + 4 * 4096;                        // vals, src, ptr
    static constexpr int bias_off = list_off + kMaxNnz * 8 + (kMaxKV + 4) * 4;   // [4][NCOLS] floats
    static constexpr int stat_off = bias_off + 4 * NCOLS * 4;                    // 2 * NCOLS doubles
    static constexpr int bar_off = stat_off + 2 * NCOLS * 8;
    static_assert(bar_off + 512 <= 232448, "shared-memory budget exceeded");
    static constexpr int kNumBars = kNC + kNH + 2 * kNB + 2 * kNX + 4;
    static constexpr int total = bar_off + kNumBars * 8 + 16;
};

template <int NCOLS>
__global__ void __launch_bounds__(kThreadsTC, 1)
gcn_tc_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap omap,
              const __grid_constant__ CUtensorMap omap_last, GcnTcParams p) {
    using L = SmemLayout<NCOLS>;
    constexpr int kNA = L::kNA, kNH = L::kNH, kNB = L::kNB, kNX = L::kNX;
    extern __shared__ __align__(1024) uint8_t smem[];
    float* As = reinterpret_cast<float*>(smem + L::a_off);
    uint8_t* Bs = smem + L::b_off;
    float* Xs = reinterpret_cast<float*>(smem + L::x_off);
    int2* s_ent = reinterpret_cast<int2*>(smem + L::list_off);     // {source row offset, value}
    int* s_ptr = reinterpret_cast<int*>(s_ent + kMaxNnz);
    float* s_bias = reinterpret_cast<float*>(smem + L::bias_off);
    double* s_sum = reinterpret_cast<double*>(smem + L::stat_off);   // double: see gcn.cu
    double* s_sq = s_sum + NCOLS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bar_off);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + kNC;
    uint64_t* b_full = a_empty + kNH;
    uint64_t* b_empty = b_full + kNB;
    uint64_t* x_full = b_empty + kNB;
    uint64_t* x_empty = x_full + kNX;
    uint64_t* t_full = x_empty + kNX;
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::kNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = p.V, K = p.K, Cin = p.Cin, Cout = p.Cout;
    const int F = (kAtomRows / V) > 8 ? 8 : (kAtomRows / V);
    const int nchunk = p.CinPad / 32;
    const int nh = (K + 1) >> 1;                    // half-slots per slice
    const int n0 = blockIdx.y * NCOLS;

    // ---- one-time setup
    for (int i = tid; i < p.nnz; i += kThreadsTC)
        s_ent[i] = make_int2(p.lsrc[i] * 32, __float_as_int(p.vals[p.lid[i]]));
    for (int i = tid; i <= K * V; i += kThreadsTC) s_ptr[i] = p.lptr[i];
    for (int i = tid; i < kNA * kAtomBytes / 4; i += kThreadsTC) As[i] = 0.f;
    for (int i = tid; i < 2 * NCOLS; i += kThreadsTC) s_sum[i] = 0.0;
    if (p.bias_k)
        for (int i = tid; i < K * NCOLS; i += kThreadsTC) {
            const int k = i / NCOLS, c = i % NCOLS;
            s_bias[i] = n0 + c < Cout ? p.bias_k[(size_t)k * Cout + n0 + c] : 0.f;
        }
    if (tid == 0) {
        for (int i = 0; i < kNC; ++i) mbar_init(&a_full[i], kAggThreads / 32);
        for (int i = 0; i < kNH; ++i) mbar_init(&a_empty[i], 1);
        for (int i = 0; i < kNB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < kNX; ++i) { mbar_init(&x_full[i], 4); mbar_init(&x_empty[i], kAggThreads / 32); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&wmap);
        if (p.tma_out) { tma_prefetch_desc(&omap); tma_prefetch_desc(&omap_last); }
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * NCOLS);
    fence_proxy_async();                       // zero-filled A atoms visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =========================== TMA producer (weights)
        if (lane == 0) {
            uint32_t it = 0;
            PROF_DECL; PROF_T0();
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x)
                for (int ch = 0; ch < nchunk; ++ch)
                    for (int k = 0; k < K; ++k, ++it) {
                        const int sb = it % kNB;
                        PROF_ADD(1);
                        mbar_wait(&b_empty[sb], ((it / kNB) & 1) ^ 1);
                        PROF_ADD(0);
                        mbar_arrive_expect_tx(&b_full[sb], L::kBAtomBytes);
                        // weight matrix rows = k*Cout + n, cols = ci
                        tma_load_2d(Bs + sb * L::kBAtomBytes, &wmap, &b_full[sb], ch * 32,
                                    k * Cout + n0);
                    }
            PROF_ADD(1);
#if ISTGCN_TC_PROF
            if (blockIdx.x == 0 && blockIdx.y == 0) { g_prof[0] += prof[0]; g_prof[1] += prof[1]; }
#endif
        }
    } else if (warp == 1) {
        // =========================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, NCOLS, false, false);
            uint32_t it = 0, tcount = 0, cc = 0;
            PROF_DECL; PROF_T0();
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tcount) {
                const int buf = tcount & 1;
                PROF_ADD(5);
                mbar_wait(&t_empty[buf], ((tcount >> 1) & 1) ^ 1);
                PROF_ADD(2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * NCOLS;
                uint32_t first = 1;
                for (int ch = 0; ch < nchunk; ++ch, ++cc) {
                    PROF_ADD(5);
                    mbar_wait(&a_full[cc % kNC], (cc / kNC) & 1);
                    PROF_ADD(3);
                    for (int k = 0; k < K; ++k, ++it) {
                        const uint32_t hidx = cc * nh + (k >> 1);
                        const int sa = (hidx % kNH) * 2 + (k & 1), sb = it % kNB;
                        PROF_ADD(5);
                        mbar_wait(&b_full[sb], (it / kNB) & 1);
                        PROF_ADD(4);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(As) + sa * kAtomBytes;
                        const uint32_t b_addr = smem_u32(Bs) + sb * L::kBAtomBytes;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            tc_mma_tf32(d_tmem, make_desc(a_addr + ks * 32, 16, 1024),
                                        make_desc(b_addr + ks * 32, 16, 1024), idesc,
                                        first ? 0u : 1u);
                            first = 0;
                        }
                        tc_commit(&b_empty[sb]);
                        if ((k & 1) || k == K - 1) tc_commit(&a_empty[hidx % kNH]);
                    }
                }
                tc_commit(&t_full[buf]);
            }
            PROF_ADD(5);
#if ISTGCN_TC_PROF
            if (blockIdx.x == 0 && blockIdx.y == 0) for (int i = 2; i <= 5; ++i) g_prof[i] += prof[i];
#endif
        }
    } else if (warp >= 4 && warp < 8) {
        // =========================== epilogue: TMEM -> registers -> global
        const int ew = warp - 4;                        // == warp % 4: TMEM lanes 32*ew .. +31
        const int r = ew * 32 + lane;
        const int w = r % V;
        uint8_t* stage = smem + L::stage_off + ew * 4096;      // [32 rows][128 B], SWIZZLE_128B
        const CUtensorMap* om = ew == 3 ? &omap_last : &omap;   // warp 3 owns F*V - 96 rows
        float cs[4] = {0.f, 0.f, 0.f, 0.f};                    // colsum(A_eff[k])[w] of this row
        if (p.bias_k && r < F * V)
            for (int k = 0; k < K; ++k) cs[k] = p.colsum[k * V + w];
        uint32_t tcount = 0;
        PROF_DECL; PROF_T0();
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tcount) {
            const int buf = tcount & 1;
            const int f0 = tile * F;
            const int valid = min(F, p.frames - f0) * V;
            const long long row0 = (long long)f0 * V;
            const bool ok = r < valid;
            const bool tma = p.tma_out && valid == F * V;       // whole tile: coalesced TMA store
            PROF_ADD(8);
            mbar_wait(&t_full[buf], (tcount >> 1) & 1);
            PROF_ADD(7);
            tc_fence_after();
            for (int c0 = 0; c0 < NCOLS; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * NCOLS + c0, v);
                const int cg = n0 + c0;                  // global output column of v[0]
                if (p.bias_k) {
                    for (int k = 0; k < K; ++k) {
                        const float ck = cs[k];
                        const float* b = s_bias + k * NCOLS + c0;    // warp-uniform: broadcast reads
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bv = *reinterpret_cast<const float4*>(b + j);
                            v[j] = fmaf(ck, bv.x, v[j]); v[j + 1] = fmaf(ck, bv.y, v[j + 1]);
                            v[j + 2] = fmaf(ck, bv.z, v[j + 2]); v[j + 3] = fmaf(ck, bv.w, v[j + 3]);
                        }
                    }
                }
                if (tma) {
                    // registers -> swizzled staging tile -> one TMA store (or reduce-add) of
                    // 32 rows x 128 B: full-line writes instead of 32 partial lines per instruction
                    if (lane == 0) bulk_wait_read();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && cg < Cout) {
                        if (p.tma_out == 2) tma_reduce_add_2d(stage, om, cg, (int)(row0 + ew * 32));
                        else tma_store_2d(stage, om, cg, (int)(row0 + ew * 32));
                        bulk_commit();
                    }
                } else if (ok && map_row(p.out_map, row0 + r, V) >= 0) {
                    const long long orow = map_row(p.out_map, row0 + r, V);
                    float* o = p.out + orow * Cout + cg;
                    if ((Cout & 3) == 0) {
                        if (p.add_rows) {
                            const float* a = p.add_rows + orow * Cout + cg;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                if (cg + j < Cout) {
                                    const float4 av = ld4(a + j);
                                    v[j] += av.x; v[j + 1] += av.y; v[j + 2] += av.z; v[j + 3] += av.w;
                                }
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (cg + j < Cout) st4(o + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (cg + j < Cout)
                                o[j] = v[j] + (p.add_rows ? p.add_rows[orow * Cout + cg + j] : 0.f);
                    }
                }
                if (p.stat_sum) {
                    float q[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        v[j] = ok ? v[j] : 0.f;
                        q[j] = v[j] * v[j];
                    }
                    const float csum = warp_column_sums(v, lane);
                    const float cq = warp_column_sums(q, lane);
                    atomicAdd(&s_sum[c0 + lane], (double)csum);
                    atomicAdd(&s_sq[c0 + lane], (double)cq);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[buf]);
        }
        if (lane == 0) bulk_wait_all();
        PROF_ADD(8);
#if ISTGCN_TC_PROF
        if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 128) { g_prof[7] += prof[7]; g_prof[8] += prof[8]; }
#endif
    } else if (warp < 4 || warp >= 16) {
        // =========================== loaders: input slice -> Xs[xb][row][32]
        // 128 threads, 8 independent 16-byte loads each per slice (all issued before the first
        // use) so that ~16 KB per SM are in flight
        const int lt = warp < 4 ? tid - 64 : tid - 16 * 32 + 64;
        uint32_t it = 0;
        PROF_DECL; PROF_T0();
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int f0 = tile * F;
            const int valid = min(F, p.frames - f0) * V;
            const long long row0 = (long long)f0 * V;
            for (int ch = 0; ch < nchunk; ++ch, ++it) {
                const int xb = it % kNX;
                PROF_ADD(10);
                mbar_wait(&x_empty[xb], ((it / kNX) & 1) ^ 1);
                PROF_ADD(9);
                float* xs = Xs + xb * kAtomRows * 32;
                const int ci0 = ch * 32;
                if ((Cin & 3) == 0) {
                    float4 v[8], zv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = lt + u * 128;
                        const int r = i >> 3, c4 = (i & 7) * 4;
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        zv[u] = v[u];
                        const long long src = (r < valid && ci0 + c4 < Cin) ? map_row(p.in_map, row0 + r, V) : -1;
                        if (src >= 0) {
                            const long long off = src * Cin + ci0 + c4;
                            v[u] = ld4(p.in + off);
                            if (p.bn.p) zv[u] = ld4(p.in2 + off);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = lt + u * 128;
                        const int r = i >> 3, c4 = (i & 7) * 4;
                        if (p.bn.p && r < valid && ci0 + c4 < Cin) {
                            const float4 pv = ld4(p.bn.p + ci0 + c4), mv = ld4(p.bn.m1 + ci0 + c4),
                                         cv = ld4(p.bn.cc + ci0 + c4), uv = ld4(p.bn.mu + ci0 + c4);
                            v[u].x = bn_back(v[u].x, zv[u].x, pv.x, mv.x, cv.x, uv.x);
                            v[u].y = bn_back(v[u].y, zv[u].y, pv.y, mv.y, cv.y, uv.y);
                            v[u].z = bn_back(v[u].z, zv[u].z, pv.z, mv.z, cv.z, uv.z);
                            v[u].w = bn_back(v[u].w, zv[u].w, pv.w, mv.w, cv.w, uv.w);
                        }
                        if (p.in_out && r < valid && ci0 + c4 < Cin)
                            st4(p.in_out + (row0 + r) * Cin + ci0 + c4, v[u]);
                        st4(xs + r * 32 + c4, v[u]);
                    }
                } else {
                    for (int i = lt; i < kAtomRows * 32; i += 128) {
                        const int r = i >> 5, c = i & 31;
                        float v = 0.f;
                        const long long src = (r < valid && ci0 + c < Cin) ? map_row(p.in_map, row0 + r, V) : -1;
                        if (src >= 0) {
                            const long long off = src * Cin + ci0 + c;
                            v = p.in[off];
                            if (p.bn.p)
                                v = bn_back(v, p.in2[off], p.bn.p[ci0 + c], p.bn.m1[ci0 + c],
                                            p.bn.cc[ci0 + c], p.bn.mu[ci0 + c]);
                            if (p.in_out) p.in_out[(row0 + r) * Cin + ci0 + c] = v;
                        }
                        xs[i] = v;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&x_full[xb]);
            }
        }
        PROF_ADD(10);
#if ISTGCN_TC_PROF
        if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 64) { g_prof[9] += prof[9]; g_prof[10] += prof[10]; }
#endif
    } else {
        // =========================== aggregators: Xs -> A atoms (SWIZZLE_128B, K-major)
        // 8 warps x 4 lane groups = 32 destination slots: group q of warp aw owns joint
        // w = aw + 8*q; its 8 lanes hold 4 channels each (128-bit shared-memory accesses), so one
        // warp instruction moves four 128-byte rows.  The entry loop is software-pipelined and
        // the F frames give F independent FMA chains per lane.
        const int aw = warp - 8;
        const int q = lane >> 3, c4 = (lane & 7) * 4;
        const int w = aw + 8 * q;
        const bool active = w < V;
        const int fstride = V * 32;
        uint32_t xit = 0;
        PROF_DECL; PROF_T0();
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            for (int ch = 0; ch < nchunk; ++ch, ++xit) {
                const int xb = xit % kNX;
                PROF_ADD(14);
                mbar_wait(&x_full[xb], (xit / kNX) & 1);
                PROF_ADD(12);
                const float* xs = Xs + xb * kAtomRows * 32 + c4;
                for (int k = 0; k < K; ++k) {
                    const uint32_t hidx = xit * nh + (k >> 1);
                    PROF_ADD(14);
                    if (!(k & 1)) mbar_wait(&a_empty[hidx % kNH], ((hidx / kNH) & 1) ^ 1);
                    PROF_ADD(13);
                    float* A = As + ((hidx % kNH) * 2 + (k & 1)) * (kAtomBytes / 4);
                    if (active) {
                        const int beg = s_ptr[k * V + w], end = s_ptr[k * V + w + 1];
                        aggregate_joint_any(F, A, xs, s_ent, beg, end, fstride, V, w, c4);
                    }
                }
                PROF_ADD(14);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&a_full[xit % kNC]);
                    mbar_arrive(&x_empty[xb]);
                }
                PROF_ADD(15);
            }
        }
#if ISTGCN_TC_PROF
        if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 256) for (int i = 12; i <= 15; ++i) g_prof[i] += prof[i];
#endif
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (p.stat_sum) {
        for (int c = tid; c < NCOLS; c += kThreadsTC) {
            if (n0 + c < Cout) {
                atomicAdd(&p.stat_sum[n0 + c], s_sum[c]);
                atomicAdd(&p.stat_sumsq[n0 + c], s_sq[c]);
            }
        }
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * NCOLS);
    }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !sym) {
            set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
            return nullptr;
        }
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

int encode_tile_map(CUtensorMap* map, const float* base, long long rows, long long cols, int box_rows,
                    bool atom32) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return ISTGCN_E_ARCH;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for [%lld x %lld] box %d", (int)r, rows, cols,
                  box_rows);
        return ISTGCN_E_ARG;
    }
    return 0;
}

// Channels-last activation [NM][T][V][C] as a 4-D tensor (C, V, T, NM): box = 32 channels x V
// joints x `frames` frames taken every `t_stride`-th frame (SWIZZLE_128B; out-of-range frames
// read as zeros = the temporal padding).
int encode_frames_map(CUtensorMap* map, const float* base, int NM, int T, int V, int C, int frames,
                      int t_stride, bool atom32) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return ISTGCN_E_ARCH;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)V, (cuuint64_t)T, (cuuint64_t)NM};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)V * C * 4, (cuuint64_t)T * V * C * 4};
    cuuint32_t box[4] = {32u, (cuuint32_t)V, (cuuint32_t)((frames - 1) * t_stride + 1), 1u};
    cuuint32_t estr[4] = {1u, 1u, (cuuint32_t)t_stride, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for frames map [%d x %d x %d x %d]", (int)r, NM, T,
                  V, C);
        return ISTGCN_E_ARG;
    }
    return 0;
}

int encode_joint_frames_map(CUtensorMap* map, const float* base, long long frames, int V, int C,
                            int box_frames) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return ISTGCN_E_ARCH;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)V, (cuuint64_t)frames};
    cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)V * C * 4};
    cuuint32_t box[3] = {32u, 1u, (cuuint32_t)box_frames};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for joint-frames map [%lld x %d x %d]", (int)r, frames, V, C);
        return ISTGCN_E_ARG;
    }
    return 0;
}

template <int NCOLS>
static int launch_tc(const CUtensorMap& map, const CUtensorMap& omap, const CUtensorMap& omap_last,
                     const GcnTcParams& p, cudaStream_t s) {
    using L = SmemLayout<NCOLS>;
    auto kern = gcn_tc_kernel<NCOLS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
    const int ny = (p.Cout + NCOLS - 1) / NCOLS;
    int nx = num_sms() / ny;
    if (nx < 1) nx = 1;
    if (nx > p.tiles) nx = p.tiles;
    kern<<<dim3(nx, ny), kThreadsTC, L::total, s>>>(map, omap, omap_last, p);
    return finish_launch("gcn_tc");
}

}  // namespace tc
}  // namespace istgcn

using namespace istgcn;

#if ISTGCN_TC_PROF
extern "C" int istgcn_debug_prof(unsigned long long* out) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, tc::g_prof, sizeof(unsigned long long) * 32);
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(tc::g_prof, z, sizeof(z));
    return 0;
}
#endif

// Tensor-core graph convolution, forward or input-gradient form (see the file header).
//   w_rows[K*Cout][CinPad]: weight with rows k*Cout + n and the 32-padded input channels as
//   columns (forward: the conv weight (K*Cout, Cin) itself, zero-padded when Cin < 32).
ISTGCN_API int istgcn_gcn_tc(const float* in, const float* in2, const float* bn_p, const float* bn_m1,
                             const float* bn_c, const float* bn_mu, const float* w_rows,
                             const float* vals, const int* lptr, const int* lsrc, const int* lid,
                             int nnz, const float* bias_k, const float* colsum, const float* add_rows,
                             float* out,
                             float* in_out, double* stat_sum, double* stat_sumsq, int frames, int V,
                             int K, int Cin, int CinPad, int Cout, int t_in, int t_out, int t_stride, int t_offset,
                             int map_side, istgcn_stream_t s) {
    ISTGCN_REQUIRE(in && w_rows && vals && lptr && lsrc && lid && out, ISTGCN_E_ARG,
                   "gcn_tc: null pointer");
    ISTGCN_REQUIRE(bn_p == nullptr || (in2 && bn_m1 && bn_c && bn_mu), ISTGCN_E_ARG,
                   "gcn_tc: bn_p needs in2, bn_m1, bn_c and bn_mu");
    ISTGCN_REQUIRE((stat_sum == nullptr) == (stat_sumsq == nullptr), ISTGCN_E_ARG,
                   "gcn_tc: pass both statistics buffers or neither");
    ISTGCN_REQUIRE(V >= 1 && V <= 32 && K >= 1 && K <= 4, ISTGCN_E_SHAPE, "gcn_tc: V=%d K=%d unsupported", V, K);
    ISTGCN_REQUIRE(CinPad % 32 == 0 && CinPad >= Cin && Cin >= 1, ISTGCN_E_SHAPE,
                   "gcn_tc: CinPad=%d must be a multiple of 32 and >= Cin=%d", CinPad, Cin);
    ISTGCN_REQUIRE(Cout >= 1 && Cout <= 256 * 4, ISTGCN_E_SHAPE, "gcn_tc: Cout=%d unsupported", Cout);
    ISTGCN_REQUIRE(nnz >= 0 && nnz <= kMaxNnz, ISTGCN_E_SHAPE, "gcn_tc: nnz=%d exceeds %d", nnz, kMaxNnz);
    ISTGCN_REQUIRE((reinterpret_cast<uintptr_t>(w_rows) & 15) == 0, ISTGCN_E_ARG,
                   "gcn_tc: weight pointer must be 16-byte aligned");
    if (frames == 0) return 0;
    ISTGCN_REQUIRE((bias_k == nullptr) == (colsum == nullptr), ISTGCN_E_ARG,
                   "gcn_tc: pass bias_k and colsum together");
    // second-generation engine (aggregation on the tensor core, gcn_tc2.cu) whenever the input can
    // be fed by TMA as it lies in HBM; ISTGCN_GCN_TC_V1=1 keeps the first-generation kernel
    static const bool force_v1 = getenv("ISTGCN_GCN_TC_V1") != nullptr;
    if (!force_v1 && bn_p == nullptr && in_out == nullptr && map_side == 0 && CinPad == Cin &&
        (add_rows == nullptr || (add_rows == out && stat_sum == nullptr)) &&
        tc::gcn_tc2_eligible(V, K, Cin, Cout, in, out)) {
        // ISTGCN_GCN_TC_V2=1 keeps the lane-mask form of the aggregation (gcn_tc2.cu)
        static const bool force_v2 = getenv("ISTGCN_GCN_TC_V2") != nullptr;
        auto launch = force_v2 ? tc::launch_gcn_tc2 : tc::launch_gcn_tc3;
        return launch(in, w_rows, vals, lptr, lsrc, lid, bias_k, colsum, out, add_rows != nullptr, stat_sum,
                      stat_sumsq, frames, V, K, Cin, Cout, (cudaStream_t)s);
    }
    tc::GcnTcParams p{in, in2, {bn_p, bn_m1, bn_c, bn_mu}, vals, lptr, lsrc, lid, bias_k, colsum,
                      add_rows, out, in_out, stat_sum, stat_sumsq, frames, V, K, Cin, CinPad, Cout,
                      nnz, 0, {0, 0, 1, 0}, {0, 0, 1, 0}, 0};
    if (map_side == 1) p.in_map = {t_in, t_out, t_stride, t_offset};
    if (map_side == 2) p.out_map = {t_in, t_out, t_stride, t_offset};
    const int F = kTileRows / V > 8 ? 8 : kTileRows / V;
    p.tiles = (frames + F - 1) / F;
    const int ncols = Cout > 128 ? 256 : (Cout > 64 ? 128 : 64);
    CUtensorMap map;
    if (int e = tc::encode_tile_map(&map, w_rows, (long long)K * Cout, CinPad, ncols)) return e;
    // whole-tile outputs leave through TMA (plain store, or reduce-add when the kernel accumulates
    // in place); strided output maps, partial tiles and odd widths use the direct path
    CUtensorMap omap = map, omap_last = map;
    const int last_rows = F * V - 96;
    if (Cout % 32 == 0 && map_side != 2 && last_rows >= 1 && (add_rows == nullptr || add_rows == out) &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        if (int e = tc::encode_tile_map(&omap, out, (long long)frames * V, Cout, 32)) return e;
        if (int e = tc::encode_tile_map(&omap_last, out, (long long)frames * V, Cout, last_rows)) return e;
        // statistics must see the accumulated value: the reduce-add path never has it in registers
        p.tma_out = add_rows ? (stat_sum ? 0 : 2) : 1;
    }
    cudaStream_t st = (cudaStream_t)s;
    if (ncols == 256) return tc::launch_tc<256>(map, omap, omap_last, p, st);
    if (ncols == 128) return tc::launch_tc<128>(map, omap, omap_last, p, st);
    return tc::launch_tc<64>(map, omap, omap_last, p, st);
}
