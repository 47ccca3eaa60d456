// Second-generation tcgen05 engine of the fused graph convolution (reference:
// net/utils/tgcn.py:76-89, net/utils/inceptionv2_gcn.py:64-89): BOTH contractions run on the
// tensor core and neither the adjacency nor the aggregated operand ever touches shared memory.
//
//     X'_k[(f,w)][ci] = sum_v A_eff[k][v][w] * IN[(f,v)][ci]            (MMA 1, per frame)
//     OUT[(f,w)][n]   = sum_k sum_ci X'_k[(f,w)][ci] * W[k][n][ci]      (MMA 2)
//
// A frame tile = 4 frames, each padded to a 32-row slot (V <= 32), so that frame f owns TMEM
// lanes 32f..32f+31 (one lane quadrant) and one epilogue warp.
//
//   MMA 1  D1_k[quadrant f][32 ci] = ADJ_k[128][32 v] * X_f[32 v][32 ci]   kind::tf32, M128 N32 K8 x4
//          A operand FROM TENSOR MEMORY: A_eff[k]^T (32 x 32, TF32-rounded) written once per CTA
//          into columns [32k, 32k+32) of all four lane quadrants; the instruction's
//          disable-output-lane mask keeps only quadrant f, so frame f's product lands in frame
//          f's lanes.  (An SS-mode MMA would re-read a 4 KB adjacency operand from shared memory
//          per instruction: measured 80 cycles per M128 N32 K8 instruction, 5x the math.)
//          B operand: the 32-channel slice of frame f exactly as it lies in HBM (rows = joints,
//          channels contiguous), dropped into its slot by TMA with the 32-byte-atom 128B
//          swizzle = the MN-major TF32 operand layout.  No CUDA-core staging at all.
//   MMA 2  D2[128][Cout] += D1_k[128][32 ci] (A operand read from tensor memory) * W_k[Cout][32 ci]
//          (weights by TMA, SWIZZLE_128B K-major; one barrier per stage of up to 4 partitions).
//   tcgen05.mma executes in issue order, so the D1 -> A-operand hand-over needs no barrier.
//
//   warp 0  TMA producer (weights)        warp 2  TMA producer (input frames)
//   warp 1  MMA issuer                    warps 4-7  epilogue (one frame each): tcgen05.ld, bias
//                                          term, BatchNorm sums (double), swizzled staging tile,
//                                          TMA store / reduce-add of [V rows][32 channels]
//   Single-thread roles run warp-convergent and elect the issuing lane per instruction group:
//   under `if (lane == 0)` the compiler wraps every uniform-datapath instruction (UTMALDG,
//   UTCHMMA) in an ELECT / BRA.U.ANY loop that costs ~60 cycles per instruction.
#include <stdlib.h>

#include "tc_common.cuh"

namespace istgcn {
namespace tc {

constexpr int kThreads2 = 256;
constexpr int kSlotRows = 32;                          // padded rows of one frame
constexpr int kFr2 = 4;                                // frames per tile
constexpr int kXStage = kFr2 * kSlotRows * 128;        // one 32-channel slice of a tile
constexpr int kAdjCol = 0;                             // TMEM: adjacency [0, 128)
constexpr int kD1Col = 128;                            //       D1 [128, 128 + 128*ND1)

template <int NCOLS>
struct Cfg2 {
    static constexpr int WU = NCOLS > 128 ? 128 : NCOLS;          // weight rows per TMA box
    static constexpr int NU = NCOLS / WU;                         // boxes per partition
    static constexpr int KG = NCOLS > 128 ? 2 : 4;                // partitions per weight stage
    static constexpr int WBYTES = WU * 128;
    static constexpr int WSTAGE = KG * NU * WBYTES;
    static constexpr int NW = NCOLS == 64 ? 3 : 2;
    static constexpr int NX = NCOLS == 64 ? 6 : 4;
    static constexpr int ND1 = NCOLS == 64 ? 2 : 1;
    static constexpr int ND2 = NCOLS == 256 ? 1 : 2;
    static constexpr int kD2Col = 512 - ND2 * NCOLS;
    static_assert(kD1Col + ND1 * 128 <= kD2Col, "tensor-memory budget exceeded");
    static constexpr int x_off = 0;
    static constexpr int w_off = x_off + NX * kXStage;
    static constexpr int stage_off = w_off + NW * WSTAGE;
    static constexpr int bias_off = stage_off + 4 * 4096;
    static constexpr int stat_off = bias_off + 4 * NCOLS * 4;
    static constexpr int bar_off = stat_off + 2 * NCOLS * 8;
    static constexpr int kNumBars = 2 * NX + 2 * NW + 4;
    static constexpr int total = bar_off + kNumBars * 8 + 16;
    static_assert(total <= 232448, "shared-memory budget exceeded");
    static_assert(NX * kXStage >= 4 * 32 * 33 * 4, "adjacency scratch lives in the input ring");
};

struct GcnTc2Params {
    const float* vals;
    const int *lptr, *lsrc, *lid;
    const float *bias_k, *colsum;
    double *stat_sum, *stat_sumsq;
    int frames, V, K, Cin, Cout, tiles, reduce;
};

template <int NCOLS>
__global__ void __launch_bounds__(kThreads2, 1)
gcn_tc2_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap wmap,
               const __grid_constant__ CUtensorMap omap, GcnTc2Params p) {
    using L = Cfg2<NCOLS>;
    constexpr int NX = L::NX, NW = L::NW, ND1 = L::ND1, ND2 = L::ND2, WU = L::WU, NU = L::NU, KG = L::KG;
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
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::kNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = p.V, K = p.K, Cout = p.Cout;
    const int nchunk = p.Cin / 32;
    const int ngrp = (K + KG - 1) / KG;                 // weight stages per slice
    const int my_tiles = p.tiles > (int)blockIdx.x
                             ? (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    // ---- one-time setup: barriers, TMEM, bias factors; the dense adjacency goes through a
    // scratch table [k][w][33] in the (still idle) input ring into tensor memory
    float* adjT = reinterpret_cast<float*>(Xs);
    for (int i = tid; i < 4 * 32 * 33; i += kThreads2) adjT[i] = 0.f;
    for (int i = tid; i < 2 * NCOLS; i += kThreads2) s_sum[i] = 0.0;
    if (p.bias_k)
        for (int i = tid; i < K * NCOLS; i += kThreads2) {
            const int k = i / NCOLS, c = i % NCOLS;
            s_bias[i] = c < Cout ? p.bias_k[(size_t)k * Cout + c] : 0.f;
        }
    if (tid == 0) {
        for (int i = 0; i < NX; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < NW; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
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
    for (int i = tid; i < K * V; i += kThreads2) {          // one thread per (k, w): no races
        const int k = i / V, w = i - k * V;
        for (int j = p.lptr[i]; j < p.lptr[i + 1]; ++j)
            adjT[(k * 32 + w) * 33 + p.lsrc[j]] += __uint_as_float(to_tf32(p.vals[p.lid[j]]));
    }
    __syncthreads();
    if (warp >= 4) {                                        // quadrant warp % 4, lane = joint w
        for (int k = 0; k < 4; ++k) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = k < K ? adjT[(k * 32 + lane) * 33 + j] : 0.f;
            tmem_st32(tmem_base + (static_cast<uint32_t>((warp - 4) * 32) << 16) + kAdjCol + k * 32, v);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // pad rows of the frame slots must be finite: zero the whole ring once (TMA never writes them)
    for (int i = tid; i < NX * kXStage / 16; i += kThreads2)
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
                    const int kn = min(KG, K - g * KG);
                    mbar_wait(&b_empty[sb], ((it / NW) & 1) ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&b_full[sb], kn * NU * L::WBYTES);
                        for (int kk = 0; kk < kn; ++kk)
#pragma unroll
                            for (int u = 0; u < NU; ++u)
                                tma_load_2d(Ws + sb * L::WSTAGE + (kk * NU + u) * L::WBYTES, &wmap,
                                            &b_full[sb], ch * 32, (g * KG + kk) * Cout + u * WU);
                    }
                    __syncwarp();
                }
    } else if (warp == 2) {
        // =========================== TMA producer: the frames of the tile, one slot each
        uint32_t it = 0;
        const uint32_t bytes = kFr2 * V * 128;
        for (int t = 0; t < my_tiles; ++t) {
            const int f0 = (blockIdx.x + t * gridDim.x) * kFr2;
            for (int ch = 0; ch < nchunk; ++ch, ++it) {
                const int xs = it % NX;
                mbar_wait(&x_empty[xs], ((it / NX) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&x_full[xs], bytes);
#pragma unroll
                    for (int f = 0; f < kFr2; ++f)
                        tma_load_3d(Xs + xs * kXStage + f * (kSlotRows * 128), &xmap, &x_full[xs],
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
            const uint32_t adj = tmem_base + kAdjCol;
            auto issue1 = [&](uint32_t s) {
                const int xs = s % NX;
                mbar_wait(&x_full[xs], (s / NX) & 1);
                tc_fence_after();
                const uint32_t d1 = tmem_base + kD1Col + (s % ND1) * 128;
                if (elect_one()) {
                    // issue order partition -> frame -> K-step: 16 consecutive instructions target
                    // the same D1 columns (the frames differ only in the lane mask).  Measured: the
                    // pipe is fastest on long same-accumulator runs; rotating over 16 accumulators
                    // (K-step outermost) is 25 % slower.
                    for (int k = 0; k < K; ++k)
#pragma unroll
                        for (int f = 0; f < kFr2; ++f) {
                            const uint32_t b_addr = xs0 + xs * kXStage + f * (kSlotRows * 128);
                            const uint32_t m0 = f == 0 ? 0u : ~0u, m1 = f == 1 ? 0u : ~0u,
                                           m2 = f == 2 ? 0u : ~0u, m3 = f == 3 ? 0u : ~0u;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                tc_mma_tf32_ts_masked(d1 + k * 32, adj + k * 32 + ks * 8,
                                                      make_desc(b_addr + ks * 1024, 4096, 512, 1), idesc1,
                                                      ks ? 1u : 0u, m0, m1, m2, m3);
                        }
                    tc_commit(&x_empty[xs]);
                }
                __syncwarp();
            };
            uint32_t it = 0;
            issue1(0);
            for (uint32_t s = 0; s < total; ++s) {
                if (ND1 == 2 && s + 1 < total) issue1(s + 1);
                const uint32_t t = s / nchunk, ch = s - t * nchunk;
                const uint32_t buf = t % ND2, use = t / ND2;
                if (ch == 0) {
                    mbar_wait(&t_empty[buf], (use & 1) ^ 1);
                    tc_fence_after();
                }
                const uint32_t d2 = tmem_base + L::kD2Col + buf * NCOLS;
                const uint32_t d1 = tmem_base + kD1Col + (s % ND1) * 128;
                for (int g = 0; g < ngrp; ++g, ++it) {
                    const int sb = it % NW;
                    const int kn = min(KG, K - g * KG);
                    mbar_wait(&b_full[sb], (it / NW) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        for (int kk = 0; kk < kn; ++kk)
#pragma unroll
                            for (int u = 0; u < NU; ++u) {
                                const uint32_t b_addr = ws0 + sb * L::WSTAGE + (kk * NU + u) * L::WBYTES;
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    tc_mma_tf32_ts(d2 + u * WU, d1 + (g * KG + kk) * 32 + ks * 8,
                                                   make_desc(b_addr + ks * 32, 16, 1024), idesc2,
                                                   (ch | g | kk | ks) ? 1u : 0u);
                            }
                        tc_commit(&b_empty[sb]);
                        if (ch == (uint32_t)nchunk - 1 && g == ngrp - 1) tc_commit(&t_full[buf]);
                    }
                    __syncwarp();
                }
                if (ND1 == 1 && s + 1 < total) issue1(s + 1);
            }
        }
    } else if (warp >= 4) {
        // =========================== epilogue: warp ew owns frame ew of the tile (TMEM lanes 32*ew..)
        const int ew = warp - 4;
        const int w = lane;
        uint8_t* stage = smem + L::stage_off + ew * 4096;      // [32 rows][128 B], SWIZZLE_128B
        double acc_s[NCOLS / 32], acc_q[NCOLS / 32];          // this warp's column sums (lane = column)
#pragma unroll
        for (int i = 0; i < NCOLS / 32; ++i) acc_s[i] = acc_q[i] = 0.0;
        float cs[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.bias_k && w < V)
            for (int k = 0; k < K; ++k) cs[k] = p.colsum[k * V + w];
        for (int t = 0; t < my_tiles; ++t) {
            const uint32_t buf = t % ND2, use = t / ND2;
            const int frame = (blockIdx.x + t * gridDim.x) * kFr2 + ew;
            const bool fok = frame < p.frames;
            mbar_wait(&t_full[buf], use & 1);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < NCOLS; c0 += 32) {
                if (c0 >= Cout) break;
                float v[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + L::kD2Col + buf * NCOLS + c0, v);
                if (p.bias_k) {
                    for (int k = 0; k < K; ++k) {
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
                if (lane == 0) bulk_wait_read();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0 && fok) {
                    if (p.reduce) tma_reduce_add_2d(stage, &omap, c0, frame * V);
                    else tma_store_2d(stage, &omap, c0, frame * V);
                    bulk_commit();
                }
                if (p.stat_sum && fok) {
                    // column sums straight from the staging tile: lane c adds column c of the V
                    // valid rows (one conflict-free wavefront per row; the TMA store reads the
                    // tile concurrently), then joins the warp's double accumulators
                    float a = 0.f, q = 0.f;
                    const int cj = lane >> 2, ce = (lane & 3) * 4;
                    for (int r = 0; r < V; ++r) {
                        const float x = *reinterpret_cast<const float*>(stage + r * 128 + ((cj ^ (r & 7)) << 4) + ce);
                        a += x;
                        q = fmaf(x, x, q);
                    }
                    acc_s[c0 / 32] += (double)a;
                    acc_q[c0 / 32] += (double)q;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[buf]);
        }
        if (lane == 0) bulk_wait_all();
        if (p.stat_sum) {
#pragma unroll
            for (int i = 0; i < NCOLS / 32; ++i) {
                atomicAdd(&s_sum[i * 32 + lane], acc_s[i]);
                atomicAdd(&s_sq[i * 32 + lane], acc_q[i]);
            }
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (p.stat_sum) {
        for (int c = tid; c < NCOLS; c += kThreads2) {
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

// 3-D (C, V, frames) view of a channels-last activation: box = 32 channels x V joints x 1 frame,
// 32-byte-atom 128B swizzle (the MN-major TF32 operand layout)
int encode_frame_slices(CUtensorMap* map, const float* base, long long frames, int V, int C) {
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Fn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !sym) {
            set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
            return ISTGCN_E_ARCH;
        }
        fn = reinterpret_cast<Fn>(sym);
    }
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)V, (cuuint64_t)frames};
    cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)V * C * 4};
    cuuint32_t box[3] = {32u, (cuuint32_t)V, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for frame slices [%lld x %d x %d]", (int)r, frames, V, C);
        return ISTGCN_E_ARG;
    }
    return 0;
}

template <int NCOLS>
static int launch_tc2_n(const CUtensorMap& xmap, const CUtensorMap& wmap, const CUtensorMap& omap,
                        const GcnTc2Params& p, cudaStream_t s) {
    using L = Cfg2<NCOLS>;
    auto kern = gcn_tc2_kernel<NCOLS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
    int nx = num_sms();
    if (nx > p.tiles) nx = p.tiles;
    kern<<<nx, kThreads2, L::total, s>>>(xmap, wmap, omap, p);
    return finish_launch("gcn_tc2");
}

// Shapes the second-generation engine takes: plain frame maps, 32-multiple channel counts (the
// input slices and output tiles move by TMA), Cout <= 256, V <= 32.
bool gcn_tc2_eligible(int V, int K, int Cin, int Cout, const float* in, const float* out) {
    return V >= 1 && V <= 32 && K >= 1 && K <= 4 && Cin % 32 == 0 && Cin >= 32 && Cout % 32 == 0 &&
           Cout >= 32 && Cout <= 256 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(out) & 15) == 0;
}

int launch_gcn_tc2(const float* in, const float* w_rows, const float* vals, const int* lptr,
                   const int* lsrc, const int* lid, const float* bias_k, const float* colsum, float* out,
                   int reduce, double* stat_sum, double* stat_sumsq, int frames, int V, int K, int Cin,
                   int Cout, cudaStream_t st) {
    GcnTc2Params p{vals, lptr, lsrc, lid, bias_k, colsum, stat_sum, stat_sumsq, frames, V, K, Cin, Cout,
                   (frames + kFr2 - 1) / kFr2, reduce};
    const int ncols = Cout > 128 ? 256 : (Cout > 64 ? 128 : 64);
    CUtensorMap xmap, wmap, omap;
    if (int e = encode_frame_slices(&xmap, in, frames, V, Cin)) return e;
    if (int e = encode_tile_map(&wmap, w_rows, (long long)K * Cout, Cin, ncols > 128 ? 128 : ncols)) return e;
    if (int e = encode_tile_map(&omap, out, (long long)frames * V, Cout, V)) return e;
    if (ncols == 256) return launch_tc2_n<256>(xmap, wmap, omap, p, st);
    if (ncols == 128) return launch_tc2_n<128>(xmap, wmap, omap, p, st);
    return launch_tc2_n<64>(xmap, wmap, omap, p, st);
}

}  // namespace tc
}  // namespace istgcn
