// tcgen05 implicit-GEMM kernel for the FULL-WIDTH temporal convolution of the baseline ST-GCN
// family (reference: net/st_gcnold.py:165-171 Conv2d(C, C, (kt,1), (stride,1), (pad,0));
// net/st_gcn_mstcn.py:189-209 merged into one 15-tap kernel):
//
//     OUT[(n,to,v)][co] = sum_tap sum_ci IN[(n, to*stride + dir*(tap - pad), v)][ci] * W[tap][co][ci]
//
// dir = +1 is the forward convolution; dir = -1 with the transposed weights is the input gradient.
// For stride 2 the input gradient splits by the parity q of the input frame ti = 2j + q: only the
// taps with q + pad - tap even reach it, from output frame j + (q + pad - tap)/2 -- two launches
// of the same kernel with a tap table, whose tiles leave through a frame-strided 4-D TMA store.
// Both operands come straight from TMA: the activation is a 4-D tensor (C, V, T, NM), a tile is F = floor(128/V) whole frames of one clip, and tap `tap` of that tile
// is simply the same box shifted by `tap - pad` frames -- frames outside [0, T) are out of bounds
// for the tensor map and read as zeros, which IS the temporal zero padding.  A temporal stride is
// the tensor map's element stride in the T dimension.
//
//   warp 0      TMA producer: per stage one activation atom [F*V rows x 32 ci] (SWIZZLE_128B,
//               K-major) and one weight atom [NCOLS x 32 ci]
//   warp 1      MMA issuer: tcgen05.mma kind::tf32, M=128, N=NCOLS, K=8; kt * Cin/32 stages per tile,
//               accumulators in TMEM, double buffered across tiles
//   warps 4-7   epilogue: tcgen05.ld, + bias, BatchNorm statistics (double), swizzled staging
//               tile -> TMA tile store
#include "tc_common.cuh"

namespace istgcn {
namespace tc {

constexpr int kThreadsTV = 256;

struct TconvParams {
    const float* bias;            // [Cout] or NULL
    float* out;                   // [NM*Tout*V][Cout]
    double *stat_sum, *stat_sumsq;
    int NM, V, Cin, Cout, tiles, tiles_per_clip;
    int in_step;                  // input frame of a tile's first output frame = to0 * in_step + tap_off
    int ntaps, tap_w[16], tap_off[16];   // weight block and frame offset of every contributing tap
    int out4d, out_q, out_step;   // 4-D strided store: output frame = out_q + out_step * to
    int frames_out;               // output frames per clip handled by this launch
};

template <int NCOLS>
struct SmemTV {
    static constexpr int kStages = NCOLS == 64 ? 8 : (NCOLS == 128 ? 6 : 4);
    static constexpr int kBAtomBytes = NCOLS * 128;
    static constexpr int stage_bytes = kAtomBytes + kBAtomBytes;
    static constexpr int ring_off = 0;
    static constexpr int out_off = ring_off + kStages * stage_bytes;        // 4 x [32 rows][128 B]
    static constexpr int bias_off = out_off + 4 * 4096;
    static constexpr int stat_off = bias_off + NCOLS * 4;
    static constexpr int bar_off = stat_off + 2 * NCOLS * 8;
    static constexpr int kNumBars = 2 * kStages + 4;
    static constexpr int total = bar_off + kNumBars * 8 + 16;
    static_assert(total <= 232448, "shared-memory budget exceeded");
};

template <int NCOLS>
__global__ void __launch_bounds__(kThreadsTV, 1)
tconv_tc_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap,
                const __grid_constant__ CUtensorMap omap, const __grid_constant__ CUtensorMap omap_last,
                const __grid_constant__ CUtensorMap omap4, TconvParams p) {
    using L = SmemTV<NCOLS>;
    constexpr int kStages = L::kStages;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* ring = smem + L::ring_off;
    float* s_bias = reinterpret_cast<float*>(smem + L::bias_off);
    double* s_sum = reinterpret_cast<double*>(smem + L::stat_off);
    double* s_sq = s_sum + NCOLS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bar_off);
    uint64_t* full = bars;
    uint64_t* empty = full + kStages;
    uint64_t* t_full = empty + kStages;
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::kNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = p.V, Cout = p.Cout;
    const int F = (kAtomRows / V) > 8 ? 8 : (kAtomRows / V);
    const int nchunk = p.Cin / 32;
    const int n0 = blockIdx.y * NCOLS;
    const uint32_t a_bytes = (uint32_t)(F * V) * 128u;

    // rows F*V .. 127 of every activation atom are never written by TMA: keep them zero
    for (int i = tid; i < kStages * L::stage_bytes / 4; i += kThreadsTV)
        reinterpret_cast<float*>(ring)[i] = 0.f;
    for (int i = tid; i < 2 * NCOLS; i += kThreadsTV) s_sum[i] = 0.0;
    for (int i = tid; i < NCOLS; i += kThreadsTV)
        s_bias[i] = (p.bias && n0 + i < Cout) ? p.bias[n0 + i] : 0.f;
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&amap); tma_prefetch_desc(&wmap);
        tma_prefetch_desc(&omap); tma_prefetch_desc(&omap_last); tma_prefetch_desc(&omap4);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * NCOLS);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // TMA producer (warp-convergent loop, elected issue: see gcn_tc2.cu)
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int n = tile / p.tiles_per_clip;
            const int to0 = (tile - n * p.tiles_per_clip) * F;
            for (int tap = 0; tap < p.ntaps; ++tap) {
                const int t_first = to0 * p.in_step + p.tap_off[tap];
                for (int ch = 0; ch < nchunk; ++ch, ++it) {
                    const int s = it % kStages;
                    mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
                    if (elect_one()) {
                        uint8_t* dst = ring + s * L::stage_bytes;
                        mbar_arrive_expect_tx(&full[s], a_bytes + L::kBAtomBytes);
                        tma_load_4d(dst, &amap, &full[s], ch * 32, 0, t_first, n);
                        tma_load_2d(dst + kAtomBytes, &wmap, &full[s], ch * 32, p.tap_w[tap] * Cout + n0);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer
        constexpr uint32_t idesc = make_idesc(128, NCOLS, false, false);
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tcount) {
            const int buf = tcount & 1;
            mbar_wait(&t_empty[buf], ((tcount >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * NCOLS;
            const int nst = p.ntaps * nchunk;
            for (int st = 0; st < nst; ++st, ++it) {
                const int s = it % kStages;
                mbar_wait(&full[s], (it / kStages) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_u32(ring + s * L::stage_bytes);
                    const uint32_t b_addr = a_addr + kAtomBytes;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        tc_mma_tf32(d_tmem, make_desc(a_addr + ks * 32, 16, 1024),
                                    make_desc(b_addr + ks * 32, 16, 1024), idesc, (st | ks) ? 1u : 0u);
                    tc_commit(&empty[s]);
                    if (st == nst - 1) tc_commit(&t_full[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int r = ew * 32 + lane;
        uint8_t* stage = smem + L::out_off + ew * 4096;
        const CUtensorMap* om = ew == 3 ? &omap_last : &omap;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tcount) {
            const int buf = tcount & 1;
            const long long row0 = (long long)tile * F * V;      // tiles never straddle clips
            const int cn = tile / p.tiles_per_clip;
            const int cto = (tile - cn * p.tiles_per_clip) * F;
            const int ct = p.out_q + p.out_step * cto;
            const bool ok = r < min(F, p.frames_out - cto) * V;     // ragged clip end: bias-only rows
            mbar_wait(&t_full[buf], (tcount >> 1) & 1);
            tc_fence_after();
            for (int c0 = 0; c0 < NCOLS; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * NCOLS + c0, v);
                const int cg = n0 + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + j);
                    v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
                }
                if (p.out4d) {
                    // the four warps' staging areas are one [128 rows][128 B] SWIZZLE_128B tile:
                    // one frame-strided 4-D store per 32 columns, issued by warp 4
                    if (ew == 0 && lane == 0) bulk_wait_read();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                } else {
                    if (lane == 0) bulk_wait_read();
                    __syncwarp();
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                fence_proxy_async();
                if (p.out4d) {
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (ew == 0 && lane == 0 && cg < Cout) {
                        tma_store_4d(smem + L::out_off, &omap4, cg, 0, ct, cn);
                        bulk_commit();
                    }
                } else {
                    __syncwarp();
                    if (lane == 0 && cg < Cout) {
                        tma_store_2d(stage, om, cg, (int)(row0 + ew * 32));
                        bulk_commit();
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
    }

    tc_fence_before();
    __syncthreads();
    if (p.stat_sum) {
        for (int c = tid; c < NCOLS; c += kThreadsTV) {
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

template <int NCOLS>
static int launch_tv(const CUtensorMap& amap, const CUtensorMap& wmap, const CUtensorMap& omap,
                     const CUtensorMap& omap_last, const CUtensorMap& omap4, const TconvParams& p,
                     cudaStream_t s) {
    using L = SmemTV<NCOLS>;
    auto kern = tconv_tc_kernel<NCOLS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
    const int ny = (p.Cout + NCOLS - 1) / NCOLS;
    int nx = num_sms() / ny;
    if (nx < 1) nx = 1;
    if (nx > p.tiles) nx = p.tiles;
    kern<<<dim3(nx, ny), kThreadsTV, L::total, s>>>(amap, wmap, omap, omap_last, omap4, p);
    return finish_launch("tconv_tc");
}

}  // namespace tc
}  // namespace istgcn

using namespace istgcn;

// Full-width temporal convolution on the tcgen05 engine (see the file header).
//   in     [NM][T][V][Cin]      channels-last activation (dir=+1) or output gradient (dir=-1)
//   w_rows [kt*Cout][Cin]       per tap: rows = output channel, columns = input channel
//   out    [NM][Tout][V][Cout]
// Requirements: Cin % 32 == 0, Cout % 32 == 0, floor(128/V)*V > 96.  Clips whose output frame
// count is not a multiple of floor(128/V) leave through 4-D stores that clip the ragged end.
// dir = -1: `in` is the output gradient [NM][Tout][V][Cin], `out` the input gradient
// [NM][T][V][Cout], w_rows the transposed weights [kt*Cout][Cin] (rows = conv input channel).
ISTGCN_API int istgcn_tconv_tc(const float* in, const float* w_rows, const float* bias, float* out,
                               double* stat_sum, double* stat_sumsq, int NM, int T, int Tout, int V,
                               int Cin, int Cout, int kt, int stride, int dir, istgcn_stream_t s) {
    ISTGCN_REQUIRE(in && w_rows && out, ISTGCN_E_ARG, "tconv_tc: null pointer");
    ISTGCN_REQUIRE((stat_sum == nullptr) == (stat_sumsq == nullptr), ISTGCN_E_ARG,
                   "tconv_tc: pass both statistics buffers or neither");
    ISTGCN_REQUIRE(V >= 1 && V <= 32 && kt >= 1 && (kt & 1) && kt <= 31, ISTGCN_E_SHAPE,
                   "tconv_tc: V=%d kt=%d unsupported", V, kt);
    ISTGCN_REQUIRE(Cin % 32 == 0 && Cout % 32 == 0 && Cin >= 32 && Cout >= 32 && Cout <= 1024,
                   ISTGCN_E_SHAPE, "tconv_tc: Cin=%d Cout=%d must be multiples of 32", Cin, Cout);
    ISTGCN_REQUIRE(dir == 1 || dir == -1, ISTGCN_E_ARG, "tconv_tc: dir=%d unsupported", dir);
    ISTGCN_REQUIRE(stride >= 1 && Tout == (T - 1) / stride + 1, ISTGCN_E_SHAPE,
                   "tconv_tc: Tout=%d does not match T=%d stride=%d", Tout, T, stride);
    ISTGCN_REQUIRE(dir == 1 || stride <= 2, ISTGCN_E_SHAPE, "tconv_tc: dir=-1 needs stride 1 or 2");
    const int F = kTileRows / V > 8 ? 8 : kTileRows / V;
    ISTGCN_REQUIRE(F * V > 96, ISTGCN_E_SHAPE, "tconv_tc: V=%d unsupported", V);
    ISTGCN_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(w_rows) |
                     reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                   ISTGCN_E_ARG, "tconv_tc: pointers must be 16-byte aligned");
    if ((long long)NM * Tout == 0) return 0;
    const int pad = (kt - 1) / 2;
    const int ncols = Cout > 128 ? 256 : (Cout > 64 ? 128 : 64);
    cudaStream_t st = (cudaStream_t)s;
    // frames on the input side / output side of this launch
    const int T_in = dir == 1 ? T : Tout, T_out = dir == 1 ? Tout : T;
    CUtensorMap amap, wmap, omap, omap_last, omap4;
    if (int e = tc::encode_tile_map(&wmap, w_rows, (long long)kt * Cout, Cin, ncols)) return e;
    const long long rows = (long long)NM * T_out * V;
    if (int e = tc::encode_tile_map(&omap, out, rows, Cout, 32)) return e;
    if (int e = tc::encode_tile_map(&omap_last, out, rows, Cout, F * V - 96)) return e;
    const bool strided_dx = dir == -1 && stride == 2;
    if (int e = tc::encode_frames_map(&amap, in, NM, T_in, V, Cin, F, dir == 1 ? stride : 1)) return e;
    if (int e = tc::encode_frames_map(&omap4, out, NM, T_out, V, Cout, F, strided_dx ? 2 : 1)) return e;
    for (int q = 0; q < (strided_dx ? 2 : 1); ++q) {
        tc::TconvParams p{};
        p.bias = bias; p.out = out; p.stat_sum = stat_sum; p.stat_sumsq = stat_sumsq;
        p.NM = NM; p.V = V; p.Cin = Cin; p.Cout = Cout;
        int frames_out = T_out;                       // output frames handled by this launch
        if (dir == 1) {
            p.in_step = stride;
            for (int tap = 0; tap < kt; ++tap) { p.tap_w[p.ntaps] = tap; p.tap_off[p.ntaps++] = tap - pad; }
        } else if (!strided_dx) {
            p.in_step = 1;
            for (int tap = 0; tap < kt; ++tap) { p.tap_w[p.ntaps] = tap; p.tap_off[p.ntaps++] = pad - tap; }
        } else {                                      // input frames ti = 2j + q
            p.in_step = 1;
            for (int tap = 0; tap < kt; ++tap)
                if (((q + pad - tap) & 1) == 0) {
                    p.tap_w[p.ntaps] = tap;
                    p.tap_off[p.ntaps++] = (q + pad - tap) / 2;
                }
            p.out4d = 1; p.out_q = q; p.out_step = 2;
            frames_out = (T_out - q + 1) / 2;
        }
        if (frames_out <= 0 || p.ntaps == 0) continue;
        if (!p.out4d && frames_out % F != 0) {        // ragged clips: the 4-D store clips the tail
            p.out4d = 1; p.out_q = 0; p.out_step = 1;
        }
        p.frames_out = frames_out;
        p.tiles_per_clip = (frames_out + F - 1) / F;
        p.tiles = NM * p.tiles_per_clip;
        int e;
        if (ncols == 256) e = tc::launch_tv<256>(amap, wmap, omap, omap_last, omap4, p, st);
        else if (ncols == 128) e = tc::launch_tv<128>(amap, wmap, omap, omap_last, omap4, p, st);
        else e = tc::launch_tv<64>(amap, wmap, omap, omap_last, omap4, p, st);
        if (e) return e;
    }
    return 0;
}
