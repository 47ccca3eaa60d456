// tcgen05 kernel for the WEIGHT gradient of the full-width temporal convolution
// (net/st_gcnold.py:165-171, backward of Conv2d(C, C, (kt,1), (stride,1), (pad,0))):
//
//     dW[tap*Cin + ci][co] = sum_{n,to,v} a[(n, to*stride + tap - pad, v)][ci] * du[(n,to,v)][co]
//
// The contraction runs over ROWS, so both operands are MN-major (32-byte-atom 128B swizzle, as in
// gcn_tc_dw.cu) and both come straight from TMA through 4-D (C, V, T, NM) tensor maps: a K-tile
// is FK whole frames of one clip (FK*V <= 64 rows), the `a` operand of tap `tap` is the same box
// shifted by tap - pad frames (out-of-range frames read as zeros = the padding; a frame past the
// end of the clip is zero in du, so ragged clip ends need no special case).
//
// The stacked gradient matrix [kt*Cin][Cout] is cut into M-blocks of 128 rows (= 4 atoms of 32
// rows, each atom one (tap, ci-slice) pair) and N-chunks of <= 128 columns.  A CTA owns up to
// 512 / N-chunk M-blocks of one N-chunk -- its accumulators stay in tensor memory for the whole
// kernel -- and a strided share of the K-tiles; the result is flushed once with fp32 atomics.
//
//   warp 0      TMA producer: per K-tile the du atoms of the N-chunk, then 4 `a` atoms per M-block
//   warp 1      MMA issuer: kind::tf32, M=128, N=N-chunk, K=8 rows per instruction
//   warps 4-7   final epilogue: TMEM -> atomics
#include "tc_common.cuh"

namespace istgcn {
namespace tc {

constexpr int kThreadsTW = 256;
constexpr int kSubRows = 64;                      // rows of a K-tile atom
constexpr int kSubBytes = kSubRows * 128;         // [64 rows][32 channels]
constexpr int kNAst = 3;                          // ring of M-block stages (4 atoms each)

struct TconvDwParams {
    float* dW;                                    // [kt*Cin][Cout]
    int NM, T, Tout, V, Cin, Cout, kt, stride, FK, ksteps, ktiles, tiles_per_clip, nb, nmb, nmb_total;
};

struct SmemTW {
    static constexpr int a_off = 0;                                   // kNAst x 4 atoms
    static constexpr int b_off = a_off + kNAst * 4 * kSubBytes;       // 2 x (nb/32 <= 4) atoms
    static constexpr int bar_off = b_off + 2 * 4 * kSubBytes;
    static constexpr int kNumBars = 2 * kNAst + 4 + 1;
    static constexpr int total = bar_off + kNumBars * 8 + 16;
};

__global__ void __launch_bounds__(kThreadsTW, 1)
tconv_dw_tc_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap dmap,
                   TconvDwParams p) {
    using L = SmemTW;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* As = smem + L::a_off;
    uint8_t* Bs = smem + L::b_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bar_off);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + kNAst;
    uint64_t* b_full = a_empty + kNAst;
    uint64_t* b_empty = b_full + 2;
    uint64_t* done = b_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::kNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = p.nb, natom = nb / 32;
    const int nchunks = p.Cout / nb;
    const int ngroup = blockIdx.y / nchunks;                 // M-block group of this CTA
    const int col0 = (blockIdx.y % nchunks) * nb;            // first output column
    const int mb0 = ngroup * p.nmb;
    const int nmb = min(p.nmb, p.nmb_total - mb0);
    const int pad = (p.kt - 1) / 2;
    const int rows_total = p.kt * p.Cin;
    const uint32_t atom_tx = (uint32_t)(p.FK * p.V) * 128u;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(p.nmb * nb)) tmem_cols <<= 1;

    // rows FK*V .. 63 of every atom are never written by TMA: they must be zero on both sides
    for (int i = tid; i < L::bar_off / 4; i += kThreadsTW) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < kNAst; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&amap); tma_prefetch_desc(&dmap); }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // TMA producer (warp-convergent loop, elected issue: see gcn_tc2.cu)
        uint32_t ita = 0, itb = 0;
        for (int kt_i = blockIdx.x; kt_i < p.ktiles; kt_i += gridDim.x, ++itb) {
            const int n = kt_i / p.tiles_per_clip;
            const int to0 = (kt_i - n * p.tiles_per_clip) * p.FK;
            const int bb = itb & 1;
            mbar_wait(&b_empty[bb], ((itb >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&b_full[bb], atom_tx * natom);
                for (int j = 0; j < natom; ++j)
                    tma_load_4d(Bs + (bb * 4 + j) * kSubBytes, &dmap, &b_full[bb], col0 + 32 * j, 0, to0, n);
            }
            __syncwarp();
            for (int mb = 0; mb < nmb; ++mb, ++ita) {
                const int sa = ita % kNAst;
                mbar_wait(&a_empty[sa], ((ita / kNAst) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&a_full[sa], atom_tx * 4);
                    for (int j = 0; j < 4; ++j) {
                        const int R = (mb0 + mb) * 128 + 32 * j;         // first stacked row of the atom
                        const int tap = R / p.Cin, ci0 = R - tap * p.Cin;
                        // atoms past the end of the matrix: an out-of-range frame reads as zeros
                        const int t_first = R < rows_total ? to0 * p.stride + tap - pad : -(1 << 20);
                        tma_load_4d(As + (sa * 4 + j) * kSubBytes, &amap, &a_full[sa],
                                    R < rows_total ? ci0 : 0, 0, t_first, n);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // MMA issuer
        const uint32_t idesc = make_idesc(128, nb, true, true);
        uint32_t ita = 0, itb = 0;
        bool first = true;
        for (int kt_i = blockIdx.x; kt_i < p.ktiles; kt_i += gridDim.x, ++itb) {
            const int bb = itb & 1;
            mbar_wait(&b_full[bb], (itb >> 1) & 1);
            const uint32_t b_addr = smem_u32(Bs + bb * 4 * kSubBytes);
            for (int mb = 0; mb < nmb; ++mb, ++ita) {
                const int sa = ita % kNAst;
                mbar_wait(&a_full[sa], (ita / kNAst) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_u32(As + sa * 4 * kSubBytes);
                    const uint32_t d_tmem = tmem_base + mb * nb;
                    for (int ks = 0; ks < p.ksteps; ++ks)
                        tc_mma_tf32(d_tmem, make_desc(a_addr + ks * 1024, kSubBytes, 512, 1),
                                    make_desc(b_addr + ks * 1024, kSubBytes, 512, 1), idesc,
                                    (first && ks == 0) ? 0u : 1u);
                    tc_commit(&a_empty[sa]);
                    if (mb == nmb - 1) tc_commit(&b_empty[bb]);
                }
                __syncwarp();
            }
            first = false;
        }
        if (elect_one()) tc_commit(done);
        __syncwarp();
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int m = ew * 32 + lane;
        mbar_wait(done, 0);
        tc_fence_after();
        if (blockIdx.x < p.ktiles) {
            for (int mb = 0; mb < nmb; ++mb) {
                const int R = (mb0 + mb) * 128 + m;
                for (int c0 = 0; c0 < nb; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + mb * nb + c0, v);
                    if (R < rows_total) {
                        float* dst = p.dW + (size_t)R * p.Cout + col0 + c0;
                        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {     // 16-byte reductions
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                red_add4(dst + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j]);
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace tc
}  // namespace istgcn

using namespace istgcn;

// dW[kt*Cin][Cout] += sum over rows of a_shifted^T du (see the file header); dbias_vc[V][Cout]
// (may be NULL) += sum over frames of du.  a [NM][T][V][Cin], du [NM][Tout][V][Cout], both
// channels-last and 16-byte aligned; outputs caller-zeroed.  Cin, Cout multiples of 32.
ISTGCN_API int istgcn_tconv_dw_tc(const float* a, const float* du, float* dW, float* dbias_vc, int NM,
                                  int T, int Tout, int V, int Cin, int Cout, int kt, int stride,
                                  istgcn_stream_t s) {
    ISTGCN_REQUIRE(a && du && dW, ISTGCN_E_ARG, "tconv_dw_tc: null pointer");
    ISTGCN_REQUIRE(V >= 1 && V <= 32 && kt >= 1 && (kt & 1) && kt <= 31, ISTGCN_E_SHAPE,
                   "tconv_dw_tc: V=%d kt=%d unsupported", V, kt);
    ISTGCN_REQUIRE(Cin % 32 == 0 && Cout % 32 == 0 && Cin >= 32 && Cout >= 32, ISTGCN_E_SHAPE,
                   "tconv_dw_tc: Cin=%d Cout=%d must be multiples of 32", Cin, Cout);
    ISTGCN_REQUIRE(Cout <= 128 || Cout % 128 == 0, ISTGCN_E_SHAPE,
                   "tconv_dw_tc: Cout=%d must be <= 128 or a multiple of 128", Cout);
    ISTGCN_REQUIRE(stride >= 1 && Tout == (T - 1) / stride + 1, ISTGCN_E_SHAPE,
                   "tconv_dw_tc: Tout=%d does not match T=%d stride=%d", Tout, T, stride);
    ISTGCN_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(du)) & 15) == 0,
                   ISTGCN_E_ARG, "tconv_dw_tc: pointers must be 16-byte aligned");
    if ((long long)NM * Tout == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    tc::TconvDwParams p{};
    p.dW = dW; p.NM = NM; p.T = T; p.Tout = Tout; p.V = V; p.Cin = Cin; p.Cout = Cout; p.kt = kt;
    p.stride = stride;
    p.FK = tc::kSubRows / V;
    p.ksteps = (p.FK * V + 7) / 8;
    p.tiles_per_clip = (Tout + p.FK - 1) / p.FK;
    p.ktiles = NM * p.tiles_per_clip;
    p.nb = Cout <= 128 ? Cout : 128;
    p.nmb_total = (kt * Cin + 127) / 128;
    p.nmb = 512 / p.nb;
    if (p.nmb > p.nmb_total) p.nmb = p.nmb_total;
    const int groups = (p.nmb_total + p.nmb - 1) / p.nmb;
    const int ny = groups * (Cout / p.nb);
    CUtensorMap amap, dmap;
    if (int e = tc::encode_frames_map(&amap, a, NM, T, V, Cin, p.FK, stride, true)) return e;
    if (int e = tc::encode_frames_map(&dmap, du, NM, Tout, V, Cout, p.FK, 1, true)) return e;
    cudaFuncSetAttribute(tc::tconv_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         tc::SmemTW::total);
    int nx = num_sms() / ny;
    if (nx < 1) nx = 1;
    if (nx > p.ktiles) nx = p.ktiles;
    tc::tconv_dw_tc_kernel<<<dim3(nx, ny), tc::kThreadsTW, tc::SmemTW::total, st>>>(amap, dmap, p);
    if (int e = finish_launch("tconv_dw_tc")) return e;
    if (dbias_vc)
        if (int e = tc::launch_frame_colsum(du, dbias_vc, NM * Tout, V * Cout, st)) return e;
    return 0;
}
