// Weight AND adjacency gradient of the fused graph convolution from ONE pass over (dz, x)
// (reference: autograd of net/utils/tgcn.py:79-86 / net/utils/inceptionv2_gcn.py:69-88):
//
//   P[(v,w)][ci][c] = sum_f x[(f,v)][ci] * dz[(f,w)][c]           for every joint pair (v, w) that is
//                                                                 non-zero in some partition
//   dWc[k*Cin+ci][c] += sum_{(v,w)} A_eff[k][v][w] * P[(v,w)][ci][c]
//   dvals[(k,v,w)]   += sum_{ci,c} Wc[k*Cin+ci][c] * P[(v,w)][ci][c]
//
// i.e. the adjacency is folded OUT of the big contraction: the tensor-core kernel is a batch of plain
// "rows-contracted" products with both operands fed by TMA exactly as they lie in HBM -- no
// aggregation MMA (whose block-diagonal structure wastes 3/4 of every M=128 instruction in
// gcn_tc_dw2 / gcn_tc_da2), no lane masks, no TMEM->smem converter warps -- and the two small
// reductions that bring the partitions back in run on the CUDA cores over the pair matrices
// (195 pairs x Cin x Cout: 3 ... 51 MB).  Replaces istgcn_gcn_tc_dw + istgcn_gcn_tc_dvals.
//
// Kernel structure = csrc/tconv_dw_tc.cu (MN-major operands, 32-byte-atom 128B swizzle): a K-tile is
// 64 consecutive frames of ONE joint (3-D tensor map (C, V, frames), box 32 x 1 x 64).  A work item is a
// BLOCK of the joint-pair pattern: ns source joints (stacked rows R = vi*Cin + ci, M-blocks of 128 rows)
// x nd destination joints (columns n = j*ncw + c, ncw channels of each), accumulators resident in tensor
// memory (M-blocks x nd*ncw <= 512 columns).  The kernel is bound by the L2 -> shared-memory operand
// stream, i.e. by (rows + columns) per K-tile for rows x columns MACs, so the host covers the pattern
// with blocks as square as tensor memory allows (the skeleton's pattern is block-structured: 8 x 2 or
// 4 x 4 blocks at Cin = 64 are ~90 % full) and marks the cells each block owns; cells of a block that
// are not in the pattern (or belong to another block) are computed and dropped.
//   warp 0      TMA producer        warp 1      MMA issuer (kind::tf32, M=128, N=nb, K=8 frames)
//   warps 4-7   final epilogue: TMEM -> fp32 atomics into P
#include "tc_common.cuh"

namespace istgcn {
namespace tc {

constexpr int kThreadsGP = 256;
constexpr int kGPRows = 64;                       // frames of a K-tile atom
constexpr int kGPBytes = kGPRows * 128;           // [64 frames][32 channels]
constexpr int kGPAst = 4;                         // ring of M-block stages (4 atoms each); 3 at N = 256

struct PairParams {
    float* P;                                     // [npairs][Cin][Cout]
    const int4* items;                            // two per item: {d0, nd, s0, ns}, {col0, ncw, owned-cell mask, 0}
    const int4* ctas;                             // per CTA: {item, first K-tile, K-tile stride, 0}
    const int* joints;                            // destination / source joints of the items (d0 / s0 index it)
    const int* pair_of;                           // [V*V] (v*V + w) -> pair index, -1 outside the pattern
    int frames, V, Cin, Cout, ktiles;
};

struct SmemGP {
    // N-chunk <= 128: 4 M-block stages (128 KB) + 2 x 4 dz atoms (64 KB); N-chunk 256: 3 stages
    // (96 KB) + 2 x 8 dz atoms (128 KB); the dz buffers start right behind the ring
    static constexpr int a_off = 0;
    static constexpr int bar_off = 3 * 4 * kGPBytes + 2 * 8 * kGPBytes;
    static constexpr int kNumBars = 2 * kGPAst + 4 + 1;
    static constexpr int total = bar_off + kNumBars * 8 + 16;
};

__global__ void __launch_bounds__(kThreadsGP, 1)
gcn_pair_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap dmap,
                   PairParams p) {
    using L = SmemGP;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* As = smem + L::a_off;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bar_off);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + kGPAst;
    uint64_t* b_full = a_empty + kGPAst;
    uint64_t* b_empty = b_full + 2;
    uint64_t* done = b_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::kNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Work split: every CTA walks the K-tiles first, first + stride, ... of ONE item; the host gives an
    // item CTAs in proportion to its M-blocks, so all CTAs advance through the frames at the same
    // pace (the x / dz K-tiles they share stay in L2) and finish together.
    const int4 cta = p.ctas[blockIdx.x];
    const int4 item = p.items[2 * cta.x], item2 = p.items[2 * cta.x + 1];
    const int kt_first = cta.y, kt_step = cta.z;
    const int d0 = item.x, nd = item.y, s0 = item.z, ns = item.w;
    const int col0 = item2.x, ncw = item2.y;
    const uint32_t owned = (uint32_t)item2.z;       // bit vi*nd + j: cell (source vi, destination j) is this item's
    const int nb = nd * ncw, natom = nb / 32;
    const int ast = nb > 128 ? 3 : kGPAst;
    uint8_t* Bs = smem + L::a_off + ast * 4 * kGPBytes;
    const int rows_total = ns * p.Cin;
    const int nmb = (rows_total + 127) / 128;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(nmb * nb)) tmem_cols <<= 1;

    if (tid == 0) {
        for (int i = 0; i < kGPAst; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&xmap); tma_prefetch_desc(&dmap); }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        uint32_t ita = 0, itb = 0;
        for (int kt_i = kt_first; kt_i < p.ktiles; kt_i += kt_step, ++itb) {
            const int f0 = kt_i * kGPRows;
            const int bb = itb & 1;
            mbar_wait(&b_empty[bb], ((itb >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&b_full[bb], (uint32_t)kGPBytes * natom);
                for (int j = 0; j < natom; ++j) {
                    const int jd = (32 * j) / ncw, c = 32 * j - jd * ncw;
                    tma_load_3d(Bs + (bb * natom + j) * kGPBytes, &dmap, &b_full[bb], col0 + c, p.joints[d0 + jd], f0);
                }
            }
            __syncwarp();
            for (int mb = 0; mb < nmb; ++mb, ++ita) {
                const int sa = ita % ast;
                mbar_wait(&a_empty[sa], ((ita / ast) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&a_full[sa], (uint32_t)kGPBytes * 4);
                    for (int j = 0; j < 4; ++j) {
                        const int R = mb * 128 + 32 * j;                 // first stacked row of the atom
                        const int vi = R / p.Cin, ci0 = R - vi * p.Cin;
                        const bool ok = R < rows_total;
                        // atoms past the end of the item: a box outside the tensor reads as zeros
                        tma_load_3d(As + (sa * 4 + j) * kGPBytes, &xmap, &a_full[sa], ok ? ci0 : 0,
                                    ok ? p.joints[s0 + vi] : 0, ok ? f0 : -(1 << 20));
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc(128, nb, true, true);
        uint32_t ita = 0, itb = 0;
        bool first = true;
        for (int kt_i = kt_first; kt_i < p.ktiles; kt_i += kt_step, ++itb) {
            const int bb = itb & 1;
            mbar_wait(&b_full[bb], (itb >> 1) & 1);
            const uint32_t b_addr = smem_u32(Bs + bb * natom * kGPBytes);
            for (int mb = 0; mb < nmb; ++mb, ++ita) {
                const int sa = ita % ast;
                mbar_wait(&a_full[sa], (ita / ast) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_u32(As + sa * 4 * kGPBytes);
                    const uint32_t d_tmem = tmem_base + mb * nb;
#pragma unroll
                    for (int ks = 0; ks < kGPRows / 8; ++ks)
                        tc_mma_tf32(d_tmem, make_desc(a_addr + ks * 1024, kGPBytes, 512, 1),
                                    make_desc(b_addr + ks * 1024, kGPBytes, 512, 1), idesc,
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
        if (kt_first < p.ktiles) {
            for (int mb = 0; mb < nmb; ++mb) {
                const int R = mb * 128 + m;
                const int vi = R / p.Cin, ci = R - vi * p.Cin;
                const bool row_ok = R < rows_total;
                const int vsrc = row_ok ? p.joints[s0 + vi] : 0;
                for (int c0 = 0; c0 < nb; c0 += 32) {
                    const int jd = c0 / ncw, c = c0 - jd * ncw;
                    // a warp's 32 rows belong to one source joint (Cin is a multiple of 32): uniform skip
                    const bool mine = row_ok && ((owned >> (vi * nd + jd)) & 1u);
                    if (!__any_sync(0xffffffffu, mine)) continue;
                    float v[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + mb * nb + c0, v);
                    if (mine) {
                        const int pid = p.pair_of[vsrc * p.V + p.joints[d0 + jd]];
                        float* dst = p.P + ((size_t)pid * p.Cin + ci) * p.Cout + col0 + c;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)     // one 16-byte reduction instead of four scalar ones
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]),
                                         "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3])
                                         : "memory");
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

// dWc[k*Cin+ci][c] += sum over the entries (k, v, w) of partition k of vals * P[pair][ci][c]
__global__ void __launch_bounds__(256) pair_reduce_dw_kernel(const float* __restrict__ P,
                                                             const float* __restrict__ vals,
                                                             const int* __restrict__ entry_pair,
                                                             const int* __restrict__ k_ptr,
                                                             float* __restrict__ dWc, int K, int CC) {
    const long long total = (long long)K * CC;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / CC), r = (int)(i - (long long)k * CC);
        float s = 0.f;
        for (int e = k_ptr[k]; e < k_ptr[k + 1]; ++e) s = fmaf(vals[e], P[(size_t)entry_pair[e] * CC + r], s);
        dWc[i] += s;
    }
}

// dvals[e] += <Wc[k(e)], P[pair(e)]>: one CTA per entry
__global__ void __launch_bounds__(256) pair_reduce_da_kernel(const float* __restrict__ P,
                                                             const float* __restrict__ Wc,
                                                             const int* __restrict__ entry_pair,
                                                             const int* __restrict__ k_ptr,
                                                             float* __restrict__ dvals, int K, int CC) {
    __shared__ float red[8];
    const int e = blockIdx.x;
    int k = 0;
    while (k + 1 < K && e >= k_ptr[k + 1]) ++k;
    const float4* p4 = reinterpret_cast<const float4*>(P + (size_t)entry_pair[e] * CC);
    const float4* w4 = reinterpret_cast<const float4*>(Wc + (size_t)k * CC);
    float s = 0.f;
    for (int i = threadIdx.x; i < CC / 4; i += blockDim.x) {
        const float4 a = p4[i], b = w4[i];
        s = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, s))));
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        dvals[e] += t;
    }
}

}  // namespace tc
}  // namespace istgcn

using namespace istgcn;

ISTGCN_API int istgcn_gcn_pair_grads(const float* dz, const float* x, const float* vals, const float* Wc,
                                     const int* items, int nitems, const int* ctas, int nctas,
                                     const int* joints, const int* pair_of, int npairs,
                                     const int* entry_pair, const int* k_ptr, int nnz, float* P_ws,
                                     float* dWc, float* dvals, int frames, int V, int K, int Cin, int Cout,
                                     istgcn_stream_t s) {
    ISTGCN_REQUIRE(dz && x && vals && Wc && items && ctas && joints && pair_of && entry_pair && k_ptr && P_ws &&
                       dWc && dvals,
                   ISTGCN_E_ARG, "gcn_pair_grads: null pointer");
    ISTGCN_REQUIRE(Cin % 32 == 0 && Cout % 32 == 0 && Cin >= 32 && Cout >= 32 && V >= 1 && V <= 64 && K >= 1,
                   ISTGCN_E_SHAPE, "gcn_pair_grads: Cin=%d Cout=%d V=%d unsupported", Cin, Cout, V);
    ISTGCN_REQUIRE(((reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(x) |
                     reinterpret_cast<uintptr_t>(P_ws) | reinterpret_cast<uintptr_t>(Wc)) & 15) == 0,
                   ISTGCN_E_ARG, "gcn_pair_grads: pointers must be 16-byte aligned");
    if (frames == 0 || nitems == 0 || nctas == 0 || nnz == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    tc::PairParams p{};
    p.P = P_ws; p.items = reinterpret_cast<const int4*>(items); p.ctas = reinterpret_cast<const int4*>(ctas);
    p.joints = joints; p.pair_of = pair_of;
    p.frames = frames; p.V = V; p.Cin = Cin; p.Cout = Cout;
    p.ktiles = (frames + tc::kGPRows - 1) / tc::kGPRows;
    CUtensorMap xmap, dmap;
    if (int e = tc::encode_joint_frames_map(&xmap, x, frames, V, Cin, tc::kGPRows)) return e;
    if (int e = tc::encode_joint_frames_map(&dmap, dz, frames, V, Cout, tc::kGPRows)) return e;
    cudaFuncSetAttribute(tc::gcn_pair_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         tc::SmemGP::total);
    tc::gcn_pair_tc_kernel<<<nctas, tc::kThreadsGP, tc::SmemGP::total, st>>>(xmap, dmap, p);
    if (int e = finish_launch("gcn_pair_tc")) return e;
    const int CC = Cin * Cout;
    long long blocks = ((long long)K * CC + 255) / 256;
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
    tc::pair_reduce_dw_kernel<<<(int)blocks, 256, 0, st>>>(P_ws, vals, entry_pair, k_ptr, dWc, K, CC);
    if (int e = finish_launch("pair_reduce_dw")) return e;
    tc::pair_reduce_da_kernel<<<nnz, 256, 0, st>>>(P_ws, Wc, entry_pair, k_ptr, dvals, K, CC);
    return finish_launch("pair_reduce_da");
}
