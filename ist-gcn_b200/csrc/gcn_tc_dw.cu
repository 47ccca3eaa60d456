// tcgen05 kernel for the weight gradient of the fused graph convolution:
//
//     dWc[k*Cin + ci][c] = sum_rows X'_k[row][ci] * dz[row][c],
//     X'_k[(f,w)][ci]    = sum_v A_eff[k][v][w] * x[(f,v)][ci]          (re-aggregated on the fly)
//
// The contraction runs over ROWS, so both operands are read "MN-major" by the tensor core: the
// very same [128 rows x 32 channels] SWIZZLE_128B atoms that are K-major operands in the forward
// kernel are here M-major (X' atoms written by the aggregator warps) and N-major (dz atoms
// dropped in by TMA), with K = the 128 rows of the frame tile (16 MMA K-steps of 8 rows).
// MN-major TF32 operands must use the 32-byte-atom 128B swizzle (UMMA SWIZZLE_128B_BASE32B /
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), so the atoms of this kernel use that pattern.
//
// Work split: a "row block" = one 32-channel slice of x times the K partitions = 128 rows of
// dWc; a CTA owns `nrb` consecutive row blocks for ALL Cout columns (nrb * Cout <= 512 TMEM
// columns) and a strided share of the frame tiles; its accumulators stay in tensor memory for
// the whole kernel and are flushed once with fp32 atomics.
//
//   warp 0        TMA producer: dz atoms [128 rows x 32 c] -> ring
//   warp 1        MMA issuer: kind::tf32, M=128 (4 X' atoms), N=32 (one dz atom), K=8 rows
//   warps 2-3,16-17  loaders: x slice -> smem
//   warps 4-7     final epilogue: TMEM -> fp32 atomics into dWc
//   warps 8-15    aggregators: x slice -> the 4 (K) X' atoms of the row block (double buffered)
#include <stdlib.h>

#include "tc_common.cuh"

namespace istgcn {
namespace tc {

constexpr int kThreadsDW = 576;
constexpr int kNBdw = 3;          // dz atom ring
constexpr int kNXdw = 2;          // x slice ring

struct GcnDwParams {
    const float* x;
    const float* vals;
    const int *lptr, *lsrc, *lid;     // grouped by (k, destination w)
    float* dWc;                       // [K*Cin][Cout]
    int frames, V, K, Cin, CinPad, Cout, nnz, tiles, nrb;
    FrameMap in_map;
};

struct SmemDW {
    static constexpr int xp_off = 0;                                   // 2 buffers x 4 atoms
    static constexpr int b_off = xp_off + 2 * 4 * kAtomBytes;
    static constexpr int x_off = b_off + kNBdw * kAtomBytes;
    static constexpr int list_off = x_off + kNXdw * kAtomRows * 32 * 4;
    static constexpr int bar_off = list_off + kMaxNnz * 8 + (kMaxKV + 4) * 4;
    static constexpr int kNumBars = 4 + 2 * kNBdw + 2 * kNXdw + 1;
    static constexpr int total = bar_off + kNumBars * 8 + 16;
};

__global__ void __launch_bounds__(kThreadsDW, 1)
gcn_tc_dw_kernel(const __grid_constant__ CUtensorMap dzmap, GcnDwParams p) {
    using L = SmemDW;
    extern __shared__ __align__(1024) uint8_t smem[];
    float* XP = reinterpret_cast<float*>(smem + L::xp_off);
    uint8_t* Bs = smem + L::b_off;
    float* Xs = reinterpret_cast<float*>(smem + L::x_off);
    int2* s_ent = reinterpret_cast<int2*>(smem + L::list_off);
    int* s_ptr = reinterpret_cast<int*>(s_ent + kMaxNnz);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bar_off);
    uint64_t* xp_full = bars;
    uint64_t* xp_empty = xp_full + 2;
    uint64_t* b_full = xp_empty + 2;
    uint64_t* b_empty = b_full + kNBdw;
    uint64_t* x_full = b_empty + kNBdw;
    uint64_t* x_empty = x_full + kNXdw;
    uint64_t* done = x_empty + kNXdw;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::kNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = p.V, K = p.K, Cin = p.Cin, Cout = p.Cout;
    const int F = (kAtomRows / V) > 8 ? 8 : (kAtomRows / V);
    const int nchunk = p.CinPad / 32, natom = Cout / 32;
    const int rb0 = blockIdx.y * p.nrb;                       // first row block (= x slice) of this CTA
    const int nrb = min(p.nrb, nchunk - rb0);
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(p.nrb * Cout)) tmem_cols <<= 1;

    for (int i = tid; i < p.nnz; i += kThreadsDW)
        s_ent[i] = make_int2(p.lsrc[i] * 32, __float_as_int(p.vals[p.lid[i]]));
    for (int i = tid; i <= K * V; i += kThreadsDW) s_ptr[i] = p.lptr[i];
    for (int i = tid; i < 2 * 4 * kAtomBytes / 4; i += kThreadsDW) XP[i] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&xp_full[i], 8); mbar_init(&xp_empty[i], 1); }
        for (int i = 0; i < kNBdw; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < kNXdw; ++i) { mbar_init(&x_full[i], 4); mbar_init(&x_empty[i], 8); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) tma_prefetch_desc(&dzmap);
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =========================== TMA producer: dz atoms, once per (tile, row block, c-atom)
        // (warp-convergent loop, elected issue: see gcn_tc2.cu)
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int row0 = tile * F * V;
            for (int rb = 0; rb < nrb; ++rb)
                for (int ca = 0; ca < natom; ++ca, ++it) {
                    const int sb = it % kNBdw;
                    mbar_wait(&b_empty[sb], ((it / kNBdw) & 1) ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&b_full[sb], kAtomBytes);
                        tma_load_2d(Bs + sb * kAtomBytes, &dzmap, &b_full[sb], ca * 32, row0);
                    }
                    __syncwarp();
                }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer
        constexpr uint32_t idesc = make_idesc(128, 32, true, true);
        uint32_t it = 0, xit = 0;
        bool first_tile = true;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            for (int rb = 0; rb < nrb; ++rb, ++xit) {
                const int b = xit & 1;
                mbar_wait(&xp_full[b], (xit >> 1) & 1);
                const uint32_t a_addr = smem_u32(XP) + b * 4 * kAtomBytes;
                for (int ca = 0; ca < natom; ++ca, ++it) {
                    const int sb = it % kNBdw;
                    mbar_wait(&b_full[sb], (it / kNBdw) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t b_addr = smem_u32(Bs) + sb * kAtomBytes;
                        const uint32_t d_tmem = tmem_base + rb * Cout + ca * 32;
#pragma unroll
                        for (int ks = 0; ks < 16; ++ks)       // 8 rows of the tile per step
                            tc_mma_tf32(d_tmem, make_desc(a_addr + ks * 1024, kAtomBytes, 512, 1),
                                        make_desc(b_addr + ks * 1024, kAtomBytes, 512, 1), idesc,
                                        (first_tile && ks == 0) ? 0u : 1u);
                        tc_commit(&b_empty[sb]);
                        if (ca == natom - 1) tc_commit(&xp_empty[b]);
                    }
                    __syncwarp();
                }
            }
            first_tile = false;
        }
        if (elect_one()) tc_commit(done);
        __syncwarp();
    } else if (warp >= 4 && warp < 8) {
        // =========================== final epilogue: accumulators -> dWc (fp32 atomics)
        const int ew = warp - 4;
        const int m = ew * 32 + lane;                          // row of the row block
        const int k = m >> 5, cil = m & 31;
        mbar_wait(done, 0);
        tc_fence_after();
        if (blockIdx.x < p.tiles) {
            for (int rb = 0; rb < nrb; ++rb) {
                const int ci = (rb0 + rb) * 32 + cil;
                for (int c0 = 0; c0 < Cout; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + rb * Cout + c0, v);
                    if (k < K && ci < Cin) {
                        float* dst = p.dWc + (size_t)(k * Cin + ci) * Cout + c0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j]);
                    }
                }
            }
        }
    } else if (warp < 4 || warp >= 16) {
        // =========================== loaders: x slice -> Xs
        const int lt = warp < 4 ? tid - 64 : tid - 16 * 32 + 64;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int f0 = tile * F;
            const int valid = min(F, p.frames - f0) * V;
            const long long row0 = (long long)f0 * V;
            for (int rb = 0; rb < nrb; ++rb, ++it) {
                const int xb = it % kNXdw;
                mbar_wait(&x_empty[xb], ((it / kNXdw) & 1) ^ 1);
                float* xs = Xs + xb * kAtomRows * 32;
                const int ci0 = (rb0 + rb) * 32;
                if ((Cin & 3) == 0) {
                    float4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = lt + u * 128;
                        const int r = i >> 3, c4 = (i & 7) * 4;
                        const long long src = (r < valid && ci0 + c4 < Cin) ? map_row(p.in_map, row0 + r, V) : -1;
                        v[u] = src >= 0 ? ld4(p.x + src * Cin + ci0 + c4)
                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = lt + u * 128;
                        st4(xs + (i >> 3) * 32 + (i & 7) * 4, v[u]);
                    }
                } else {
                    for (int i = lt; i < kAtomRows * 32; i += 128) {
                        const int r = i >> 5, c = i & 31;
                        const long long src = (r < valid && ci0 + c < Cin) ? map_row(p.in_map, row0 + r, V) : -1;
                        xs[i] = src >= 0 ? p.x[src * Cin + ci0 + c] : 0.f;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&x_full[xb]);
            }
        }
    } else {
        // =========================== aggregators: Xs -> the K atoms of XP buffer b
        const int aw = warp - 8;
        const int q = lane >> 3, c4 = (lane & 7) * 4;
        const int w = aw + 8 * q;
        const bool active = w < V;
        const int fstride = V * 32;
        uint32_t xit = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            for (int rb = 0; rb < nrb; ++rb, ++xit) {
                const int xb = xit % kNXdw, b = xit & 1;
                mbar_wait(&x_full[xb], (xit / kNXdw) & 1);
                mbar_wait(&xp_empty[b], ((xit >> 1) & 1) ^ 1);
                const float* xs = Xs + xb * kAtomRows * 32 + c4;
                if (active) {
                    for (int k = 0; k < K; ++k) {
                        float* A = XP + (b * 4 + k) * (kAtomBytes / 4);
                        aggregate_joint_any<true>(F, A, xs, s_ent, s_ptr[k * V + w], s_ptr[k * V + w + 1],
                                            fstride, V, w, c4);
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&xp_full[b]); mbar_arrive(&x_empty[xb]); }
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

// column sums over frames: out[j] += sum_f in[f][j], j < n (n = V*C, a multiple of 4)
__global__ void frame_colsum_kernel(const float* __restrict__ in, float* __restrict__ out, int frames,
                                    int n, int frames_per_cta) {
    const int f0 = blockIdx.y * frames_per_cta, f1 = min(frames, f0 + frames_per_cta);
    for (int j4 = blockIdx.x * blockDim.x + threadIdx.x; j4 < n / 4; j4 += gridDim.x * blockDim.x) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int f = f0; f < f1; ++f) {
            const float4 v = ld4(in + (size_t)f * n + j4 * 4);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        atomicAdd(out + j4 * 4 + 0, a.x);
        atomicAdd(out + j4 * 4 + 1, a.y);
        atomicAdd(out + j4 * 4 + 2, a.z);
        atomicAdd(out + j4 * 4 + 3, a.w);
    }
}

int launch_frame_colsum(const float* in, float* out, int frames, int n, cudaStream_t st) {
    int slabs = (num_sms() * 4 * 256) / (n / 4 > 0 ? n / 4 : 1);
    if (slabs < 1) slabs = 1;
    if (slabs > frames) slabs = frames;
    const int fpc = (frames + slabs - 1) / slabs;
    dim3 grid((n / 4 + 255) / 256, (frames + fpc - 1) / fpc);
    frame_colsum_kernel<<<grid, 256, 0, st>>>(in, out, frames, n, fpc);
    return finish_launch("frame_colsum");
}

}  // namespace tc
}  // namespace istgcn

using namespace istgcn;

// dWc[K*Cin][Cout] += X'^T dz on the tcgen05 engine; dbiasterm[V][Cout] += sum over frames of dz.
// dz [frames*V][Cout] row-major (16-byte aligned, Cout % 32 == 0), both outputs caller-zeroed.
ISTGCN_API int istgcn_gcn_tc_dw(const float* dz, const float* x, const float* vals, const int* lptr,
                                const int* lsrc, const int* lid, int nnz, float* dWc,
                                float* dbiasterm, int frames, int V, int K, int Cin, int Cout,
                                int t_in, int t_out, int t_stride, int t_offset, istgcn_stream_t s) {
    ISTGCN_REQUIRE(dz && x && vals && lptr && lsrc && lid && dWc, ISTGCN_E_ARG, "gcn_tc_dw: null pointer");
    ISTGCN_REQUIRE(V >= 1 && V <= 32 && K >= 1 && K <= 4, ISTGCN_E_SHAPE, "gcn_tc_dw: V=%d K=%d", V, K);
    ISTGCN_REQUIRE(Cout % 32 == 0 && Cout >= 32 && Cout <= 512, ISTGCN_E_SHAPE,
                   "gcn_tc_dw: Cout=%d must be a multiple of 32 (<= 512)", Cout);
    ISTGCN_REQUIRE(Cin >= 1 && (Cin < 32 || Cin % 32 == 0), ISTGCN_E_SHAPE, "gcn_tc_dw: Cin=%d unsupported", Cin);
    ISTGCN_REQUIRE(nnz >= 0 && nnz <= kMaxNnz, ISTGCN_E_SHAPE, "gcn_tc_dw: nnz=%d", nnz);
    if (frames == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    // second-generation kernel (aggregation on the tensor core, gcn_tc_dw2.cu) for plain frame maps
    static const bool force_v1 = getenv("ISTGCN_GCN_TC_V1") != nullptr;
    if (!force_v1 && t_out == 0 && tc::gcn_tc_dw2_eligible(V, K, Cin, Cout) &&
        ((reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(x)) & 15) == 0) {
        if (int e = tc::launch_gcn_tc_dw2(dz, x, vals, lptr, lsrc, lid, dWc, frames, V, K, Cin, Cout, st))
            return e;
        if (dbiasterm)
            if (int e = tc::launch_frame_colsum(dz, dbiasterm, frames, V * Cout, st)) return e;
        return 0;
    }
    tc::GcnDwParams p{x, vals, lptr, lsrc, lid, dWc, frames, V, K, Cin, (Cin + 31) / 32 * 32, Cout, nnz, 0, 0,
                       {t_in, t_out, t_stride, t_offset}};
    const int F = kTileRows / V > 8 ? 8 : kTileRows / V;
    p.tiles = (frames + F - 1) / F;
    const int nchunk = p.CinPad / 32;
    p.nrb = 512 / Cout;
    if (p.nrb > nchunk) p.nrb = nchunk;
    const int groups = (nchunk + p.nrb - 1) / p.nrb;
    CUtensorMap dzmap;
    if (int e = tc::encode_tile_map(&dzmap, dz, (long long)frames * V, Cout, 128, true)) return e;
    cudaFuncSetAttribute(tc::gcn_tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemDW::total);
    int nx = num_sms() / groups;
    if (nx < 1) nx = 1;
    if (nx > p.tiles) nx = p.tiles;
    tc::gcn_tc_dw_kernel<<<dim3(nx, groups), tc::kThreadsDW, tc::SmemDW::total, st>>>(dzmap, p);
    if (int e = finish_launch("gcn_tc_dw")) return e;
    if (dbiasterm)
        if (int e = tc::launch_frame_colsum(dz, dbiasterm, frames, V * Cout, st)) return e;
    return 0;
}
