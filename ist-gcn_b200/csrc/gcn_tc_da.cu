// tcgen05 kernel for the adjacency gradient of the fused graph convolution:
//
//     G_k[(f,w)][ci]   = sum_c dz[(f,w)][c] * Wc[k*Cin + ci][c]            (tensor cores)
//     dvals[(k,v,w)]  += sum_{f,ci} x[(f,v)][ci] * G_k[(f,w)][ci]          (CUDA cores)
//
// only at the static non-zero positions of A_eff.  dz (the gradient w.r.t. the graph-conv
// output, materialised by the input-gradient kernel) and Wc are both plain row-major matrices,
// so BOTH tcgen05 operands come straight from TMA: no thread ever touches an operand tile.
//
//   warp 0      TMA producer: per stage one dz atom [128 rows x 32 c] and K weight boxes
//               [32 ci x 32 c] (SWIZZLE_128B, K-major for both operands)
//   warp 1      MMA issuer: M=128, N=K*32, K=8 per instruction; accumulator = G for one
//               32-channel slice of ci and all K partitions, double buffered in TMEM
//   warps 4-7, 12-23  epilogue, four teams of four warps; team t takes every fourth entry of
//               each (k, w) list (the list lengths are skewed, max 12 vs mean 2.4, and a warp
//               runs max-over-lanes iterations):
//               tcgen05.ld (thread = row), 32-long dot products against the x slice
//               in shared memory (128-bit reads, four independent partial sums), shared-memory
//               atomics per non-zero
//   warps 8-11  loaders: x slice -> shared memory, [row][32 ci] with a 36-float pitch
#include <stdlib.h>

#include "tc_common.cuh"

namespace istgcn {
namespace tc {

#ifndef ISTGCN_DA_PROF
#define ISTGCN_DA_PROF 0      // 1: per-role clock64 accounting of CTA 0 (tools/dbg_da_time.py)
#endif
#if ISTGCN_DA_PROF
__device__ unsigned long long g_prof_da[32];
#define DPROF_DECL long long prof[32] = {0}
#define DPROF_T0() long long _t0 = clock64()
#define DPROF_ADD(i) do { long long _t1 = clock64(); prof[i] += _t1 - _t0; _t0 = _t1; } while (0)
#define DPROF_OUT(lo, hi) do { if (blockIdx.x == 0) for (int _i = lo; _i <= hi; ++_i) g_prof_da[_i] += prof[_i]; } while (0)
#else
#define DPROF_DECL
#define DPROF_T0()
#define DPROF_ADD(i)
#define DPROF_OUT(lo, hi)
#endif

constexpr int kTeamsDA = 4;                // epilogue warp teams (4 warps each)
constexpr int kThreadsDA = (12 + 4 * (kTeamsDA - 1)) * 32;   // TMA, MMA, 2 idle, team 0, loaders, teams 1..
constexpr int kStagesDA = 4;
constexpr int kXP = 36;                        // row pitch of the x slice (floats)

struct GcnDaParams {
    const float* x;                            // [rows][Cin]
    const float* vals_unused;
    const int *lptr, *lsrc, *lid;              // grouped by (k, destination w): source v, id
    float* dvals;
    int frames, V, K, Cin, CinPad, Cout, nnz, tiles;
};

struct SmemDA {
    static constexpr int stage_bytes = 2 * kAtomBytes;               // dz atom + weight atom
    static constexpr int ring_off = 0;
    static constexpr int xt_off = ring_off + kStagesDA * stage_bytes;
    static constexpr int list_off = xt_off + 2 * kAtomRows * kXP * 4;
    static constexpr int dv_off = list_off + kMaxNnz * 8 + (kMaxKV + 4) * 4;
    static constexpr int bar_off = dv_off + kMaxNnz * 4;
    static constexpr int kNumBars = 2 * kStagesDA + 4 + 4;
    static constexpr int total = bar_off + kNumBars * 8 + 16;
};

__global__ void __launch_bounds__(kThreadsDA, 1)
gcn_tc_da_kernel(const __grid_constant__ CUtensorMap dzmap, const __grid_constant__ CUtensorMap wmap,
                 GcnDaParams p) {
    using L = SmemDA;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* ring = smem + L::ring_off;
    float* XT = reinterpret_cast<float*>(smem + L::xt_off);
    int2* s_ent = reinterpret_cast<int2*>(smem + L::list_off);        // {source joint v, id}
    int* s_ptr = reinterpret_cast<int*>(s_ent + kMaxNnz);
    float* s_dv = reinterpret_cast<float*>(smem + L::dv_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bar_off);
    uint64_t* full = bars;
    uint64_t* empty = full + kStagesDA;
    uint64_t* t_full = empty + kStagesDA;
    uint64_t* t_empty = t_full + 2;
    uint64_t* x_full = t_empty + 2;
    uint64_t* x_empty = x_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::kNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = p.V, K = p.K, Cin = p.Cin, Cout = p.Cout;
    const int F = (kAtomRows / V) > 8 ? 8 : (kAtomRows / V);
    const int nchunk = p.CinPad / 32, natom = Cout / 32;
    const int NG = K * 32;                                            // MMA N

    for (int i = tid; i < p.nnz; i += kThreadsDA) {
        s_ent[i] = make_int2(p.lsrc[i], p.lid[i]);
        s_dv[i] = 0.f;
    }
    for (int i = tid; i <= K * V; i += kThreadsDA) s_ptr[i] = p.lptr[i];
    if (tid == 0) {
        for (int i = 0; i < kStagesDA; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4 * kTeamsDA);
            mbar_init(&x_full[i], 4); mbar_init(&x_empty[i], 4 * kTeamsDA);
        }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&dzmap); tma_prefetch_desc(&wmap); }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // TMA producer (warp-convergent loop, elected issue: see gcn_tc2.cu)
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int row0 = tile * F * V;
            for (int ch = 0; ch < nchunk; ++ch)
                for (int ca = 0; ca < natom; ++ca, ++it) {
                    const int s = it % kStagesDA;
                    mbar_wait(&empty[s], ((it / kStagesDA) & 1) ^ 1);
                    if (elect_one()) {
                        uint8_t* dst = ring + s * L::stage_bytes;
                        mbar_arrive_expect_tx(&full[s], kAtomBytes + K * 32 * 128);
                        tma_load_2d(dst, &dzmap, &full[s], ca * 32, row0);
                        for (int k = 0; k < K; ++k)
                            tma_load_2d(dst + kAtomBytes + k * 32 * 128, &wmap, &full[s], ca * 32,
                                        k * Cin + ch * 32);
                    }
                    __syncwarp();
                }
        }
    } else if (warp == 1) {
        // MMA issuer
        const uint32_t idesc = make_idesc(128, NG, false, false);
        uint32_t it = 0, cit = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x)
            for (int ch = 0; ch < nchunk; ++ch, ++cit) {
                const int buf = cit & 1;
                mbar_wait(&t_empty[buf], ((cit >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 128;
                for (int ca = 0; ca < natom; ++ca, ++it) {
                    const int s = it % kStagesDA;
                    mbar_wait(&full[s], (it / kStagesDA) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_addr = smem_u32(ring + s * L::stage_bytes);
                        const uint32_t b_addr = a_addr + kAtomBytes;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc_mma_tf32(d_tmem, make_desc(a_addr + ks * 32, 16, 1024),
                                        make_desc(b_addr + ks * 32, 16, 1024), idesc,
                                        (ca | ks) ? 1u : 0u);
                        tc_commit(&empty[s]);
                        if (ca == natom - 1) tc_commit(&t_full[buf]);
                    }
                    __syncwarp();
                }
            }
    } else if ((warp >= 4 && warp < 8) || warp >= 12) {
        // ---- epilogue: G rows from TMEM, dots against the x slice; team t takes the entries
        // beg + t, beg + t + kTeamsDA, ... of every (k, w) list
        const int ew = warp & 3;
        const int team = warp >= 12 ? 1 + ((warp - 12) >> 2) : 0;
        const int r = ew * 32 + lane;
        uint32_t cit = 0;
        DPROF_DECL; DPROF_T0();
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int f0 = tile * F;
            const int valid = min(F, p.frames - f0) * V;
            const bool ok = r < valid;
            const int fr = r / V, w = r - fr * V;
            for (int ch = 0; ch < nchunk; ++ch, ++cit) {
                const int buf = cit & 1;
                DPROF_ADD(7);
                mbar_wait(&t_full[buf], (cit >> 1) & 1);
                DPROF_ADD(5);
                mbar_wait(&x_full[buf], (cit >> 1) & 1);
                DPROF_ADD(6);
                tc_fence_after();
                const float* xt = XT + buf * kAtomRows * kXP + fr * V * kXP;
                for (int k = 0; k < K; ++k) {
                    float g[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * 128 + k * 32, g);
                    if (ok) {
                        for (int j = s_ptr[k * V + w] + ((team + k) % kTeamsDA); j < s_ptr[k * V + w + 1];
                             j += kTeamsDA) {
                            const int2 e = s_ent[j];
                            const float* xr = xt + e.x * kXP;
                            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                            for (int c4 = 0; c4 < 32; c4 += 4) {
                                const float4 xv = *reinterpret_cast<const float4*>(xr + c4);
                                a0 = fmaf(g[c4], xv.x, a0); a1 = fmaf(g[c4 + 1], xv.y, a1);
                                a2 = fmaf(g[c4 + 2], xv.z, a2); a3 = fmaf(g[c4 + 3], xv.w, a3);
                            }
                            atomicAdd(&s_dv[e.y], (a0 + a1) + (a2 + a3));
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&t_empty[buf]); mbar_arrive(&x_empty[buf]); }
            }
        }
        DPROF_ADD(7);
        if (tid == 128) DPROF_OUT(5, 7);
    } else if (warp >= 8 && warp < 12) {
        // ---- loaders: x slice [rows][32 ci] -> XT[ci][row]
        const int lt = tid - 8 * 32;
        uint32_t cit = 0;
        DPROF_DECL; DPROF_T0();
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int f0 = tile * F;
            const int valid = min(F, p.frames - f0) * V;
            const long long row0 = (long long)f0 * V;
            for (int ch = 0; ch < nchunk; ++ch, ++cit) {
                const int buf = cit & 1;
                DPROF_ADD(9);
                mbar_wait(&x_empty[buf], ((cit >> 1) & 1) ^ 1);
                DPROF_ADD(8);
                float* xt = XT + buf * kAtomRows * kXP;
                const int ci0 = ch * 32;
                if ((Cin & 3) == 0) {
                    float4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = lt + u * 128;
                        const int rr = i >> 3, c4 = (i & 7) * 4;
                        v[u] = (rr < valid && ci0 + c4 < Cin) ? ld4(p.x + (row0 + rr) * Cin + ci0 + c4)
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = lt + u * 128;
                        const int rr = i >> 3, c4 = (i & 7) * 4;
                        st4(xt + rr * kXP + c4, v[u]);
                    }
                } else {
                    for (int i = lt; i < kAtomRows * 32; i += 128) {
                        const int rr = i >> 5, c = i & 31;
                        xt[rr * kXP + c] = (rr < valid && ci0 + c < Cin) ? p.x[(row0 + rr) * Cin + ci0 + c] : 0.f;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&x_full[buf]);
            }
        }
        DPROF_ADD(9);
        if (tid == 256) DPROF_OUT(8, 9);
    }

    tc_fence_before();
    __syncthreads();
    for (int i = tid; i < p.nnz; i += kThreadsDA) atomicAdd(&p.dvals[i], s_dv[i]);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace tc
}  // namespace istgcn

using namespace istgcn;

#if ISTGCN_DA_PROF
extern "C" int istgcn_debug_prof_da(unsigned long long* out) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, tc::g_prof_da, sizeof(unsigned long long) * 32);
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(tc::g_prof_da, z, sizeof(z));
    return 0;
}
#endif

// dvals[id] += sum_{f,ci} x[(f,v)][ci] * (dz Wc_k^T)[(f,w)][ci] over the non-zeros (k,v,w).
// dz [frames*V][Cout] and Wc [K*Cin][Cout] row-major, 16-byte aligned, Cout % 32 == 0.
ISTGCN_API int istgcn_gcn_tc_dvals(const float* dz, const float* x, const float* Wc, const int* lptr,
                                   const int* lsrc, const int* lid, int nnz, float* dvals, int frames,
                                   int V, int K, int Cin, int Cout, istgcn_stream_t s) {
    ISTGCN_REQUIRE(dz && x && Wc && lptr && lsrc && lid && dvals, ISTGCN_E_ARG, "gcn_tc_dvals: null pointer");
    ISTGCN_REQUIRE(V >= 1 && V <= 32 && K >= 1 && K <= 4, ISTGCN_E_SHAPE, "gcn_tc_dvals: V=%d K=%d", V, K);
    ISTGCN_REQUIRE(Cout % 32 == 0 && Cout >= 32, ISTGCN_E_SHAPE, "gcn_tc_dvals: Cout=%d must be a multiple of 32", Cout);
    ISTGCN_REQUIRE(Cin >= 1 && (Cin < 32 || Cin % 32 == 0), ISTGCN_E_SHAPE, "gcn_tc_dvals: Cin=%d unsupported", Cin);
    ISTGCN_REQUIRE(nnz >= 0 && nnz <= kMaxNnz, ISTGCN_E_SHAPE, "gcn_tc_dvals: nnz=%d", nnz);
    if (frames == 0) return 0;
    // second-generation kernel (both contractions on the tensor core, gcn_tc_da2.cu)
    static const bool force_v1 = getenv("ISTGCN_GCN_TC_V1") != nullptr;
    if (!force_v1 && tc::gcn_tc_da2_eligible(V, K, Cin, Cout) &&
        ((reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(Wc)) & 15) == 0)
        return tc::launch_gcn_tc_da2(dz, x, Wc, lptr, lsrc, lid, nnz, dvals, frames, V, K, Cin, Cout,
                                     (cudaStream_t)s);
    tc::GcnDaParams p{x, nullptr, lptr, lsrc, lid, dvals, frames, V, K, Cin, (Cin + 31) / 32 * 32, Cout, nnz, 0};
    const int F = kTileRows / V > 8 ? 8 : kTileRows / V;
    p.tiles = (frames + F - 1) / F;
    CUtensorMap dzmap, wmap;
    if (int e = tc::encode_tile_map(&dzmap, dz, (long long)frames * V, Cout, 128)) return e;
    if (int e = tc::encode_tile_map(&wmap, Wc, (long long)K * Cin, Cout, 32)) return e;
    cudaFuncSetAttribute(tc::gcn_tc_da_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemDA::total);
    int nx = num_sms();
    if (nx > p.tiles) nx = p.tiles;
    tc::gcn_tc_da_kernel<<<nx, tc::kThreadsDA, tc::SmemDA::total, (cudaStream_t)s>>>(dzmap, wmap, p);
    return finish_launch("gcn_tc_dvals");
}
