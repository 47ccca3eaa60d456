// Second-generation tcgen05 kernel for the weight gradient of the fused graph convolution
// (reference forward: net/utils/tgcn.py:76-89):
//
//     dWc[k*Cin + ci][n] = sum_rows X'_k[row][ci] * dz[row][n],
//     X'_k[(f,w)][ci]    = sum_v A_eff[k][v][w] * x[(f,v)][ci]
//
// The aggregation runs on the tensor core exactly as in gcn_tc2.cu (adjacency in tensor memory,
// lane-masked MMA 1 per frame, input frames by TMA).  The contraction of the weight gradient runs
// over ROWS, so X' has to be a shared-memory operand: four converter warps move each aggregated
// slice TMEM -> registers -> the MN-major (32-byte-atom 128B swizzle) operand tile, 16 KB per
// partition, and MMA 2 multiplies it with the dz tile that TMA dropped in the same layout:
//
//     ACC[rb][(k,ci)][n] += XP[rows][(k,ci)]^T * DZ[rows][n]      M = 128, N = Cout, 16 x K = 8 rows
//
// A frame tile = 4 frames in 32-row slots (pad rows: zero lanes of MMA 1 / zero-filled dz rows).
// Accumulators stay in tensor memory for the whole kernel (nrb row blocks x Cout columns <= 256
// next to the adjacency and D1) and are flushed once with fp32 atomics; the row blocks that do not
// fit are spread over up to eight CTA groups (blockIdx.y), which walk the same tile sequence in
// step so that the dz tiles they all read stay in L2 (Cout = 256: one row block per group).
//
//   warp 0  TMA producer (x slices)          warp 2  TMA producer (dz tiles, one per frame tile)
//   warp 1  MMA issuer                        warps 4-7  converters, then the final epilogue
#include "tc_common.cuh"

namespace istgcn {
namespace tc {

constexpr int kThreadsDw2 = 256;
constexpr int kDwSlot = 32, kDwFr = 4;
constexpr int kDwXStage = kDwFr * kDwSlot * 128;         // 16 KB: one x slice of a tile
constexpr int kDwXP = 4 * kAtomBytes;                    // 64 KB: X' operand, 4 partitions
constexpr int kDwAdjCol = 0, kDwD1Col = 128, kDwAccCol = 256;

struct Dw2Params {
    const float* vals;
    const int *lptr, *lsrc, *lid;
    float* dWc;                       // [K*Cin][Cout]
    int frames, V, K, Cin, Cout, tiles, nrb;
};

struct Dw2Layout {                    // runtime layout (depends on Cout)
    int nx, ndz, nxp, dz_bytes;
    int x_off, dz_off, xp_off, bar_off, total;
    __host__ __device__ Dw2Layout(int Cout) {
        dz_bytes = (Cout / 32) * kAtomBytes;
        nxp = Cout <= 64 ? 2 : 1;
        ndz = 1;
        nx = 2;
        x_off = 0;
        dz_off = x_off + nx * kDwXStage;
        xp_off = dz_off + ndz * dz_bytes;
        bar_off = xp_off + nxp * kDwXP;
        total = bar_off + 32 * 8 + 16;
    }
};

__global__ void __launch_bounds__(kThreadsDw2, 1)
gcn_tc_dw2_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap dzmap,
                  Dw2Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Dw2Layout L(p.Cout);
    uint8_t* Xs = smem + L.x_off;
    uint8_t* DZ = smem + L.dz_off;
    uint8_t* XP = smem + L.xp_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* x_full = bars;              // [2]
    uint64_t* x_empty = bars + 2;         // [2]
    uint64_t* dz_full = bars + 4;         // [1]
    uint64_t* dz_empty = bars + 5;        // [1]
    uint64_t* d1_full = bars + 6;
    uint64_t* d1_empty = bars + 7;
    uint64_t* xp_full = bars + 8;         // [2]
    uint64_t* xp_empty = bars + 10;       // [2]
    uint64_t* done = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = p.V, K = p.K, Cin = p.Cin, Cout = p.Cout;
    const int nchunk = Cin / 32, natom = Cout / 32;
    const int rb0 = blockIdx.y * p.nrb;                       // first x slice (= row block) of this CTA
    const int nrb = min(p.nrb, nchunk - rb0);
    const int my_tiles = p.tiles > (int)blockIdx.x
                             ? (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t total = (uint32_t)my_tiles * nrb;

    // ---- setup: barriers, TMEM, dense adjacency -> tensor memory (scratch in the XP area)
    float* adjT = reinterpret_cast<float*>(XP);
    for (int i = tid; i < 4 * 32 * 33; i += kThreadsDw2) adjT[i] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1);
            mbar_init(&xp_full[i], 4); mbar_init(&xp_empty[i], 1);
        }
        mbar_init(dz_full, 1); mbar_init(dz_empty, 1);
        mbar_init(d1_full, 1); mbar_init(d1_empty, 4);
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) tma_prefetch_desc(&xmap);
    if (warp == 2 && lane == 0) tma_prefetch_desc(&dzmap);
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    for (int i = tid; i < K * V; i += kThreadsDw2) {
        const int k = i / V, w = i - k * V;
        for (int j = p.lptr[i]; j < p.lptr[i + 1]; ++j)
            adjT[(k * 32 + w) * 33 + p.lsrc[j]] += __uint_as_float(to_tf32(p.vals[p.lid[j]]));
    }
    __syncthreads();
    if (warp >= 4) {
        for (int k = 0; k < 4; ++k) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = k < K ? adjT[(k * 32 + lane) * 33 + j] : 0.f;
            tmem_st32(tmem_base + (static_cast<uint32_t>((warp - 4) * 32) << 16) + kDwAdjCol + k * 32, v);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // pad rows / unused partitions must be finite zeros: clear every operand buffer once
    for (int i = tid; i < L.bar_off / 16; i += kThreadsDw2)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    __syncthreads();

    if (warp == 0) {
        // =========================== TMA producer: x slices (rb0 .. rb0+nrb-1) of every tile
        uint32_t it = 0;
        const uint32_t bytes = kDwFr * V * 128;
        for (int t = 0; t < my_tiles; ++t) {
            const int f0 = (blockIdx.x + t * gridDim.x) * kDwFr;
            for (int rb = 0; rb < nrb; ++rb, ++it) {
                const int xs = it & 1;
                mbar_wait(&x_empty[xs], ((it >> 1) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&x_full[xs], bytes);
#pragma unroll
                    for (int f = 0; f < kDwFr; ++f)
                        tma_load_3d(Xs + xs * kDwXStage + f * (kDwSlot * 128), &xmap, &x_full[xs],
                                    (rb0 + rb) * 32, 0, f0 + f);
                }
                __syncwarp();
            }
        }
    } else if (warp == 2) {
        // =========================== TMA producer: the dz tile [Cout/32 atoms][128 rows][128 B]
        const uint32_t bytes = kDwFr * V * 128 * natom;
        for (int t = 0; t < my_tiles; ++t) {
            const int f0 = (blockIdx.x + t * gridDim.x) * kDwFr;
            mbar_wait(dz_empty, (t & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(dz_full, bytes);
                for (int a = 0; a < natom; ++a)
#pragma unroll
                    for (int f = 0; f < kDwFr; ++f)
                        tma_load_3d(DZ + a * kAtomBytes + f * (kDwSlot * 128), &dzmap, dz_full, a * 32, 0,
                                    f0 + f);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // =========================== MMA issuer
        if (total > 0) {
            constexpr uint32_t idesc1 = make_idesc(128, 32, false, true);
            const uint32_t idesc2 = make_idesc(128, Cout, true, true);
            const uint32_t xs0 = smem_u32(Xs), dz0 = smem_u32(DZ), xp0 = smem_u32(XP);
            const uint32_t adj = tmem_base + kDwAdjCol, d1 = tmem_base + kDwD1Col;
            auto issue1 = [&](uint32_t s) {
                const int xs = s & 1;
                mbar_wait(&x_full[xs], (s >> 1) & 1);
                mbar_wait(d1_empty, (s & 1) ^ 1);           // converters have read slice s-1
                tc_fence_after();
                if (elect_one()) {
                    // partition -> frame -> K-step: long same-accumulator runs (see gcn_tc2.cu)
                    for (int k = 0; k < K; ++k)
#pragma unroll
                        for (int f = 0; f < kDwFr; ++f) {
                            const uint32_t b_addr = xs0 + xs * kDwXStage + f * (kDwSlot * 128);
                            const uint32_t m0 = f == 0 ? 0u : ~0u, m1 = f == 1 ? 0u : ~0u,
                                           m2 = f == 2 ? 0u : ~0u, m3 = f == 3 ? 0u : ~0u;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                tc_mma_tf32_ts_masked(d1 + k * 32, adj + k * 32 + ks * 8,
                                                      make_desc(b_addr + ks * 1024, 4096, 512, 1), idesc1,
                                                      ks ? 1u : 0u, m0, m1, m2, m3);
                        }
                    tc_commit(&x_empty[xs]);
                    tc_commit(d1_full);
                }
                __syncwarp();
            };
            issue1(0);
            for (uint32_t s = 0; s < total; ++s) {
                if (s + 1 < total) issue1(s + 1);
                const uint32_t t = s / nrb, rb = s - t * nrb;
                const uint32_t xb = L.nxp == 2 ? (s & 1) : 0, xuse = L.nxp == 2 ? (s >> 1) : s;
                if (rb == 0) mbar_wait(dz_full, t & 1);
                mbar_wait(&xp_full[xb], xuse & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = xp0 + xb * kDwXP;
                    const uint32_t acc = tmem_base + kDwAccCol + rb * Cout;
#pragma unroll 4
                    for (int ks = 0; ks < 16; ++ks)             // 8 rows of the tile per step
                        tc_mma_tf32(acc, make_desc(a_addr + ks * 1024, kAtomBytes, 512, 1),
                                    make_desc(dz0 + ks * 1024, kAtomBytes, 512, 1), idesc2,
                                    (t == 0 && ks == 0) ? 0u : 1u);
                    tc_commit(&xp_empty[xb]);
                    if (rb == (uint32_t)nrb - 1) tc_commit(dz_empty);
                    if (s == total - 1) tc_commit(done);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // =========================== converters: D1 (TMEM) -> XP (MN-major operand atoms)
        const int q = warp - 4;
        const int row = q * 32 + lane;
        for (uint32_t s = 0; s < total; ++s) {
            const uint32_t xb = L.nxp == 2 ? (s & 1) : 0, xuse = L.nxp == 2 ? (s >> 1) : s;
            mbar_wait(d1_full, s & 1);
            mbar_wait(&xp_empty[xb], (xuse & 1) ^ 1);
            tc_fence_after();
            float* xp = reinterpret_cast<float*>(XP + xb * kDwXP);
            for (int k = 0; k < K; ++k) {
                float v[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kDwD1Col + k * 32, v);
                float* dst = xp + k * (kAtomBytes / 4) + row * 32;
#pragma unroll
                for (int c = 0; c < 4; ++c) {                   // 32-byte chunk c -> chunk c ^ (row & 3)
                    float* d8 = dst + ((c ^ (row & 3)) << 3);
                    st4(d8, make_float4(v[8 * c], v[8 * c + 1], v[8 * c + 2], v[8 * c + 3]));
                    st4(d8 + 4, make_float4(v[8 * c + 4], v[8 * c + 5], v[8 * c + 6], v[8 * c + 7]));
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(d1_empty);
                mbar_arrive(&xp_full[xb]);
            }
        }
        // ---- final epilogue: accumulators -> dWc (fp32 atomics)
        if (total > 0) {
            mbar_wait(done, 0);
            tc_fence_after();
            const int k = q, cil = lane;
            for (int rb = 0; rb < nrb; ++rb) {
                const int ci = (rb0 + rb) * 32 + cil;
                for (int c0 = 0; c0 < Cout; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kDwAccCol + rb * Cout + c0, v);
                    if (k < K) {
                        float* dst = p.dWc + (size_t)(k * Cin + ci) * Cout + c0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Shapes the second-generation weight-gradient kernel takes (plain frame maps only).
bool gcn_tc_dw2_eligible(int V, int K, int Cin, int Cout) {
    return V >= 1 && V <= 32 && K >= 1 && K <= 4 && Cin % 32 == 0 && Cin >= 32 && Cout % 32 == 0 &&
           Cout >= 32 && Cout <= 256 && (256 / Cout) * 8 >= Cin / 32;     // at most eight CTA groups
}

int launch_gcn_tc_dw2(const float* dz, const float* x, const float* vals, const int* lptr,
                      const int* lsrc, const int* lid, float* dWc, int frames, int V, int K, int Cin,
                      int Cout, cudaStream_t st) {
    Dw2Params p{vals, lptr, lsrc, lid, dWc, frames, V, K, Cin, Cout, (frames + kDwFr - 1) / kDwFr, 0};
    const int nchunk = Cin / 32;
    p.nrb = 256 / Cout;
    if (p.nrb > nchunk) p.nrb = nchunk;
    const int groups = (nchunk + p.nrb - 1) / p.nrb;
    CUtensorMap xmap, dzmap;
    if (int e = encode_frame_slices(&xmap, x, frames, V, Cin)) return e;
    if (int e = encode_frame_slices(&dzmap, dz, frames, V, Cout)) return e;
    const Dw2Layout L(Cout);
    cudaFuncSetAttribute(gcn_tc_dw2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    int nx = num_sms() / groups;
    if (nx < 1) nx = 1;
    if (nx > p.tiles) nx = p.tiles;
    gcn_tc_dw2_kernel<<<dim3(nx, groups), kThreadsDw2, L.total, st>>>(xmap, dzmap, p);
    return finish_launch("gcn_tc_dw2");
}

}  // namespace tc
}  // namespace istgcn
