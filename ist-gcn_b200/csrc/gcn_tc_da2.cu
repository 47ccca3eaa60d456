// Second-generation tcgen05 kernel for the adjacency gradient of the fused graph convolution
// (reference forward: net/utils/tgcn.py:76-89):
//
//     dvals[(k,v,w)] = sum_f sum_ci x[(f,v)][ci] * G_k[(f,w)][ci],      G_k = dz * Wc_k^T
//
// Both contractions on the tensor core, nothing but the final read-out on CUDA cores:
//
//   MMA G   D_G[128 rows][kg*32] = DZ[128 rows][Cout] * Wc_{k,slice}[32 ci][Cout]^T   (SS, per c-atom)
//   MMA S   D_S_k[128 (f,w)][128 (f',v)] += D_G,k[128][32 ci] (A operand FROM TENSOR MEMORY)
//                                           * X[128 (f',v)][32 ci]^T                 (TS, N = 128)
//
// D_S_k accumulates over the channel slices AND over all frame tiles of the CTA (the tile-local
// frame index of a row is the same in every tile), so the kernel has no per-tile epilogue at all;
// the wanted entries are the diagonal blocks f = f', read out once at the end.  (The first-
// generation kernel spent 98 % of its time in per-tile CUDA-core dot products.)  Two partitions
// per CTA (2 x 128 + 64 TMEM columns), the partition pairs are spread over blockIdx.y.
//
//   warp 0  TMA producer: dz atom + Wc rows per stage     warp 2  TMA producer: x slices
//   warp 1  MMA issuer (tcgen05.mma executes in issue order: D_G -> A operand needs no barrier)
//   warps 4-7  final read-out: tcgen05.ld, diagonal-block entries -> shared atomics -> dvals
#include "tc_common.cuh"

namespace istgcn {
namespace tc {

constexpr int kThreadsDa2 = 256;
constexpr int kDa2Stages = 4;
constexpr int kDa2KG = 2;                                  // partitions per CTA
constexpr int kDa2StageBytes = kAtomBytes + kDa2KG * 32 * 128;
constexpr int kDa2GCol = 0, kDa2SCol = 128;

struct Da2Params {
    const int *lptr, *lsrc, *lid;
    float* dvals;
    int frames, V, K, Cin, Cout, nnz, tiles;
};

struct SmemDa2 {
    static constexpr int ring_off = 0;
    static constexpr int x_off = ring_off + kDa2Stages * kDa2StageBytes;
    static constexpr int buf_off = x_off + 2 * kAtomBytes;             // [4 warps][32][33] floats
    static constexpr int dv_off = buf_off + 4 * 32 * 33 * 4;
    static constexpr int bar_off = (dv_off + kMaxNnz * 4 + 7) / 8 * 8;
    static constexpr int kNumBars = 2 * kDa2Stages + 4 + 1;
    static constexpr int total = bar_off + kNumBars * 8 + 16;
};

__global__ void __launch_bounds__(kThreadsDa2, 1)
gcn_tc_da2_kernel(const __grid_constant__ CUtensorMap dzmap, const __grid_constant__ CUtensorMap wmap,
                  const __grid_constant__ CUtensorMap xmap, Da2Params p) {
    using L = SmemDa2;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* ring = smem + L::ring_off;
    uint8_t* Xs = smem + L::x_off;
    float* s_buf = reinterpret_cast<float*>(smem + L::buf_off);
    float* s_dv = reinterpret_cast<float*>(smem + L::dv_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bar_off);
    uint64_t* full = bars;
    uint64_t* empty = full + kDa2Stages;
    uint64_t* x_full = empty + kDa2Stages;
    uint64_t* x_empty = x_full + 2;
    uint64_t* done = x_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L::kNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = p.V, K = p.K, Cin = p.Cin, Cout = p.Cout;
    const int F = (kAtomRows / V) > 8 ? 8 : (kAtomRows / V);
    const int nchunk = Cin / 32, natom = Cout / 32;
    const int k0 = blockIdx.y * kDa2KG;
    const int kg = min(kDa2KG, K - k0);
    const int my_tiles = p.tiles > (int)blockIdx.x
                             ? (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    for (int i = tid; i < p.nnz; i += kThreadsDa2) s_dv[i] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < kDa2Stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&dzmap); tma_prefetch_desc(&wmap); }
    if (warp == 2 && lane == 0) tma_prefetch_desc(&xmap);
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =========================== TMA producer: dz atom [128 rows][32 c] + Wc rows of kg partitions
        uint32_t it = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int row0 = (blockIdx.x + t * gridDim.x) * F * V;
            for (int ch = 0; ch < nchunk; ++ch)
                for (int ca = 0; ca < natom; ++ca, ++it) {
                    const int s = it % kDa2Stages;
                    mbar_wait(&empty[s], ((it / kDa2Stages) & 1) ^ 1);
                    if (elect_one()) {
                        uint8_t* dst = ring + s * kDa2StageBytes;
                        mbar_arrive_expect_tx(&full[s], kAtomBytes + kg * 32 * 128);
                        tma_load_2d(dst, &dzmap, &full[s], ca * 32, row0);
                        for (int kk = 0; kk < kg; ++kk)
                            tma_load_2d(dst + kAtomBytes + kk * 32 * 128, &wmap, &full[s], ca * 32,
                                        (k0 + kk) * Cin + ch * 32);
                    }
                    __syncwarp();
                }
        }
    } else if (warp == 2) {
        // =========================== TMA producer: x slices [128 rows][32 ci] (K-major operand)
        uint32_t it = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int row0 = (blockIdx.x + t * gridDim.x) * F * V;
            for (int ch = 0; ch < nchunk; ++ch, ++it) {
                const int xb = it & 1;
                mbar_wait(&x_empty[xb], ((it >> 1) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&x_full[xb], kAtomBytes);
                    tma_load_2d(Xs + xb * kAtomBytes, &xmap, &x_full[xb], ch * 32, row0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer
        if (my_tiles > 0) {
            const uint32_t idesc_g = make_idesc(128, kg * 32, false, false);
            constexpr uint32_t idesc_s = make_idesc(128, 128, false, false);
            const uint32_t d_g = tmem_base + kDa2GCol, d_s = tmem_base + kDa2SCol;
            uint32_t it = 0, xit = 0;
            for (int t = 0; t < my_tiles; ++t)
                for (int ch = 0; ch < nchunk; ++ch, ++xit) {
                    for (int ca = 0; ca < natom; ++ca, ++it) {
                        const int s = it % kDa2Stages;
                        mbar_wait(&full[s], (it / kDa2Stages) & 1);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t a_addr = smem_u32(ring + s * kDa2StageBytes);
                            const uint32_t b_addr = a_addr + kAtomBytes;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                tc_mma_tf32(d_g, make_desc(a_addr + ks * 32, 16, 1024),
                                            make_desc(b_addr + ks * 32, 16, 1024), idesc_g,
                                            (ca | ks) ? 1u : 0u);
                            tc_commit(&empty[s]);
                        }
                        __syncwarp();
                    }
                    const int xb = xit & 1;
                    mbar_wait(&x_full[xb], (xit >> 1) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t x_addr = smem_u32(Xs + xb * kAtomBytes);
                        for (int kk = 0; kk < kg; ++kk)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                tc_mma_tf32_ts(d_s + kk * 128, d_g + kk * 32 + ks * 8,
                                               make_desc(x_addr + ks * 32, 16, 1024), idesc_s,
                                               (t | ch | ks) ? 1u : 0u);
                        tc_commit(&x_empty[xb]);
                        if (t == my_tiles - 1 && ch == nchunk - 1) tc_commit(done);
                    }
                    __syncwarp();
                }
        }
    } else if (warp >= 4) {
        // =========================== read-out: diagonal blocks of D_S -> dvals
        if (my_tiles > 0) {
            const int q = warp - 4;
            const int r = q * 32 + lane;
            const int f = r / V, w = r - f * V;
            const bool ok = r < F * V;
            float* buf = s_buf + (q * 32 + lane) * 33;
            mbar_wait(done, 0);
            tc_fence_after();
            for (int kk = 0; kk < kg; ++kk) {
                const int kw = (k0 + kk) * V + w;
                const int beg = ok ? p.lptr[kw] : 0, end = ok ? p.lptr[kw + 1] : 0;
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kDa2SCol + kk * 128 + c0, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) buf[j] = v[j];
                    for (int j = beg; j < end; ++j) {
                        const int col = f * V + p.lsrc[j] - c0;
                        if (col >= 0 && col < 32) atomicAdd(&s_dv[j], buf[col]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (my_tiles > 0)
        for (int j = tid; j < p.nnz; j += kThreadsDa2) {
            const float v = s_dv[j];
            if (v != 0.f) atomicAdd(&p.dvals[p.lid[j]], v);
        }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Cout <= 128 only: every 32-channel slice of x re-streams the whole dz tile (Cout/32 atoms) from
// L2, and at Cout = 256 that stream (measured 0.47 ms vs 0.34 ms) costs more than the
// first-generation kernel's CUDA-core dots.
bool gcn_tc_da2_eligible(int V, int K, int Cin, int Cout) {
    return V >= 1 && V <= 32 && K >= 1 && K <= 4 && Cin % 32 == 0 && Cin >= 32 && Cout % 32 == 0 &&
           Cout >= 32 && Cout <= 128;
}

int launch_gcn_tc_da2(const float* dz, const float* x, const float* Wc, const int* lptr, const int* lsrc,
                      const int* lid, int nnz, float* dvals, int frames, int V, int K, int Cin, int Cout,
                      cudaStream_t st) {
    Da2Params p{lptr, lsrc, lid, dvals, frames, V, K, Cin, Cout, nnz, 0};
    const int F = kTileRows / V > 8 ? 8 : kTileRows / V;
    p.tiles = (frames + F - 1) / F;
    CUtensorMap dzmap, wmap, xmap;
    if (int e = encode_tile_map(&dzmap, dz, (long long)frames * V, Cout, 128)) return e;
    if (int e = encode_tile_map(&wmap, Wc, (long long)K * Cin, Cout, 32)) return e;
    if (int e = encode_tile_map(&xmap, x, (long long)frames * V, Cin, 128)) return e;
    cudaFuncSetAttribute(gcn_tc_da2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemDa2::total);
    const int groups = (K + kDa2KG - 1) / kDa2KG;
    int nx = num_sms() / groups;
    if (nx < 1) nx = 1;
    if (nx > p.tiles) nx = p.tiles;
    gcn_tc_da2_kernel<<<dim3(nx, groups), kThreadsDa2, SmemDa2::total, st>>>(dzmap, wmap, xmap, p);
    return finish_launch("gcn_tc_dvals2");
}

}  // namespace tc
}  // namespace istgcn
