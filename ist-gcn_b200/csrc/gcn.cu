// Fused graph convolution (reference: net/utils/tgcn.py:76-89, net/utils/inceptionv2_gcn.py:64-89)
// -- "aggregate first" formulation:
//
//     X'[(f,w)][k*Cin+ci] = sum_v A_eff[k][v][w] * x[(f,v)][ci]        (sparse, CUDA cores, smem)
//     z [(f,w)][c]        = sum_{k,ci} X'[(f,w)][k*Cin+ci] * Wc[k*Cin+ci][c] + biasterm[w][c]
//
// so the K*Cout-wide intermediate of the reference (conv1x1 output, 11-31 MB per clip and layer)
// never exists: a CTA owns a tile of whole frames (floor(128/V) frames = 125 rows for NTU),
// stages 32-channel slices of x in shared memory, aggregates them per partition with the static
// non-zero list of A_eff, and feeds the result straight into the channel-mix GEMM.  The epilogue
// adds the aggregated bias, accumulates the BatchNorm statistics and stores z once.
//
// This file holds the mma.sync (TF32 / 3xTF32) engine; gcn_tc.cu holds the tcgen05 engine for the
// wide layers.  Algorithmic HBM bytes per row: (Cin + Cout) * 4 forward.
#include "common.cuh"

namespace istgcn {

constexpr int kLdA = 36;   // 32-wide operand slices, padded so fragment loads are conflict-free

__device__ __forceinline__ int frames_in_tile(int V) {
    const int f = kTileRows / V;
    return f > 8 ? 8 : f;
}

// xs[128][32] <- x rows map(row0 + r), r < valid_rows, channels [ci0, ci0+32) (zero beyond)
__device__ __forceinline__ void load_x_slice(float* xs, const float* __restrict__ x, long long row0,
                                             int valid_rows, int Cin, int ci0, int tid,
                                             const FrameMap& fm, int V) {
    if ((Cin & 3) == 0) {
        stage4<kTileRows * 8 / kThreads>(
            tid, 0,
            [&](int i) {
                const int r = i >> 3, c4 = (i & 7) * 4;
                const long long src = (r < valid_rows && ci0 + c4 < Cin) ? map_row(fm, row0 + r, V) : -1;
                return src >= 0 ? ld4(x + src * Cin + ci0 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            },
            [&](int i, const float4& v) { st4(xs + (i >> 3) * 32 + (i & 7) * 4, v); });
    } else {
        for (int i = tid; i < kTileRows * 32; i += kThreads) {
            const int r = i >> 5, c = i & 31;
            const long long src = (r < valid_rows && ci0 + c < Cin) ? map_row(fm, row0 + r, V) : -1;
            xs[i] = src >= 0 ? x[src * Cin + ci0 + c] : 0.f;
        }
    }
}

// As[(f*V+w)*lda + lane] = sum_j val_j * xs[(f*V+v_j)*32 + lane] for the entries of partition k
template <int FMAX>
__device__ __forceinline__ void aggregate_partition(float* As, int lda, const float* xs,
                                                    const int* s_ptr, const int* s_src,
                                                    const float* s_val, int k, int V, int F,
                                                    int warp, int lane) {
    for (int w = warp; w < V; w += kWarps) {
        const int beg = s_ptr[k * V + w], end = s_ptr[k * V + w + 1];
        float acc[FMAX];
#pragma unroll
        for (int f = 0; f < FMAX; ++f) acc[f] = 0.f;
        for (int j = beg; j < end; ++j) {
            const int v = s_src[j];
            const float a = s_val[j];
#pragma unroll
            for (int f = 0; f < FMAX; ++f)
                if (f < F) acc[f] = fmaf(a, xs[(f * V + v) * 32 + lane], acc[f]);
        }
#pragma unroll
        for (int f = 0; f < FMAX; ++f)
            if (f < F) As[(f * V + w) * lda + lane] = acc[f];
    }
}

// ------------------------------------------------------------------------------------ forward
struct GcnFwdParams {
    const float *x, *Wc, *biasterm, *vals, *add_rows;
    const int *dst_ptr, *dst_src, *dst_id;
    float* z;
    double *stat_sum, *stat_sumsq;
    int frames, V, K, Cin, Cout, nnz, tiles;
    FrameMap fm;
};

template <int NCOLS, bool PRECISE>
__global__ void __launch_bounds__(kThreads) gcn_fwd_kernel(GcnFwdParams p) {
    constexpr int LDB = NCOLS + 8;
    constexpr int NT = NCOLS / 16;
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                          // [128][32]
    float* As = xs + kTileRows * 32;           // [128][36]
    float* Bs = As + kTileRows * kLdA;         // [32][LDB]
    // BatchNorm partial sums in double: sum(x^2) - sum(x)^2/n cancels badly for channels whose
    // mean dwarfs their spread, and fp32 partials made the variance order-dependent at 1e-4
    double* s_sum = reinterpret_cast<double*>(Bs + 32 * LDB);   // [NCOLS]
    double* s_sq = s_sum + NCOLS;                                // [NCOLS]
    float* s_val = reinterpret_cast<float*>(s_sq + NCOLS);       // [nnz]
    int* s_src = reinterpret_cast<int*>(s_val + kMaxNnz);
    int* s_ptr = s_src + kMaxNnz;              // [K*V+1]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int warp_m = warp & 3, warp_n = warp >> 2;
    const int V = p.V, K = p.K, Cin = p.Cin, Cout = p.Cout;
    const int F = frames_in_tile(V);
    const int n0 = blockIdx.y * NCOLS;

    for (int i = tid; i < p.nnz; i += kThreads) {
        s_src[i] = p.dst_src[i];
        s_val[i] = p.vals[p.dst_id[i]];
    }
    for (int i = tid; i <= K * V; i += kThreads) s_ptr[i] = p.dst_ptr[i];
    for (int i = tid; i < kTileRows * kLdA; i += kThreads) As[i] = 0.f;
    for (int i = tid; i < 2 * NCOLS; i += kThreads) s_sum[i] = 0.0;
    __syncthreads();

    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const int f0 = tile * F;
        const int nf = min(F, p.frames - f0);
        const int valid_rows = nf * V;
        const long long row0 = (long long)f0 * V;
        float acc[2][NT][4];
        zero_acc<2, NT>(acc);

        for (int ci0 = 0; ci0 < Cin; ci0 += 32) {
            __syncthreads();                   // previous readers of xs / As / Bs are done
            load_x_slice(xs, p.x, row0, valid_rows, Cin, ci0, tid, p.fm, V);
            for (int k = 0; k < K; ++k) {
                if (k > 0) __syncthreads();    // mma of partition k-1 finished with As / Bs
                // Bs[i][n] = Wc[(k*Cin + ci0 + i)][n0 + n]
                stage4<32 * (NCOLS / 4) / kThreads>(
                    tid, 0,
                    [&](int i) {
                        const int r = i / (NCOLS / 4), c4 = (i % (NCOLS / 4)) * 4;
                        return ci0 + r < Cin ? ld4(p.Wc + (size_t)(k * Cin + ci0 + r) * Cout + n0 + c4)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                    },
                    [&](int i, const float4& v) {
                        st4(Bs + (i / (NCOLS / 4)) * LDB + (i % (NCOLS / 4)) * 4, v);
                    });
                if (k == 0) __syncthreads();   // xs visible before the first aggregation
                aggregate_partition<8>(As, kLdA, xs, s_ptr, s_src, s_val, k, V, F, warp, lane);
                __syncthreads();
                warp_mma<2, NT, false, false, PRECISE>(acc, As + warp_m * 32 * kLdA, kLdA,
                                                       Bs + warp_n * (NCOLS / 2), LDB, 32, lane);
            }
        }
        // epilogue: + biasterm, store, BatchNorm statistics
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int cl = warp_n * (NCOLS / 2) + nt * 8 + 2 * t;     // local column of c0
            double cs0 = 0.0, cs1 = 0.0, cq0 = 0.0, cq1 = 0.0;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = warp_m * 32 + mt * 16 + g + 8 * h;
                    if (r < valid_rows) {
                        const int w = r % V;
                        float2 b = make_float2(0.f, 0.f);
                        if (p.biasterm)
                            b = *reinterpret_cast<const float2*>(p.biasterm + (size_t)w * Cout + n0 + cl);
                        if (p.add_rows) {                       // partial sums of a previous tap
                            const float2 a = *reinterpret_cast<const float2*>(
                                p.add_rows + (row0 + r) * Cout + n0 + cl);
                            b.x += a.x; b.y += a.y;
                        }
                        const float v0 = acc[mt][nt][2 * h] + b.x;
                        const float v1 = acc[mt][nt][2 * h + 1] + b.y;
                        *reinterpret_cast<float2*>(p.z + (row0 + r) * Cout + n0 + cl) =
                            make_float2(v0, v1);
                        cs0 += v0; cs1 += v1; cq0 += (double)v0 * v0; cq1 += (double)v1 * v1;
                    }
                }
            }
            if (p.stat_sum) {
                cs0 = group_sum_g(cs0); cs1 = group_sum_g(cs1);
                cq0 = group_sum_g(cq0); cq1 = group_sum_g(cq1);
                if (g == 0) {
                    atomicAdd(&s_sum[cl], cs0); atomicAdd(&s_sum[cl + 1], cs1);
                    atomicAdd(&s_sq[cl], cq0); atomicAdd(&s_sq[cl + 1], cq1);
                }
            }
        }
    }
    if (p.stat_sum) {
        __syncthreads();
        for (int c = tid; c < NCOLS; c += kThreads) {
            atomicAdd(&p.stat_sum[n0 + c], s_sum[c]);
            atomicAdd(&p.stat_sumsq[n0 + c], s_sq[c]);
        }
    }
}

template <int NCOLS>
static size_t gcn_fwd_smem() {
    return sizeof(float) * (kTileRows * 32 + kTileRows * kLdA + 32 * (NCOLS + 8) + 4 * NCOLS +
                            kMaxNnz) +
           sizeof(int) * (kMaxNnz + kMaxKV + 4);
}

template <int NCOLS, bool PRECISE>
static int launch_gcn_fwd(const GcnFwdParams& p, cudaStream_t s) {
    const size_t smem = gcn_fwd_smem<NCOLS>();
    auto kern = gcn_fwd_kernel<NCOLS, PRECISE>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ny = p.Cout / NCOLS;
    int nx = num_sms() * 3 / ny;
    if (nx < 1) nx = 1;
    if (nx > p.tiles) nx = p.tiles;
    kern<<<dim3(nx, ny), kThreads, smem, s>>>(p);
    return finish_launch("gcn_fwd");
}

// --------------------------------------------------------------------------- backward: input
constexpr int kLdG = 132;   // Gs[128][4*32 + 4]

// BatchNorm backward folded into the load of the upstream gradient:
// dz[r][c] = p[c] * ((g[r][c] - m1[c]) - cc[c] * (z[r][c] - mu[c]))     (p == NULL: dz = g)
struct BnBack {
    const float *p, *m1, *cc, *mu;
};
__device__ __forceinline__ float4 make_dz(const float* __restrict__ g, const float* __restrict__ z,
                                          const BnBack& b, long long off, int c) {
    float4 gv = ld4(g + off);
    if (b.p) {
        const float4 zv = ld4(z + off), pv = ld4(b.p + c), mv = ld4(b.m1 + c), cv = ld4(b.cc + c),
                     uv = ld4(b.mu + c);
        gv.x = bn_back(gv.x, zv.x, pv.x, mv.x, cv.x, uv.x);
        gv.y = bn_back(gv.y, zv.y, pv.y, mv.y, cv.y, uv.y);
        gv.z = bn_back(gv.z, zv.z, pv.z, mv.z, cv.z, uv.z);
        gv.w = bn_back(gv.w, zv.w, pv.w, mv.w, cv.w, uv.w);
    }
    return gv;
}

struct GcnBwdXParams {
    const float *g, *z;
    BnBack bn;
    const float *x, *Wc, *vals, *add_in;
    const int *src_ptr, *src_kw, *src_id;
    float *gin, *dvals;
    int frames, V, K, Cin, Cout, nnz, tiles;
    FrameMap fm;
};

template <bool PRECISE>
__global__ void __launch_bounds__(kThreads) gcn_bwd_x_kernel(GcnBwdXParams p) {
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                           // [128][32]
    float* DZs = xs + kTileRows * 32;           // [128][36]
    float* Wt = DZs + kTileRows * kLdA;         // [128 = (k, ci)][36]
    float* Gs = Wt + kTileRows * kLdA;          // [128][132]
    float* s_val = Gs + kTileRows * kLdG;       // [nnz] source order
    int* s_goff = reinterpret_cast<int*>(s_val + kMaxNnz);   // w*kLdG + k*32
    int* s_v = s_goff + kMaxNnz;                // source joint of entry
    int* s_id = s_v + kMaxNnz;
    int* s_ptr = s_id + kMaxNnz;                // [V+1]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int warp_m = warp & 3, warp_n = warp >> 2;
    const int V = p.V, K = p.K, Cin = p.Cin, Cout = p.Cout;
    const int F = frames_in_tile(V);

    for (int i = tid; i <= V; i += kThreads) s_ptr[i] = p.src_ptr[i];
    __syncthreads();
    for (int i = tid; i < p.nnz; i += kThreads) {
        const int kw = p.src_kw[i], id = p.src_id[i];
        s_goff[i] = (kw % V) * kLdG + (kw / V) * 32;
        s_val[i] = p.vals[id];
        s_id[i] = id;
        int v = 0;                               // source joint: the group that contains i
        while (s_ptr[v + 1] <= i) ++v;
        s_v[i] = v;
    }
    float dval_acc[kMaxNnz / kThreads];
#pragma unroll
    for (int i = 0; i < kMaxNnz / kThreads; ++i) dval_acc[i] = 0.f;
    __syncthreads();

    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const int f0 = tile * F;
        const int nf = min(F, p.frames - f0);
        const int valid_rows = nf * V;
        const long long row0 = (long long)f0 * V;

        for (int ci0 = 0; ci0 < Cin; ci0 += 32) {
            float acc[2][8][4];
            zero_acc<2, 8>(acc);
            for (int c0 = 0; c0 < Cout; c0 += 32) {
                __syncthreads();
                // DZs[r][c] for c in [c0, c0+32)
                stage4<kTileRows * 8 / kThreads>(
                    tid, 0,
                    [&](int i) {
                        const int r = i >> 3, c4 = (i & 7) * 4;
                        return r < valid_rows
                                   ? make_dz(p.g, p.z, p.bn, (row0 + r) * Cout + c0 + c4, c0 + c4)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                    },
                    [&](int i, const float4& v) { st4(DZs + (i >> 3) * kLdA + (i & 7) * 4, v); });
                // Wt[k*32 + i][c] = Wc[(k*Cin + ci0 + i)][c0 + c]
                stage4<kTileRows * 8 / kThreads>(
                    tid, 0,
                    [&](int i) {
                        const int r = i >> 3, c4 = (i & 7) * 4;
                        const int k = r >> 5, ci = ci0 + (r & 31);
                        return (k < K && ci < Cin) ? ld4(p.Wc + (size_t)(k * Cin + ci) * Cout + c0 + c4)
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                    },
                    [&](int i, const float4& v) { st4(Wt + (i >> 3) * kLdA + (i & 7) * 4, v); });
                __syncthreads();
                warp_mma<2, 8, false, true, PRECISE>(acc, DZs + warp_m * 32 * kLdA, kLdA,
                                                     Wt + warp_n * 64 * kLdA, kLdA, 32, lane);
            }
            // G tile -> shared memory, x slice next to it
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = warp_m * 32 + mt * 16 + g + 8 * h;
                        const int c = warp_n * 64 + nt * 8 + 2 * t;
                        *reinterpret_cast<float2*>(Gs + r * kLdG + c) =
                            make_float2(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1]);
                    }
            load_x_slice(xs, p.x, row0, valid_rows, Cin, ci0, tid, p.fm, V);
            __syncthreads();
            // (a) gin[(f,v)][ci] = sum_j val_j * G[(f,w_j)][k_j*32 + ci]  (+ add_in)
            if (p.gin && ci0 + lane < Cin) {
                for (int r = warp; r < valid_rows; r += kWarps) {
                    const int f = r / V, v = r - f * V;
                    const float* grow = Gs + f * V * kLdG + lane;
                    float a = 0.f;
                    for (int j = s_ptr[v]; j < s_ptr[v + 1]; ++j) a = fmaf(s_val[j], grow[s_goff[j]], a);
                    const long long orow = map_row(p.fm, row0 + r, V);
                    if (orow < 0) continue;                    // tap falls into the zero padding
                    const long long o = orow * Cin + ci0 + lane;
                    if (p.add_in) a += p.add_in[o];
                    p.gin[o] = a;
                }
            }
            // (b) dvals[j] += sum_{f,ci} x[(f,v_j)][ci] * G[(f,w_j)][k_j*32 + ci]
            if (p.dvals) {
#pragma unroll
                for (int jj = 0; jj < kMaxNnz / kThreads; ++jj) {
                    const int j = tid + jj * kThreads;
                    if (j < p.nnz) {
                        const float* xr = xs + s_v[j] * 32;
                        const float* gr = Gs + s_goff[j];
                        float a = 0.f;
                        for (int f = 0; f < nf; ++f) {
#pragma unroll 8
                            for (int i = 0; i < 32; ++i) {
                                const int c = (i + lane) & 31;
                                a = fmaf(xr[f * V * 32 + c], gr[f * V * kLdG + c], a);
                            }
                        }
                        dval_acc[jj] += a;
                    }
                }
            }
        }
    }
    if (p.dvals) {
#pragma unroll
        for (int jj = 0; jj < kMaxNnz / kThreads; ++jj) {
            const int j = tid + jj * kThreads;
            if (j < p.nnz) atomicAdd(&p.dvals[s_id[j]], dval_acc[jj]);
        }
    }
}

static size_t gcn_bwd_x_smem() {
    return sizeof(float) * (kTileRows * 32 + 2 * kTileRows * kLdA + kTileRows * kLdG + kMaxNnz) +
           sizeof(int) * (3 * kMaxNnz + 64 + 4);
}

// -------------------------------------------------------------------------- backward: weights
struct GcnBwdWParams {
    const float *g, *z;
    BnBack bn;
    const float *x, *vals;
    const int *dst_ptr, *dst_src, *dst_id;
    float *dWc, *dbiasterm;
    int frames, V, K, Cin, Cout, nnz, tiles, mblocks, nblocks;
    FrameMap fm;
};

// CTA = one (partition k, MB input channels, NB output channels) block of dWc, looping over a
// strided share of the frame tiles with the accumulators in registers (split-K over rows).
template <int MB, int NB, bool PRECISE>
__global__ void __launch_bounds__(kThreads) gcn_bwd_w_kernel(GcnBwdWParams p) {
    constexpr int WM = MB >= 64 ? 4 : 2, WN = kWarps / WM;
    constexpr int MT = MB / (16 * WM), NT = NB / (8 * WN);
    constexpr int LDX = MB + 8, LDZ = NB + 8;
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                           // [MB/32][128][32]
    float* Xs = xs + kTileRows * MB;            // [128][LDX]   aggregated X' for partition k
    float* DZs = Xs + kTileRows * LDX;          // [128][LDZ]
    float* s_db = DZs + kTileRows * LDZ;        // [V][NB] (only used by the k==0, mblock==0 CTAs)
    float* s_val = s_db + 32 * NB;
    int* s_src = reinterpret_cast<int*>(s_val + kMaxNnz);
    int* s_ptr = s_src + kMaxNnz;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int warp_m = warp % WM, warp_n = warp / WM;
    const int V = p.V, K = p.K, Cin = p.Cin, Cout = p.Cout;
    const int F = frames_in_tile(V);
    int b = blockIdx.x;
    const int nblk = b % p.nblocks; b /= p.nblocks;
    const int mblk = b % p.mblocks; b /= p.mblocks;
    const int k = b;
    const int ci_base = mblk * MB, n0 = nblk * NB;
    const bool do_bias = p.dbiasterm != nullptr && k == 0 && mblk == 0;

    for (int i = tid; i < p.nnz; i += kThreads) {
        s_src[i] = p.dst_src[i];
        s_val[i] = p.vals[p.dst_id[i]];
    }
    for (int i = tid; i <= K * V; i += kThreads) s_ptr[i] = p.dst_ptr[i];
    for (int i = tid; i < kTileRows * LDX; i += kThreads) Xs[i] = 0.f;
    for (int i = tid; i < 32 * NB; i += kThreads) s_db[i] = 0.f;
    float acc[MT][NT][4];
    zero_acc<MT, NT>(acc);
    __syncthreads();

    for (int tile = blockIdx.y; tile < p.tiles; tile += gridDim.y) {
        const int f0 = tile * F;
        const int nf = min(F, p.frames - f0);
        const int valid_rows = nf * V;
        const long long row0 = (long long)f0 * V;
        __syncthreads();                         // previous tile's mma is done with Xs / DZs
        for (int sl = 0; sl < MB / 32; ++sl)
            load_x_slice(xs + sl * kTileRows * 32, p.x, row0, valid_rows, Cin, ci_base + sl * 32, tid,
                         p.fm, V);
#pragma unroll 1
        for (int base = 0; base < kTileRows * (NB / 4); base += 8 * kThreads)
            stage4<8>(
                tid, base,
                [&](int i) {
                    const int r = i / (NB / 4), c4 = (i % (NB / 4)) * 4;
                    return r < valid_rows
                               ? make_dz(p.g, p.z, p.bn, (row0 + r) * Cout + n0 + c4, n0 + c4)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
                },
                [&](int i, const float4& v) { st4(DZs + (i / (NB / 4)) * LDZ + (i % (NB / 4)) * 4, v); });
        __syncthreads();
        for (int sl = 0; sl < MB / 32; ++sl)
            aggregate_partition<8>(Xs + sl * 32, LDX, xs + sl * kTileRows * 32, s_ptr, s_src, s_val,
                                   k, V, F, warp, lane);
        if (do_bias) {
            for (int i = tid; i < V * NB; i += kThreads) {
                const int w = i / NB, c = i - w * NB;
                float a = 0.f;
                for (int f = 0; f < nf; ++f) a += DZs[(f * V + w) * LDZ + c];
                s_db[i] += a;
            }
        }
        __syncthreads();
        // dW block (MB x NB) += Xs^T (MB x 128) * DZs (128 x NB)
        warp_mma<MT, NT, true, false, PRECISE>(acc, Xs + warp_m * (MT * 16), LDX,
                                               DZs + warp_n * (NT * 8), LDZ, kTileRows, lane);
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ci = ci_base + warp_m * (MT * 16) + mt * 16 + g + 8 * (i >> 1);
                const int c = n0 + warp_n * (NT * 8) + nt * 8 + 2 * t + (i & 1);
                if (ci < Cin) atomicAdd(&p.dWc[(size_t)(k * Cin + ci) * Cout + c], acc[mt][nt][i]);
            }
    if (do_bias) {
        __syncthreads();
        for (int i = tid; i < V * NB; i += kThreads) {
            const int w = i / NB, c = i - w * NB;
            atomicAdd(&p.dbiasterm[(size_t)w * Cout + n0 + c], s_db[i]);
        }
    }
}

template <int MB, int NB>
static size_t gcn_bwd_w_smem() {
    return sizeof(float) * (kTileRows * MB + kTileRows * (MB + 8) + kTileRows * (NB + 8) + 32 * NB +
                            kMaxNnz) +
           sizeof(int) * (kMaxNnz + kMaxKV + 4);
}

template <int MB, int NB, bool PRECISE>
static int launch_gcn_bwd_w(GcnBwdWParams p, cudaStream_t s) {
    const size_t smem = gcn_bwd_w_smem<MB, NB>();
    auto kern = gcn_bwd_w_kernel<MB, NB, PRECISE>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    p.mblocks = (p.Cin + MB - 1) / MB;
    p.nblocks = p.Cout / NB;
    const int combos = p.K * p.mblocks * p.nblocks;
    const int per_sm = smem > 110 * 1024 ? 1 : 2;
    int splits = (num_sms() * per_sm + combos - 1) / combos;
    if (splits < 1) splits = 1;
    if (splits > p.tiles) splits = p.tiles;
    kern<<<dim3(combos, splits), kThreads, smem, s>>>(p);
    return finish_launch("gcn_bwd_w");
}

static int check_gcn_dims(const char* who, int frames, int V, int K, int Cin, int Cout, int nnz) {
    ISTGCN_REQUIRE(frames >= 0 && V >= 1 && V <= 32, ISTGCN_E_SHAPE, "%s: V=%d unsupported (1..32)", who, V);
    ISTGCN_REQUIRE(K >= 1 && K <= 4, ISTGCN_E_SHAPE, "%s: K=%d unsupported (1..4)", who, K);
    ISTGCN_REQUIRE(Cin >= 1 && (Cin <= 32 || Cin % 32 == 0), ISTGCN_E_SHAPE,
                   "%s: Cin=%d unsupported (<=32 or a multiple of 32)", who, Cin);
    ISTGCN_REQUIRE(Cout >= 64 && Cout % 64 == 0, ISTGCN_E_SHAPE,
                   "%s: Cout=%d unsupported (multiple of 64)", who, Cout);
    ISTGCN_REQUIRE(nnz >= 0 && nnz <= kMaxNnz, ISTGCN_E_SHAPE, "%s: nnz=%d exceeds %d", who, nnz, kMaxNnz);
    return 0;
}

}  // namespace istgcn

using namespace istgcn;

ISTGCN_API int istgcn_gcn_fwd(const float* x, const float* Wc, const float* biasterm,
                              const float* vals, const int* dst_ptr, const int* dst_src,
                              const int* dst_id, int nnz, const float* add_rows, float* z, double* stat_sum,
                              double* stat_sumsq, int frames, int V, int K, int Cin, int Cout,
                              int t_in, int t_out, int t_stride, int t_offset, int math, istgcn_stream_t s) {
    ISTGCN_REQUIRE(x && Wc && vals && dst_ptr && dst_src && dst_id && z, ISTGCN_E_ARG,
                   "gcn_fwd: null pointer");
    ISTGCN_REQUIRE((stat_sum == nullptr) == (stat_sumsq == nullptr), ISTGCN_E_ARG,
                   "gcn_fwd: pass both statistics buffers or neither");
    if (int e = check_gcn_dims("gcn_fwd", frames, V, K, Cin, Cout, nnz)) return e;
    if (frames == 0) return 0;
    GcnFwdParams p{x, Wc, biasterm, vals, add_rows, dst_ptr, dst_src, dst_id, z, stat_sum, stat_sumsq,
                   frames, V, K, Cin, Cout, nnz, 0, {t_in, t_out, t_stride, t_offset}};
    const int F = kTileRows / V > 8 ? 8 : kTileRows / V;
    p.tiles = (frames + F - 1) / F;
    cudaStream_t st = (cudaStream_t)s;
    const bool precise = math == ISTGCN_MATH_3XTF32;
    if (Cout % 128 == 0)
        return precise ? launch_gcn_fwd<128, true>(p, st) : launch_gcn_fwd<128, false>(p, st);
    return precise ? launch_gcn_fwd<64, true>(p, st) : launch_gcn_fwd<64, false>(p, st);
}

ISTGCN_API int istgcn_gcn_bwd_x(const float* g, const float* z, const float* bn_p,
                                const float* bn_m1, const float* bn_c, const float* bn_mu,
                                const float* x, const float* Wc, const float* vals,
                                const int* src_ptr, const int* src_kw, const int* src_id, int nnz,
                                const float* add_in, float* gin, float* dvals, int frames, int V,
                                int K, int Cin, int Cout, int t_in, int t_out, int t_stride, int t_offset,
                                int math, istgcn_stream_t s) {
    ISTGCN_REQUIRE(g && x && Wc && vals && src_ptr && src_kw && src_id && (gin || dvals), ISTGCN_E_ARG,
                   "gcn_bwd_x: null pointer");
    ISTGCN_REQUIRE(bn_p == nullptr || (z && bn_m1 && bn_c && bn_mu), ISTGCN_E_ARG,
                   "gcn_bwd_x: bn_p needs z, bn_m1, bn_c and bn_mu");
    if (int e = check_gcn_dims("gcn_bwd_x", frames, V, K, Cin, Cout, nnz)) return e;
    if (frames == 0) return 0;
    GcnBwdXParams pr{g, z, {bn_p, bn_m1, bn_c, bn_mu}, x, Wc, vals, add_in, src_ptr, src_kw, src_id, gin, dvals,
                     frames, V, K, Cin, Cout, nnz, 0, {t_in, t_out, t_stride, t_offset}};
    const int F = kTileRows / V > 8 ? 8 : kTileRows / V;
    pr.tiles = (frames + F - 1) / F;
    const size_t smem = gcn_bwd_x_smem();
    int nx = num_sms();
    if (nx > pr.tiles) nx = pr.tiles;
    if (math == ISTGCN_MATH_3XTF32) {
        cudaFuncSetAttribute(gcn_bwd_x_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        gcn_bwd_x_kernel<true><<<nx, kThreads, smem, (cudaStream_t)s>>>(pr);
    } else {
        cudaFuncSetAttribute(gcn_bwd_x_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        gcn_bwd_x_kernel<false><<<nx, kThreads, smem, (cudaStream_t)s>>>(pr);
    }
    return finish_launch("gcn_bwd_x");
}

ISTGCN_API int istgcn_gcn_bwd_w(const float* g, const float* z, const float* bn_p,
                                const float* bn_m1, const float* bn_c, const float* bn_mu,
                                const float* x, const float* vals,
                                const int* dst_ptr, const int* dst_src, const int* dst_id, int nnz,
                                float* dWc, float* dbiasterm, int frames, int V, int K, int Cin,
                                int Cout, int t_in, int t_out, int t_stride, int t_offset, int math,
                                istgcn_stream_t s) {
    ISTGCN_REQUIRE(g && x && vals && dst_ptr && dst_src && dst_id && dWc, ISTGCN_E_ARG,
                   "gcn_bwd_w: null pointer");
    ISTGCN_REQUIRE(bn_p == nullptr || (z && bn_m1 && bn_c && bn_mu), ISTGCN_E_ARG,
                   "gcn_bwd_w: bn_p needs z, bn_m1, bn_c and bn_mu");
    if (int e = check_gcn_dims("gcn_bwd_w", frames, V, K, Cin, Cout, nnz)) return e;
    if (frames == 0) return 0;
    GcnBwdWParams pr{g, z, {bn_p, bn_m1, bn_c, bn_mu}, x, vals, dst_ptr, dst_src, dst_id, dWc, dbiasterm,
                     frames, V, K, Cin, Cout, nnz, 0, 0, 0, {t_in, t_out, t_stride, t_offset}};
    const int F = kTileRows / V > 8 ? 8 : kTileRows / V;
    pr.tiles = (frames + F - 1) / F;
    cudaStream_t st = (cudaStream_t)s;
    const bool pc = math == ISTGCN_MATH_3XTF32;
    if (Cin <= 32) {
        if (Cout % 128 == 0)
            return pc ? launch_gcn_bwd_w<32, 128, true>(pr, st) : launch_gcn_bwd_w<32, 128, false>(pr, st);
        return pc ? launch_gcn_bwd_w<32, 64, true>(pr, st) : launch_gcn_bwd_w<32, 64, false>(pr, st);
    }
    if (Cin % 128 == 0 && Cout % 128 == 0)
        return pc ? launch_gcn_bwd_w<128, 128, true>(pr, st) : launch_gcn_bwd_w<128, 128, false>(pr, st);
    if (Cout % 128 == 0)
        return pc ? launch_gcn_bwd_w<64, 128, true>(pr, st) : launch_gcn_bwd_w<64, 128, false>(pr, st);
    if (Cin % 64 == 0)
        return pc ? launch_gcn_bwd_w<64, 64, true>(pr, st) : launch_gcn_bwd_w<64, 64, false>(pr, st);
    return pc ? launch_gcn_bwd_w<32, 64, true>(pr, st) : launch_gcn_bwd_w<32, 64, false>(pr, st);
}
