// Inception TCN with 1x1 bottlenecks (reference: net/st_gcn_mstcn_1x1.py:186-266):
//     a  = relu(BN1(z))                  tcn_start
//     h1 = a Wd + bd                     conv_1x1_start   C -> b = int(sqrt(C)) in {8, 11, 16}
//     h2 = sum_i imp[i] * tcn_i(h1)      3x1 / 9x1 / 15x1 temporal convs, stride s  == ONE 15-tap
//                                        conv with Weff = imp0*W3 (+6) + imp1*W9 (+3) + imp2*W15
//     u  = h2 Wu + bu                    conv_1x1_end     b -> C   (BN2 statistics in the epilogue)
// The block is ~13 FLOP/byte, i.e. HBM-bound: forward reads z once and writes u once; the
// b-channel intermediates h1/h2 (1/8..1/16 of an activation) are the only extra traffic and are
// kept for the backward pass.  The skinny GEMMs (N or M = b padded to bp in {8, 16}) run on the
// legacy mma.sync TF32 path - tcgen05 tiles (M=128, N>=16 from shared-memory descriptors) buy
// nothing at K=bp<=16 / N=bp<=16.
//
//   tcn_down    z -> h1                         rows x C x bp GEMM, BN1+ReLU applied on load
//   tcn_up      h1 -> h2 -> u                   15-tap implicit GEMM from a smem halo tile, then
//                                               rows x bp x C GEMM, BN2 sums
//   tcn_bwd_up  go,u -> du -> dh2, dWu, dbu, dbeff
//   tcn_bwd_t   dh2,h1 -> dh1, dWeff, dbd
//   tcn_bwd_dn  dh1,z -> g1 (ReLU-masked), BN1-backward sums, dWd
#include "common.cuh"

namespace istgcn {

constexpr int kTaps = 15, kHalf = 7;

__host__ __device__ inline int ld_g(int bp) { return bp + 4; }                 // g-indexed rows
__host__ __device__ inline int ld_t(int bp) { return bp == 8 ? 8 : 24; }       // t-indexed rows

// Register prefetch of one [128 rows x 32 channels] slice of a row-major [rows][C] matrix (4
// float4 per thread).  The slice for step i+1 is requested before the math of step i starts, so
// the DRAM latency hides behind the tensor-core work instead of sitting in front of it.
__device__ __forceinline__ void prefetch_slice(float4 (&v)[4], const float* __restrict__ base,
                                               long long row0, long long rows, int C, int c0,
                                               int tid) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = tid + u * kThreads;
        const long long r = row0 + (i >> 3);
        v[u] = r < rows ? ld4(base + r * C + c0 + (i & 7) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ----------------------------------------------------------------------------------- down
struct TcnDownParams {
    const float *z, *mean1, *scale1, *beta1, *Wd, *bd;
    float* h1;
    long long rows;
    int C, bp;
};

template <int NT, bool PRECISE>
__global__ void __launch_bounds__(kThreads) tcn_down_kernel(TcnDownParams p) {
    constexpr int BP = NT * 8;
    constexpr int LDW = BP == 8 ? 8 : 24;
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                       // [128][36]
    float* Ws = As + kTileRows * 36;        // [C][LDW]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int C = p.C;
    for (int i = tid; i < C * BP; i += kThreads) Ws[(i / BP) * LDW + (i % BP)] = p.Wd[i];
    const long long tiles = (p.rows + kTileRows - 1) / kTileRows;
    float4 zr[4];
    if (blockIdx.x < tiles) prefetch_slice(zr, p.z, (long long)blockIdx.x * kTileRows, p.rows, C, 0, tid);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * kTileRows;
        const int valid = (int)min((long long)kTileRows, p.rows - row0);
        float acc[1][NT][4];
        zero_acc<1, NT>(acc);
        for (int c0 = 0; c0 < C; c0 += 32) {
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = tid + u * kThreads;
                const int r = i >> 3, c4 = (i & 7) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < valid) {
                    const float4 mu = ld4(p.mean1 + c0 + c4), sc = ld4(p.scale1 + c0 + c4),
                                 be = ld4(p.beta1 + c0 + c4);
                    v.x = fmaxf(bn_apply(zr[u].x, mu.x, sc.x, be.x), 0.f);
                    v.y = fmaxf(bn_apply(zr[u].y, mu.y, sc.y, be.y), 0.f);
                    v.z = fmaxf(bn_apply(zr[u].z, mu.z, sc.z, be.z), 0.f);
                    v.w = fmaxf(bn_apply(zr[u].w, mu.w, sc.w, be.w), 0.f);
                }
                st4(As + r * 36 + c4, v);
            }
            {   // next slice: same tile, or the first slice of this CTA's next tile
                const bool same = c0 + 32 < C;
                const long long nrow0 = same ? row0 : (tile + gridDim.x) * kTileRows;
                if (same || tile + gridDim.x < tiles)
                    prefetch_slice(zr, p.z, nrow0, p.rows, C, same ? c0 + 32 : 0, tid);
            }
            __syncthreads();
            warp_mma<1, NT, false, false, PRECISE>(acc, As + warp * 16 * 36, 36, Ws + c0 * LDW, LDW,
                                                   32, lane);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = warp * 16 + g + 8 * h;
                const int c = nt * 8 + 2 * t;
                if (r < valid)
                    *reinterpret_cast<float2*>(p.h1 + (row0 + r) * BP + c) =
                        make_float2(acc[0][nt][2 * h] + p.bd[c], acc[0][nt][2 * h + 1] + p.bd[c + 1]);
            }
    }
}

// ----------------------------------------------------------------------------------- up
struct TcnUpParams {
    const float *h1, *Weff, *beff, *Wu, *bu;
    float *h2, *u;
    double *stat_sum, *stat_sumsq;
    int NM, T, Tout, V, C, bp, stride, TT, tiles_per_sample;
};

constexpr int kUpRows = 128;   // output rows (frames*V) of one temporal tile

template <int NT, bool PRECISE>
__global__ void __launch_bounds__(kThreads) tcn_up_kernel(TcnUpParams p) {
    constexpr int BP = NT * 8;
    constexpr int LDH = BP + 4;                 // halo / h2 tiles (g-indexed)
    constexpr int LDW = BP == 8 ? 8 : 24;       // Weff[(tap, ci)][co]  (t-indexed)
    extern __shared__ __align__(16) float smem[];
    const int V = p.V, C = p.C, s = p.stride, TT = p.TT;
    const int TI = (TT - 1) * s + kTaps;        // input frames of the halo tile
    const int LDU = C + 8;
    float* hs = smem;                           // [TI*V][LDH]
    float* h2s = hs + TI * V * LDH;             // [256][LDH]
    float* Wts = h2s + kUpRows * LDH;           // [15*BP][LDW]
    float* Wus = Wts + kTaps * BP * LDW;        // [BP][LDU]
    // BatchNorm sums: fp32 column sums of each warp's 16 rows -> s_col[warp][C][2], then thread c
    // adds the (at most 8) partials of channel c to its DOUBLE register accumulators once per
    // tile (the one-pass variance cancels in fp32, see gcn.cu; no shared-memory atomics: the
    // emulated fp64 atomics were ~40 % of this kernel's instructions)
    float* s_col = Wus + BP * LDU;                               // [8][C][2]
    float* bus = s_col + 16 * C;                                 // [C] conv_1x1_end bias
    double acc_s = 0.0, acc_q = 0.0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    for (int i = tid; i < kTaps * BP * BP; i += kThreads) Wts[(i / BP) * LDW + (i % BP)] = p.Weff[i];
    for (int i = tid; i < BP * C; i += kThreads) Wus[(i / C) * LDU + (i % C)] = p.Wu[i];
    for (int i = tid; i < C; i += kThreads) bus[i] = p.bu[i];

    const int total = p.NM * p.tiles_per_sample;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int n = tile / p.tiles_per_sample;
        const int to0 = (tile - n * p.tiles_per_sample) * TT;
        const int nto = min(TT, p.Tout - to0);
        const int valid = nto * V;
        const int ti0 = to0 * s - kHalf;        // first input frame of the halo (may be < 0)
        __syncthreads();                        // previous tile finished with hs / h2s
        // halo tile: TI frames x V joints x BP channels, zero outside [0, T)
        for (int i = tid; i < TI * V * (BP / 4); i += kThreads) {
            const int r = i / (BP / 4), c4 = (i % (BP / 4)) * 4;
            const int ti = ti0 + r / V;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ti >= 0 && ti < p.T)
                v = ld4(p.h1 + (((size_t)n * p.T + ti) * V + (r % V)) * BP + c4);
            st4(hs + r * LDH + c4, v);
        }
        __syncthreads();
        // ---- temporal 15-tap conv: h2[(to,v)][co] = sum_tap sum_ci hs[(to*s+tap, v)][ci] Weff
        for (int mt = warp; mt * 16 < valid; mt += kWarps) {
            int base[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int r = mt * 16 + g + 8 * h;
                if (r >= valid) r = 0;
                const int to_l = r / V, v = r - to_l * V;
                base[h] = ((to_l * s) * V + v) * LDH;
            }
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
            for (int tap = 0; tap < kTaps; ++tap) {
#pragma unroll
                for (int kk = 0; kk < NT; ++kk) {
                    float a[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        a[i] = hs[base[i & 1] + tap * V * LDH + kk * 8 + t + 4 * (i >> 1)];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        float b[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i)
                            b[i] = Wts[(tap * BP + kk * 8 + t + 4 * i) * LDW + nt * 8 + g];
                        mma_step<PRECISE>(acc[nt], a, b);
                    }
                }
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = mt * 16 + g + 8 * h;
                    const int c = nt * 8 + 2 * t;
                    const float v0 = acc[nt][2 * h] + p.beff[c], v1 = acc[nt][2 * h + 1] + p.beff[c + 1];
                    const bool ok = r < valid;
                    *reinterpret_cast<float2*>(h2s + r * LDH + c) = ok ? make_float2(v0, v1)
                                                                      : make_float2(0.f, 0.f);
                    if (ok)
                        *reinterpret_cast<float2*>(
                            p.h2 + (((size_t)n * p.Tout + to0) * V + r) * BP + c) = make_float2(v0, v1);
                }
        }
        __syncthreads();
        // ---- up projection u = h2 Wu + bu, BN2 statistics
        for (int mt = warp; mt * 16 < valid; mt += kWarps) {
            float a[NT][4];
#pragma unroll
            for (int kk = 0; kk < NT; ++kk)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    a[kk][i] = h2s[(mt * 16 + g + 8 * (i & 1)) * LDH + kk * 8 + t + 4 * (i >> 1)];
            // row pointers and validity hoisted out of the column loop; the bias seeds the accumulator
            const int r0 = mt * 16 + g, r1 = r0 + 8;
            const bool ok0 = r0 < valid, ok1 = r1 < valid;
            float* u0 = p.u + (((size_t)n * p.Tout + to0) * V + r0) * C + 2 * t;
            float* u1 = u0 + (size_t)8 * C;
            const float* wb = Wus + t * LDU + g;
            for (int nt = 0; nt < C / 8; ++nt) {
                const float2 bb = *reinterpret_cast<const float2*>(bus + nt * 8 + 2 * t);
                float acc[4] = {bb.x, bb.y, bb.x, bb.y};
#pragma unroll
                for (int kk = 0; kk < NT; ++kk) {
                    float b[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) b[i] = wb[(kk * 8 + 4 * i) * LDU + nt * 8];
                    mma_step<PRECISE>(acc, a[kk], b);
                }
                const int c = nt * 8 + 2 * t;
                float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                if (ok0) {
                    *reinterpret_cast<float2*>(u0 + nt * 8) = make_float2(acc[0], acc[1]);
                    s0 = acc[0]; s1 = acc[1]; q0 = acc[0] * acc[0]; q1 = acc[1] * acc[1];
                }
                if (ok1) {
                    *reinterpret_cast<float2*>(u1 + nt * 8) = make_float2(acc[2], acc[3]);
                    s0 += acc[2]; s1 += acc[3]; q0 = fmaf(acc[2], acc[2], q0); q1 = fmaf(acc[3], acc[3], q1);
                }
                if (p.stat_sum) {
                    s0 = group_sum_g(s0); s1 = group_sum_g(s1);
                    q0 = group_sum_g(q0); q1 = group_sum_g(q1);
                    if (g == 0)
                        *reinterpret_cast<float4*>(s_col + ((size_t)mt * C + c) * 2) =
                            make_float4(s0, q0, s1, q1);
                }
            }
        }
        if (p.stat_sum) {
            __syncthreads();
            if (tid < C) {
                for (int m = 0; m * 16 < valid; ++m) {
                    const float2 v = *reinterpret_cast<const float2*>(s_col + ((size_t)m * C + tid) * 2);
                    acc_s += (double)v.x;
                    acc_q += (double)v.y;
                }
            }
        }
    }
    if (p.stat_sum && tid < C) {
        atomicAdd(&p.stat_sum[tid], acc_s);
        atomicAdd(&p.stat_sumsq[tid], acc_q);
    }
}

// --------------------------------------------------------------------------- backward: up
struct TcnBwdUpParams {
    const float *go, *u, *p2, *m12, *c2, *mean2, *h2, *Wu;
    float *dh2, *dWu, *dbu, *dbeff;
    long long rows;
    int C, bp;
    float drop_p, keep_scale;
    uint64_t seed;
    const unsigned long long* step;
};

template <int NT, bool PRECISE>
__global__ void __launch_bounds__(kThreads, 4) tcn_bwd_up_kernel(TcnBwdUpParams p) {
    constexpr int BP = NT * 8;
    constexpr int LDT = BP == 8 ? 8 : 24;       // h2s as A^T (t-indexed rows)
    constexpr int MAXCH = 8;                    // C <= 256 -> at most 8 column chunks of 32
    extern __shared__ __align__(16) float smem[];
    const int C = p.C, LDWU = C + 4;
    float* DUs = smem;                          // [128][36]
    float* h2s = DUs + kTileRows * 36;          // [128][LDT] (+8 pad)
    float* Wus = h2s + kTileRows * LDT + 8;     // [BP][LDWU]   B^T: n = j, k = c
    float* s_dbu = Wus + BP * LDWU;             // [C]
    float* s_dbe = s_dbu + C;                   // [BP]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    for (int i = tid; i < BP * C; i += kThreads) Wus[(i / C) * LDWU + (i % C)] = p.Wu[i];
    for (int i = tid; i < C + BP; i += kThreads) s_dbu[i] = 0.f;
    float acc_w[MAXCH][4];                      // dWu[j][c]: m = j (16, rows >= bp ignored)
#pragma unroll
    for (int ch = 0; ch < MAXCH; ++ch)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc_w[ch][i] = 0.f;
    const int wn = warp & 3, wk = warp >> 2;    // n-tile of the chunk, K-half (64 rows)
    const uint64_t eseed = effective_seed(p.seed, p.step);

    const long long tiles = (p.rows + kTileRows - 1) / kTileRows;
    float4 gr[4], ur[4];
    if (blockIdx.x < tiles) {
        prefetch_slice(gr, p.go, (long long)blockIdx.x * kTileRows, p.rows, C, 0, tid);
        prefetch_slice(ur, p.u, (long long)blockIdx.x * kTileRows, p.rows, C, 0, tid);
    }
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * kTileRows;
        const int valid = (int)min((long long)kTileRows, p.rows - row0);
        __syncthreads();
        for (int i = tid; i < kTileRows * (BP / 4); i += kThreads) {
            const int r = i / (BP / 4), c4 = (i % (BP / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < valid) v = ld4(p.h2 + (row0 + r) * BP + c4);
            st4(h2s + r * LDT + c4, v);
        }
        float acc_h[1][NT][4];
        zero_acc<1, NT>(acc_h);
#pragma unroll
        for (int ch = 0; ch < MAXCH; ++ch) {
            const int c0 = ch * 32;
            if (c0 < C) {
                __syncthreads();
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
                    const int i = tid + uu * kThreads;
                    const int r = i >> 3, c4 = (i & 7) * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < valid) {
                        const long long off = (row0 + r) * C + c0 + c4;
                        const float4 gv = gr[uu], uv = ur[uu];
                        const float4 pv = ld4(p.p2 + c0 + c4), mv = ld4(p.m12 + c0 + c4),
                                     cv = ld4(p.c2 + c0 + c4), nv = ld4(p.mean2 + c0 + c4);
                        float gy[4] = {gv.x, gv.y, gv.z, gv.w};
                        if (p.drop_p > 0.f) {
                            bool keep[4];
                            dropout_keep4(eseed, (uint64_t)(off >> 2), p.drop_p, keep);
#pragma unroll
                            for (int j = 0; j < 4; ++j) gy[j] = keep[j] ? gy[j] * p.keep_scale : 0.f;
                        }
                        v.x = bn_back(gy[0], uv.x, pv.x, mv.x, cv.x, nv.x);
                        v.y = bn_back(gy[1], uv.y, pv.y, mv.y, cv.y, nv.y);
                        v.z = bn_back(gy[2], uv.z, pv.z, mv.z, cv.z, nv.z);
                        v.w = bn_back(gy[3], uv.w, pv.w, mv.w, cv.w, nv.w);
                    }
                    st4(DUs + r * 36 + c4, v);
                }
                {   // request the next slice of go / u before the math of this one
                    const bool same = c0 + 32 < C;
                    const long long nrow0 = same ? row0 : (tile + gridDim.x) * kTileRows;
                    if (same || tile + gridDim.x < tiles) {
                        prefetch_slice(gr, p.go, nrow0, p.rows, C, same ? c0 + 32 : 0, tid);
                        prefetch_slice(ur, p.u, nrow0, p.rows, C, same ? c0 + 32 : 0, tid);
                    }
                }
                __syncthreads();
                // dh2[rows][j] += DU[rows][c0..] * Wu[j][c0..]^T
                warp_mma<1, NT, false, true, PRECISE>(acc_h, DUs + warp * 16 * 36, 36, Wus + c0, LDWU,
                                                      32, lane);
                // dWu[j][c0 + wn*8 ..] += h2^T (K = 64 rows of this warp's half) * DU
                {
                    float accw[1][1][4] = {{{acc_w[ch][0], acc_w[ch][1], acc_w[ch][2], acc_w[ch][3]}}};
                    warp_mma<1, 1, true, false, PRECISE>(accw, h2s + wk * 64 * LDT, LDT,
                                                         DUs + wk * 64 * 36 + wn * 8, 36, 64, lane);
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc_w[ch][i] = accw[0][0][i];
                }
                // dbu[c] += column sums of DU
                {
                    const int c = tid & 31, r0 = (tid >> 5) * 16;
                    float sacc = 0.f;
#pragma unroll 4
                    for (int r = 0; r < 16; ++r) sacc += DUs[(r0 + r) * 36 + c];
                    atomicAdd(&s_dbu[c0 + c], sacc);
                }
            }
        }
        // dh2 tile -> global, dbeff = column sums of dh2
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int c = nt * 8 + 2 * t;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = warp * 16 + g + 8 * h;
                if (r < valid) {
                    *reinterpret_cast<float2*>(p.dh2 + (row0 + r) * BP + c) =
                        make_float2(acc_h[0][nt][2 * h], acc_h[0][nt][2 * h + 1]);
                    s0 += acc_h[0][nt][2 * h];
                    s1 += acc_h[0][nt][2 * h + 1];
                }
            }
            s0 = group_sum_g(s0); s1 = group_sum_g(s1);
            if (g == 0) { atomicAdd(&s_dbe[c], s0); atomicAdd(&s_dbe[c + 1], s1); }
        }
    }
    __syncthreads();
#pragma unroll
    for (int ch = 0; ch < MAXCH; ++ch) {
        if (ch * 32 < C) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = g + 8 * (i >> 1);
                const int c = ch * 32 + wn * 8 + 2 * t + (i & 1);
                if (j < BP) atomicAdd(&p.dWu[(size_t)j * C + c], acc_w[ch][i]);
            }
        }
    }
    for (int c = tid; c < C; c += kThreads) atomicAdd(&p.dbu[c], s_dbu[c]);
    for (int c = tid; c < BP; c += kThreads) atomicAdd(&p.dbeff[c], s_dbe[c]);
}

// ----------------------------------------------------------------------- backward: temporal
struct TcnBwdTParams {
    const float *dh2, *h1, *Weff;
    float *dh1, *dWeff, *dbd;
    int NM, T, Tout, V, bp, stride, TT, tiles_per_sample;
};

// Tile = TT input frames of one sample.  dh1[(ti,v)][ci] = sum_tap sum_co Weff[tap][ci][co] *
// dh2[(to,v)][co] with to*s + tap - 7 = ti;  dWeff[tap][ci][co] += h1[(ti,v)][ci] * dh2[(to,v)][co].
template <int NT, bool PRECISE>
__global__ void __launch_bounds__(kThreads) tcn_bwd_t_kernel(TcnBwdTParams p) {
    constexpr int BP = NT * 8;
    constexpr int LDH = BP + 4;                 // dh2 halo (g-indexed gathers)
    constexpr int LDT = BP == 8 ? 8 : 24;       // h1 tile as A^T
    constexpr int LDWT = BP + 4;                // Weff[tap][ci][co] as B^T (n = ci, k = co)
    extern __shared__ __align__(16) float smem[];
    const int V = p.V, s = p.stride, TT = p.TT;
    const int TO = (TT - 1 + 2 * kHalf) / s + 2;   // output frames that can touch the tile
    float* ds = smem;                           // [TO*V][LDH]   dh2 halo
    float* h1s = ds + TO * V * LDH;             // [256][LDT] (+8)
    float* Wts = h1s + kUpRows * LDT + 8;       // [15*BP][LDWT]
    float* s_dbd = Wts + kTaps * BP * LDWT;     // [BP]
    int2* s_row = reinterpret_cast<int2*>(s_dbd + BP + (BP & 1));   // [256] {frame in tile, v*LDH}
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    for (int i = tid; i < kTaps * BP * BP; i += kThreads) Wts[(i / BP) * LDWT + (i % BP)] = p.Weff[i];
    for (int i = tid; i < BP; i += kThreads) s_dbd[i] = 0.f;
    float acc_w[2][NT][4];                      // this warp's taps: warp and warp + 8
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < NT; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc_w[a][b][c] = 0.f;

    const int total = p.NM * p.tiles_per_sample;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int n = tile / p.tiles_per_sample;
        const int ti0 = (tile - n * p.tiles_per_sample) * TT;
        const int nti = min(TT, p.T - ti0);
        const int valid = nti * V;
        // first output frame that can contribute: to*s + 14 - 7 >= ti0  ->  to >= (ti0 - 7)/s
        int to_lo = ti0 - kHalf;
        to_lo = to_lo <= 0 ? 0 : (to_lo + s - 1) / s;
        __syncthreads();
        for (int i = tid; i < TO * V * (BP / 4); i += kThreads) {
            const int r = i / (BP / 4), c4 = (i % (BP / 4)) * 4;
            const int to = to_lo + r / V;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (to < p.Tout) v = ld4(p.dh2 + (((size_t)n * p.Tout + to) * V + (r % V)) * BP + c4);
            st4(ds + r * LDH + c4, v);
        }
        for (int i = tid; i < kUpRows * (BP / 4); i += kThreads) {
            const int r = i / (BP / 4), c4 = (i % (BP / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < valid) v = ld4(p.h1 + (((size_t)n * p.T + ti0) * V + r) * BP + c4);
            st4(h1s + r * LDT + c4, v);
        }
        // per row: {n0 = ti0 + frame + 7, joint offset | valid-tap mask << 16}: tap `tap` reads output
        // frame (n0 - tap) / s when bit `tap` of the mask is set (in range, right parity)
        const int sh = s == 2 ? 1 : 0;
        for (int r = tid; r < kUpRows; r += kThreads) {
            int2 e = make_int2(0, 0);
            if (r < valid) {
                const int ti_l = r / V;
                const int n0 = ti0 + ti_l + kHalf;
                unsigned mask = 0;
                for (int tap = 0; tap < kTaps; ++tap) {
                    const int num = n0 - tap;
                    if (num >= 0 && (num & sh) == 0 && (num >> sh) < p.Tout) mask |= 1u << tap;
                }
                e = make_int2(n0, ((r - ti_l * V) * LDH) | (int)(mask << 16));
            }
            s_row[r] = e;
        }
        __syncthreads();
        const int VL = V * LDH;
        auto src_off = [&](int2 ri, int tap) -> int {
            const int off = (((ri.x - tap) >> sh) - to_lo) * VL + (ri.y & 0xFFFF);
            return ((ri.y >> (16 + tap)) & 1) ? off : -1;
        };
        // ---- dh1
        for (int mt = warp; mt * 16 < valid; mt += kWarps) {
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
            const int2 ri0 = s_row[mt * 16 + g], ri1 = s_row[mt * 16 + g + 8];
            for (int tap = 0; tap < kTaps; ++tap) {
                const int o0 = src_off(ri0, tap), o1 = src_off(ri1, tap);
#pragma unroll
                for (int kk = 0; kk < NT; ++kk) {       // k = co
                    float a[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int o = (i & 1) ? o1 : o0;
                        a[i] = o < 0 ? 0.f : ds[o + kk * 8 + t + 4 * (i >> 1)];
                    }
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {   // n = ci
                        float b[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i)
                            b[i] = Wts[(tap * BP + nt * 8 + g) * LDWT + kk * 8 + t + 4 * i];
                        mma_step<PRECISE>(acc[nt], a, b);
                    }
                }
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int c = nt * 8 + 2 * t;
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = mt * 16 + g + 8 * h;
                    if (r < valid) {
                        *reinterpret_cast<float2*>(
                            p.dh1 + (((size_t)n * p.T + ti0) * V + r) * BP + c) =
                            make_float2(acc[nt][2 * h], acc[nt][2 * h + 1]);
                        s0 += acc[nt][2 * h];
                        s1 += acc[nt][2 * h + 1];
                    }
                }
                s0 = group_sum_g(s0); s1 = group_sum_g(s1);
                if (g == 0) { atomicAdd(&s_dbd[c], s0); atomicAdd(&s_dbd[c + 1], s1); }
            }
        }
        // ---- dWeff[tap][ci][co]: m = ci (A^T from h1s), n = co, k = rows of the tile
#pragma unroll
        for (int sel = 0; sel < 2; ++sel) {
            const int tap = warp + sel * kWarps;
            if (tap < kTaps) {
                for (int k0 = 0; k0 < valid; k0 += 8) {
                    float a[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        a[i] = h1s[(k0 + t + 4 * (i >> 1)) * LDT + g + 8 * (i & 1)];
                    const int o0 = src_off(s_row[k0 + t], tap), o1 = src_off(s_row[k0 + t + 4], tap);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        float b[2];
                        b[0] = o0 < 0 ? 0.f : ds[o0 + nt * 8 + g];
                        b[1] = o1 < 0 ? 0.f : ds[o1 + nt * 8 + g];
                        mma_step<PRECISE>(acc_w[sel][nt], a, b);
                    }
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int sel = 0; sel < 2; ++sel) {
        const int tap = warp + sel * kWarps;
        if (tap < kTaps) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ci = g + 8 * (i >> 1), co = nt * 8 + 2 * t + (i & 1);
                    if (ci < BP) atomicAdd(&p.dWeff[(tap * BP + ci) * BP + co], acc_w[sel][nt][i]);
                }
        }
    }
    for (int c = tid; c < BP; c += kThreads) atomicAdd(&p.dbd[c], s_dbd[c]);
}

// --------------------------------------------------------------------------- backward: down
struct TcnBwdDownParams {
    const float *dh1, *z, *scale1, *beta1, *mean1, *rstd1, *Wd;
    float *g1, *dWd;
    double *sg1, *sg1x;
    long long rows;
    int C, bp;
};

template <int NT, bool PRECISE>
__global__ void __launch_bounds__(kThreads, 3) tcn_bwd_down_kernel(TcnBwdDownParams p) {
    constexpr int BP = NT * 8;
    constexpr int LDD = BP + 4;                 // dh1s as A (g-indexed), Wds as B^T (g-indexed)
    constexpr int LDA = 40;                     // a-tile as A^T (t-indexed)
    constexpr int MAXCH = 8;
    extern __shared__ __align__(16) float smem[];
    const int C = p.C;
    float* Zs = smem;                           // [128][36]   raw z chunk
    float* As = Zs + kTileRows * 36;            // [128][40]   relu(bn1(z)) chunk
    float* dhs = As + kTileRows * LDA;          // [128][LDD]
    float* Wds = dhs + kTileRows * LDD;         // [C][LDD]
    float* s_g = Wds + C * LDD;                 // [C]
    float* s_gx = s_g + C;                      // [C]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    for (int i = tid; i < C * BP; i += kThreads) Wds[(i / BP) * LDD + (i % BP)] = p.Wd[i];
    for (int i = tid; i < 2 * C; i += kThreads) s_g[i] = 0.f;
    float acc_w[MAXCH][NT][4];                  // dWd[c][j]: m = c (this warp's m-tile), n = j
#pragma unroll
    for (int a = 0; a < MAXCH; ++a)
#pragma unroll
        for (int b = 0; b < NT; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc_w[a][b][c] = 0.f;
    const int wm = warp & 1, wk = warp >> 1;    // m-tile (16 channels) of the chunk, K-quarter

    const long long tiles = (p.rows + kTileRows - 1) / kTileRows;
    float4 zr[4];
    if (blockIdx.x < tiles) prefetch_slice(zr, p.z, (long long)blockIdx.x * kTileRows, p.rows, C, 0, tid);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * kTileRows;
        const int valid = (int)min((long long)kTileRows, p.rows - row0);
        __syncthreads();
        for (int i = tid; i < kTileRows * (BP / 4); i += kThreads) {
            const int r = i / (BP / 4), c4 = (i % (BP / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < valid) v = ld4(p.dh1 + (row0 + r) * BP + c4);
            st4(dhs + r * LDD + c4, v);
        }
#pragma unroll
        for (int ch = 0; ch < MAXCH; ++ch) {
            const int c0 = ch * 32;
            if (c0 < C) {
                __syncthreads();
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
                    const int i = tid + uu * kThreads;
                    const int r = i >> 3, c4 = (i & 7) * 4;
                    const float4 zv = r < valid ? zr[uu] : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < valid) {
                        const float4 mu = ld4(p.mean1 + c0 + c4), sc = ld4(p.scale1 + c0 + c4),
                                     be = ld4(p.beta1 + c0 + c4);
                        av.x = fmaxf(bn_apply(zv.x, mu.x, sc.x, be.x), 0.f);
                        av.y = fmaxf(bn_apply(zv.y, mu.y, sc.y, be.y), 0.f);
                        av.z = fmaxf(bn_apply(zv.z, mu.z, sc.z, be.z), 0.f);
                        av.w = fmaxf(bn_apply(zv.w, mu.w, sc.w, be.w), 0.f);
                    }
                    st4(Zs + r * 36 + c4, zv);
                    st4(As + r * LDA + c4, av);
                }
                {
                    const bool same = c0 + 32 < C;
                    const long long nrow0 = same ? row0 : (tile + gridDim.x) * kTileRows;
                    if (same || tile + gridDim.x < tiles)
                        prefetch_slice(zr, p.z, nrow0, p.rows, C, same ? c0 + 32 : 0, tid);
                }
                __syncthreads();
                // da[rows][c0..c0+32) = dh1[rows][j] * Wd[c][j]^T  (K = bp)
                float acc[1][4][4];
                zero_acc<1, 4>(acc);
                warp_mma<1, 4, false, true, PRECISE>(acc, dhs + warp * 16 * LDD, LDD, Wds + c0 * LDD,
                                                     LDD, BP, lane);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int cl = nt * 8 + 2 * t;
                    const float mu0 = p.mean1[c0 + cl], mu1 = p.mean1[c0 + cl + 1];
                    const float rs0 = p.rstd1[c0 + cl], rs1 = p.rstd1[c0 + cl + 1];
                    float s0 = 0.f, s1 = 0.f, x0 = 0.f, x1 = 0.f;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = warp * 16 + g + 8 * h;
                        if (r < valid) {
                            const float a0 = As[r * LDA + cl], a1 = As[r * LDA + cl + 1];
                            const float g0 = a0 > 0.f ? acc[0][nt][2 * h] : 0.f;
                            const float g1v = a1 > 0.f ? acc[0][nt][2 * h + 1] : 0.f;
                            *reinterpret_cast<float2*>(p.g1 + (row0 + r) * C + c0 + cl) =
                                make_float2(g0, g1v);
                            s0 += g0; s1 += g1v;
                            x0 += g0 * (Zs[r * 36 + cl] - mu0) * rs0;
                            x1 += g1v * (Zs[r * 36 + cl + 1] - mu1) * rs1;
                        }
                    }
                    s0 = group_sum_g(s0); s1 = group_sum_g(s1);
                    x0 = group_sum_g(x0); x1 = group_sum_g(x1);
                    if (g == 0) {
                        atomicAdd(&s_g[c0 + cl], s0); atomicAdd(&s_g[c0 + cl + 1], s1);
                        atomicAdd(&s_gx[c0 + cl], x0); atomicAdd(&s_gx[c0 + cl + 1], x1);
                    }
                }
                // dWd[c0 + wm*16 ..][j] += a^T (K = 32 rows of this warp's quarter) * dh1
                {
                    float accw[1][NT][4];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int i = 0; i < 4; ++i) accw[0][nt][i] = acc_w[ch][nt][i];
                    // B (k = row, n = j) read from dhs with k as the row index: ld = LDD is
                    // g-friendly, not t-friendly -> a few 2-way conflicts on a tiny operand.
                    warp_mma<1, NT, true, false, PRECISE>(accw, As + wk * 32 * LDA + wm * 16, LDA,
                                                          dhs + wk * 32 * LDD, LDD, 32, lane);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc_w[ch][nt][i] = accw[0][nt][i];
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int ch = 0; ch < MAXCH; ++ch) {
        if (ch * 32 < C) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = ch * 32 + wm * 16 + g + 8 * (i >> 1);
                    const int j = nt * 8 + 2 * t + (i & 1);
                    atomicAdd(&p.dWd[(size_t)c * BP + j], acc_w[ch][nt][i]);
                }
        }
    }
    for (int c = tid; c < C; c += kThreads) {
        atomicAdd(&p.sg1[c], (double)s_g[c]);
        atomicAdd(&p.sg1x[c], (double)s_gx[c]);
    }
}

// The same stage as a plain streaming kernel on the CUDA cores.  The contraction of this stage
// runs over the bp <= 16 bottleneck channels only, so there is nothing for a tensor core to do:
// thread = (row sub-group, CH channels), one vector of z in / one vector of g1 out per row, the
// weight-gradient and BatchNorm sums stay in registers over the thread's rows and are flushed
// once per CTA (CH = 4 for bp = 8, 2 for bp = 16: CH*bp accumulators must stay in registers).
// Exact fp32 (serves both math modes); HBM-bound like the block-tail kernels.
template <int BP, int CH>
__global__ void __launch_bounds__(256) tcn_bwd_down_stream_kernel(TcnBwdDownParams p) {
    extern __shared__ __align__(16) float smem[];
    const int C = p.C, cgn = C / CH;            // channel groups = threads per row
    float* Wt = smem;                           // [BP][C]  (Wd transposed: conflict-free reads)
    float* red = Wt + BP * C;                   // [256][2*CH] reduction scratch
    const int tid = threadIdx.x;
    const int rows_per_iter = 256 / cgn;
    const int col = tid % cgn, rsub = tid / cgn, c = col * CH;
    for (int i = tid; i < C * BP; i += 256) Wt[(i % BP) * C + (i / BP)] = p.Wd[i];
    __syncthreads();
    float mu[CH], sc[CH], be[CH], rs[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        mu[i] = p.mean1[c + i]; sc[i] = p.scale1[c + i]; be[i] = p.beta1[c + i]; rs[i] = p.rstd1[c + i];
    }
    float accw[CH][BP];
#pragma unroll
    for (int i = 0; i < CH; ++i)
#pragma unroll
        for (int j = 0; j < BP; ++j) accw[i][j] = 0.f;
    float sg[CH], sgx[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) sg[i] = sgx[i] = 0.f;
    if (rsub < rows_per_iter) {
        for (long long r = (long long)blockIdx.x * rows_per_iter + rsub; r < p.rows;
             r += (long long)gridDim.x * rows_per_iter) {
            float z4[CH];
            if (CH == 4) {
                const float4 zv = ld4(p.z + r * C + c);
                z4[0] = zv.x; z4[1] = zv.y; z4[CH - 2] = zv.z; z4[CH - 1] = zv.w;
            } else {
                const float2 zv = *reinterpret_cast<const float2*>(p.z + r * C + c);
                z4[0] = zv.x; z4[1] = zv.y;
            }
            float dh[BP];
#pragma unroll
            for (int j = 0; j < BP; j += 4) {
                const float4 d = ld4(p.dh1 + r * BP + j);
                dh[j] = d.x; dh[j + 1] = d.y; dh[j + 2] = d.z; dh[j + 3] = d.w;
            }
            float a4[CH], da[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                a4[i] = fmaxf(bn_apply(z4[i], mu[i], sc[i], be[i]), 0.f);
                da[i] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < BP; ++j) {
#pragma unroll
                for (int i = 0; i < CH; ++i) da[i] = fmaf(dh[j], Wt[j * C + c + i], da[i]);
            }
            float g[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                g[i] = a4[i] > 0.f ? da[i] : 0.f;
                sg[i] += g[i];
                sgx[i] += g[i] * (z4[i] - mu[i]) * rs[i];
#pragma unroll
                for (int j = 0; j < BP; ++j) accw[i][j] = fmaf(a4[i], dh[j], accw[i][j]);
            }
            if (CH == 4)
                st4(p.g1 + r * C + c, make_float4(g[0], g[1], g[CH - 2], g[CH - 1]));
            else
                *reinterpret_cast<float2*>(p.g1 + r * C + c) = make_float2(g[0], g[1]);
        }
    }
    // ---- CTA reduction over the row sub-groups, then one atomic per output and CTA
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        red[tid * 2 * CH + i] = sg[i];
        red[tid * 2 * CH + CH + i] = sgx[i];
    }
    __syncthreads();
    for (int ch = tid; ch < C; ch += 256) {
        const int cc = ch / CH, i = ch % CH;
        double s0 = 0.0, s1 = 0.0;
        for (int k = 0; k < rows_per_iter; ++k) {
            s0 += red[(k * cgn + cc) * 2 * CH + i];
            s1 += red[(k * cgn + cc) * 2 * CH + CH + i];
        }
        atomicAdd(&p.sg1[ch], s0);
        atomicAdd(&p.sg1x[ch], s1);
    }
    // dWd two bottleneck columns at a time through the same scratch
#pragma unroll
    for (int j0 = 0; j0 < BP; j0 += 2) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            red[tid * 2 * CH + i * 2] = accw[i][j0];
            red[tid * 2 * CH + i * 2 + 1] = accw[i][j0 + 1];
        }
        __syncthreads();
        for (int o = tid; o < C * 2; o += 256) {
            const int ch = o >> 1, jj = o & 1;
            const int cc = ch / CH, i = ch % CH;
            float s0 = 0.f;
            for (int k = 0; k < rows_per_iter; ++k) s0 += red[(k * cgn + cc) * 2 * CH + i * 2 + jj];
            atomicAdd(&p.dWd[(size_t)ch * BP + j0 + jj], s0);
        }
    }
}

static int grid_for(long long tiles, int per_sm) {
    long long n = (long long)num_sms() * per_sm;
    if (n > tiles) n = tiles;
    return (int)(n < 1 ? 1 : n);
}

template <typename K>
static void set_smem(K kern, size_t bytes) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

static int check_tcn(const char* who, int NM, int T, int V, int C, int bp, int stride) {
    ISTGCN_REQUIRE(NM >= 0 && T >= 1 && V >= 1 && V <= 32, ISTGCN_E_SHAPE, "%s: bad NM/T/V (%d,%d,%d)", who, NM, T, V);
    ISTGCN_REQUIRE(C % 32 == 0 && C >= 32 && C <= 256, ISTGCN_E_SHAPE, "%s: C=%d unsupported (32..256, multiple of 32)", who, C);
    ISTGCN_REQUIRE(bp == 8 || bp == 16, ISTGCN_E_SHAPE, "%s: padded bottleneck bp=%d must be 8 or 16", who, bp);
    ISTGCN_REQUIRE(stride == 1 || stride == 2, ISTGCN_E_SHAPE, "%s: stride=%d unsupported", who, stride);
    return 0;
}

}  // namespace istgcn

using namespace istgcn;

ISTGCN_API int istgcn_tcn_fwd(const float* z, const float* mean1, const float* scale1,
                              const float* beta1, const float* Wd, const float* bd, const float* Weff, const float* beff,
                              const float* Wu, const float* bu, float* h1, float* h2, float* u,
                              double* stat_sum, double* stat_sumsq, int NM, int T, int V, int C,
                              int bp, int stride, int math, istgcn_stream_t s) {
    ISTGCN_REQUIRE(z && mean1 && scale1 && beta1 && Wd && bd && Weff && beff && Wu && bu && h1 && h2 && u,
                   ISTGCN_E_ARG, "tcn_fwd: null pointer");
    if (int e = check_tcn("tcn_fwd", NM, T, V, C, bp, stride)) return e;
    if (NM == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    const bool pc = math == ISTGCN_MATH_3XTF32;
    const int Tout = (T - 1) / stride + 1;
    {
        TcnDownParams p{z, mean1, scale1, beta1, Wd, bd, h1, (long long)NM * T * V, C, bp};
        const size_t smem = sizeof(float) * (kTileRows * 36 + C * ld_t(bp));
        const int grid = grid_for((p.rows + kTileRows - 1) / kTileRows, 4);
#define LAUNCH_DOWN(NT, PC)                                          \
    set_smem(tcn_down_kernel<NT, PC>, smem);                         \
    tcn_down_kernel<NT, PC><<<grid, kThreads, smem, st>>>(p)
        if (bp == 8) { if (pc) { LAUNCH_DOWN(1, true); } else { LAUNCH_DOWN(1, false); } }
        else { if (pc) { LAUNCH_DOWN(2, true); } else { LAUNCH_DOWN(2, false); } }
#undef LAUNCH_DOWN
        if (int e = finish_launch("tcn_down")) return e;
    }
    {
        TcnUpParams p{h1, Weff, beff, Wu, bu, h2, u, stat_sum, stat_sumsq,
                      NM, T, Tout, V, C, bp, stride, 0, 0};
        p.TT = kUpRows / V;
        if (p.TT > Tout) p.TT = Tout;
        p.tiles_per_sample = (Tout + p.TT - 1) / p.TT;
        const int TI = (p.TT - 1) * stride + kTaps;
        const size_t smem = sizeof(float) * ((size_t)TI * V * ld_g(bp) + kUpRows * ld_g(bp) +
                                             kTaps * bp * ld_t(bp) + bp * (C + 8) + 17 * C);
        const int grid = grid_for((long long)NM * p.tiles_per_sample, 4);
#define LAUNCH_UP(NT, PC)                                            \
    set_smem(tcn_up_kernel<NT, PC>, smem);                           \
    tcn_up_kernel<NT, PC><<<grid, kThreads, smem, st>>>(p)
        if (bp == 8) { if (pc) { LAUNCH_UP(1, true); } else { LAUNCH_UP(1, false); } }
        else { if (pc) { LAUNCH_UP(2, true); } else { LAUNCH_UP(2, false); } }
#undef LAUNCH_UP
        if (int e = finish_launch("tcn_up")) return e;
    }
    return 0;
}

ISTGCN_API int istgcn_tcn_bwd(const float* go, const float* u, const float* p2, const float* m12,
                              const float* c2, const float* mean2, const float* z,
                              const float* scale1, const float* beta1, const float* mean1,
                              const float* rstd1,
                              const float* h1, const float* h2, const float* Wd, const float* Weff,
                              const float* Wu, float* dh2_ws, float* dh1_ws, float* g1, double* sg1,
                              double* sg1x, float* dWd, float* dbd, float* dWeff, float* dbeff,
                              float* dWu, float* dbu, int NM, int T, int V, int C, int bp,
                              int stride, float drop_p, uint64_t drop_seed,
                              const unsigned long long* drop_step, int math, istgcn_stream_t s) {
    ISTGCN_REQUIRE(go && u && p2 && m12 && c2 && mean2 && z && scale1 && beta1 && mean1 && rstd1 && h1 && h2 &&
                       Wd && Weff && Wu && dh2_ws && dh1_ws && g1 && sg1 && sg1x && dWd && dbd &&
                       dWeff && dbeff && dWu && dbu,
                   ISTGCN_E_ARG, "tcn_bwd: null pointer");
    if (int e = check_tcn("tcn_bwd", NM, T, V, C, bp, stride)) return e;
    ISTGCN_REQUIRE(drop_p >= 0.f && drop_p < 1.f, ISTGCN_E_ARG, "tcn_bwd: dropout p=%f", drop_p);
    if (NM == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    const bool pc = math == ISTGCN_MATH_3XTF32;
    const int Tout = (T - 1) / stride + 1;
    {
        TcnBwdUpParams p{go, u, p2, m12, c2, mean2, h2, Wu, dh2_ws, dWu, dbu, dbeff,
                         (long long)NM * Tout * V, C, bp, drop_p, 1.f / (1.f - drop_p), drop_seed,
                         drop_step};
        const size_t smem = sizeof(float) * (kTileRows * 36 + kTileRows * ld_t(bp) + 8 + bp * (C + 4) + C + bp);
        const int grid = grid_for((p.rows + kTileRows - 1) / kTileRows, 4);
#define LAUNCH_BU(NT, PC)                                            \
    set_smem(tcn_bwd_up_kernel<NT, PC>, smem);                       \
    tcn_bwd_up_kernel<NT, PC><<<grid, kThreads, smem, st>>>(p)
        if (bp == 8) { if (pc) { LAUNCH_BU(1, true); } else { LAUNCH_BU(1, false); } }
        else { if (pc) { LAUNCH_BU(2, true); } else { LAUNCH_BU(2, false); } }
#undef LAUNCH_BU
        if (int e = finish_launch("tcn_bwd_up")) return e;
    }
    {
        TcnBwdTParams p{dh2_ws, h1, Weff, dh1_ws, dWeff, dbd, NM, T, Tout, V, bp, stride, 0, 0};
        p.TT = kUpRows / V;
        if (p.TT > T) p.TT = T;
        p.tiles_per_sample = (T + p.TT - 1) / p.TT;
        const int TO = (p.TT - 1 + 2 * kHalf) / stride + 2;
        const size_t smem = sizeof(float) * ((size_t)TO * V * ld_g(bp) + kUpRows * ld_t(bp) + 8 +
                                             kTaps * bp * ld_g(bp) + bp + 2) + kUpRows * sizeof(int2);
        const int grid = grid_for((long long)NM * p.tiles_per_sample, 4);
#define LAUNCH_BT(NT, PC)                                            \
    set_smem(tcn_bwd_t_kernel<NT, PC>, smem);                        \
    tcn_bwd_t_kernel<NT, PC><<<grid, kThreads, smem, st>>>(p)
        if (bp == 8) { if (pc) { LAUNCH_BT(1, true); } else { LAUNCH_BT(1, false); } }
        else { if (pc) { LAUNCH_BT(2, true); } else { LAUNCH_BT(2, false); } }
#undef LAUNCH_BT
        if (int e = finish_launch("tcn_bwd_t")) return e;
    }
    {
        TcnBwdDownParams p{dh1_ws, z, scale1, beta1, mean1, rstd1, Wd, g1, dWd, sg1, sg1x,
                           (long long)NM * T * V, C, bp};
        const int cgn = C / (bp == 8 ? 4 : 2);
        // streaming CUDA-core version: wins for bp = 8 (0.18 vs 0.23 ms at C = 64); at bp = 16 the
        // 2*C*bp FMAs per row make it instruction-bound and the mma.sync kernel below is faster
        if (bp == 8 && C % 4 == 0 && cgn <= 256 && 256 % cgn == 0) {
            const size_t sm = sizeof(float) * ((size_t)bp * C + 256 * 8);
            const int rpi = 256 / cgn;
            long long blocks = (p.rows + rpi * 8 - 1) / (rpi * 8);
            const long long cap = (long long)num_sms() * 4;
            if (blocks > cap) blocks = cap;
            if (bp == 8) {
                set_smem(tcn_bwd_down_stream_kernel<8, 4>, sm);
                tcn_bwd_down_stream_kernel<8, 4><<<(int)blocks, 256, sm, st>>>(p);
            } else {
                set_smem(tcn_bwd_down_stream_kernel<16, 2>, sm);
                tcn_bwd_down_stream_kernel<16, 2><<<(int)blocks, 256, sm, st>>>(p);
            }
            return finish_launch("tcn_bwd_down");
        }
        const size_t smem = sizeof(float) * (kTileRows * 36 + kTileRows * 40 + kTileRows * ld_g(bp) +
                                             C * ld_g(bp) + 2 * C);
        const int grid = grid_for((p.rows + kTileRows - 1) / kTileRows, 3);
#define LAUNCH_BD(NT, PC)                                            \
    set_smem(tcn_bwd_down_kernel<NT, PC>, smem);                     \
    tcn_bwd_down_kernel<NT, PC><<<grid, kThreads, smem, st>>>(p)
        if (bp == 8) { if (pc) { LAUNCH_BD(1, true); } else { LAUNCH_BD(1, false); } }
        else { if (pc) { LAUNCH_BD(2, true); } else { LAUNCH_BD(2, false); } }
#undef LAUNCH_BD
        if (int e = finish_launch("tcn_bwd_down")) return e;
    }
    return 0;
}
