// Inception TCN with 1x1 bottlenecks, second generation ('tf32' math mode): streaming kernels.
// reference: net/st_gcn_mstcn_1x1.py:186-266 (see csrc/tcn.cu for the algebra; that file keeps the
// error-compensated 3xTF32 parity mode).
//
// The chain is HBM-bound (C channels in, C out per row, 2*C*bp FLOP/row with bp = 8 / 16), so the
// kernels are built around the memory stream, not around the matrix unit:
//   * no shared-memory staging of the activation tiles and no __syncthreads in the main loops:
//     every warp owns 16 consecutive rows; the rows go from global memory STRAIGHT into mma.sync
//     A fragments.  A thread loads 16-byte (or 8-byte) vectors of consecutive channels and the
//     contraction index k of the m16n8k8 fragment is simply re-labelled: fragment slot (k = t,
//     k = t + 4) of step h holds channels (4t + 2h, 4t + 2h + 1) of the 16-channel chunk, and the
//     (tiny, pre-arranged) weight fragments use the same labelling;
//   * the accumulator layout of one product IS the A-fragment layout of the next one under the same
//     re-labelling, so 15-tap conv -> 1x1 needs no data movement at all;
//   * every big tensor is touched once: z is read by down, u written by up, (go, u) read by bwd_up,
//     (z) read and g1 written by bwd_down.  Only the bp-wide intermediates (h1, h2, dh2, dh1 =
//     1/8 .. 1/16 of an activation, L2-resident) make round trips, through the two small
//     temporal kernels;
//   * the products whose contraction runs over ROWS (weight gradients) take their second operand
//     from a per-warp 16 x 64 staging tile in shared memory (XOR-swizzled, __syncwarp only);
//   * per-channel sums (BatchNorm statistics, bias gradients) stay in registers over all the
//     tiles of a warp and are flushed once: shuffles -> shared -> one double atomic per CTA.
// Kernels (one C-ABI entry point each, so that every launch can be timed on its own):
//   tcn2_down      z -> h1 = relu(BN1(z)) Wd + bd
//   tcn2_conv      h1 -> h2 = sum_tap Weff[tap] h1[to*s + tap - 7] + beff        (csrc/tcn2_small.cu)
//   tcn2_up        h2 -> u = h2 Wu + bu, BN2 sums
//   tcn2_bwd_up    go, u, h2 -> du -> dh2 = du Wu^T, dWu, dbu, dbeff
//   tcn2_bwd_conv  dh2, h1 -> dh1, dWeff, dbd                                    (csrc/tcn2_small.cu)
//   tcn2_bwd_down  dh1, z -> g1 = (dh1 Wd^T) masked by the ReLU, BN1-backward sums, dWd
#include "common.cuh"

namespace istgcn {
namespace {

constexpr int kT2Threads = 256;
constexpr int kT2HeavyWarps = 4;      // warps per CTA of the register-heavy backward kernels

__device__ __forceinline__ uint32_t tf(float x) { return to_tf32_fast(x); }
// value rounded to TF32 (stored form of the bottleneck tensors h1, h2, dh2, dh1: they are only ever
// tensor-core operands, so their consumers need no conversion instruction)
__device__ __forceinline__ float rnd_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ uint32_t raw(float x) { return __float_as_uint(x); }
__device__ __forceinline__ float tff(float x) { return __uint_as_float(to_tf32_fast(x)); }

__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], float2 w) {
    const uint32_t b[2] = {__float_as_uint(w.x), __float_as_uint(w.y)};
    mma_m16n8k8(d, a, b);
}

// XOR swizzle of the per-warp [16][64] staging tiles: rows that are read together as the k index of
// a B fragment (t, t + 4 -> row & 3 = t) land in different 8-bank groups, rows that are written
// together (g, g + 1) in different 16-bank halves.
__device__ __forceinline__ int swz(int row) { return ((row & 1) << 4) | (((row >> 1) & 1) << 3); }

__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 lds4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

// A fragments of a 16-row tile of a [rows][BP] matrix (BP = 8 * NT), rows g / g + 8 of the tile,
// under the channel re-labelling: NT = 1: one k-step, slots (t, t+4) = channels (2t, 2t+1);
// NT = 2: two k-steps, slots of step kk = channels (4t + 2kk, 4t + 2kk + 1).
template <int NT>
__device__ __forceinline__ void load_small_a(uint32_t (&a)[NT][4], const float* __restrict__ base,
                                             long long off0, long long off1, bool ok0, bool ok1,
                                             int t) {
    if (NT == 1) {
        const float2 x = ok0 ? ldg2(base + off0 + 2 * t) : make_float2(0.f, 0.f);
        const float2 y = ok1 ? ldg2(base + off1 + 2 * t) : make_float2(0.f, 0.f);
        a[0][0] = raw(x.x); a[0][1] = raw(y.x); a[0][2] = raw(x.y); a[0][3] = raw(y.y);
    } else {
        const float4 x = ok0 ? ldg4(base + off0 + 4 * t) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 y = ok1 ? ldg4(base + off1 + 4 * t) : make_float4(0.f, 0.f, 0.f, 0.f);
        a[0][0] = raw(x.x); a[0][1] = raw(y.x); a[0][2] = raw(x.y); a[0][3] = raw(y.y);
        a[NT - 1][0] = raw(x.z); a[NT - 1][1] = raw(y.z); a[NT - 1][2] = raw(x.w); a[NT - 1][3] = raw(y.w);
    }
}
// first channel of slot (k = t) of k-step kk under that re-labelling (slot k = t + 4 is the next one)
template <int NT>
__device__ __forceinline__ int small_ch(int t, int kk) { return NT == 1 ? 2 * t : 4 * t + 2 * kk; }

// Labelling of the OUTPUT columns of the products whose result is a wide activation tile (u, da):
// column g of n-tile nt is channel 16*(nt/2) + 4*(g/2) + 2*(nt%2) + g%2 of the 64-wide slice, so the
// accumulators a thread holds for n-tiles 2jj and 2jj+1 (columns 2t, 2t+1 of each) are the four
// CONSECUTIVE channels 16jj + 4t .. +3: one 16-byte store / load per row instead of two 8-byte ones
// (half the L1 wavefronts -- the 8-byte form kept the LSU data pipe at 65 %).
__device__ __forceinline__ int out_col(int nt, int g) {
    return 16 * (nt >> 1) + 4 * (g >> 1) + 2 * (nt & 1) + (g & 1);
}

// A fragments with the contraction over ROWS: the transposed [16 rows][BP] tile, m = channel j (g,
// g + 8), k = row (t, t + 4) of the 8-row half `ks`.  Rows >= rows read as zero.
template <int NT>
__device__ __forceinline__ void load_small_at(uint32_t (&a)[4], const float* __restrict__ base,
                                              long long row0, long long rows, int ks, int g, int t) {
    constexpr int BP = NT * 8;
    const long long ra = row0 + 8 * ks + t, rb = ra + 4;
    const float* pa = base + ra * BP + g;
    const float* pb = base + rb * BP + g;
    const bool oka = ra < rows, okb = rb < rows;
    a[0] = raw(oka ? __ldg(pa) : 0.f);
    a[2] = raw(okb ? __ldg(pb) : 0.f);
    if (NT == 2) {
        a[1] = raw(oka ? __ldg(pa + 8) : 0.f);
        a[3] = raw(okb ? __ldg(pb + 8) : 0.f);
    } else {
        a[1] = 0u;
        a[3] = 0u;
    }
}

// acc[nt] (m = j, n = channel 8nt + g of the 64-wide slice) += A^T-fragment x staging tile
__device__ __forceinline__ void rows_mma(float (&acc)[8][4], const uint32_t (&a)[4],
                                         const float* __restrict__ tile, int ks, int g, int t) {
    const float* r0 = tile + (8 * ks + t) * 64;
    const float* r1 = r0 + 4 * 64;
    const int sw = swz(t);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int col = (8 * nt + g) ^ sw;
        const uint32_t b[2] = {tf(r0[col]), tf(r1[col])};
        mma_m16n8k8(acc[nt], a, b);
    }
}

// ============================================================================ down
struct DownP {
    const float *z, *mean, *scale, *beta, *Wd, *bd;
    float* h1;
    long long rows;
    BnFold fold;                // fold.sum != nullptr: mean / scale come from the raw BN1 sums
};

template <int C, int NT>
__global__ void __launch_bounds__(kT2Threads, 2) tcn2_down_kernel(DownP p) {
    constexpr int BP = NT * 8, CH = C / 16;
    constexpr int HALVES = C > 128 ? 2 : 1, CHH = CH / HALVES;
    extern __shared__ __align__(16) float smem[];
    float* s_mean = smem;
    float* s_scale = smem + C;
    float* s_beta = smem + 2 * C;
    float2* s_w = reinterpret_cast<float2*>(smem + 3 * C);          // [CH][2][NT][32]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    for (int i = tid; i < C; i += kT2Threads) {
        if (p.fold.sum) bn_fold(p.fold, i, blockIdx.x == 0, s_mean[i], s_scale[i]);
        else { s_mean[i] = p.mean[i]; s_scale[i] = p.scale[i]; }
        s_beta[i] = p.beta[i];
    }
    for (int i = tid; i < CH * 2 * NT * 32; i += kT2Threads) {
        const int l = i & 31, q = i >> 5;
        const int nt = q % NT, h = (q / NT) & 1, ch = q / (NT * 2);
        const int c0 = ch * 16 + 4 * (l & 3) + 2 * h, n = nt * 8 + (l >> 2);
        s_w[i] = make_float2(tff(p.Wd[c0 * BP + n]), tff(p.Wd[(c0 + 1) * BP + n]));
    }
    __syncthreads();
    const long long ntiles = (p.rows + 15) >> 4;
    for (long long tile = (long long)blockIdx.x * 8 + warp; tile < ntiles; tile += (long long)gridDim.x * 8) {
        const long long r0 = tile * 16 + g, r1 = r0 + 8;
        const float* z0 = p.z + (r0 < p.rows ? r0 : p.rows - 1) * C + 4 * t;
        const float* z1 = p.z + (r1 < p.rows ? r1 : p.rows - 1) * C + 4 * t;
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll
        for (int half = 0; half < HALVES; ++half) {
            float4 va[CHH], vb[CHH];
#pragma unroll
            for (int j = 0; j < CHH; ++j) {
                va[j] = lds4(z0 + (half * CHH + j) * 16);
                vb[j] = lds4(z1 + (half * CHH + j) * 16);
            }
#pragma unroll
            for (int j = 0; j < CHH; ++j) {
                const int ch = half * CHH + j;
                const float4 mu = *reinterpret_cast<const float4*>(s_mean + ch * 16 + 4 * t);
                const float4 sc = *reinterpret_cast<const float4*>(s_scale + ch * 16 + 4 * t);
                const float4 be = *reinterpret_cast<const float4*>(s_beta + ch * 16 + 4 * t);
                const float x[4] = {fmaxf(bn_apply(va[j].x, mu.x, sc.x, be.x), 0.f),
                                    fmaxf(bn_apply(va[j].y, mu.y, sc.y, be.y), 0.f),
                                    fmaxf(bn_apply(va[j].z, mu.z, sc.z, be.z), 0.f),
                                    fmaxf(bn_apply(va[j].w, mu.w, sc.w, be.w), 0.f)};
                const float y[4] = {fmaxf(bn_apply(vb[j].x, mu.x, sc.x, be.x), 0.f),
                                    fmaxf(bn_apply(vb[j].y, mu.y, sc.y, be.y), 0.f),
                                    fmaxf(bn_apply(vb[j].z, mu.z, sc.z, be.z), 0.f),
                                    fmaxf(bn_apply(vb[j].w, mu.w, sc.w, be.w), 0.f)};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t a[4] = {tf(x[2 * h]), tf(y[2 * h]), tf(x[2 * h + 1]), tf(y[2 * h + 1])};
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma(acc[nt], a, s_w[((ch * 2 + h) * NT + nt) * 32 + lane]);
                }
            }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int c = nt * 8 + 2 * t;
            const float2 b = ldg2(p.bd + c);
            if (r0 < p.rows)
                *reinterpret_cast<float2*>(p.h1 + r0 * BP + c) =
                    make_float2(rnd_tf32(acc[nt][0] + b.x), rnd_tf32(acc[nt][1] + b.y));
            if (r1 < p.rows)
                *reinterpret_cast<float2*>(p.h1 + r1 * BP + c) =
                    make_float2(rnd_tf32(acc[nt][2] + b.x), rnd_tf32(acc[nt][3] + b.y));
        }
    }
}

// ============================================================================ up
struct UpP {
    const float *h2, *Wu, *bu;
    float* u;
    double *ssum, *ssq;
    long long rows;
    int C;
};

template <int NT>
__global__ void __launch_bounds__(kT2Threads, 2) tcn2_up_kernel(UpP p) {
    constexpr int BP = NT * 8;
    extern __shared__ __align__(16) float smem[];
    const int C = p.C, S = C >> 6;
    float2* s_w = reinterpret_cast<float2*>(smem);                  // [S][8 nt][NT kk][32]
    float* s_bu = smem + S * 8 * NT * 64;                           // [C]
    float* s_sum = s_bu + C;                                        // [C]
    float* s_sq = s_sum + C;                                        // [C]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    for (int i = tid; i < S * 8 * NT * 32; i += kT2Threads) {
        const int l = i & 31, q = i >> 5;
        const int kk = q % NT, nt = (q / NT) & 7, sl = q / (NT * 8);
        const int j = small_ch<NT>(l & 3, kk), c = sl * 64 + out_col(nt, l >> 2);
        s_w[i] = make_float2(tff(p.Wu[j * C + c]), tff(p.Wu[(j + 1) * C + c]));
    }
    for (int i = tid; i < C; i += kT2Threads) { s_bu[i] = p.bu[i]; s_sum[i] = 0.f; s_sq[i] = 0.f; }
    __syncthreads();
    const int slice = warp % S, tsub = warp / S, tpc = 8 / S;
    float st[4][8];                              // per 16-channel chunk jj: sums [0..3], squares [4..7]
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int i = 0; i < 8; ++i) st[jj][i] = 0.f;
    const long long ntiles = (p.rows + 15) >> 4;
    for (long long tile = (long long)blockIdx.x * tpc + tsub; tile < ntiles; tile += (long long)gridDim.x * tpc) {
        const long long r0 = tile * 16 + g, r1 = r0 + 8;
        const bool ok0 = r0 < p.rows, ok1 = r1 < p.rows;
        uint32_t a[NT][4];
        load_small_a<NT>(a, p.h2, r0 * BP, r1 * BP, ok0, ok1, t);
        float* u0 = p.u + r0 * C + slice * 64 + 4 * t;
        float* u1 = p.u + r1 * C + slice * 64 + 4 * t;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            // n-tiles 2jj, 2jj+1 under the out_col labelling: this thread's accumulators are the four
            // consecutive channels 16jj + 4t .. +3 of rows g (v0) and g + 8 (v1)
            const float4 bb = *reinterpret_cast<const float4*>(s_bu + slice * 64 + jj * 16 + 4 * t);
            float lo[4] = {bb.x, bb.y, bb.x, bb.y}, hi[4] = {bb.z, bb.w, bb.z, bb.w};
#pragma unroll
            for (int kk = 0; kk < NT; ++kk) {
                mma(lo, a[kk], s_w[((slice * 8 + 2 * jj) * NT + kk) * 32 + lane]);
                mma(hi, a[kk], s_w[((slice * 8 + 2 * jj + 1) * NT + kk) * 32 + lane]);
            }
            float v0[4] = {lo[0], lo[1], hi[0], hi[1]}, v1[4] = {lo[2], lo[3], hi[2], hi[3]};
            if (ok0) st4(u0 + jj * 16, make_float4(v0[0], v0[1], v0[2], v0[3]));
            if (ok1) st4(u1 + jj * 16, make_float4(v1[0], v1[1], v1[2], v1[3]));
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (!ok0) v0[e] = 0.f;
                if (!ok1) v1[e] = 0.f;
                st[jj][e] += v0[e] + v1[e];
                st[jj][4 + e] = fmaf(v0[e], v0[e], fmaf(v1[e], v1[e], st[jj][4 + e]));
            }
        }
    }
    if (p.ssum) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float sv = group_sum_g(st[jj][e]), qv = group_sum_g(st[jj][4 + e]);
                if (g == 0) {
                    atomicAdd(&s_sum[slice * 64 + jj * 16 + 4 * t + e], sv);
                    atomicAdd(&s_sq[slice * 64 + jj * 16 + 4 * t + e], qv);
                }
            }
        __syncthreads();
        for (int c = tid; c < C; c += kT2Threads) {
            atomicAdd(&p.ssum[c], (double)s_sum[c]);
            atomicAdd(&p.ssq[c], (double)s_sq[c]);
        }
    }
}

// ============================================================================ backward: up
struct BwdUpP {
    const float *go, *u, *p2, *m12, *c2, *mean2, *h2, *Wu;
    float *dh2, *dWu, *dbu, *dbeff;
    long long rows;
    int C;
    float drop_p, keep_scale;
    uint64_t seed;
    const unsigned long long* step;
    BnBwdFold fold;             // fold.sg != nullptr: p2 / m12 / c2 come from the raw BN2-backward sums
};

// FOLD: the BN2-backward coefficients come from the raw sums (a separate instantiation: the kernel sits at
// its register cap and the extra prologue code cost the bp = 8 variant 9 % when it shared one body)
template <int NT, bool FOLD>
__global__ void __launch_bounds__(kT2HeavyWarps * 32, 3) tcn2_bwd_up_kernel(BwdUpP p) {
    constexpr int BP = NT * 8, W = kT2HeavyWarps, NTHR = W * 32;
    extern __shared__ __align__(16) float smem[];
    const int C = p.C, S = C >> 6;
    float* s_p = smem;                                              // p2, m12, c2, mean2: [4][C]
    float2* s_w = reinterpret_cast<float2*>(smem + 4 * C);          // [S][4 jj][2 h][NT][32]
    float* s_tile = smem + 4 * C + S * 8 * NT * 64;                 // [W warps][16][64]
    float* s_part = s_tile + W * 16 * 64;                           // [2][W warps][32][NT*4]
    float* s_dW = s_part + 2 * W * 32 * NT * 4;                     // [BP][C]
    float* s_dbu = s_dW + BP * C;                                   // [C]
    float* s_dbe = s_dbu + C;                                       // [BP]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    for (int i = tid; i < C; i += NTHR) {
        if (FOLD) bn_bwd_fold(p.fold, i, blockIdx.x == 0, s_p[i], s_p[C + i], s_p[2 * C + i]);
        else { s_p[i] = p.p2[i]; s_p[C + i] = p.m12[i]; s_p[2 * C + i] = p.c2[i]; }
        s_p[3 * C + i] = p.mean2[i];
    }
    for (int i = tid; i < S * 8 * NT * 32; i += NTHR) {
        const int l = i & 31, q = i >> 5;
        const int nt = q % NT, h = (q / NT) & 1, jj = (q / (NT * 2)) & 3, sl = q / (NT * 8);
        const int j = nt * 8 + (l >> 2), c0 = sl * 64 + jj * 16 + 4 * (l & 3) + 2 * h;
        s_w[i] = make_float2(tff(p.Wu[j * C + c0]), tff(p.Wu[j * C + c0 + 1]));
    }
    for (int i = tid; i < BP * C + C + BP; i += NTHR) s_dW[i] = 0.f;
    __syncthreads();
    const int slice = warp % S, tsub = warp / S, tpc = W / S;
    float* tile_s = s_tile + warp * 16 * 64;
    const uint64_t eseed = effective_seed(p.seed, p.step);
    float accW[8][4], dbu[4][4], dbe[NT][2];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) accW[a][b] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dbu[a][b] = 0.f;
#pragma unroll
    for (int a = 0; a < NT; ++a) dbe[a][0] = dbe[a][1] = 0.f;

    const long long ntiles = (p.rows + 15) >> 4;
    const long long stride_t = (long long)gridDim.x * tpc;
    const long long iters = (ntiles - (long long)blockIdx.x * tpc + stride_t - 1) / stride_t;   // uniform per CTA
    for (long long it = 0; it < iters; ++it) {
        const long long tile = (long long)blockIdx.x * tpc + it * stride_t + tsub;
        const bool active = tile < ntiles;
        float acch[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acch[nt][i] = 0.f;
        if (active) {
            const long long r0 = tile * 16 + g, r1 = r0 + 8;
            const bool ok0 = r0 < p.rows, ok1 = r1 < p.rows;
            const long long o0 = (ok0 ? r0 : p.rows - 1) * C + slice * 64 + 4 * t;
            const long long o1 = (ok1 ? r1 : p.rows - 1) * C + slice * 64 + 4 * t;
            // (a register prefetch of the next tile here, as in tcn2_bwd_down_kernel, measured 10 % SLOWER:
            // 64 more live registers during the weight-gradient MMAs at the 168-register cap)
            float4 g0[4], g1[4], u0[4], u1[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                g0[jj] = lds4(p.go + o0 + jj * 16); u0[jj] = lds4(p.u + o0 + jj * 16);
                g1[jj] = lds4(p.go + o1 + jj * 16); u1[jj] = lds4(p.u + o1 + jj * 16);
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int c = slice * 64 + jj * 16 + 4 * t;
                const float4 pv = *reinterpret_cast<const float4*>(s_p + c);
                const float4 mv = *reinterpret_cast<const float4*>(s_p + C + c);
                const float4 cv = *reinterpret_cast<const float4*>(s_p + 2 * C + c);
                const float4 nv = *reinterpret_cast<const float4*>(s_p + 3 * C + c);
                float ga[4] = {g0[jj].x, g0[jj].y, g0[jj].z, g0[jj].w};
                float gb[4] = {g1[jj].x, g1[jj].y, g1[jj].z, g1[jj].w};
                if (p.drop_p > 0.f) {
                    bool ka[4], kb[4];
                    dropout_keep4(eseed, (uint64_t)((o0 + jj * 16) >> 2), p.drop_p, ka);
                    dropout_keep4(eseed, (uint64_t)((o1 + jj * 16) >> 2), p.drop_p, kb);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        ga[e] = ka[e] ? ga[e] * p.keep_scale : 0.f;
                        gb[e] = kb[e] ? gb[e] * p.keep_scale : 0.f;
                    }
                }
                float da[4], db[4];
                da[0] = bn_back(ga[0], u0[jj].x, pv.x, mv.x, cv.x, nv.x);
                da[1] = bn_back(ga[1], u0[jj].y, pv.y, mv.y, cv.y, nv.y);
                da[2] = bn_back(ga[2], u0[jj].z, pv.z, mv.z, cv.z, nv.z);
                da[3] = bn_back(ga[3], u0[jj].w, pv.w, mv.w, cv.w, nv.w);
                db[0] = bn_back(gb[0], u1[jj].x, pv.x, mv.x, cv.x, nv.x);
                db[1] = bn_back(gb[1], u1[jj].y, pv.y, mv.y, cv.y, nv.y);
                db[2] = bn_back(gb[2], u1[jj].z, pv.z, mv.z, cv.z, nv.z);
                db[3] = bn_back(gb[3], u1[jj].w, pv.w, mv.w, cv.w, nv.w);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (!ok0) da[e] = 0.f;
                    if (!ok1) db[e] = 0.f;
                    dbu[jj][e] += da[e] + db[e];
                }
                // dh2[rows][j] += du[rows][c] Wu[j][c]
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t a[4] = {tf(da[2 * h]), tf(db[2 * h]), tf(da[2 * h + 1]), tf(db[2 * h + 1])};
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        mma(acch[nt], a, s_w[(((slice * 4 + jj) * 2 + h) * NT + nt) * 32 + lane]);
                }
                const int pc = (jj * 16 + 4 * t) ^ swz(g);
                st4(tile_s + g * 64 + pc, make_float4(da[0], da[1], da[2], da[3]));
                st4(tile_s + (g + 8) * 64 + pc, make_float4(db[0], db[1], db[2], db[3]));
            }
            __syncwarp();
            // dWu[j][c] += h2[rows][j] du[rows][c]
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t a[4];
                load_small_at<NT>(a, p.h2, tile * 16, p.rows, ks, g, t);
                rows_mma(accW, a, tile_s, ks, g, t);
            }
            __syncwarp();
        }
        // ---- dh2 tile: sum over the channel slices of the row tile, then store; dbeff partials
        if (S == 1) {
            if (active) {
                const long long r0 = tile * 16 + g, r1 = r0 + 8;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int c = nt * 8 + 2 * t;
                    if (r0 < p.rows)
                        *reinterpret_cast<float2*>(p.dh2 + r0 * BP + c) = make_float2(rnd_tf32(acch[nt][0]), rnd_tf32(acch[nt][1]));
                    if (r1 < p.rows)
                        *reinterpret_cast<float2*>(p.dh2 + r1 * BP + c) = make_float2(rnd_tf32(acch[nt][2]), rnd_tf32(acch[nt][3]));
                    dbe[nt][0] += acch[nt][0] + acch[nt][2];
                    dbe[nt][1] += acch[nt][1] + acch[nt][3];
                }
            }
        } else {
            float* part = s_part + (size_t)(it & 1) * W * 32 * NT * 4;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                st4(part + ((size_t)warp * 32 + lane) * NT * 4 + nt * 4,
                    make_float4(acch[nt][0], acch[nt][1], acch[nt][2], acch[nt][3]));
            __syncthreads();
            if (slice == 0 && active) {
                const long long r0 = tile * 16 + g, r1 = r0 + 8;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    float4 sum = make_float4(acch[nt][0], acch[nt][1], acch[nt][2], acch[nt][3]);
                    for (int sl = 1; sl < S; ++sl) {
                        const float4 o = ld4(part + ((size_t)(warp + sl) * 32 + lane) * NT * 4 + nt * 4);
                        sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w;
                    }
                    const int c = nt * 8 + 2 * t;
                    if (r0 < p.rows)
                        *reinterpret_cast<float2*>(p.dh2 + r0 * BP + c) = make_float2(rnd_tf32(sum.x), rnd_tf32(sum.y));
                    if (r1 < p.rows)
                        *reinterpret_cast<float2*>(p.dh2 + r1 * BP + c) = make_float2(rnd_tf32(sum.z), rnd_tf32(sum.w));
                    dbe[nt][0] += sum.x + sum.z;
                    dbe[nt][1] += sum.y + sum.w;
                }
            }
        }
    }
    // ---- flush: weight / bias gradients -> shared -> one atomic per value and CTA
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = g + 8 * (i >> 1), c = slice * 64 + nt * 8 + 2 * t + (i & 1);
            if (j < BP) atomicAdd(&s_dW[j * C + c], accW[nt][i]);
        }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float v = group_sum_g(dbu[jj][e]);
            if (g == 0) atomicAdd(&s_dbu[slice * 64 + jj * 16 + 4 * t + e], v);
        }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float v = group_sum_g(dbe[nt][e]);
            if (g == 0 && slice == 0) atomicAdd(&s_dbe[nt * 8 + 2 * t + e], v);
        }
    __syncthreads();
    if ((reinterpret_cast<uintptr_t>(p.dWu) & 15) == 0) {          // 16-byte reductions: a quarter of the L2 traffic
        for (int i = tid; i < BP * C / 4; i += NTHR) red_add4(p.dWu + 4 * i, *reinterpret_cast<const float4*>(s_dW + 4 * i));
    } else {
        for (int i = tid; i < BP * C; i += NTHR) atomicAdd(&p.dWu[i], s_dW[i]);
    }
    for (int i = tid; i < C; i += NTHR) atomicAdd(&p.dbu[i], s_dbu[i]);
    for (int i = tid; i < BP; i += NTHR) atomicAdd(&p.dbeff[i], s_dbe[i]);
}

// ============================================================================ backward: down
struct BwdDownP {
    const float *dh1, *z, *mean, *scale, *beta, *rstd, *Wd;
    float *g1, *dWd;
    double *sg, *sgx;
    long long rows;
    int C;
};

template <int NT>
__global__ void __launch_bounds__(kT2HeavyWarps * 32, 3) tcn2_bwd_down_kernel(BwdDownP p) {
    constexpr int BP = NT * 8, W = kT2HeavyWarps, NTHR = W * 32;
    extern __shared__ __align__(16) float smem[];
    const int C = p.C, S = C >> 6;
    float* s_c = smem;                                              // mean, scale, beta, rstd: [4][C]
    float2* s_w = reinterpret_cast<float2*>(smem + 4 * C);          // [S][8 nt][NT kk][32]
    float* s_tile = smem + 4 * C + S * 8 * NT * 64;                 // [W warps][16][64]
    float* s_dW = s_tile + W * 16 * 64;                             // [BP][C]
    float* s_sg = s_dW + BP * C;                                    // [C]
    float* s_sgx = s_sg + C;                                        // [C]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    for (int i = tid; i < C; i += NTHR) {
        s_c[i] = p.mean[i]; s_c[C + i] = p.scale[i]; s_c[2 * C + i] = p.beta[i]; s_c[3 * C + i] = p.rstd[i];
    }
    for (int i = tid; i < S * 8 * NT * 32; i += NTHR) {
        const int l = i & 31, q = i >> 5;
        const int kk = q % NT, nt = (q / NT) & 7, sl = q / (NT * 8);
        const int j = small_ch<NT>(l & 3, kk), c = sl * 64 + out_col(nt, l >> 2);
        s_w[i] = make_float2(tff(p.Wd[c * BP + j]), tff(p.Wd[c * BP + j + 1]));
    }
    for (int i = tid; i < BP * C + 2 * C; i += NTHR) s_dW[i] = 0.f;
    __syncthreads();
    const int slice = warp % S, tsub = warp / S, tpc = W / S;
    float* tile_s = s_tile + warp * 16 * 64;
    float accW[8][4], st[4][8];                  // st: per 16-channel chunk, sum g1 [0..3], sum g1*zhat [4..7]
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) accW[a][b] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) st[a][b] = 0.f;
    const long long ntiles = (p.rows + 15) >> 4;
    const long long stride_t = (long long)gridDim.x * tpc;
    // register prefetch of the next tile's z rows and dh1 fragments (see tcn2_bwd_up_kernel)
    float4 z0[4], z1[4];
    uint32_t a[NT][4];
    auto issue_loads = [&](long long tl) {
        const long long r0 = tl * 16 + g, r1 = r0 + 8;
        const bool ok0 = r0 < p.rows, ok1 = r1 < p.rows;
        const long long o0 = (ok0 ? r0 : p.rows - 1) * C + slice * 64 + 4 * t;
        const long long o1 = (ok1 ? r1 : p.rows - 1) * C + slice * 64 + 4 * t;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) { z0[jj] = lds4(p.z + o0 + jj * 16); z1[jj] = lds4(p.z + o1 + jj * 16); }
        load_small_a<NT>(a, p.dh1, r0 * BP, r1 * BP, ok0, ok1, t);
    };
    if ((long long)blockIdx.x * tpc + tsub < ntiles) issue_loads((long long)blockIdx.x * tpc + tsub);
    for (long long tile = (long long)blockIdx.x * tpc + tsub; tile < ntiles; tile += stride_t) {
        const long long r0 = tile * 16 + g, r1 = r0 + 8;
        const bool ok0 = r0 < p.rows, ok1 = r1 < p.rows;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            float lo[4] = {0.f, 0.f, 0.f, 0.f}, hi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int kk = 0; kk < NT; ++kk) {
                mma(lo, a[kk], s_w[((slice * 8 + 2 * jj) * NT + kk) * 32 + lane]);
                mma(hi, a[kk], s_w[((slice * 8 + 2 * jj + 1) * NT + kk) * 32 + lane]);
            }
            const float d0[4] = {lo[0], lo[1], hi[0], hi[1]}, d1[4] = {lo[2], lo[3], hi[2], hi[3]};
            const int c = slice * 64 + jj * 16 + 4 * t;
            const float4 mu = *reinterpret_cast<const float4*>(s_c + c);
            const float4 sc = *reinterpret_cast<const float4*>(s_c + C + c);
            const float4 be = *reinterpret_cast<const float4*>(s_c + 2 * C + c);
            const float4 rs = *reinterpret_cast<const float4*>(s_c + 3 * C + c);
            const float zz0[4] = {z0[jj].x, z0[jj].y, z0[jj].z, z0[jj].w}, zz1[4] = {z1[jj].x, z1[jj].y, z1[jj].z, z1[jj].w};
            const float mua[4] = {mu.x, mu.y, mu.z, mu.w}, sca[4] = {sc.x, sc.y, sc.z, sc.w};
            const float bea[4] = {be.x, be.y, be.z, be.w}, rsa[4] = {rs.x, rs.y, rs.z, rs.w};
            float a0[4], a1[4], g0[4], g1v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a0[e] = ok0 ? fmaxf(bn_apply(zz0[e], mua[e], sca[e], bea[e]), 0.f) : 0.f;
                a1[e] = ok1 ? fmaxf(bn_apply(zz1[e], mua[e], sca[e], bea[e]), 0.f) : 0.f;
                g0[e] = a0[e] > 0.f ? d0[e] : 0.f;
                g1v[e] = a1[e] > 0.f ? d1[e] : 0.f;
                st[jj][e] += g0[e] + g1v[e];
                st[jj][4 + e] += (g0[e] * (zz0[e] - mua[e]) + g1v[e] * (zz1[e] - mua[e])) * rsa[e];
            }
            if (ok0) st4(p.g1 + r0 * C + c, make_float4(g0[0], g0[1], g0[2], g0[3]));
            if (ok1) st4(p.g1 + r1 * C + c, make_float4(g1v[0], g1v[1], g1v[2], g1v[3]));
            const int pc = (jj * 16 + 4 * t) ^ swz(g);
            st4(tile_s + g * 64 + pc, make_float4(a0[0], a0[1], a0[2], a0[3]));
            st4(tile_s + (g + 8) * 64 + pc, make_float4(a1[0], a1[1], a1[2], a1[3]));
        }
        __syncwarp();
        if (tile + stride_t < ntiles) issue_loads(tile + stride_t);
        // dWd[c][j] += a[rows][c] dh1[rows][j]   (m = j, n = c, k = rows)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t at[4];
            load_small_at<NT>(at, p.dh1, tile * 16, p.rows, ks, g, t);
            rows_mma(accW, at, tile_s, ks, g, t);
        }
        __syncwarp();
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = g + 8 * (i >> 1), c = slice * 64 + nt * 8 + 2 * t + (i & 1);
            if (j < BP) atomicAdd(&s_dW[j * C + c], accW[nt][i]);
        }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float sv = group_sum_g(st[jj][e]), xv = group_sum_g(st[jj][4 + e]);
            if (g == 0) {
                atomicAdd(&s_sg[slice * 64 + jj * 16 + 4 * t + e], sv);
                atomicAdd(&s_sgx[slice * 64 + jj * 16 + 4 * t + e], xv);
            }
        }
    __syncthreads();
    if ((reinterpret_cast<uintptr_t>(p.dWd) & 15) == 0) {          // dWd[c][j .. j+3] in one 16-byte reduction
        for (int i = tid; i < C * BP / 4; i += NTHR) {
            const int c = i / (BP / 4), j = (i - c * (BP / 4)) * 4;
            red_add4(p.dWd + c * BP + j, make_float4(s_dW[j * C + c], s_dW[(j + 1) * C + c], s_dW[(j + 2) * C + c],
                                                     s_dW[(j + 3) * C + c]));
        }
    } else {
        for (int i = tid; i < BP * C; i += NTHR) {
            const int j = i / C, c = i - j * C;
            atomicAdd(&p.dWd[c * BP + j], s_dW[i]);
        }
    }
    for (int c = tid; c < C; c += NTHR) {
        atomicAdd(&p.sg[c], (double)s_sg[c]);
        atomicAdd(&p.sgx[c], (double)s_sgx[c]);
    }
}

template <typename K>
void set_smem2(K kern, size_t bytes) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int grid2(long long units, int per_sm) {
    long long n = (long long)num_sms() * per_sm;
    if (n > units) n = units;
    return (int)(n < 1 ? 1 : n);
}

int check2(const char* who, long long rows, int C, int bp) {
    ISTGCN_REQUIRE(rows >= 0, ISTGCN_E_SHAPE, "%s: negative row count", who);
    ISTGCN_REQUIRE(C == 64 || C == 128 || C == 256, ISTGCN_E_SHAPE, "%s: C=%d unsupported (64, 128, 256)", who, C);
    ISTGCN_REQUIRE(bp == 8 || bp == 16, ISTGCN_E_SHAPE, "%s: padded bottleneck bp=%d must be 8 or 16", who, bp);
    return 0;
}

}  // namespace
}  // namespace istgcn

using namespace istgcn;

static int launch_tcn2_down(const float* z, const float* mean1, const float* scale1, const float* beta1,
                            const float* Wd, const float* bd, float* h1, long long rows, int C, int bp,
                            const BnFold& fold, istgcn_stream_t s);

ISTGCN_API int istgcn_tcn2_down(const float* z, const float* mean1, const float* scale1,
                                const float* beta1, const float* Wd, const float* bd, float* h1,
                                long long rows, int C, int bp, istgcn_stream_t s) {
    ISTGCN_REQUIRE(z && mean1 && scale1 && beta1 && Wd && bd && h1, ISTGCN_E_ARG, "tcn2_down: null pointer");
    return launch_tcn2_down(z, mean1, scale1, beta1, Wd, bd, h1, rows, C, bp, BnFold{}, s);
}

// The same with BatchNorm-1's bookkeeping folded in (training): mean1 / scale1 / rstd1 are OUTPUTS
// derived from the raw sums of z (what istgcn_bn_finalize would have written), running statistics
// updated in place (may be NULL).
ISTGCN_API int istgcn_tcn2_down_bn(const float* z, const double* stat_sum, const double* stat_sumsq,
                                   double count, const float* gamma1, const float* beta1,
                                   float* running_mean, float* running_var, float momentum, float eps,
                                   float* mean1, float* scale1, float* rstd1, const float* Wd,
                                   const float* bd, float* h1, long long rows, int C, int bp,
                                   istgcn_stream_t s) {
    ISTGCN_REQUIRE(z && stat_sum && stat_sumsq && gamma1 && beta1 && mean1 && scale1 && rstd1 && Wd && bd && h1,
                   ISTGCN_E_ARG, "tcn2_down_bn: null pointer");
    ISTGCN_REQUIRE(count > 0 && (running_mean == nullptr) == (running_var == nullptr), ISTGCN_E_ARG,
                   "tcn2_down_bn: empty batch or one running-statistics pointer");
    const BnFold fold{stat_sum, stat_sumsq, 1.0 / count, count > 1 ? count / (count - 1.0) : 1.0, gamma1,
                      running_mean, running_var, momentum, eps, mean1, scale1, rstd1};
    return launch_tcn2_down(z, mean1, scale1, beta1, Wd, bd, h1, rows, C, bp, fold, s);
}

static int launch_tcn2_down(const float* z, const float* mean1, const float* scale1, const float* beta1,
                            const float* Wd, const float* bd, float* h1, long long rows, int C, int bp,
                            const BnFold& fold, istgcn_stream_t s) {
    if (int e = check2("tcn2_down", rows, C, bp)) return e;
    if (rows == 0) return 0;
    DownP p{z, mean1, scale1, beta1, Wd, bd, h1, rows, fold};
    const int nt = bp / 8;
    const size_t smem = sizeof(float) * (3 * C + (C / 16) * 2 * nt * 64);
    const int grid = grid2((rows + 127) / 128, 2);
#define T2_DOWN(CC, NT)                                              \
    set_smem2(tcn2_down_kernel<CC, NT>, smem);                       \
    tcn2_down_kernel<CC, NT><<<grid, kT2Threads, smem, (cudaStream_t)s>>>(p)
    if (C == 64) { if (nt == 1) { T2_DOWN(64, 1); } else { T2_DOWN(64, 2); } }
    else if (C == 128) { if (nt == 1) { T2_DOWN(128, 1); } else { T2_DOWN(128, 2); } }
    else { if (nt == 1) { T2_DOWN(256, 1); } else { T2_DOWN(256, 2); } }
#undef T2_DOWN
    return finish_launch("tcn2_down");
}

ISTGCN_API int istgcn_tcn2_up(const float* h2, const float* Wu, const float* bu, float* u,
                              double* stat_sum, double* stat_sumsq, long long rows, int C, int bp,
                              istgcn_stream_t s) {
    ISTGCN_REQUIRE(h2 && Wu && bu && u, ISTGCN_E_ARG, "tcn2_up: null pointer");
    ISTGCN_REQUIRE((stat_sum == nullptr) == (stat_sumsq == nullptr), ISTGCN_E_ARG, "tcn2_up: one statistics pointer");
    if (int e = check2("tcn2_up", rows, C, bp)) return e;
    if (rows == 0) return 0;
    UpP p{h2, Wu, bu, u, stat_sum, stat_sumsq, rows, C};
    const int nt = bp / 8, S = C / 64;
    const size_t smem = sizeof(float) * (S * 8 * nt * 64 + 3 * C);
    const int grid = grid2(((rows + 15) / 16 + 8 / S - 1) / (8 / S), 2);
    if (nt == 1) {
        set_smem2(tcn2_up_kernel<1>, smem);
        tcn2_up_kernel<1><<<grid, kT2Threads, smem, (cudaStream_t)s>>>(p);
    } else {
        set_smem2(tcn2_up_kernel<2>, smem);
        tcn2_up_kernel<2><<<grid, kT2Threads, smem, (cudaStream_t)s>>>(p);
    }
    return finish_launch("tcn2_up");
}

static int launch_tcn2_bwd_up(const float* go, const float* u, const float* p2, const float* m12,
                              const float* c2, const float* mean2, const float* h2, const float* Wu,
                              float* dh2, float* dWu, float* dbu, float* dbeff, long long rows, int C,
                              int bp, float drop_p, uint64_t drop_seed, const unsigned long long* drop_step,
                              const BnBwdFold& fold, istgcn_stream_t s);

ISTGCN_API int istgcn_tcn2_bwd_up(const float* go, const float* u, const float* p2, const float* m12,
                                  const float* c2, const float* mean2, const float* h2,
                                  const float* Wu, float* dh2, float* dWu, float* dbu, float* dbeff,
                                  long long rows, int C, int bp, float drop_p, uint64_t drop_seed,
                                  const unsigned long long* drop_step, istgcn_stream_t s) {
    ISTGCN_REQUIRE(go && u && p2 && m12 && c2 && mean2 && h2 && Wu && dh2 && dWu && dbu && dbeff,
                   ISTGCN_E_ARG, "tcn2_bwd_up: null pointer");
    return launch_tcn2_bwd_up(go, u, p2, m12, c2, mean2, h2, Wu, dh2, dWu, dbu, dbeff, rows, C, bp, drop_p,
                              drop_seed, drop_step, BnBwdFold{}, s);
}

// The same with istgcn_bn_bwd_coeffs folded in: p2 / m12 / c2 (and dgamma2 / dbeta2, may be NULL) are
// OUTPUTS derived from the raw sums sg = sum gy, sgx = sum gy * uhat over `count` rows.
ISTGCN_API int istgcn_tcn2_bwd_up_bn(const float* go, const float* u, const double* sg, const double* sgx,
                                     double count, const float* gamma2, const float* rstd2, float* p2,
                                     float* m12, float* c2, float* dgamma2, float* dbeta2,
                                     const float* mean2, const float* h2, const float* Wu, float* dh2,
                                     float* dWu, float* dbu, float* dbeff, long long rows, int C, int bp,
                                     float drop_p, uint64_t drop_seed, const unsigned long long* drop_step,
                                     istgcn_stream_t s) {
    ISTGCN_REQUIRE(go && u && sg && sgx && gamma2 && rstd2 && p2 && m12 && c2 && mean2 && h2 && Wu && dh2 &&
                       dWu && dbu && dbeff,
                   ISTGCN_E_ARG, "tcn2_bwd_up_bn: null pointer");
    ISTGCN_REQUIRE(count > 0, ISTGCN_E_ARG, "tcn2_bwd_up_bn: empty batch");
    const BnBwdFold fold{sg, sgx, 1.0 / count, gamma2, rstd2, p2, m12, c2, dgamma2, dbeta2};
    return launch_tcn2_bwd_up(go, u, p2, m12, c2, mean2, h2, Wu, dh2, dWu, dbu, dbeff, rows, C, bp, drop_p,
                              drop_seed, drop_step, fold, s);
}

static int launch_tcn2_bwd_up(const float* go, const float* u, const float* p2, const float* m12,
                              const float* c2, const float* mean2, const float* h2, const float* Wu,
                              float* dh2, float* dWu, float* dbu, float* dbeff, long long rows, int C,
                              int bp, float drop_p, uint64_t drop_seed, const unsigned long long* drop_step,
                              const BnBwdFold& fold, istgcn_stream_t s) {
    ISTGCN_REQUIRE(drop_p >= 0.f && drop_p < 1.f, ISTGCN_E_ARG, "tcn2_bwd_up: dropout p=%f", drop_p);
    if (int e = check2("tcn2_bwd_up", rows, C, bp)) return e;
    if (rows == 0) return 0;
    BwdUpP p{go, u, p2, m12, c2, mean2, h2, Wu, dh2, dWu, dbu, dbeff, rows, C, drop_p,
             1.f / (1.f - drop_p), drop_seed, drop_step, fold};
    const int nt = bp / 8, S = C / 64;
    constexpr int W = kT2HeavyWarps;
    const size_t smem = sizeof(float) * (4 * C + S * 8 * nt * 64 + W * 16 * 64 + 2 * W * 32 * nt * 4 +
                                         bp * C + C + bp);
    const int grid = grid2(((rows + 15) / 16 + W / S - 1) / (W / S), 3);
#define T2_BWD_UP(NT, FOLD)                                            \
    set_smem2(tcn2_bwd_up_kernel<NT, FOLD>, smem);                     \
    tcn2_bwd_up_kernel<NT, FOLD><<<grid, W * 32, smem, (cudaStream_t)s>>>(p)
    if (nt == 1) { if (fold.sg) { T2_BWD_UP(1, true); } else { T2_BWD_UP(1, false); } }
    else { if (fold.sg) { T2_BWD_UP(2, true); } else { T2_BWD_UP(2, false); } }
#undef T2_BWD_UP
    return finish_launch("tcn2_bwd_up");
}

ISTGCN_API int istgcn_tcn2_bwd_down(const float* dh1, const float* z, const float* mean1,
                                    const float* scale1, const float* beta1, const float* rstd1,
                                    const float* Wd, float* g1, float* dWd, double* sg1, double* sg1x,
                                    long long rows, int C, int bp, istgcn_stream_t s) {
    ISTGCN_REQUIRE(dh1 && z && mean1 && scale1 && beta1 && rstd1 && Wd && g1 && dWd && sg1 && sg1x,
                   ISTGCN_E_ARG, "tcn2_bwd_down: null pointer");
    if (int e = check2("tcn2_bwd_down", rows, C, bp)) return e;
    if (rows == 0) return 0;
    BwdDownP p{dh1, z, mean1, scale1, beta1, rstd1, Wd, g1, dWd, sg1, sg1x, rows, C};
    const int nt = bp / 8, S = C / 64;
    constexpr int W = kT2HeavyWarps;
    const size_t smem = sizeof(float) * (4 * C + S * 8 * nt * 64 + W * 16 * 64 + bp * C + 2 * C);
    const int grid = grid2(((rows + 15) / 16 + W / S - 1) / (W / S), 3);
    if (nt == 1) {
        set_smem2(tcn2_bwd_down_kernel<1>, smem);
        tcn2_bwd_down_kernel<1><<<grid, W * 32, smem, (cudaStream_t)s>>>(p);
    } else {
        set_smem2(tcn2_bwd_down_kernel<2>, smem);
        tcn2_bwd_down_kernel<2><<<grid, W * 32, smem, (cudaStream_t)s>>>(p);
    }
    return finish_launch("tcn2_bwd_down");
}
