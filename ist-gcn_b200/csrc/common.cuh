// Shared device helpers for the istgcn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/istgcn_b200.h"

#define ISTGCN_API extern "C" __attribute__((visibility("default")))

namespace istgcn {

void set_error(const char* fmt, ...);
int finish_launch(const char* what);   // cudaGetLastError -> error text, returns code

#define ISTGCN_REQUIRE(cond, code, ...)            \
    do {                                           \
        if (!(cond)) {                             \
            ::istgcn::set_error(__VA_ARGS__);      \
            return (code);                         \
        }                                          \
    } while (0)

constexpr int kThreads = 256;     // every tiled kernel runs 8 warps
constexpr int kWarps = 8;
constexpr int kMaxNnz = 1024;     // non-zeros of A_eff (197 NTU-sym, 152 OpenPose-sym)
constexpr int kMaxKV = 4 * 32;    // K*V upper bound for the CSR pointer arrays
constexpr int kTileRows = 128;    // rows of a frame tile (floor(128 / V) frames)

__host__ __device__ inline int frames_per_tile(int V) { return kTileRows / V; }

int num_sms();

// ---------------------------------------------------------------- tensor-core primitives
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// Round-to-nearest TF32 for the fast mode in ONE integer add: the tensor core ignores the low 13
// mantissa bits, so adding half a TF32 ulp to the fp32 bit pattern turns its truncation into
// rounding.  (cvt.rna.tf32.f32 is emulated on sm_100a: FSETP + IADD + LOP3 + a constant move per
// value, six values per mma -- a quarter of the TCN kernels' instructions.  Inf/NaN inputs are not
// special-cased here; the 3xTF32 parity mode keeps the exact conversion.)
__device__ __forceinline__ uint32_t to_tf32_fast(float x) { return __float_as_uint(x) + 0x1000u; }

// D(16x8) += A(16x8, row) * B(8x8, col), TF32 inputs, fp32 accumulate (legacy mma.sync path;
// used for the skinny GEMMs whose N or M is the 8..16-wide bottleneck).
__device__ __forceinline__ void mma_m16n8k8(float (&d)[4], const uint32_t (&a)[4],
                                            const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Split x into a TF32 head and a TF32 tail (x ~= hi + lo to ~21 mantissa bits).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}

// Warp-level tile product on shared-memory operands.
//   acc[mt][nt] (16x8 each) += A[16*MT x kdim] * B[kdim x 8*NT]
//   A element (m, k): A_T ? A[k*lda + m] : A[m*lda + k]      (A points at the warp's row 0)
//   B element (k, n): B_T ? B[n*ldb + k] : B[k*ldb + n]      (B points at the warp's column 0)
// Bank-conflict-free fragment loads: row-indexed-by-g operands (A, B_T) need ld % 8 == 4,
// row-indexed-by-t operands (A_T, B) need ld % 16 == 8.
// PRECISE = 3xTF32 error compensation (small cross terms first).
template <int MT, int NT, bool A_T, bool B_T, bool PRECISE>
__device__ __forceinline__ void warp_mma_raw(float (&acc)[MT][NT][4], const float* __restrict__ A,
                                             int lda, const float* __restrict__ B, int ldb,
                                             int kdim, int lane) {
    const int g = lane >> 2, t = lane & 3;
    for (int k0 = 0; k0 < kdim; k0 += 8) {
        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = mt * 16 + g + 8 * (i & 1);
                const int k = k0 + t + 4 * (i >> 1);
                const float v = A_T ? A[k * lda + m] : A[m * lda + k];
                if (PRECISE) split_tf32(v, ah[mt][i], al[mt][i]);
                else ah[mt][i] = to_tf32_fast(v);
            }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            uint32_t bh[2], bl[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int k = k0 + t + 4 * i;
                const int n = nt * 8 + g;
                const float v = B_T ? B[n * ldb + k] : B[k * ldb + n];
                if (PRECISE) split_tf32(v, bh[i], bl[i]);
                else bh[i] = to_tf32_fast(v);
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                if (PRECISE) {
                    mma_m16n8k8(acc[mt][nt], al[mt], bh);
                    mma_m16n8k8(acc[mt][nt], ah[mt], bl);
                }
                mma_m16n8k8(acc[mt][nt], ah[mt], bh);
            }
        }
    }
}

// The tensor core adds products into the accumulator with truncation, so a long K loop drifts
// (~K * 2^-24).  In PRECISE mode every 32-deep slice accumulates into a fresh fragment that is
// then added to the running sum with ordinary round-to-nearest fp32 adds.
template <int MT, int NT, bool A_T, bool B_T, bool PRECISE>
__device__ __forceinline__ void warp_mma(float (&acc)[MT][NT][4], const float* __restrict__ A,
                                         int lda, const float* __restrict__ B, int ldb,
                                         int kdim, int lane) {
    if (PRECISE) {
        for (int k0 = 0; k0 < kdim; k0 += 32) {
            float part[MT][NT][4];
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int b = 0; b < NT; ++b)
#pragma unroll
                    for (int c = 0; c < 4; ++c) part[a][b][c] = 0.f;
            const int kk = kdim - k0 < 32 ? kdim - k0 : 32;
            warp_mma_raw<MT, NT, A_T, B_T, true>(part, A_T ? A + k0 * lda : A + k0, lda,
                                                 B_T ? B + k0 : B + k0 * ldb, ldb, kk, lane);
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int b = 0; b < NT; ++b)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[a][b][c] += part[a][b][c];
        }
    } else {
        warp_mma_raw<MT, NT, A_T, B_T, false>(acc, A, lda, B, ldb, kdim, lane);
    }
}

// one 16x8x8 step on already-loaded fp32 fragment values
template <bool PRECISE>
__device__ __forceinline__ void mma_step(float (&acc)[4], const float (&a)[4], const float (&b)[2]) {
    uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (PRECISE) split_tf32(a[i], ah[i], al[i]);
        else ah[i] = to_tf32_fast(a[i]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        if (PRECISE) split_tf32(b[i], bh[i], bl[i]);
        else bh[i] = to_tf32_fast(b[i]);
    }
    if (PRECISE) {
        float part[4] = {0.f, 0.f, 0.f, 0.f};
        mma_m16n8k8(part, al, bh);
        mma_m16n8k8(part, ah, bl);
        mma_m16n8k8(part, ah, bh);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += part[i];
    } else {
        mma_m16n8k8(acc, ah, bh);
    }
}

template <int MT, int NT>
__device__ __forceinline__ void zero_acc(float (&acc)[MT][NT][4]) {
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < NT; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
}

// ---------------------------------------------------------------- reductions / misc
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the 8 "g" lanes that share the same (lane & 3): xor 4, 8, 16
__device__ __forceinline__ float group_sum_g(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}

__device__ __forceinline__ double group_sum_g(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}

// Counter-based dropout keep decisions: ONE 64-bit mix per group of four consecutive elements
// (every kernel walks float4 groups), 16 bits of it per element; identical in the forward and
// backward kernels (nn.Dropout's Philox stream cannot be reproduced from a custom kernel; parity
// tests read the mask back through istgcn_dropout_mask).  The per-element 64-bit mix this
// replaces was ~30 integer instructions per element -- 25 % of tcn_bwd_up's instruction count.
__device__ __forceinline__ uint64_t dropout_bits(uint64_t seed, uint64_t group) {
    uint64_t x = seed + group * 0x9E3779B97F4A7C15ull;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
__device__ __forceinline__ uint32_t dropout_threshold(float p) {      // keep iff 16 bits >= threshold
    return (uint32_t)(p * 65536.0f + 0.5f);
}
// keep[j] for elements 4*group + j
__device__ __forceinline__ void dropout_keep4(uint64_t seed, uint64_t group, float p, bool (&keep)[4]) {
    const uint64_t x = dropout_bits(seed, group);
    const uint32_t thr = dropout_threshold(p);
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    keep[0] = (lo & 0xFFFFu) >= thr; keep[1] = (lo >> 16) >= thr;
    keep[2] = (hi & 0xFFFFu) >= thr; keep[3] = (hi >> 16) >= thr;
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t idx, float p) {
    const uint64_t x = dropout_bits(seed, idx >> 2);
    return (uint32_t)((x >> (16 * (idx & 3))) & 0xFFFFu) >= dropout_threshold(p);
}

// BatchNorm forward on one element, cancellation-free: (x - mean) * scale + beta
__device__ __forceinline__ float bn_apply(float x, float mean, float scale, float beta) {
    return fmaf(x - mean, scale, beta);
}
// BatchNorm backward on one element: p * ((g - m1) - c * (x - mean)), with p = gamma*rstd,
// m1 = mean(g), c = rstd * mean(g * xhat).  Written with explicit differences: when g is
// nearly constant over the batch (e.g. right behind a global average pool) g - m1 is exact,
// whereas a precomputed p*g + q*x + r form carries a systematic 2^-24 * |p*m1| bias per channel.
__device__ __forceinline__ float bn_back(float g, float x, float p, float m1, float c, float mean) {
    return p * ((g - m1) - c * (x - mean));
}

// ---- BatchNorm bookkeeping folded into the consumer kernel (the one-block bn_finalize /
// bn_bwd_coeffs launches cost ~6 us each with their gaps, ~75 of them per training step).
// Every CTA of the consumer derives the per-channel coefficients it needs from the raw sums; the
// threads named by `write` also store them (for the backward pass / later consumers) and update the
// running statistics.  sum == nullptr: the consumer takes ready-made coefficients as before.
struct BnFold {                 // training forward: y = (x - mean) * scale + beta
    const double *sum, *sumsq;
    double inv_count, unbias;   // 1 / count, count / (count - 1)
    const float* gamma;
    float *running_mean, *running_var;
    float momentum, eps;
    float *mean_out, *scale_out, *rstd_out;
};
__device__ __forceinline__ void bn_fold(const BnFold& f, int c, bool write, float& mean, float& scale) {
    const double mu = f.sum[c] * f.inv_count;
    double var = fma(-mu, mu, f.sumsq[c] * f.inv_count);
    if (var < 0) var = 0;
    const float v = (float)var + f.eps;
    float rs = rsqrtf(v);
    rs = rs * (1.5f - 0.5f * v * rs * rs);                    // one Newton step: float-exact
    mean = (float)mu;
    scale = f.gamma[c] * rs;
    if (write) {
        f.mean_out[c] = mean;
        f.scale_out[c] = scale;
        f.rstd_out[c] = rs;
        if (f.running_mean) {
            f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mean;
            f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)(var * f.unbias);
        }
    }
}
struct BnBwdFold {              // backward: dx = p * ((g - m1) - c * (x - mean))
    const double *sg, *sgx;     // sum g, sum g * xhat
    double inv_count;
    const float *gamma, *rstd;
    float *p_out, *m1_out, *c_out, *dgamma, *dbeta;
};
__device__ __forceinline__ void bn_bwd_fold(const BnBwdFold& f, int c, bool write, float& p, float& m1,
                                            float& cc) {
    const double sg = f.sg[c], sgx = f.sgx[c];
    const float rs = f.rstd[c];
    p = f.gamma[c] * rs;
    m1 = (float)(sg * f.inv_count);
    cc = (float)((double)rs * (sgx * f.inv_count));
    if (write) {
        f.p_out[c] = p;
        f.m1_out[c] = m1;
        f.c_out[c] = cc;
        if (f.dgamma) f.dgamma[c] = (float)sgx;
        if (f.dbeta) f.dbeta[c] = (float)sg;
    }
}

// Dropout seed = host value mixed with an optional device-resident step counter, so that a
// captured CUDA graph draws a fresh mask on every replay.
__device__ __forceinline__ uint64_t effective_seed(uint64_t seed, const unsigned long long* step) {
    return step ? seed + static_cast<uint64_t>(*step) * 0xD1B54A32D192ED03ull : seed;
}

// one 16-byte reduction into global memory (four scalar atomicAdds cost four L2 transactions)
__device__ __forceinline__ void red_add4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ float4 ld4(const float* p) {
    return *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ void st4(float* p, const float4& v) {
    *reinterpret_cast<float4*>(p) = v;
}

// Temporal stride of the residual branch (st_gcnold.py:186-193): output frame f = n*t_out + to
// reads input frame n*t_in + to*stride.  t_out == 0 means "same frames" (the graph convolution).
// Temporal row map of a strided / shifted 1x1 convolution: frame `to` of the t_out-frame side
// pairs with frame to*stride + offset of the t_in-frame side (the residual branch uses offset 0;
// the taps of a temporal convolution use offset = tap - pad).  A frame outside [0, t_in) is the
// zero padding: map_row returns -1 and the caller reads zeros / skips the store.
struct FrameMap {
    int t_in, t_out, stride, offset;
};
__device__ __forceinline__ long long map_row(const FrameMap& m, long long row, int V) {
    if (m.t_out == 0) return row;
    const long long f = row / V;
    const int v = (int)(row - f * V);
    const long long n = f / m.t_out;
    const int to = (int)(f - n * m.t_out);
    const int ti = to * m.stride + m.offset;
    if (ti < 0 || ti >= m.t_in) return -1;
    return (n * m.t_in + ti) * V + v;
}

// Cooperative tile staging with memory-level parallelism: every thread first issues COUNT
// independent 16-byte loads (items tid + u*kThreads, u < COUNT, offset by `base`) and only then
// stores them, so COUNT*kThreads*16 bytes are in flight per CTA instead of one load per thread.
// ld(i) -> float4 (returns zeros / the transformed value for item i), st(i, v).
template <int COUNT, typename LD, typename ST>
__device__ __forceinline__ void stage4(int tid, int base, LD ld, ST st) {
    float4 v[COUNT];
#pragma unroll
    for (int u = 0; u < COUNT; ++u) v[u] = ld(base + tid + u * kThreads);
#pragma unroll
    for (int u = 0; u < COUNT; ++u) st(base + tid + u * kThreads, v[u]);
}

}  // namespace istgcn
