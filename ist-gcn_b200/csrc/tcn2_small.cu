// The two temporal kernels of the fast-mode Inception TCN (csrc/tcn2.cu): the merged 15-tap
// convolution over the bp-wide bottleneck tensors (net/st_gcn_mstcn_1x1.py:197-217, 257-261) and
// its backward.  The tensors are 1/8 .. 1/16 of an activation and L2-resident, so these kernels
// are bound by instruction issue, not by HBM: the first version (rows gathered from global memory
// per tap with predicated loads and 64-bit addressing) spent ~530 instructions per 16-row tile for
// 15 MMAs.  Here a CTA stages the halo of one (sample, 16-frame) tile in shared memory with plain
// coalesced 16-byte copies (zero frames = the temporal padding; for the stride-2 backward the
// staged array is the zero-upsampled gradient, so every tap reads unconditionally) and the warps
// read their mma.sync fragments from it with one shared-memory load per operand register.
//   tcn2_conv      h2[(n,to,v)] = sum_tap Weff[tap]^T h1[(n, to*s + tap - 7, v)] + beff
//   tcn2_bwd_conv  dh1[(n,ti,v)] = sum_tap Weff[tap] dh2[(n, (ti + 7 - tap)/s, v)]     (data kernel)
//                  dWeff[tap][ci][co] += sum_rows h1[(n, to*s + tap - 7, v)][ci] dh2[(n,to,v)][co]
//                  dbd[ci] += sum_rows dh1                                             (weight kernel)
// All four bottleneck tensors are written pre-rounded to TF32 by their producers (they are only
// ever tensor-core operands), so no conversion instruction sits in front of any MMA here.
#include "common.cuh"

namespace istgcn {
namespace {

constexpr int kTaps = 15, kHalf = 7, kTT = 16, kThr = 256;

__device__ __forceinline__ float rnd_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void mma_f(float (&d)[4], const float (&a)[4], float b0, float b1) {
    const uint32_t aa[4] = {__float_as_uint(a[0]), __float_as_uint(a[1]), __float_as_uint(a[2]), __float_as_uint(a[3])};
    const uint32_t bb[2] = {__float_as_uint(b0), __float_as_uint(b1)};
    mma_m16n8k8(d, aa, bb);
}

struct SmallP {
    const float* in;     // data kernel FWD: h1 (Tin = T frames); BWD: dh2 (Tin = Tout frames)
    const float* in2;    // weight kernel: dh2 [NM][Tf][V][BP]
    const float* W;      // Weff [15][BP(in)][BP(out)]
    const float* bias;   // FWD: beff, BWD: nullptr
    float* out;          // FWD: h2, BWD: dh1  ([NM][Tf][V][BP]);  weight kernel: dWeff
    float* colsum;       // BWD: dbd[BP] += column sums of out; else nullptr
    int NM, Tin, Tf, V, stride, inv16;
    unsigned inv_per;    // ceil(2^32 / (V*BP/4))
};

// Stage Q frames of sample n into shared memory; staged frame q comes from source frame src_of(q)
// (-1 = zeros: temporal padding / the zero frames of the upsampled stride-2 gradient).  The copy is
// flat over the 16-byte elements of the staged array -- element -> (frame, offset) by a multiply-high
// with the precomputed reciprocal of the frame size -- and every thread first issues a batch of
// independent loads, then the stores: one L2 round trip per batch instead of one per element.
template <int BP, typename F>
__device__ __forceinline__ void stage_frames(float* __restrict__ dst, const float* __restrict__ in,
                                             long long n, int Tin, int V, int Q, unsigned inv_per,
                                             F src_of, int tid) {
    constexpr int kBatch = 6;
    const int per = V * BP / 4, total = Q * per;
    const float4* base = reinterpret_cast<const float4*>(in + (n * Tin) * V * BP);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i0 = tid; i0 < total; i0 += kBatch * kThr) {
        float4 v[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int idx = i0 + k * kThr;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < total) {
                const int q = (int)__umulhi((unsigned)idx, inv_per);
                const int src = src_of(q);
                if (src >= 0) v[k] = __ldg(base + src * per + (idx - q * per));
            }
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int idx = i0 + k * kThr;
            if (idx < total) d4[idx] = v[k];
        }
    }
}

// ---------------------------------------------------------------------------- data kernel
// VV = number of joints as a compile-time constant (25 NTU, 18 Kinetics; 0 = run-time): the tap
// offsets of the fragment loads become immediates instead of one multiply-add per tap and row.
template <int NT, bool BWD, int VV>
__global__ void __launch_bounds__(kThr, 3) tcn2_small_conv_kernel(SmallP p) {
    constexpr int BP = NT * 8;
    extern __shared__ __align__(16) float smem[];
    float2* s_w = reinterpret_cast<float2*>(smem);                  // [15][NT kk][NT nt][32]
    float* s_col = smem + kTaps * NT * NT * 64;                     // [BP] (+ padding to 16 floats)
    float* s_in = s_col + 16;                                       // [Q][V][BP]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int V = VV ? VV : p.V, s = p.stride;
    const int sq = BWD ? 1 : s;
    const int Q = BWD ? kTT + 2 * kHalf : (kTT - 1) * s + kTaps;
    for (int i = tid; i < kTaps * NT * NT * 32; i += kThr) {
        const int l = i & 31, q = i >> 5;
        const int nt = q % NT, kk = (q / NT) % NT, tp = q / (NT * NT);
        const int ksel = NT == 1 ? 2 * (l & 3) : 4 * (l & 3) + 2 * kk, nsel = nt * 8 + (l >> 2);
        float w0, w1;
        if (!BWD) {     // B(k = ci, n = co) = Weff[tap][ci][co]
            w0 = p.W[(tp * BP + ksel) * BP + nsel];
            w1 = p.W[(tp * BP + ksel + 1) * BP + nsel];
        } else {        // staged index q = f + tap' reads tap = 14 - tap';  B(k = co, n = ci) = Weff[tap][ci][co]
            const int tap = kTaps - 1 - tp;
            w0 = p.W[(tap * BP + nsel) * BP + ksel];
            w1 = p.W[(tap * BP + nsel) * BP + ksel + 1];
        }
        s_w[i] = make_float2(rnd_tf32(w0), rnd_tf32(w1));
    }
    if (tid < 16) s_col[tid] = 0.f;
    float col[NT][2];
#pragma unroll
    for (int a = 0; a < NT; ++a) col[a][0] = col[a][1] = 0.f;
    const int tps = (p.Tf + kTT - 1) / kTT;
    const int items = p.NM * tps;
    const int tapstep = V * BP;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int n = item / tps;
        const int f0 = (item - n * tps) * kTT;
        const int nfr = p.Tf - f0 < kTT ? p.Tf - f0 : kTT;
        const int valid = nfr * V;
        __syncthreads();                         // the previous tile's readers are done (and s_w is filled)
        if (!BWD) {
            const int first = f0 * s - kHalf;
            stage_frames<BP>(s_in, p.in, n, p.Tin, V, Q, p.inv_per,
                             [&](int q) { const int f = first + q; return f >= 0 && f < p.Tin ? f : -1; }, tid);
        } else {
            const int first = f0 - kHalf;        // staged frame q <-> num = to*s = first + q
            stage_frames<BP>(s_in, p.in, n, p.Tin, V, Q, p.inv_per,
                             [&](int q) {
                                 const int num = first + q;
                                 if (num < 0 || (s == 2 && (num & 1))) return -1;
                                 const int to = s == 2 ? num >> 1 : num;
                                 return to < p.Tin ? to : -1;
                             }, tid);
        }
        __syncthreads();
        for (int wt = warp; wt * 16 < valid; wt += kThr / 32) {
            const int rl0 = wt * 16 + g, rl1 = rl0 + 8;
            const int rc0 = rl0 < valid ? rl0 : valid - 1, rc1 = rl1 < valid ? rl1 : valid - 1;
            const int fl0 = (rc0 * p.inv16) >> 16, fl1 = (rc1 * p.inv16) >> 16;
            const float* a0p = s_in + ((fl0 * sq) * V + (rc0 - fl0 * V)) * BP + (NT == 1 ? 2 : 4) * t;
            const float* a1p = s_in + ((fl1 * sq) * V + (rc1 - fl1 * V)) * BP + (NT == 1 ? 2 : 4) * t;
            // two accumulator sets (even / odd taps): the dependent-MMA chain is the critical path
            float acc[NT][4], acc2[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[nt][i] = acc2[nt][i] = 0.f;
#pragma unroll
            for (int tap = 0; tap < kTaps; ++tap) {
                float (&ac)[NT][4] = (tap & 1) ? acc2 : acc;
                if (NT == 1) {
                    const float2 x = *reinterpret_cast<const float2*>(a0p + tap * tapstep);
                    const float2 y = *reinterpret_cast<const float2*>(a1p + tap * tapstep);
                    const float a[4] = {x.x, y.x, x.y, y.y};
                    const float2 w = s_w[tap * 32 + lane];
                    mma_f(ac[0], a, w.x, w.y);
                } else {
                    const float4 x = *reinterpret_cast<const float4*>(a0p + tap * tapstep);
                    const float4 y = *reinterpret_cast<const float4*>(a1p + tap * tapstep);
                    const float a[2][4] = {{x.x, y.x, x.y, y.y}, {x.z, y.z, x.w, y.w}};
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) {
                            const float2 w = s_w[((tap * NT + kk) * NT + nt) * 32 + lane];
                            mma_f(ac[nt], a[kk], w.x, w.y);
                        }
                }
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[nt][i] += acc2[nt][i];
            float* o0 = p.out + (((long long)n * p.Tf + f0) * V + rl0) * BP + 2 * t;
            float* o1 = o0 + 8 * BP;
            const bool ok0 = rl0 < valid, ok1 = rl1 < valid;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                float2 b = make_float2(0.f, 0.f);
                if (!BWD) b = __ldg(reinterpret_cast<const float2*>(p.bias + nt * 8 + 2 * t));
                const float v00 = acc[nt][0] + b.x, v01 = acc[nt][1] + b.y;
                const float v10 = acc[nt][2] + b.x, v11 = acc[nt][3] + b.y;
                if (ok0) *reinterpret_cast<float2*>(o0 + nt * 8) = make_float2(rnd_tf32(v00), rnd_tf32(v01));
                if (ok1) *reinterpret_cast<float2*>(o1 + nt * 8) = make_float2(rnd_tf32(v10), rnd_tf32(v11));
                if (BWD) {
                    col[nt][0] += (ok0 ? v00 : 0.f) + (ok1 ? v10 : 0.f);
                    col[nt][1] += (ok0 ? v01 : 0.f) + (ok1 ? v11 : 0.f);
                }
            }
        }
    }
    if (BWD && p.colsum) {
        __syncthreads();
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float v = group_sum_g(col[nt][e]);
                if (g == 0) atomicAdd(&s_col[nt * 8 + 2 * t + e], v);
            }
        __syncthreads();
        if (tid < BP) atomicAdd(&p.colsum[tid], s_col[tid]);
    }
}

// ---------------------------------------------------------------------------- weight kernel
// m = ci, n = co, k = output rows.  Warps 0..3 take taps 0..7, warps 4..7 taps 8..14; inside a group
// the four warps split the 16-row tiles of the staged (sample, 16-frame) tile.
template <int NT, int VV>
__global__ void __launch_bounds__(kThr, 2) tcn2_small_dw_kernel(SmallP p) {
    constexpr int BP = NT * 8;
    extern __shared__ __align__(16) float smem[];
    const int V = VV ? VV : p.V, s = p.stride;
    const int Q = (kTT - 1) * s + kTaps;
    float* s_dW = smem;                                             // [15][BP][BP]
    float* s_d = s_dW + kTaps * BP * BP;                            // [kTT][V][BP]   dh2 tile
    float* s_h = s_d + kTT * V * BP;                                // [Q][V][BP]     h1 halo
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    for (int i = tid; i < kTaps * BP * BP; i += kThr) s_dW[i] = 0.f;
    const int tg = warp >> 2, wq = warp & 3;
    float acc[8][NT][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < NT; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
    const int tps = (p.Tf + kTT - 1) / kTT;
    const int items = p.NM * tps;
    const int tapstep = V * BP;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int n = item / tps;
        const int f0 = (item - n * tps) * kTT;
        const int nfr = p.Tf - f0 < kTT ? p.Tf - f0 : kTT;
        const int valid = nfr * V;
        __syncthreads();
        {
            const int first = f0 * s - kHalf;
            stage_frames<BP>(s_h, p.in, n, p.Tin, V, Q, p.inv_per,
                             [&](int q) { const int f = first + q; return f >= 0 && f < p.Tin ? f : -1; }, tid);
            stage_frames<BP>(s_d, p.in2, n, p.Tf, V, kTT, p.inv_per,
                             [&](int q) { return q < nfr ? f0 + q : -1; }, tid);
        }
        __syncthreads();
        for (int wt = wq; wt * 16 < valid; wt += 4) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const float* ap[2];
                float b[NT][2];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    int rl = wt * 16 + 8 * ks + t + 4 * hh;
                    if (rl > kTT * V - 1) rl = kTT * V - 1;          // rows >= valid read zeros from s_d
                    const int fl = (rl * p.inv16) >> 16;
                    ap[hh] = s_h + ((fl * s) * V + (rl - fl * V)) * BP + g;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) b[nt][hh] = s_d[rl * BP + nt * 8 + g];
                }
                if (NT == 2) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int tap = tg * 8 + j;                  // warp-uniform
                        if (tap < kTaps) {
                            const float a[4] = {ap[0][tap * tapstep], ap[0][tap * tapstep + 8],
                                                ap[1][tap * tapstep], ap[1][tap * tapstep + 8]};
#pragma unroll
                            for (int nt = 0; nt < NT; ++nt) mma_f(acc[j][nt], a, b[nt][0], b[nt][1]);
                        }
                    }
                } else {
                    // bp = 8: the 16 rows of the A fragment hold TWO taps (m = g: tap 2j, m = g + 8:
                    // tap 2j + 1), so 15 taps cost 8 MMAs instead of 15 half-empty ones
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int ta = tg * 8 + 2 * j, tb = ta + 1;
                        const bool hb = tb < kTaps;
                        const float a[4] = {ap[0][ta * tapstep], hb ? ap[0][(hb ? tb : ta) * tapstep] : 0.f,
                                            ap[1][ta * tapstep], hb ? ap[1][(hb ? tb : ta) * tapstep] : 0.f};
                        mma_f(acc[j][0], a, b[0][0], b[0][1]);
                    }
                }
            }
        }
    }
    __syncthreads();
    if (NT == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int tap = tg * 8 + j;
            if (tap < kTaps) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int ci = g + 8 * (i >> 1), co = nt * 8 + 2 * t + (i & 1);
                        atomicAdd(&s_dW[(tap * BP + ci) * BP + co], acc[j][nt][i]);
                    }
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int tap = tg * 8 + 2 * j + (i >> 1), co = 2 * t + (i & 1);
                if (tap < kTaps) atomicAdd(&s_dW[(tap * BP + g) * BP + co], acc[j][0][i]);
            }
    }
    __syncthreads();
    if ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
        for (int i = tid; i < kTaps * BP * BP / 4; i += kThr)
            red_add4(p.out + 4 * i, make_float4(s_dW[4 * i], s_dW[4 * i + 1], s_dW[4 * i + 2], s_dW[4 * i + 3]));
    } else {
        for (int i = tid; i < kTaps * BP * BP; i += kThr) atomicAdd(&p.out[i], s_dW[i]);
    }
}

template <typename K>
void set_smem3(K kern, size_t bytes) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int check_small(const char* who, int NM, int T, int V, int bp, int stride) {
    ISTGCN_REQUIRE(NM >= 0 && T >= 1 && V >= 1 && V <= 48, ISTGCN_E_SHAPE, "%s: bad NM/T/V (%d,%d,%d)", who, NM, T, V);
    ISTGCN_REQUIRE(bp == 8 || bp == 16, ISTGCN_E_SHAPE, "%s: padded bottleneck bp=%d must be 8 or 16", who, bp);
    ISTGCN_REQUIRE(stride == 1 || stride == 2, ISTGCN_E_SHAPE, "%s: stride=%d unsupported", who, stride);
    ISTGCN_REQUIRE((long long)NM * T * V * bp < (1ll << 31), ISTGCN_E_SHAPE, "%s: tensor too large", who);
    return 0;
}

template <typename K>
int grid_small(K kern, size_t smem, int NM, int Tf) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThr, smem) != cudaSuccess || occ < 1) occ = 1;
    const long long items = (long long)NM * ((Tf + kTT - 1) / kTT);
    long long n = (long long)num_sms() * occ;
    if (n > items) n = items;
    return (int)(n < 1 ? 1 : n);
}

}  // namespace
}  // namespace istgcn

using namespace istgcn;

#define T2S_DISPATCH_V(LAUNCH)                  \
    if (V == 25) { LAUNCH(25); }                \
    else if (V == 18) { LAUNCH(18); }           \
    else { LAUNCH(0); }

ISTGCN_API int istgcn_tcn2_conv(const float* h1, const float* Weff, const float* beff, float* h2,
                                int NM, int T, int V, int bp, int stride, istgcn_stream_t s) {
    ISTGCN_REQUIRE(h1 && Weff && beff && h2, ISTGCN_E_ARG, "tcn2_conv: null pointer");
    if (int e = check_small("tcn2_conv", NM, T, V, bp, stride)) return e;
    if (NM == 0) return 0;
    const int Tout = (T - 1) / stride + 1, nt = bp / 8;
    const unsigned inv_per = (unsigned)(((1ull << 32) + (unsigned)(V * bp / 4) - 1) / (unsigned)(V * bp / 4));
    SmallP p{h1, nullptr, Weff, beff, h2, nullptr, NM, T, Tout, V, stride, (65536 + V - 1) / V, inv_per};
    const int Q = (kTT - 1) * stride + kTaps;
    const size_t smem = sizeof(float) * ((size_t)kTaps * nt * nt * 64 + 16 + (size_t)Q * V * bp);
#define T2S_FWD(VV)                                                                                  \
    if (nt == 1) {                                                                                    \
        set_smem3(tcn2_small_conv_kernel<1, false, VV>, smem);                                        \
        tcn2_small_conv_kernel<1, false, VV><<<grid_small(tcn2_small_conv_kernel<1, false, VV>, smem, NM, Tout), \
                                               kThr, smem, (cudaStream_t)s>>>(p);                    \
    } else {                                                                                          \
        set_smem3(tcn2_small_conv_kernel<2, false, VV>, smem);                                        \
        tcn2_small_conv_kernel<2, false, VV><<<grid_small(tcn2_small_conv_kernel<2, false, VV>, smem, NM, Tout), \
                                               kThr, smem, (cudaStream_t)s>>>(p);                    \
    }
    T2S_DISPATCH_V(T2S_FWD)
#undef T2S_FWD
    return finish_launch("tcn2_conv");
}

ISTGCN_API int istgcn_tcn2_bwd_conv(const float* dh2, const float* h1, const float* Weff, float* dh1,
                                    float* dWeff, float* dbd, int NM, int T, int V, int bp, int stride,
                                    istgcn_stream_t s) {
    // dh1 == NULL: weight gradient only; dWeff == NULL: input gradient (+ dbd) only -- the two kernels are
    // independent, a caller may put the weight gradient on another stream
    ISTGCN_REQUIRE(dh2 && h1 && Weff && (dh1 || dWeff) && (dh1 == nullptr || dbd), ISTGCN_E_ARG,
                   "tcn2_bwd_conv: null pointer");
    if (int e = check_small("tcn2_bwd_conv", NM, T, V, bp, stride)) return e;
    if (NM == 0) return 0;
    const int Tout = (T - 1) / stride + 1, nt = bp / 8;
    const int inv16 = (65536 + V - 1) / V;
    const unsigned inv_per = (unsigned)(((1ull << 32) + (unsigned)(V * bp / 4) - 1) / (unsigned)(V * bp / 4));
    if (dh1) {   // dh1 (frames of the output = T) from the zero-upsampled dh2, column sums -> dbd
        SmallP p{dh2, nullptr, Weff, nullptr, dh1, dbd, NM, Tout, T, V, stride, inv16, inv_per};
        const int Q = kTT + 2 * kHalf;
        const size_t smem = sizeof(float) * ((size_t)kTaps * nt * nt * 64 + 16 + (size_t)Q * V * bp);
#define T2S_BWD(VV)                                                                                  \
    if (nt == 1) {                                                                                    \
        set_smem3(tcn2_small_conv_kernel<1, true, VV>, smem);                                         \
        tcn2_small_conv_kernel<1, true, VV><<<grid_small(tcn2_small_conv_kernel<1, true, VV>, smem, NM, T), \
                                              kThr, smem, (cudaStream_t)s>>>(p);                     \
    } else {                                                                                          \
        set_smem3(tcn2_small_conv_kernel<2, true, VV>, smem);                                         \
        tcn2_small_conv_kernel<2, true, VV><<<grid_small(tcn2_small_conv_kernel<2, true, VV>, smem, NM, T), \
                                              kThr, smem, (cudaStream_t)s>>>(p);                     \
    }
        T2S_DISPATCH_V(T2S_BWD)
#undef T2S_BWD
        if (int e = finish_launch("tcn2_bwd_conv (data)")) return e;
    }
    if (dWeff) {
        SmallP p{h1, dh2, Weff, nullptr, dWeff, nullptr, NM, T, Tout, V, stride, inv16, inv_per};
        const int Q = (kTT - 1) * stride + kTaps;
        const size_t smem = sizeof(float) * ((size_t)kTaps * bp * bp + (size_t)kTT * V * bp + (size_t)Q * V * bp);
#define T2S_DW(VV)                                                                                   \
    if (nt == 1) {                                                                                    \
        set_smem3(tcn2_small_dw_kernel<1, VV>, smem);                                                 \
        tcn2_small_dw_kernel<1, VV><<<grid_small(tcn2_small_dw_kernel<1, VV>, smem, NM, Tout), kThr, smem, \
                                      (cudaStream_t)s>>>(p);                                         \
    } else {                                                                                          \
        set_smem3(tcn2_small_dw_kernel<2, VV>, smem);                                                 \
        tcn2_small_dw_kernel<2, VV><<<grid_small(tcn2_small_dw_kernel<2, VV>, smem, NM, Tout), kThr, smem, \
                                      (cudaStream_t)s>>>(p);                                         \
    }
        T2S_DISPATCH_V(T2S_DW)
#undef T2S_DW
    }
    return finish_launch("tcn2_bwd_conv (weights)");
}
