// Graph convolution of the network's first block (reference: net/utils/tgcn.py:76-89 with
// in_channels = 3, net/st_gcnold.py:46): with Cin <= 4 input channels the layer is a pure
// streaming problem (12 B in, 256 B out per row), so it runs on CUDA cores in full fp32 instead
// of padding the input to a 32-channel tensor-core slice.
//
//   forward   out[(f,w)][n] = sum_{k,c} X'_k[(f,w)][c] * Wc[k*Cin+c][n] + biasterm[w][n],
//             X'_k[(f,w)][c] = sum_{j in dst(k,w)} vals_j * x[(f,v_j)][c]
//             + BatchNorm sums of out (double)
//   backward  ONE kernel for everything behind dz = BN1-backward(g1, z) (never materialised):
//             G_k[(f,w)][c] = sum_n dz[(f,w)][n] * Wc[k*Cin+c][n]
//             dx[(f,v)][c]  = sum_k sum_{j in t(k,v)} vals_j * G_k[(f,w_j)][c]
//             dvals[j]     += sum_{f,c} x[(f,v_j)][c] * G_{k_j}[(f,w_j)][c]
//             dWc[k*Cin+c][n] += sum_rows X'_k[row][c] * dz[row][n]
//             dbt[w][n]    += sum_f dz[(f,w)][n]
//
// Thread layout (both kernels): thread = (joint w, 4-channel group cg); a warp covers two joints
// = 512 contiguous bytes of an activation row pair, a frame is V*Cout*4 contiguous bytes.
#include "common.cuh"

namespace istgcn {

constexpr int kSmallFT = 8;          // frames per tile
constexpr int kSmallKC = 16;         // K * 4 (input channels padded to 4)

struct SmallLists {
    const float* vals;
    const int *lptr, *lsrc, *lid;    // grouped by (k, destination w): source v
    const int *tptr, *tsrc, *tid;    // grouped by (k, source v): destination w   (backward only)
    int nnz;
};

// x tile -> xs[f][v][4], aggregated xa[f][w][k*4 + c]
__device__ __forceinline__ void small_stage(const float* __restrict__ x, float* xs, float* xa,
                                            const SmallLists& L, long long f0, int nf, int V, int K,
                                            int Cin, int tid, int nthreads) {
    for (int i = tid; i < nf * V * 4; i += nthreads) {
        const int c = i & 3, r = i >> 2;
        xs[i] = c < Cin ? x[(f0 * V + r) * Cin + c] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < nf * V * K; i += nthreads) {
        const int k = i % K, r = i / K;
        const int f = r / V, w = r - f * V;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = L.lptr[k * V + w]; j < L.lptr[k * V + w + 1]; ++j) {
            const float av = L.vals[L.lid[j]];
            const float4 xv = *reinterpret_cast<const float4*>(xs + (f * V + L.lsrc[j]) * 4);
            a.x = fmaf(av, xv.x, a.x); a.y = fmaf(av, xv.y, a.y);
            a.z = fmaf(av, xv.z, a.z); a.w = fmaf(av, xv.w, a.w);
        }
        *reinterpret_cast<float4*>(xa + r * kSmallKC + k * 4) = a;
    }
    __syncthreads();
}

// ----------------------------------------------------------------------------- forward
template <int COUT>
__global__ void __launch_bounds__(512, 1)
gcn_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ Wc,
                     const float* __restrict__ biasterm, SmallLists L, float* __restrict__ out,
                     double* stat_sum, double* stat_sumsq, float* __restrict__ xagg_out,
                     float* __restrict__ zsum_out, long long frames, int V, int K, int Cin) {
    constexpr int CG = COUT / 4;                 // channel groups per row
    constexpr int WPB = 512 / CG;                // joints handled per pass
    __shared__ __align__(16) float xs[kSmallFT * 32 * 4];
    __shared__ __align__(16) float xa[kSmallFT * 32 * kSmallKC];
    __shared__ double s_stat[2 * COUT];
    const int tid = threadIdx.x;
    const int cg = tid % CG, wslot = tid / CG;
    for (int i = tid; i < 2 * COUT; i += 512) s_stat[i] = 0.0;
    // this thread's weight slice: Wc[kc][4cg .. 4cg+3] for the K*Cin real rows
    float4 wr[kSmallKC];
#pragma unroll
    for (int kc = 0; kc < kSmallKC; ++kc) {
        const int k = kc >> 2, c = kc & 3;
        wr[kc] = (k < K && c < Cin) ? ld4(Wc + (size_t)(k * Cin + c) * COUT + 4 * cg)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
    const long long tiles = (frames + kSmallFT - 1) / kSmallFT;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long f0 = tile * kSmallFT;
        const int nf = (int)min((long long)kSmallFT, frames - f0);
        __syncthreads();
        small_stage(x, xs, xa, L, f0, nf, V, K, Cin, tid, 512);
        if (xagg_out)       // the aggregated input X'[(f,w)][k*4 + c], TF32-rounded: the backward's operand
            for (int i = tid; i < nf * V * kSmallKC; i += 512)
                xagg_out[(size_t)f0 * V * kSmallKC + i] =
                    (i & (kSmallKC - 1)) < 4 * K ? __uint_as_float((__float_as_uint(xa[i]) + 0x1000u) & 0xFFFFE000u)
                                                 : 0.f;
        float s4[4] = {0, 0, 0, 0}, q4[4] = {0, 0, 0, 0};
        for (int w = wslot; w < V; w += WPB) {
            const float4 bt = ld4(biasterm + (size_t)w * COUT + 4 * cg);
            for (int f = 0; f < nf; ++f) {
                const float* a = xa + (f * V + w) * kSmallKC;
                float4 o = bt;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 av = *reinterpret_cast<const float4*>(a + 4 * q);
                    const float ar[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 wv = wr[q * 4 + c];
                        o.x = fmaf(ar[c], wv.x, o.x); o.y = fmaf(ar[c], wv.y, o.y);
                        o.z = fmaf(ar[c], wv.z, o.z); o.w = fmaf(ar[c], wv.w, o.w);
                    }
                }
                st4(out + ((f0 + f) * V + w) * COUT + 4 * cg, o);
                s4[0] += o.x; s4[1] += o.y; s4[2] += o.z; s4[3] += o.w;
                q4[0] = fmaf(o.x, o.x, q4[0]); q4[1] = fmaf(o.y, o.y, q4[1]);
                q4[2] = fmaf(o.z, o.z, q4[2]); q4[3] = fmaf(o.w, o.w, q4[3]);
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) { ds[c] += (double)s4[c]; dq[c] += (double)q4[c]; }
    }
    if (zsum_out && wslot < V && WPB >= V) {      // per-joint sums of the output over the frames
#pragma unroll
        for (int c = 0; c < 4; ++c) atomicAdd(&zsum_out[(size_t)wslot * COUT + 4 * cg + c], (float)ds[c]);
    }
    if (stat_sum) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            atomicAdd(&s_stat[4 * cg + c], ds[c]);
            atomicAdd(&s_stat[COUT + 4 * cg + c], dq[c]);
        }
        __syncthreads();
        for (int i = tid; i < COUT; i += 512) {
            atomicAdd(&stat_sum[i], s_stat[i]);
            atomicAdd(&stat_sumsq[i], s_stat[COUT + i]);
        }
    }
}

// ----------------------------------------------------------------------------- backward
// COUT = 64 only: one row = 16 lanes, so the 16 partial dot products of a row are reduced by a
// 15-shuffle butterfly inside a half warp.
template <int CIN_T>      // 3: the (k, c = 3) rows are compile-time zeros (saves 16 accumulators); 4: generic
__global__ void __launch_bounds__(512, 1)
gcn_small_bwd_kernel(const float* __restrict__ g1, const float* __restrict__ z,
                     const float* __restrict__ bn_p, const float* __restrict__ bn_m1,
                     const float* __restrict__ bn_c, const float* __restrict__ bn_mu,
                     const float* __restrict__ x, const float* __restrict__ Wc, SmallLists L,
                     float* __restrict__ dx, float* dvals, float* dWc, float* dbt, long long frames,
                     int V, int K, int Cin) {
    constexpr int COUT = 64, CG = 16;
    __shared__ __align__(16) float xs[kSmallFT * 32 * 4];
    __shared__ __align__(16) float xa[kSmallFT * 32 * kSmallKC];
    __shared__ __align__(16) float G[kSmallFT * 32 * kSmallKC];
    __shared__ __align__(16) float s_W[kSmallKC * COUT];
    __shared__ float s_dW[kSmallKC * COUT];
    const int tid = threadIdx.x, lane = tid & 31;
    const int cg = tid % CG, w = tid / CG;            // one joint per thread slot (V <= 32)
    const bool active = w < V;
    for (int i = tid; i < kSmallKC * COUT; i += 512) {
        const int kc = i / COUT, n = i - kc * COUT;
        const int k = kc >> 2, c = kc & 3;
        s_W[i] = (k < K && c < Cin) ? Wc[(size_t)(k * Cin + c) * COUT + n] : 0.f;
        s_dW[i] = 0.f;
    }
    __shared__ __align__(16) float s_bn[4 * COUT];   // p, m1, c, mu
    for (int i = tid; i < COUT; i += 512) {
        s_bn[i] = bn_p[i]; s_bn[COUT + i] = bn_m1[i]; s_bn[2 * COUT + i] = bn_c[i]; s_bn[3 * COUT + i] = bn_mu[i];
    }
    float4 dw[kSmallKC];
#pragma unroll
    for (int i = 0; i < kSmallKC; ++i) dw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 db = make_float4(0.f, 0.f, 0.f, 0.f);
    float dv0 = 0.f, dv1 = 0.f;                       // entries tid and tid + 512 of the (k, w) lists
    int ent_kw[2] = {-1, -1}, ent_v[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int e = tid + 512 * h;
        if (e < L.nnz) {                              // largest kw with lptr[kw] <= e
            int lo = 0, hi = K * V;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (L.lptr[mid] <= e) lo = mid; else hi = mid;
            }
            ent_kw[h] = lo;
            ent_v[h] = L.lsrc[e];
        }
    }
    const long long tiles = (frames + kSmallFT - 1) / kSmallFT;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long f0 = tile * kSmallFT;
        const int nf = (int)min((long long)kSmallFT, frames - f0);
        __syncthreads();
        small_stage(x, xs, xa, L, f0, nf, V, K, Cin, tid, 512);
        // ---- phase A: dz on the fly, dbt / dWc accumulators, G by half-warp butterfly
        for (int fb = 0; fb < nf; fb += 2) {           // 2 frames of loads in flight per thread
            float4 gv[2], zv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                gv[u] = zv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (active && fb + u < nf) {
                    const size_t off = ((size_t)(f0 + fb + u) * V + w) * COUT + 4 * cg;
                    gv[u] = ld4(g1 + off);
                    zv[u] = ld4(z + off);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int f = fb + u;
                if (f >= nf) break;                    // block-uniform
                float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                if (active) {
                    const float4 bp = ld4(s_bn + 4 * cg), bm = ld4(s_bn + COUT + 4 * cg),
                                 bc = ld4(s_bn + 2 * COUT + 4 * cg), bu = ld4(s_bn + 3 * COUT + 4 * cg);
                    d.x = bn_back(gv[u].x, zv[u].x, bp.x, bm.x, bc.x, bu.x);
                    d.y = bn_back(gv[u].y, zv[u].y, bp.y, bm.y, bc.y, bu.y);
                    d.z = bn_back(gv[u].z, zv[u].z, bp.z, bm.z, bc.z, bu.z);
                    d.w = bn_back(gv[u].w, zv[u].w, bp.w, bm.w, bc.w, bu.w);
                }
                db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
                const float* a = xa + (f * V + (active ? w : 0)) * kSmallKC;
                float part[kSmallKC];
#pragma unroll
                for (int kc = 0; kc < kSmallKC; ++kc) {
                    if ((kc & 3) >= CIN_T) { part[kc] = 0.f; continue; }
                    const float av = a[kc];
                    dw[kc].x = fmaf(av, d.x, dw[kc].x); dw[kc].y = fmaf(av, d.y, dw[kc].y);
                    dw[kc].z = fmaf(av, d.z, dw[kc].z); dw[kc].w = fmaf(av, d.w, dw[kc].w);
                    const float4 wv = *reinterpret_cast<const float4*>(s_W + kc * COUT + 4 * cg);
                    part[kc] = fmaf(d.x, wv.x, fmaf(d.y, wv.y, fmaf(d.z, wv.z, d.w * wv.w)));
                }
                // butterfly over the 16 lanes of the row: lane l ends with the full sum of part[l]
#pragma unroll
                for (int step = 8; step >= 1; step >>= 1) {
                    const bool upper = (lane & step) != 0;
#pragma unroll
                    for (int i = 0; i < step; ++i) {
                        const float send = upper ? part[i] : part[i + step];
                        const float keep = upper ? part[i + step] : part[i];
                        part[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
                    }
                }
                if (active) G[(f * V + w) * kSmallKC + cg] = part[0];
            }
        }
        __syncthreads();
        // ---- phase B: dx by the transposed lists, dvals per entry
        for (int i = tid; i < nf * V * Cin; i += 512) {
            const int c = i % Cin, r = i / Cin;
            const int f = r / V, v = r - f * V;
            float acc = 0.f;
            for (int k = 0; k < K; ++k)
                for (int j = L.tptr[k * V + v]; j < L.tptr[k * V + v + 1]; ++j)
                    acc = fmaf(L.vals[L.tid[j]], G[(f * V + L.tsrc[j]) * kSmallKC + k * 4 + c], acc);
            dx[((f0 + f) * V + v) * Cin + c] = acc;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (ent_kw[h] < 0) continue;
            const int k = ent_kw[h] / V, wd = ent_kw[h] - k * V, v = ent_v[h];
            float acc = 0.f;
            for (int f = 0; f < nf; ++f) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + (f * V + v) * 4);
                const float4 gq = *reinterpret_cast<const float4*>(G + (f * V + wd) * kSmallKC + k * 4);
                acc += xv.x * gq.x + xv.y * gq.y + xv.z * gq.z + xv.w * gq.w;
            }
            if (h == 0) dv0 += acc; else dv1 += acc;
        }
    }
    // ---- flush
    __syncthreads();
    if (active) {
#pragma unroll
        for (int kc = 0; kc < kSmallKC; ++kc) {
            if ((kc & 3) >= CIN_T) continue;
            atomicAdd(&s_dW[kc * COUT + 4 * cg + 0], dw[kc].x);
            atomicAdd(&s_dW[kc * COUT + 4 * cg + 1], dw[kc].y);
            atomicAdd(&s_dW[kc * COUT + 4 * cg + 2], dw[kc].z);
            atomicAdd(&s_dW[kc * COUT + 4 * cg + 3], dw[kc].w);
        }
        if (dbt) {
            float* o = dbt + (size_t)w * COUT + 4 * cg;
            atomicAdd(o + 0, db.x); atomicAdd(o + 1, db.y); atomicAdd(o + 2, db.z); atomicAdd(o + 3, db.w);
        }
    }
    __syncthreads();
    for (int i = tid; i < kSmallKC * COUT; i += 512) {
        const int kc = i / COUT, n = i - kc * COUT;
        const int k = kc >> 2, c = kc & 3;
        if (k < K && c < Cin) atomicAdd(&dWc[(size_t)(k * Cin + c) * COUT + n], s_dW[i]);
    }
    if (dvals) {
        if (tid < L.nnz) atomicAdd(&dvals[L.lid[tid]], dv0);
        if (tid + 512 < L.nnz) atomicAdd(&dvals[L.lid[tid + 512]], dv1);
    }
}

// ----------------------------------------------------------------------------- backward, second form
// The heavy part of the first block's backward has the shape of the temporal chain's up-projection
// backward: with X' (aggregated input, 16 columns = K x 4) in the role of h2 and Wc (padded to 16 rows)
// in the role of Wu, istgcn_tcn2_bwd_up computes dz = BN1-backward(g1, z) on the fly and from it
//     G[(f,w)][kc] = sum_n dz[(f,w)][n] Wc16[kc][n]         (its dh2)
//     dWc16[kc][n] += sum_rows X'[row][kc] dz[row][n]       (its dWu)
// on the tensor core in ONE pass over (g1, z) at 0.6 of the HBM rate (the CUDA-core kernel above: 0.15).
// What is left is cheap: the per-joint sums of g1 (below; together with the per-joint sums of z from the
// forward they give dbt, because dz is affine in (g1, z)), and dx / dvals from the 16-wide G.

// sums[w][c] += sum_f a[(f,w)][c]: thread = one float4 column of the [V*C] frame vector
__global__ void joint_colsum_kernel(const float* __restrict__ a, float* __restrict__ sums, int frames, int n,
                                    int frames_per_cta) {
    const int n4 = n / 4;
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int slab = (int)(gid / n4);
    const int j4 = (int)(gid - (long long)slab * n4);
    const int f0 = slab * frames_per_cta, f1 = min(frames, f0 + frames_per_cta);
    if (f0 >= f1) return;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int f = f0;
    for (; f + 8 <= f1; f += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ld4(a + (size_t)(f + u) * n + j4 * 4);
#pragma unroll
        for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    for (; f < f1; ++f) {
        const float4 v = ld4(a + (size_t)f * n + j4 * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    red_add4(sums + j4 * 4, s);
}

// dx[(f,v)][c] = sum_k sum_{j in t(k,v)} vals_j G[(f,w_j)][k*4+c]  (written),
// dvals[j] += sum_{f,c} x[(f,v_j)][c] G[(f,w_j)][k_j*4+c], and the bias-term gradient
// dbt[w][n] += p[n] * ((sg1[w][n] - F m1[n]) - c[n] * (sz[w][n] - F mu[n]))   (block 0 only)
__global__ void __launch_bounds__(256)
gcn_small_post_kernel(const float* __restrict__ Gg, const float* __restrict__ x, SmallLists L,
                      float* __restrict__ dx, float* dvals, const float* __restrict__ sg1,
                      const float* __restrict__ sz, const float* __restrict__ bn_p,
                      const float* __restrict__ bn_m1, const float* __restrict__ bn_c,
                      const float* __restrict__ bn_mu, float* dbt, long long frames, int V, int K, int Cin,
                      int Cout) {
    constexpr int FT = 16;                                   // frames per tile
    __shared__ __align__(16) float xs[FT * 32 * 4];
    __shared__ __align__(16) float G[FT * 32 * kSmallKC];
    __shared__ int s_tptr[4 * 32 + 1];                       // the transposed lists, once per CTA
    __shared__ unsigned char s_tsrc[1024];                   // joints < 32; 48 KB of static shared memory in all
    __shared__ float s_tval[1024];
    const int tid = threadIdx.x;
    __shared__ int s_lptr[4 * 32 + 1];
    for (int i = tid; i <= K * V; i += 256) { s_tptr[i] = L.tptr[i]; s_lptr[i] = L.lptr[i]; }
    for (int i = tid; i < L.nnz; i += 256) { s_tsrc[i] = (unsigned char)L.tsrc[i]; s_tval[i] = L.vals[L.tid[i]]; }
    __syncthreads();
    if (blockIdx.x == 0 && dbt)
        for (int i = tid; i < V * Cout; i += 256) {
            const int n = i % Cout;
            const float F = (float)frames;
            dbt[i] += bn_p[n] * ((sg1[i] - F * bn_m1[n]) - bn_c[n] * (sz[i] - F * bn_mu[n]));
        }
    float dv[4] = {0.f, 0.f, 0.f, 0.f};                      // entries tid, tid + 256, ... of the (k, w) lists
    int ent_kw[4] = {-1, -1, -1, -1}, ent_v[4] = {0, 0, 0, 0};
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const int e = tid + 256 * h;
        if (e < L.nnz) {                                     // largest kw with lptr[kw] <= e
            int lo = 0, hi = K * V;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_lptr[mid] <= e) lo = mid; else hi = mid;
            }
            ent_kw[h] = lo;
            ent_v[h] = L.lsrc[e];
        }
    }
    const long long tiles = (frames + FT - 1) / FT;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long f0 = tile * FT;
        const int nf = (int)min((long long)FT, frames - f0);
        __syncthreads();
        for (int i = tid; i < nf * V * 4; i += 256) {
            const int c = i & 3, r = i >> 2;
            xs[i] = c < Cin ? x[(f0 * V + r) * Cin + c] : 0.f;
        }
        for (int i = tid; i < nf * V * (kSmallKC / 4); i += 256)
            reinterpret_cast<float4*>(G)[i] = ld4(Gg + (size_t)f0 * V * kSmallKC + 4 * (size_t)i);
        __syncthreads();
        for (int i = tid; i < nf * V * Cin; i += 256) {
            const int c = i % Cin, r = i / Cin;
            const int f = r / V, v = r - f * V;
            float acc = 0.f;
            for (int k = 0; k < K; ++k)
                for (int j = s_tptr[k * V + v]; j < s_tptr[k * V + v + 1]; ++j)
                    acc = fmaf(s_tval[j], G[(f * V + s_tsrc[j]) * kSmallKC + k * 4 + c], acc);
            dx[((f0 + f) * V + v) * Cin + c] = acc;
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            if (ent_kw[h] < 0) continue;
            const int k = ent_kw[h] / V, wd = ent_kw[h] - k * V, v = ent_v[h];
            float acc = 0.f;
            for (int f = 0; f < nf; ++f) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + (f * V + v) * 4);
                const float4 gq = *reinterpret_cast<const float4*>(G + (f * V + wd) * kSmallKC + k * 4);
                acc += xv.x * gq.x + xv.y * gq.y + xv.z * gq.z + xv.w * gq.w;
            }
            dv[h] += acc;
        }
    }
    if (dvals) {
#pragma unroll
        for (int h = 0; h < 4; ++h)
            if (tid + 256 * h < L.nnz) atomicAdd(&dvals[L.lid[tid + 256 * h]], dv[h]);
    }
}

}  // namespace istgcn

using namespace istgcn;

// Forward graph convolution for Cin <= 4 (see the file header).  Wc [K*Cin][Cout], biasterm
// [V][Cout] = sum_k colsum(A_eff[k])[w] * bias_k[n]; lists grouped by (k, destination w).
ISTGCN_API int istgcn_gcn_small_fwd(const float* x, const float* Wc, const float* biasterm,
                                    const float* vals, const int* lptr, const int* lsrc, const int* lid,
                                    int nnz, float* out, double* stat_sum, double* stat_sumsq,
                                    float* xagg_out, float* zsum_out, int frames, int V, int K, int Cin,
                                    int Cout, istgcn_stream_t s) {
    ISTGCN_REQUIRE(x && Wc && biasterm && vals && lptr && lsrc && lid && out, ISTGCN_E_ARG,
                   "gcn_small_fwd: null pointer");
    ISTGCN_REQUIRE((stat_sum == nullptr) == (stat_sumsq == nullptr), ISTGCN_E_ARG,
                   "gcn_small_fwd: pass both statistics buffers or neither");
    ISTGCN_REQUIRE(V >= 1 && V <= 32 && K >= 1 && K <= 4 && Cin >= 1 && Cin <= 4, ISTGCN_E_SHAPE,
                   "gcn_small_fwd: V=%d K=%d Cin=%d unsupported", V, K, Cin);
    ISTGCN_REQUIRE(Cout == 64 || Cout == 128, ISTGCN_E_SHAPE, "gcn_small_fwd: Cout=%d unsupported", Cout);
    ISTGCN_REQUIRE(nnz >= 0 && nnz <= kMaxNnz, ISTGCN_E_SHAPE, "gcn_small_fwd: nnz=%d", nnz);
    ISTGCN_REQUIRE(((reinterpret_cast<uintptr_t>(Wc) | reinterpret_cast<uintptr_t>(biasterm) |
                     reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                   ISTGCN_E_ARG, "gcn_small_fwd: pointers must be 16-byte aligned");
    if (frames == 0) return 0;
    SmallLists L{vals, lptr, lsrc, lid, nullptr, nullptr, nullptr, nnz};
    const long long tiles = ((long long)frames + kSmallFT - 1) / kSmallFT;
    int grid = num_sms() * 2;
    if (grid > tiles) grid = (int)tiles;
    cudaStream_t st = (cudaStream_t)s;
    if (Cout == 64)
        gcn_small_fwd_kernel<64><<<grid, 512, 0, st>>>(x, Wc, biasterm, L, out, stat_sum, stat_sumsq,
                                                       xagg_out, zsum_out, frames, V, K, Cin);
    else
        gcn_small_fwd_kernel<128><<<grid, 512, 0, st>>>(x, Wc, biasterm, L, out, stat_sum, stat_sumsq,
                                                        xagg_out, nullptr, frames, V, K, Cin);
    return finish_launch("gcn_small_fwd");
}

// Whole backward of that layer behind dz = p*((g1 - m1) - c*(z - mu)): dx [frames*V][Cin]
// (written), dvals / dWc [K*Cin][64] / dbt [V][64] (accumulated, caller-zeroed; dvals and dbt may
// be NULL).  Lists: (lptr, lsrc, lid) grouped by (k, destination w); (tptr, tsrc, tid) grouped by
// (k, source v).  Cout == 64.
ISTGCN_API int istgcn_gcn_small_bwd(const float* g1, const float* z, const float* bn_p,
                                    const float* bn_m1, const float* bn_c, const float* bn_mu,
                                    const float* x, const float* Wc, const float* vals, const int* lptr,
                                    const int* lsrc, const int* lid, const int* tptr, const int* tsrc,
                                    const int* tid, int nnz, float* dx, float* dvals, float* dWc,
                                    float* dbt, int frames, int V, int K, int Cin, int Cout,
                                    istgcn_stream_t s) {
    ISTGCN_REQUIRE(g1 && z && bn_p && bn_m1 && bn_c && bn_mu && x && Wc && vals && lptr && lsrc && lid &&
                       tptr && tsrc && tid && dx && dWc,
                   ISTGCN_E_ARG, "gcn_small_bwd: null pointer");
    ISTGCN_REQUIRE(V >= 1 && V <= 32 && K >= 1 && K <= 4 && Cin >= 1 && Cin <= 4, ISTGCN_E_SHAPE,
                   "gcn_small_bwd: V=%d K=%d Cin=%d unsupported", V, K, Cin);
    ISTGCN_REQUIRE(Cout == 64, ISTGCN_E_SHAPE, "gcn_small_bwd: Cout=%d unsupported (64 only)", Cout);
    ISTGCN_REQUIRE(nnz >= 0 && nnz <= kMaxNnz, ISTGCN_E_SHAPE, "gcn_small_bwd: nnz=%d", nnz);
    ISTGCN_REQUIRE(((reinterpret_cast<uintptr_t>(g1) | reinterpret_cast<uintptr_t>(z)) & 15) == 0,
                   ISTGCN_E_ARG, "gcn_small_bwd: pointers must be 16-byte aligned");
    if (frames == 0) return 0;
    SmallLists L{vals, lptr, lsrc, lid, tptr, tsrc, tid, nnz};
    const long long tiles = ((long long)frames + kSmallFT - 1) / kSmallFT;
    int grid = num_sms();
    if (grid > tiles) grid = (int)tiles;
    if (Cin <= 3)
        gcn_small_bwd_kernel<3><<<grid, 512, 0, (cudaStream_t)s>>>(g1, z, bn_p, bn_m1, bn_c, bn_mu, x, Wc, L,
                                                                  dx, dvals, dWc, dbt, frames, V, K, Cin);
    else
        gcn_small_bwd_kernel<4><<<grid, 512, 0, (cudaStream_t)s>>>(g1, z, bn_p, bn_m1, bn_c, bn_mu, x, Wc, L,
                                                                  dx, dvals, dWc, dbt, frames, V, K, Cin);
    return finish_launch("gcn_small_bwd");
}

// sums[V][C] += sum over frames of a[(f,v)][c] (caller-zeroed; C % 4 == 0)
ISTGCN_API int istgcn_joint_colsum(const float* a, float* sums, int frames, int V, int C, istgcn_stream_t s) {
    ISTGCN_REQUIRE(a && sums, ISTGCN_E_ARG, "joint_colsum: null pointer");
    ISTGCN_REQUIRE(C % 4 == 0 && V >= 1, ISTGCN_E_SHAPE, "joint_colsum: C=%d V=%d", C, V);
    ISTGCN_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(sums)) & 15) == 0, ISTGCN_E_ARG,
                   "joint_colsum: pointers must be 16-byte aligned");
    if (frames == 0) return 0;
    const int n = V * C;
    int slabs = (num_sms() * 4 * 256) / (n / 4);      // long slabs: one atomic per thread and column at the end
    if (slabs < 1) slabs = 1;
    if (slabs > frames) slabs = frames;
    const int fpc = (frames + slabs - 1) / slabs;
    const long long items = (long long)(n / 4) * ((frames + fpc - 1) / fpc);
    joint_colsum_kernel<<<(int)((items + 255) / 256), 256, 0, (cudaStream_t)s>>>(a, sums, frames, n, fpc);
    return finish_launch("joint_colsum");
}

// Tail of the first block's backward behind the tensor-core pass (see above): G [frames*V][16] is the
// dh2 of istgcn_tcn2_bwd_up run on (g1, z, X', Wc16).  dx [frames*V][Cin] written; dvals, dbt [V][Cout]
// accumulated (may be NULL); sg1 / sz [V][Cout] = per-joint sums over the frames of g1 / z.
ISTGCN_API int istgcn_gcn_small_bwd_post(const float* G, const float* x, const float* vals, const int* lptr,
                                         const int* lsrc, const int* lid, const int* tptr, const int* tsrc,
                                         const int* tid, int nnz, float* dx, float* dvals, const float* sg1,
                                         const float* sz, const float* bn_p, const float* bn_m1,
                                         const float* bn_c, const float* bn_mu, float* dbt, int frames,
                                         int V, int K, int Cin, int Cout, istgcn_stream_t s) {
    ISTGCN_REQUIRE(G && x && vals && lptr && lsrc && lid && tptr && tsrc && tid && dx, ISTGCN_E_ARG,
                   "gcn_small_bwd_post: null pointer");
    ISTGCN_REQUIRE(dbt == nullptr || (sg1 && sz && bn_p && bn_m1 && bn_c && bn_mu), ISTGCN_E_ARG,
                   "gcn_small_bwd_post: dbt needs the per-joint sums and the BatchNorm coefficients");
    ISTGCN_REQUIRE(V >= 1 && V <= 32 && K >= 1 && K <= 4 && Cin >= 1 && Cin <= 4, ISTGCN_E_SHAPE,
                   "gcn_small_bwd_post: V=%d K=%d Cin=%d unsupported", V, K, Cin);
    ISTGCN_REQUIRE(nnz >= 0 && nnz <= 1024, ISTGCN_E_SHAPE, "gcn_small_bwd_post: nnz=%d", nnz);
    ISTGCN_REQUIRE((reinterpret_cast<uintptr_t>(G) & 15) == 0, ISTGCN_E_ARG, "gcn_small_bwd_post: G alignment");
    if (frames == 0) return 0;
    SmallLists L{vals, lptr, lsrc, lid, tptr, tsrc, tid, nnz};
    const long long tiles = ((long long)frames + 15) / 16;
    int grid = num_sms() * 4;
    if (grid > tiles) grid = (int)tiles;
    gcn_small_post_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(G, x, L, dx, dvals, sg1, sz, bn_p, bn_m1, bn_c, bn_mu,
                                                             dbt, frames, V, K, Cin, Cout);
    return finish_launch("gcn_small_bwd_post");
}
