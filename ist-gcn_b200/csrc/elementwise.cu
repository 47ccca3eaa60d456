// Bandwidth-bound kernels of the IST-GCN block: data_bn (+layout change), BatchNorm
// bookkeeping, the block tail (BN2 -> dropout -> +residual -> ReLU) and the pooling head.
// All HBM-bound: coalesced float4 traffic, per-channel reductions kept in registers and
// flushed once per CTA with double-precision atomics.
#include <stdarg.h>

#include "common.cuh"

namespace istgcn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int finish_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// ------------------------------------------------------------------------------ data_bn
// x: (N, C, T, V, M) contiguous.  One CTA per (n, c) slab and t-slice; thread = (t-group, v*M+m).
__global__ void data_bn_stats_kernel(const float* __restrict__ x, double* __restrict__ sum,
                                     double* __restrict__ sumsq, int C, int T, int V, int M,
                                     int t_per_cta) {
    extern __shared__ double shd[];           // [2][V]
    const int VM = V * M;
    const int groups = blockDim.x / VM;
    const int nc = blockIdx.x;               // n*C + c
    const int c = nc % C;
    const int t0 = blockIdx.y * t_per_cta;
    const int t1 = min(T, t0 + t_per_cta);
    for (int i = threadIdx.x; i < 2 * V; i += blockDim.x) shd[i] = 0.0;
    __syncthreads();
    const int grp = threadIdx.x / VM, vm = threadIdx.x % VM;
    if (grp < groups) {
        const float* slab = x + (size_t)nc * T * VM;
        double s = 0.0, q = 0.0;
        for (int t = t0 + grp; t < t1; t += groups) {
            const float v = slab[(size_t)t * VM + vm];
            s += v;
            q += (double)v * v;
        }
        atomicAdd(&shd[vm / M], s);
        atomicAdd(&shd[V + vm / M], q);
    }
    __syncthreads();
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        atomicAdd(&sum[v * C + c], shd[v]);
        atomicAdd(&sumsq[v * C + c], shd[V + v]);
    }
}

// thread per output row (n, m, t, v): gathers C strided inputs, writes C contiguous outputs.
__global__ void data_bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                     const float* __restrict__ scale,
                                     const float* __restrict__ beta, float* __restrict__ y,
                                     int N, int C, int T, int V, int M) {
    const long long rows = (long long)N * M * T * V;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows;
         r += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(r % V);
        long long q = r / V;
        const int t = (int)(q % T);
        q /= T;
        const int m = (int)(q % M);
        const int n = (int)(q / M);
        for (int c = 0; c < C; ++c) {
            const float val = x[((((size_t)n * C + c) * T + t) * V + v) * M + m];
            y[r * C + c] = bn_apply(val, mean[v * C + c], scale[v * C + c], beta[v * C + c]);
        }
    }
}

__global__ void data_bn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                   double* __restrict__ dgamma, double* __restrict__ dbeta, int C,
                                   int T, int V, int M, int t_per_cta) {
    extern __shared__ float sh[];            // [2][V]
    const int VM = V * M;
    const int groups = blockDim.x / VM;
    const int nc = blockIdx.x;
    const int n = nc / C, c = nc % C;
    const int t0 = blockIdx.y * t_per_cta;
    const int t1 = min(T, t0 + t_per_cta);
    for (int i = threadIdx.x; i < 2 * V; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    const int grp = threadIdx.x / VM, vm = threadIdx.x % VM;
    if (grp < groups) {
        const int v = vm / M, m = vm % M;
        const float mu = mean[v * C + c], rs = rstd[v * C + c];
        const float* slab = x + (size_t)nc * T * VM;
        float sb = 0.f, sg = 0.f;
        for (int t = t0 + grp; t < t1; t += groups) {
            const float xv = slab[(size_t)t * VM + vm];
            const float gv = g[((((size_t)n * M + m) * T + t) * V + v) * C + c];
            sb += gv;
            sg += gv * (xv - mu) * rs;
        }
        atomicAdd(&sh[v], sg);
        atomicAdd(&sh[V + v], sb);
    }
    __syncthreads();
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        atomicAdd(&dgamma[v * C + c], (double)sh[v]);
        atomicAdd(&dbeta[v * C + c], (double)sh[V + v]);
    }
}

// ------------------------------------------------------------------------------ BN coefficients
__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq,
                                   double count, const float* __restrict__ gamma,
                                   float* running_mean, float* running_var, float momentum,
                                   float eps, float* scale, float* mean, float* rstd, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mu = sum[c] / count;
    double var = sumsq[c] / count - mu * mu;
    if (var < 0) var = 0;
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    scale[c] = gamma[c] * rs;
    if (mean) mean[c] = (float)mu;
    if (rstd) rstd[c] = rs;
    if (running_mean) {
        const double unbiased = count > 1 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

__global__ void bn_eval_coeffs_kernel(const float* gamma, const float* rv, float eps, float* scale,
                                      int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    scale[c] = gamma[c] / sqrtf(rv[c] + eps);
}

// y = gamma*xhat + beta, xhat = (x-mean)*rstd:  dx = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat))
//    = p*((g - m1) - c*(x - mean)),  p = gamma*rstd, m1 = mean(g), c = rstd*mean(g*xhat)
__global__ void bn_bwd_coeffs_kernel(const double* sg, const double* sgx, double count,
                                     const float* gamma, const float* rstd, float* p, float* m1,
                                     float* cc, float* dgamma, float* dbeta, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    p[c] = gamma[c] * rstd[c];
    m1[c] = (float)(sg[c] / count);
    cc[c] = (float)((double)rstd[c] * (sgx[c] / count));
    if (dgamma) dgamma[c] = (float)sgx[c];
    if (dbeta) dbeta[c] = (float)sg[c];
}

// ------------------------------------------------------------------------------ block tail
// mode: 0 no residual, 1 identity (res = block input), 2 conv+BN (res*scale_r + shift_r)
__global__ void block_tail_fwd_kernel(const float* __restrict__ u, const float* __restrict__ mean2,
                                      const float* __restrict__ scale2,
                                      const float* __restrict__ beta2, const float* __restrict__ res,
                                      const float* __restrict__ mean_r,
                                      const float* __restrict__ scale_r,
                                      const float* __restrict__ beta_r, float* __restrict__ out,
                                      long long n4, int C, int mode, float drop_p, float keep_scale,
                                      uint64_t seed0, const unsigned long long* step, BnFold fold2,
                                      BnFold foldr) {
    const uint64_t seed = effective_seed(seed0, step);
    const int c4 = C >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    // The grid stride is a multiple of C/4 (the launcher sees to it), so a thread stays on ONE group
    // of four channels: the BatchNorm coefficients live in registers and no 64-bit modulo sits in the
    // loop (the per-iteration `i % c4` + six coefficient loads made this kernel issue-bound at 0.78 of
    // the HBM peak).
    const int c = (int)(i % c4) * 4;
    const bool wr = i < c4;              // the first C/4 threads of the grid store folded coefficients
    float4 mu, sc;
    if (fold2.sum) {
        bn_fold(fold2, c, wr, mu.x, sc.x); bn_fold(fold2, c + 1, wr, mu.y, sc.y);
        bn_fold(fold2, c + 2, wr, mu.z, sc.z); bn_fold(fold2, c + 3, wr, mu.w, sc.w);
    } else {
        mu = ld4(mean2 + c); sc = ld4(scale2 + c);
    }
    const float4 be = ld4(beta2 + c);
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f), a = m, b = m;
    if (mode == 2) {
        if (foldr.sum) {
            bn_fold(foldr, c, wr, m.x, a.x); bn_fold(foldr, c + 1, wr, m.y, a.y);
            bn_fold(foldr, c + 2, wr, m.z, a.z); bn_fold(foldr, c + 3, wr, m.w, a.w);
        } else {
            m = ld4(mean_r + c); a = ld4(scale_r + c);
        }
        b = ld4(beta_r + c);
    }
    for (; i < n4; i += stride) {
        const float4 uu = ld4(u + i * 4);
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mode != 0) r = ld4(res + i * 4);
        float y[4] = {bn_apply(uu.x, mu.x, sc.x, be.x), bn_apply(uu.y, mu.y, sc.y, be.y),
                      bn_apply(uu.z, mu.z, sc.z, be.z), bn_apply(uu.w, mu.w, sc.w, be.w)};
        if (drop_p > 0.f) {
            bool keep[4];
            dropout_keep4(seed, (uint64_t)i, drop_p, keep);
#pragma unroll
            for (int j = 0; j < 4; ++j) y[j] = keep[j] ? y[j] * keep_scale : 0.f;
        }
        if (mode == 1) {
            y[0] += r.x; y[1] += r.y; y[2] += r.z; y[3] += r.w;
        } else if (mode == 2) {
            y[0] += bn_apply(r.x, m.x, a.x, b.x); y[1] += bn_apply(r.y, m.y, a.y, b.y);
            y[2] += bn_apply(r.z, m.z, a.z, b.z); y[3] += bn_apply(r.w, m.w, a.w, b.w);
        }
        st4(out + i * 4, make_float4(fmaxf(y[0], 0.f), fmaxf(y[1], 0.f), fmaxf(y[2], 0.f),
                                     fmaxf(y[3], 0.f)));
    }
}

// Thread = one float4 column group, loops over rows; per-channel sums stay in registers.
// blockDim = 256 -> rows_per_iter = 256 / (C/4).
template <bool HAS_R>
__global__ void block_tail_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ out,
                                      const float* __restrict__ u, const float* __restrict__ mean2,
                                      const float* __restrict__ rstd2, const float* __restrict__ rres,
                                      const float* __restrict__ mean_r,
                                      const float* __restrict__ rstd_r, float* __restrict__ go,
                                      double* __restrict__ sg2, double* __restrict__ sg2x,
                                      double* __restrict__ sgr, double* __restrict__ sgrx,
                                      long long rows, int C, float drop_p, float keep_scale,
                                      uint64_t seed0, const unsigned long long* step) {
    const uint64_t seed = effective_seed(seed0, step);
    __shared__ float red[4][256 * 4 / 4 * 4];   // [quantity][thread*4 + j] -> 4 KB each
    const int c4 = C >> 2;
    const int rows_per_iter = blockDim.x / c4;
    const int col = threadIdx.x % c4, rsub = threadIdx.x / c4;
    const int c = col * 4;
    float a_g[4] = {0, 0, 0, 0}, a_gx[4] = {0, 0, 0, 0}, a_r[4] = {0, 0, 0, 0},
          a_rx[4] = {0, 0, 0, 0};
    float mu[4], rs[4], mur[4], rsr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        mu[j] = mean2[c + j];
        rs[j] = rstd2[c + j];
        mur[j] = HAS_R ? mean_r[c + j] : 0.f;
        rsr[j] = HAS_R ? rstd_r[c + j] : 0.f;
    }
    if (rsub < rows_per_iter) {
        for (long long r = (long long)blockIdx.x * rows_per_iter + rsub; r < rows;
             r += (long long)gridDim.x * rows_per_iter) {
            const long long off = r * C + c;
            const float4 gv = ld4(gout + off), ov = ld4(out + off), uv = ld4(u + off);
            float g4[4] = {gv.x, gv.y, gv.z, gv.w};
            const float o4[4] = {ov.x, ov.y, ov.z, ov.w};
            const float u4[4] = {uv.x, uv.y, uv.z, uv.w};
            float r4[4] = {0, 0, 0, 0};
            if (HAS_R) {
                const float4 rv = ld4(rres + off);
                r4[0] = rv.x; r4[1] = rv.y; r4[2] = rv.z; r4[3] = rv.w;
            }
            bool keep[4] = {true, true, true, true};
            if (drop_p > 0.f) dropout_keep4(seed, (uint64_t)(off >> 2), drop_p, keep);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float gj = o4[j] > 0.f ? g4[j] : 0.f;
                g4[j] = gj;
                float gy = gj;
                if (drop_p > 0.f) gy = keep[j] ? gj * keep_scale : 0.f;
                a_g[j] += gy;
                a_gx[j] += gy * (u4[j] - mu[j]) * rs[j];
                if (HAS_R) {
                    a_r[j] += gj;
                    a_rx[j] += gj * (r4[j] - mur[j]) * rsr[j];
                }
            }
            st4(go + off, make_float4(g4[0], g4[1], g4[2], g4[3]));
        }
    }
    // reduce over the row sub-groups of this CTA, then one double atomic per channel
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        red[0][threadIdx.x * 4 + j] = a_g[j];
        red[1][threadIdx.x * 4 + j] = a_gx[j];
        red[2][threadIdx.x * 4 + j] = a_r[j];
        red[3][threadIdx.x * 4 + j] = a_rx[j];
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
        const int cc = ch >> 2, j = ch & 3;
        float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int k = 0; k < rows_per_iter; ++k) {
            const int th = k * c4 + cc;
            s0 += red[0][th * 4 + j];
            s1 += red[1][th * 4 + j];
            s2 += red[2][th * 4 + j];
            s3 += red[3][th * 4 + j];
        }
        atomicAdd(&sg2[ch], (double)s0);
        atomicAdd(&sg2x[ch], (double)s1);
        if (HAS_R) {
            atomicAdd(&sgr[ch], (double)s2);
            atomicAdd(&sgrx[ch], (double)s3);
        }
    }
}

__global__ void dropout_mask_kernel(unsigned char* mask, long long n, float p, uint64_t seed0,
                                    const unsigned long long* step) {
    const uint64_t seed = effective_seed(seed0, step);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        mask[i] = dropout_keep(seed, (uint64_t)i, p) ? 1 : 0;
}

// ------------------------------------------------------------------------------ pooling head
// x rows [(n*M+m)*TV + tv][C]; pooled[n][c] = mean over (m, tv).  grid (N*M, slices).
__global__ void pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ pooled, int M,
                                int TV, int C, int rows_per_cta, float inv) {
    __shared__ float red[256 * 4];
    const int c4 = C >> 2;
    const int rows_per_iter = blockDim.x / c4;
    const int col = threadIdx.x % c4, rsub = threadIdx.x / c4;
    const int nm = blockIdx.x;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(TV, r0 + rows_per_cta);
    float a[4] = {0, 0, 0, 0};
    if (rsub < rows_per_iter) {
        for (int r = r0 + rsub; r < r1; r += rows_per_iter) {
            const float4 v = ld4(x + ((size_t)nm * TV + r) * C + col * 4);
            a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) red[threadIdx.x * 4 + j] = a[j];
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
        float s = 0;
        for (int k = 0; k < rows_per_iter; ++k) s += red[(k * c4 + (ch >> 2)) * 4 + (ch & 3)];
        atomicAdd(&pooled[(size_t)(nm / M) * C + ch], s * inv);
    }
}

__global__ void pool_bwd_kernel(const float* __restrict__ gp, float* __restrict__ gx, int M, int TV,
                                int C, long long n4, float inv) {
    const int c4 = C >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) * 4;
        const long long row = i / c4;
        const int n = (int)(row / ((long long)M * TV));
        const float4 g = ld4(gp + (size_t)n * C + c);
        st4(gx + i * 4, make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv));
    }
}

static inline int ew_grid(long long work_items, int threads) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 16;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace istgcn

using namespace istgcn;

ISTGCN_API const char* istgcn_last_error(void) { return g_err; }
ISTGCN_API int istgcn_version(void) { return 100; }

ISTGCN_API int istgcn_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return (int)e;
    }
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    ISTGCN_REQUIRE(major == 10, ISTGCN_E_ARCH,
                   "istgcn_b200 is built for sm_100a only; device has compute capability %d.x",
                   major);
    return 0;
}

ISTGCN_API int istgcn_data_bn_stats(const float* x, double* sum, double* sumsq, int N, int C, int T,
                                    int V, int M, istgcn_stream_t s) {
    ISTGCN_REQUIRE(x && sum && sumsq, ISTGCN_E_ARG, "data_bn_stats: null pointer");
    ISTGCN_REQUIRE(V * M <= 256 && V <= 64, ISTGCN_E_SHAPE, "data_bn_stats: V*M=%d too large", V * M);
    if ((long long)N * C * T == 0) return 0;
    const int slices = T >= 64 ? 4 : 1;
    const int tpc = (T + slices - 1) / slices;
    dim3 grid(N * C, slices);
    data_bn_stats_kernel<<<grid, 256, 2 * V * sizeof(double), (cudaStream_t)s>>>(x, sum, sumsq, C, T,
                                                                                V, M, tpc);
    return finish_launch("data_bn_stats");
}

ISTGCN_API int istgcn_data_bn_apply(const float* x, const float* mean, const float* scale,
                                    const float* beta, float* y, int N, int C, int T, int V, int M,
                                    istgcn_stream_t s) {
    ISTGCN_REQUIRE(x && mean && scale && beta && y, ISTGCN_E_ARG, "data_bn_apply: null pointer");
    const long long rows = (long long)N * M * T * V;
    if (rows == 0) return 0;
    data_bn_apply_kernel<<<ew_grid(rows, 256), 256, 0, (cudaStream_t)s>>>(x, mean, scale, beta, y, N,
                                                                          C, T, V, M);
    return finish_launch("data_bn_apply");
}

ISTGCN_API int istgcn_data_bn_bwd(const float* x, const float* g, const float* mean,
                                  const float* rstd, double* dgamma, double* dbeta, int N, int C,
                                  int T, int V, int M, istgcn_stream_t s) {
    ISTGCN_REQUIRE(x && g && mean && rstd && dgamma && dbeta, ISTGCN_E_ARG,
                   "data_bn_bwd: null pointer");
    ISTGCN_REQUIRE(V * M <= 256 && V <= 64, ISTGCN_E_SHAPE, "data_bn_bwd: V*M=%d too large", V * M);
    if ((long long)N * C * T == 0) return 0;
    const int slices = T >= 64 ? 4 : 1;
    const int tpc = (T + slices - 1) / slices;
    dim3 grid(N * C, slices);
    data_bn_bwd_kernel<<<grid, 256, 2 * V * sizeof(float), (cudaStream_t)s>>>(
        x, g, mean, rstd, dgamma, dbeta, C, T, V, M, tpc);
    return finish_launch("data_bn_bwd");
}

ISTGCN_API int istgcn_bn_finalize(const double* sum, const double* sumsq, double count,
                                  const float* gamma, float* running_mean, float* running_var,
                                  float momentum, float eps, float* scale, float* mean, float* rstd,
                                  int C, istgcn_stream_t s) {
    ISTGCN_REQUIRE(sum && sumsq && gamma && scale && mean && rstd, ISTGCN_E_ARG,
                   "bn_finalize: null pointer");
    ISTGCN_REQUIRE(count > 0, ISTGCN_E_SHAPE, "bn_finalize: empty batch");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)s>>>(
        sum, sumsq, count, gamma, running_mean, running_var, momentum, eps, scale, mean, rstd, C);
    return finish_launch("bn_finalize");
}

ISTGCN_API int istgcn_bn_eval_coeffs(const float* gamma, const float* running_var, float eps,
                                     float* scale, int C, istgcn_stream_t s) {
    ISTGCN_REQUIRE(gamma && running_var && scale, ISTGCN_E_ARG, "bn_eval_coeffs: null pointer");
    bn_eval_coeffs_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)s>>>(gamma, running_var, eps,
                                                                        scale, C);
    return finish_launch("bn_eval_coeffs");
}

ISTGCN_API int istgcn_bn_bwd_coeffs(const double* sg, const double* sgx, double count,
                                    const float* gamma, const float* rstd, float* p, float* m1,
                                    float* c, float* dgamma, float* dbeta, int C, istgcn_stream_t s) {
    ISTGCN_REQUIRE(sg && sgx && gamma && rstd && p && m1 && c, ISTGCN_E_ARG,
                   "bn_bwd_coeffs: null pointer");
    bn_bwd_coeffs_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)s>>>(sg, sgx, count, gamma, rstd, p,
                                                                       m1, c, dgamma, dbeta, C);
    return finish_launch("bn_bwd_coeffs");
}

static int launch_block_tail_fwd(const float* u, const float* mean2, const float* scale2, const float* beta2,
                                 const float* res, const float* mean_r, const float* scale_r,
                                 const float* beta_r, float* out, long long rows, int C, int mode,
                                 float drop_p, uint64_t drop_seed, const unsigned long long* drop_step,
                                 const BnFold& fold2, const BnFold& foldr, istgcn_stream_t s);

ISTGCN_API int istgcn_block_tail_fwd(const float* u, const float* mean2, const float* scale2,
                                     const float* beta2, const float* res, const float* mean_r,
                                     const float* scale_r, const float* beta_r, float* out,
                                     long long rows, int C, float drop_p, uint64_t drop_seed,
                                     const unsigned long long* drop_step, istgcn_stream_t s) {
    ISTGCN_REQUIRE(u && mean2 && scale2 && beta2 && out, ISTGCN_E_ARG, "block_tail_fwd: null pointer");
    ISTGCN_REQUIRE(scale_r == nullptr || (res && mean_r && beta_r), ISTGCN_E_ARG,
                   "block_tail_fwd: residual BatchNorm needs res, mean_r and beta_r");
    const int mode = res == nullptr ? 0 : (scale_r == nullptr ? 1 : 2);
    return launch_block_tail_fwd(u, mean2, scale2, beta2, res, mean_r, scale_r, beta_r, out, rows, C, mode,
                                 drop_p, drop_seed, drop_step, BnFold{}, BnFold{}, s);
}

// The same with the BatchNorm bookkeeping folded in (training): mean2 / scale2 / rstd2 -- and, when the
// residual is a conv + BatchNorm (sum_r != NULL), mean_r / scale_r / rstd_r -- are OUTPUTS derived from
// the raw sums over `count` rows; running statistics updated in place (may be NULL).
ISTGCN_API int istgcn_block_tail_fwd_bn(const float* u, const double* sum2, const double* sumsq2, double count,
                                        const float* gamma2, const float* beta2, float* rmean2, float* rvar2,
                                        float momentum2, float eps2, float* mean2, float* scale2,
                                        float* rstd2, const float* res, const double* sum_r,
                                        const double* sumsq_r, const float* gamma_r, const float* beta_r,
                                        float* rmean_r, float* rvar_r, float momentum_r, float eps_r,
                                        float* mean_r, float* scale_r, float* rstd_r, float* out,
                                        long long rows, int C, float drop_p, uint64_t drop_seed,
                                        const unsigned long long* drop_step, istgcn_stream_t s) {
    ISTGCN_REQUIRE(u && sum2 && sumsq2 && gamma2 && beta2 && mean2 && scale2 && rstd2 && out, ISTGCN_E_ARG,
                   "block_tail_fwd_bn: null pointer");
    ISTGCN_REQUIRE(count > 0, ISTGCN_E_ARG, "block_tail_fwd_bn: empty batch");
    ISTGCN_REQUIRE(sum_r == nullptr || (res && sumsq_r && gamma_r && beta_r && mean_r && scale_r && rstd_r),
                   ISTGCN_E_ARG, "block_tail_fwd_bn: residual BatchNorm needs res, its sums, parameters and outputs");
    const double unb = count > 1 ? count / (count - 1.0) : 1.0;
    const BnFold f2{sum2, sumsq2, 1.0 / count, unb, gamma2, rmean2, rvar2, momentum2, eps2, mean2, scale2, rstd2};
    BnFold fr{};
    if (sum_r) fr = BnFold{sum_r, sumsq_r, 1.0 / count, unb, gamma_r, rmean_r, rvar_r, momentum_r, eps_r,
                           mean_r, scale_r, rstd_r};
    const int mode = res == nullptr ? 0 : (sum_r == nullptr ? 1 : 2);
    return launch_block_tail_fwd(u, mean2, scale2, beta2, res, mean_r, scale_r, beta_r, out, rows, C, mode,
                                 drop_p, drop_seed, drop_step, f2, fr, s);
}

static int launch_block_tail_fwd(const float* u, const float* mean2, const float* scale2, const float* beta2,
                                 const float* res, const float* mean_r, const float* scale_r,
                                 const float* beta_r, float* out, long long rows, int C, int mode,
                                 float drop_p, uint64_t drop_seed, const unsigned long long* drop_step,
                                 const BnFold& fold2, const BnFold& foldr, istgcn_stream_t s) {
    ISTGCN_REQUIRE(C % 4 == 0, ISTGCN_E_SHAPE, "block_tail_fwd: C=%d not a multiple of 4", C);
    ISTGCN_REQUIRE(drop_p >= 0.f && drop_p < 1.f, ISTGCN_E_ARG, "block_tail_fwd: dropout p=%f", drop_p);
    const long long n4 = rows * C / 4;
    if (n4 == 0) return 0;
    // grid stride = blocks * threads must be a multiple of C/4: a thread then keeps its channel group
    const int c4 = C / 4;
    int threads = 256;
    while (threads % c4 != 0 && threads < 1024) threads += 32;
    ISTGCN_REQUIRE(threads % c4 == 0, ISTGCN_E_SHAPE, "block_tail_fwd: C=%d has no block size that is a multiple of C/4", C);
    block_tail_fwd_kernel<<<ew_grid(n4, threads), threads, 0, (cudaStream_t)s>>>(
        u, mean2, scale2, beta2, res, mean_r, scale_r, beta_r, out, n4, C, mode, drop_p,
        1.f / (1.f - drop_p), drop_seed, drop_step, fold2, foldr);
    return finish_launch("block_tail_fwd");
}

ISTGCN_API int istgcn_block_tail_bwd(const float* gout, const float* out, const float* u,
                                     const float* mean2, const float* rstd2, const float* rres,
                                     const float* mean_r, const float* rstd_r, float* go,
                                     double* sg2, double* sg2x, double* sgr, double* sgrx,
                                     long long rows, int C, float drop_p, uint64_t drop_seed,
                                     const unsigned long long* drop_step, istgcn_stream_t s) {
    ISTGCN_REQUIRE(gout && out && u && mean2 && rstd2 && go && sg2 && sg2x, ISTGCN_E_ARG,
                   "block_tail_bwd: null pointer");
    ISTGCN_REQUIRE(C % 4 == 0 && C <= 1024, ISTGCN_E_SHAPE, "block_tail_bwd: C=%d unsupported", C);
    if (rows == 0) return 0;
    const int rows_per_iter = 256 / (C / 4);
    long long blocks = (rows + rows_per_iter * 8 - 1) / (rows_per_iter * 8);
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const float ks = 1.f / (1.f - drop_p);
    if (rres)
        block_tail_bwd_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)s>>>(
            gout, out, u, mean2, rstd2, rres, mean_r, rstd_r, go, sg2, sg2x, sgr, sgrx, rows, C,
            drop_p, ks, drop_seed, drop_step);
    else
        block_tail_bwd_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)s>>>(
            gout, out, u, mean2, rstd2, nullptr, nullptr, nullptr, go, sg2, sg2x, nullptr, nullptr,
            rows, C, drop_p, ks, drop_seed, drop_step);
    return finish_launch("block_tail_bwd");
}

ISTGCN_API int istgcn_dropout_mask(unsigned char* mask, long long n, float p, uint64_t seed,
                                   const unsigned long long* drop_step, istgcn_stream_t s) {
    ISTGCN_REQUIRE(mask, ISTGCN_E_ARG, "dropout_mask: null pointer");
    if (n == 0) return 0;
    dropout_mask_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)s>>>(mask, n, p, seed, drop_step);
    return finish_launch("dropout_mask");
}

ISTGCN_API int istgcn_pool_fwd(const float* x, float* pooled, int N, int M, int TV, int C,
                               istgcn_stream_t s) {
    ISTGCN_REQUIRE(x && pooled, ISTGCN_E_ARG, "pool_fwd: null pointer");
    ISTGCN_REQUIRE(C % 4 == 0 && C <= 1024, ISTGCN_E_SHAPE, "pool_fwd: C=%d unsupported", C);
    if ((long long)N * M * TV == 0) return 0;
    cudaError_t e = cudaMemsetAsync(pooled, 0, sizeof(float) * (size_t)N * C, (cudaStream_t)s);
    if (e != cudaSuccess) { set_error("pool_fwd memset: %s", cudaGetErrorString(e)); return (int)e; }
    int slices = (num_sms() * 4 + N * M - 1) / (N * M);
    if (slices < 1) slices = 1;
    if (slices > TV) slices = TV;
    const int rpc = (TV + slices - 1) / slices;
    slices = (TV + rpc - 1) / rpc;
    dim3 grid(N * M, slices);
    pool_fwd_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(x, pooled, M, TV, C, rpc,
                                                        1.f / ((float)M * (float)TV));
    return finish_launch("pool_fwd");
}

ISTGCN_API int istgcn_pool_bwd(const float* gpooled, float* gx, int N, int M, int TV, int C,
                               istgcn_stream_t s) {
    ISTGCN_REQUIRE(gpooled && gx, ISTGCN_E_ARG, "pool_bwd: null pointer");
    ISTGCN_REQUIRE(C % 4 == 0, ISTGCN_E_SHAPE, "pool_bwd: C=%d unsupported", C);
    const long long n4 = (long long)N * M * TV * C / 4;
    if (n4 == 0) return 0;
    pool_bwd_kernel<<<ew_grid(n4, 256), 256, 0, (cudaStream_t)s>>>(gpooled, gx, M, TV, C, n4,
                                                                   1.f / ((float)M * (float)TV));
    return finish_launch("pool_bwd");
}
