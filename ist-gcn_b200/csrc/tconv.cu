// Element-wise stages around the full-width temporal convolution of the baseline ST-GCN block
// (reference: net/st_gcnold.py:160-174, net/st_gcn_msgcn.py, net/st_gcn_mstcn.py):
//
//     a  = relu(BN1(z))                         istgcn_bn_relu_apply   (forward)
//     u  = sum_tap shift_tap(a) * W_tap         kt launches of the strided / shifted 1x1 engine
//     du = BN2-backward(dropout-mask(go), u)    istgcn_bn_back_apply   (backward)
//     da = sum_tap shift_tap^T(du) * W_tap^T    kt launches, accumulated
//     g1 = da * (a > 0), sums for BN1-backward  istgcn_relu_bn_bwd
//
// All three are HBM-bound float4 streams over channels-last rows.
#include "common.cuh"

namespace istgcn {

__global__ void bn_relu_apply_kernel(const float* __restrict__ z, const float* __restrict__ mean,
                                     const float* __restrict__ scale, const float* __restrict__ beta,
                                     float* __restrict__ a, long long n4, int C) {
    const int c4 = C >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) * 4;
        const float4 zv = ld4(z + i * 4);
        const float4 mu = ld4(mean + c), sc = ld4(scale + c), be = ld4(beta + c);
        st4(a + i * 4, make_float4(fmaxf(bn_apply(zv.x, mu.x, sc.x, be.x), 0.f),
                                   fmaxf(bn_apply(zv.y, mu.y, sc.y, be.y), 0.f),
                                   fmaxf(bn_apply(zv.z, mu.z, sc.z, be.z), 0.f),
                                   fmaxf(bn_apply(zv.w, mu.w, sc.w, be.w), 0.f)));
    }
}

__global__ void bn_back_apply_kernel(const float* __restrict__ go, const float* __restrict__ u,
                                     const float* __restrict__ p, const float* __restrict__ m1,
                                     const float* __restrict__ cc, const float* __restrict__ mean,
                                     float* __restrict__ du, long long n4, int C, float drop_p,
                                     float keep_scale, uint64_t seed0, const unsigned long long* step) {
    const uint64_t seed = effective_seed(seed0, step);
    const int c4 = C >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) * 4;
        const float4 gv = ld4(go + i * 4), uv = ld4(u + i * 4);
        const float4 pv = ld4(p + c), mv = ld4(m1 + c), cv = ld4(cc + c), nv = ld4(mean + c);
        float g[4] = {gv.x, gv.y, gv.z, gv.w};
        if (drop_p > 0.f) {
            bool keep[4];
            dropout_keep4(seed, (uint64_t)i, drop_p, keep);
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] = keep[j] ? g[j] * keep_scale : 0.f;
        }
        st4(du + i * 4, make_float4(bn_back(g[0], uv.x, pv.x, mv.x, cv.x, nv.x),
                                    bn_back(g[1], uv.y, pv.y, mv.y, cv.y, nv.y),
                                    bn_back(g[2], uv.z, pv.z, mv.z, cv.z, nv.z),
                                    bn_back(g[3], uv.w, pv.w, mv.w, cv.w, nv.w)));
    }
}

// dz = BatchNorm-backward(g, z) written out AND summed over frames in the same pass:
// colsum[v][c] += sum_f dz[(f,v)][c] (the gradient of the graph convolution's bias term,
// tgcn.py:79).  Thread = one float4 column of the [V*C] frame vector, looping over the frames of
// its slab: the frame_colsum launch that re-read dz (0.06 ms per layer) is gone.
__global__ void bn_back_colsum_kernel(const float* __restrict__ g, const float* __restrict__ z,
                                      const float* __restrict__ p, const float* __restrict__ m1,
                                      const float* __restrict__ cc, const float* __restrict__ mean,
                                      float* __restrict__ dz, float* __restrict__ colsum, int frames,
                                      int n, int C, int frames_per_cta, BnBwdFold fold) {
    // flat (slab, column) index: no idle lanes when V*C/4 is not a multiple of the block size
    const int n4 = n / 4;
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int slab = (int)(gid / n4);
    const int f0 = slab * frames_per_cta, f1 = min(frames, f0 + frames_per_cta);
    for (int j4 = (int)(gid - (long long)slab * n4); f0 < f1 && j4 < n4; j4 += n4) {
        const int c = (j4 * 4) % C;
        float4 pv, mv, cv;
        if (fold.sg) {                   // coefficients from the raw sums; the first C/4 threads store them
            const bool wr = gid < C / 4;
            bn_bwd_fold(fold, c, wr, pv.x, mv.x, cv.x);
            bn_bwd_fold(fold, c + 1, wr, pv.y, mv.y, cv.y);
            bn_bwd_fold(fold, c + 2, wr, pv.z, mv.z, cv.z);
            bn_bwd_fold(fold, c + 3, wr, pv.w, mv.w, cv.w);
        } else {
            pv = ld4(p + c); mv = ld4(m1 + c); cv = ld4(cc + c);
        }
        const float4 nv = ld4(mean + c);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        int f = f0;
        for (; f + 4 <= f1; f += 4) {                 // eight independent loads in flight
            float4 gv[4], zv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const size_t off = (size_t)(f + u) * n + j4 * 4;
                gv[u] = ld4(g + off);
                zv[u] = ld4(z + off);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 d = make_float4(bn_back(gv[u].x, zv[u].x, pv.x, mv.x, cv.x, nv.x),
                                             bn_back(gv[u].y, zv[u].y, pv.y, mv.y, cv.y, nv.y),
                                             bn_back(gv[u].z, zv[u].z, pv.z, mv.z, cv.z, nv.z),
                                             bn_back(gv[u].w, zv[u].w, pv.w, mv.w, cv.w, nv.w));
                st4(dz + (size_t)(f + u) * n + j4 * 4, d);
                a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
            }
        }
        for (; f < f1; ++f) {
            const size_t off = (size_t)f * n + j4 * 4;
            const float4 gv = ld4(g + off), zv = ld4(z + off);
            const float4 d = make_float4(bn_back(gv.x, zv.x, pv.x, mv.x, cv.x, nv.x),
                                         bn_back(gv.y, zv.y, pv.y, mv.y, cv.y, nv.y),
                                         bn_back(gv.z, zv.z, pv.z, mv.z, cv.z, nv.z),
                                         bn_back(gv.w, zv.w, pv.w, mv.w, cv.w, nv.w));
            st4(dz + off, d);
            a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
        }
        red_add4(colsum + j4 * 4, a);
    }
}

// thread = (row sub-group, 4 channels); per-thread partial sums, one double atomic per channel
// and CTA at the end
__global__ void relu_bn_bwd_kernel(const float* __restrict__ da, const float* __restrict__ a,
                                   const float* __restrict__ z, const float* __restrict__ mean1,
                                   const float* __restrict__ rstd1, float* __restrict__ g1,
                                   double* __restrict__ sg, double* __restrict__ sgx, long long rows,
                                   int C) {
    __shared__ double red[2][256 * 4];
    const int c4 = C >> 2;
    const int rows_per_iter = blockDim.x / c4;
    const int col = threadIdx.x % c4, rsub = threadIdx.x / c4;
    const int c = col * 4;
    double a_g[4] = {0, 0, 0, 0}, a_gx[4] = {0, 0, 0, 0};
    float mu[4], rs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        mu[j] = mean1[c + j];
        rs[j] = rstd1[c + j];
    }
    if (rsub < rows_per_iter) {
        for (long long r = (long long)blockIdx.x * rows_per_iter + rsub; r < rows;
             r += (long long)gridDim.x * rows_per_iter) {
            const long long off = r * C + c;
            const float4 dv = ld4(da + off), av = ld4(a + off), zv = ld4(z + off);
            const float d4[4] = {dv.x, dv.y, dv.z, dv.w}, a4[4] = {av.x, av.y, av.z, av.w},
                        z4[4] = {zv.x, zv.y, zv.z, zv.w};
            float g[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                g[j] = a4[j] > 0.f ? d4[j] : 0.f;
                a_g[j] += g[j];
                a_gx[j] += g[j] * ((z4[j] - mu[j]) * rs[j]);
            }
            st4(g1 + off, make_float4(g[0], g[1], g[2], g[3]));
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        red[0][threadIdx.x * 4 + j] = a_g[j];
        red[1][threadIdx.x * 4 + j] = a_gx[j];
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
        const int cc = ch >> 2, j = ch & 3;
        double s0 = 0, s1 = 0;
        for (int k = 0; k < rows_per_iter; ++k) {
            const int th = k * c4 + cc;
            s0 += red[0][th * 4 + j];
            s1 += red[1][th * 4 + j];
        }
        atomicAdd(&sg[ch], s0);
        atomicAdd(&sgx[ch], s1);
    }
}

static inline int stream_grid(long long work_items, int threads) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 16;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace istgcn

using namespace istgcn;

ISTGCN_API int istgcn_bn_relu_apply(const float* z, const float* mean, const float* scale,
                                    const float* beta, float* a, long long rows, int C,
                                    istgcn_stream_t s) {
    ISTGCN_REQUIRE(z && mean && scale && beta && a, ISTGCN_E_ARG, "bn_relu_apply: null pointer");
    ISTGCN_REQUIRE(C % 4 == 0, ISTGCN_E_SHAPE, "bn_relu_apply: C=%d not a multiple of 4", C);
    const long long n4 = rows * C / 4;
    if (n4 == 0) return 0;
    bn_relu_apply_kernel<<<stream_grid(n4, 256), 256, 0, (cudaStream_t)s>>>(z, mean, scale, beta, a, n4, C);
    return finish_launch("bn_relu_apply");
}

ISTGCN_API int istgcn_bn_back_apply(const float* go, const float* u, const float* p, const float* m1,
                                    const float* c, const float* mean, float* du, long long rows,
                                    int C, float drop_p, uint64_t drop_seed,
                                    const unsigned long long* drop_step, istgcn_stream_t s) {
    ISTGCN_REQUIRE(go && u && p && m1 && c && mean && du, ISTGCN_E_ARG, "bn_back_apply: null pointer");
    ISTGCN_REQUIRE(C % 4 == 0, ISTGCN_E_SHAPE, "bn_back_apply: C=%d not a multiple of 4", C);
    ISTGCN_REQUIRE(drop_p >= 0.f && drop_p < 1.f, ISTGCN_E_ARG, "bn_back_apply: dropout p=%f", drop_p);
    const long long n4 = rows * C / 4;
    if (n4 == 0) return 0;
    bn_back_apply_kernel<<<stream_grid(n4, 256), 256, 0, (cudaStream_t)s>>>(
        go, u, p, m1, c, mean, du, n4, C, drop_p, 1.f / (1.f - drop_p), drop_seed, drop_step);
    return finish_launch("bn_back_apply");
}

// dz[frames*V][C] = p*((g - m1) - c*(z - mean)) and colsum[V][C] += sum over frames of dz
// (caller-zeroed) in one pass.  C % 4 == 0.
static int launch_bn_back_colsum(const float* g, const float* z, const float* p, const float* m1,
                                 const float* c, const float* mean, float* dz, float* colsum, int frames,
                                 int V, int C, const BnBwdFold& fold, istgcn_stream_t s);

ISTGCN_API int istgcn_bn_back_colsum(const float* g, const float* z, const float* p, const float* m1,
                                     const float* c, const float* mean, float* dz, float* colsum,
                                     int frames, int V, int C, istgcn_stream_t s) {
    ISTGCN_REQUIRE(g && z && p && m1 && c && mean && dz && colsum, ISTGCN_E_ARG,
                   "bn_back_colsum: null pointer");
    return launch_bn_back_colsum(g, z, p, m1, c, mean, dz, colsum, frames, V, C, BnBwdFold{}, s);
}

// The same with istgcn_bn_bwd_coeffs folded in: p / m1 / c (and dgamma / dbeta, may be NULL) are OUTPUTS
// derived from sg = sum g, sgx = sum g * zhat over `count` rows.
ISTGCN_API int istgcn_bn_back_colsum_bn(const float* g, const float* z, const double* sg, const double* sgx,
                                        double count, const float* gamma, const float* rstd, float* p,
                                        float* m1, float* c, float* dgamma, float* dbeta, const float* mean,
                                        float* dz, float* colsum, int frames, int V, int C,
                                        istgcn_stream_t s) {
    ISTGCN_REQUIRE(g && z && sg && sgx && gamma && rstd && p && m1 && c && mean && dz && colsum, ISTGCN_E_ARG,
                   "bn_back_colsum_bn: null pointer");
    ISTGCN_REQUIRE(count > 0, ISTGCN_E_ARG, "bn_back_colsum_bn: empty batch");
    const BnBwdFold fold{sg, sgx, 1.0 / count, gamma, rstd, p, m1, c, dgamma, dbeta};
    return launch_bn_back_colsum(g, z, p, m1, c, mean, dz, colsum, frames, V, C, fold, s);
}

static int launch_bn_back_colsum(const float* g, const float* z, const float* p, const float* m1,
                                 const float* c, const float* mean, float* dz, float* colsum, int frames,
                                 int V, int C, const BnBwdFold& fold, istgcn_stream_t s) {
    ISTGCN_REQUIRE(C % 4 == 0 && V >= 1, ISTGCN_E_SHAPE, "bn_back_colsum: C=%d V=%d", C, V);
    ISTGCN_REQUIRE((reinterpret_cast<uintptr_t>(colsum) & 15) == 0, ISTGCN_E_ARG,
                   "bn_back_colsum: colsum must be 16-byte aligned");
    if (frames == 0) return 0;
    const int n = V * C;
    int slabs = (num_sms() * 8 * 256) / (n / 4);
    if (slabs < 1) slabs = 1;
    if (slabs > frames) slabs = frames;
    const int fpc = (frames + slabs - 1) / slabs;
    const long long items = (long long)(n / 4) * ((frames + fpc - 1) / fpc);
    const int grid = (int)((items + 255) / 256);
    bn_back_colsum_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(g, z, p, m1, c, mean, dz, colsum, frames, n,
                                                             C, fpc, fold);
    return finish_launch("bn_back_colsum");
}

ISTGCN_API int istgcn_relu_bn_bwd(const float* da, const float* a, const float* z, const float* mean1,
                                  const float* rstd1, float* g1, double* sg, double* sgx,
                                  long long rows, int C, istgcn_stream_t s) {
    ISTGCN_REQUIRE(da && a && z && mean1 && rstd1 && g1 && sg && sgx, ISTGCN_E_ARG,
                   "relu_bn_bwd: null pointer");
    ISTGCN_REQUIRE(C % 4 == 0 && C <= 1024, ISTGCN_E_SHAPE, "relu_bn_bwd: C=%d unsupported", C);
    if (rows == 0) return 0;
    const int rows_per_iter = 256 / (C / 4);
    long long blocks = (rows + rows_per_iter * 8 - 1) / (rows_per_iter * 8);
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    relu_bn_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(da, a, z, mean1, rstd1, g1, sg, sgx,
                                                                 rows, C);
    return finish_launch("relu_bn_bwd");
}
