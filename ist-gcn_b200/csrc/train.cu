// Training-step bookkeeping kernels: the optimiser step of the reference's
// processor/recognition.py:152-159 (torch.optim.SGD with momentum, nesterov, weight decay) over
// the FLAT parameter / gradient / momentum buffers of istgcn/dp.py, one launch per bucket.
// The learning rate is read from device memory, so a captured CUDA graph follows the step
// schedule (recognition.py:168-176) without re-capture, and the 1/world gradient average of the
// data-parallel all-reduce is folded into the same pass.
#include "common.cuh"

namespace istgcn {

//   g' = g*grad_scale + wd*p ;  buf = momentum*buf + g' ;  p -= lr * (nesterov ? g' + momentum*buf : buf)
// (torch/optim/sgd.py _single_tensor_sgd with dampening 0: a zero-initialised momentum buffer
// gives the same first step as its `buf = clone(g')` special case.)
__global__ void __launch_bounds__(256) sgd_step_kernel(float* __restrict__ p,
                                                       const float* __restrict__ g,
                                                       float* __restrict__ buf, long long n,
                                                       const float* __restrict__ lr_ptr,
                                                       float momentum, float wd, int nesterov,
                                                       float grad_scale) {
    const float lr = *lr_ptr;
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pv = ld4(p + 4 * i), bv = ld4(buf + 4 * i);
        const float4 gv = ld4(g + 4 * i);
        float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ba[4] = {bv.x, bv.y, bv.z, bv.w};
        const float ga[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gg = fmaf(wd, pa[j], ga[j] * grad_scale);
            ba[j] = fmaf(momentum, ba[j], gg);
            pa[j] -= lr * (nesterov ? fmaf(momentum, ba[j], gg) : ba[j]);
        }
        st4(p + 4 * i, make_float4(pa[0], pa[1], pa[2], pa[3]));
        st4(buf + 4 * i, make_float4(ba[0], ba[1], ba[2], ba[3]));
    }
    // tail (n not a multiple of 4)
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gg = fmaf(wd, p[i], g[i] * grad_scale);
        const float b = fmaf(momentum, buf[i], gg);
        buf[i] = b;
        p[i] -= lr * (nesterov ? fmaf(momentum, b, gg) : b);
    }
}

// Training-time augmentation of the reference's feeder on the device (feeder/feeder.py:70-85 ->
// feeder/tools.py:32-102): temporal window (random_choose / auto_pading: out frame t reads input
// frame t + shift[n], zeros outside) followed by random_move (x, y of EVERY frame of the window,
// padding included, rotated / scaled / translated by the per-frame matrix the host drew with the
// reference's own random calls: move[n][t] = {cos(a)*s, sin(a)*s, t_x, t_y}).
__global__ void feeder_augment_kernel(const float* __restrict__ in, const int* __restrict__ shift,
                                      const float* __restrict__ move, float* __restrict__ out,
                                      int N, int C, int Tin, int Tout, int VM) {
    const long long total = (long long)N * Tout * VM;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int vm = (int)(i % VM);
        const long long nt = i / VM;
        const int t = (int)(nt % Tout), n = (int)(nt / Tout);
        const int ts = t + (shift ? shift[n] : 0);
        const bool ok = ts >= 0 && ts < Tin;
        float x = 0.f, y = 0.f;
        const float* src = in + ((long long)n * C * Tin + ts) * VM + vm;
        float* dst = out + ((long long)n * C * Tout + t) * VM + vm;
        if (ok) {
            x = src[0];
            if (C > 1) y = src[(long long)Tin * VM];
        }
        if (move && C > 1) {
            const float4 m = *reinterpret_cast<const float4*>(move + ((long long)n * Tout + t) * 4);
            const float nx = m.x * x - m.y * y + m.z, ny = m.y * x + m.x * y + m.w;
            x = nx;
            y = ny;
        }
        dst[0] = x;
        if (C > 1) dst[(long long)Tout * VM] = y;
        for (int c = 2; c < C; ++c) dst[(long long)c * Tout * VM] = ok ? src[(long long)c * Tin * VM] : 0.f;
    }
}

}  // namespace istgcn

using namespace istgcn;

ISTGCN_API int istgcn_feeder_augment(const float* in, const int* shift, const float* move, float* out,
                                     int N, int C, int Tin, int Tout, int V, int M,
                                     istgcn_stream_t s) {
    ISTGCN_REQUIRE(in && out, ISTGCN_E_ARG, "feeder_augment: null pointer");
    ISTGCN_REQUIRE(N >= 0 && C >= 1 && Tin >= 1 && Tout >= 1 && V >= 1 && M >= 1, ISTGCN_E_SHAPE,
                   "feeder_augment: bad shape");
    ISTGCN_REQUIRE(move == nullptr || (uintptr_t)move % 16 == 0, ISTGCN_E_ARG, "feeder_augment: move must be 16-byte aligned");
    const long long total = (long long)N * Tout * V * M;
    if (total == 0) return 0;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    feeder_augment_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(in, shift, move, out, N, C, Tin, Tout, V * M);
    return finish_launch("feeder_augment");
}

ISTGCN_API int istgcn_sgd_step(float* p, const float* g, float* buf, long long n, const float* lr,
                               float momentum, float weight_decay, int nesterov, float grad_scale,
                               istgcn_stream_t s) {
    ISTGCN_REQUIRE(p && g && buf && lr, ISTGCN_E_ARG, "sgd_step: null pointer");
    ISTGCN_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)buf) % 16 == 0, ISTGCN_E_ARG,
                   "sgd_step: buffers must be 16-byte aligned");
    if (n <= 0) return 0;
    long long blocks = ((n + 3) / 4 + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    sgd_step_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(p, g, buf, n, lr, momentum,
                                                             weight_decay, nesterov, grad_scale);
    return finish_launch("sgd_step");
}
