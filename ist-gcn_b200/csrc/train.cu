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


// ------------------------------------------------------------------------------------------
// Parameter regrouping of one IST-GCN block in ONE kernel each way (instead of ~50 tiny ATen ops
// and as many again in their autograd): reference-layout parameters -> the operands of the fused
// kernels (istgcn/modules.py:graph_conv_operands / bottleneck_tcn_operands state the algebra).
struct PrepP {
    // graph conv
    const float *W, *bias, *A[3], *imp[3];
    const long long* flat_idx;
    const int *dst_ptr, *dst_id;
    float *vals, *colsum, *Wc, *biasterm;
    // bottleneck TCN
    const float *Ws, *bs, *Wt[3], *bt[3], *We, *m_imp;
    float *Wd, *bd, *Weff, *beff, *Wu;
    // strided-conv residual
    const float *Wres, *bres;
    float *Wr, *btr;
    int K, V, Cin, Cout, b, bp, nnz, ns;
};

__device__ __forceinline__ float a_eff_at(const PrepP& p, long long idx) {
    float v = 0.f;
    for (int i = 0; i < p.ns; ++i) v += p.A[i][idx] * (p.imp[i] ? p.imp[i][idx] : 1.f);
    return v;
}

__global__ void __launch_bounds__(256) block_prep_fwd_kernel(PrepP p) {
    const int gsz = gridDim.x * blockDim.x, gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int K = p.K, V = p.V, Cin = p.Cin, Cout = p.Cout, b = p.b, bp = p.bp;
    for (int i = gid; i < p.nnz; i += gsz) p.vals[i] = a_eff_at(p, p.flat_idx[i]);
    for (int i = gid; i < K * V; i += gsz) {                          // colsum[k][w]
        float s = 0.f;
        for (int j = p.dst_ptr[i]; j < p.dst_ptr[i + 1]; ++j) s += a_eff_at(p, p.flat_idx[p.dst_id[j]]);
        p.colsum[i] = s;
    }
    for (int i = gid; i < K * Cin * Cout; i += gsz) {                 // Wc[k*Cin+ci][c] = W[k*Cout+c][ci]
        const int ci = i % Cin, c = (i / Cin) % Cout, k = i / (Cout * Cin);      // coalesced reads, strided writes
        p.Wc[((long long)k * Cin + ci) * Cout + c] = p.W[i];
    }
    for (int i = gid; i < V * Cout; i += gsz) {                       // biasterm[w][c]
        const int c = i % Cout, w = i / Cout;
        float s = 0.f;
        if (p.bias)
            for (int k = 0; k < K; ++k) {
                float cs = 0.f;
                for (int j = p.dst_ptr[k * V + w]; j < p.dst_ptr[k * V + w + 1]; ++j)
                    cs += a_eff_at(p, p.flat_idx[p.dst_id[j]]);
                s = fmaf(p.bias[k * Cout + c], cs, s);
            }
        p.biasterm[i] = s;
    }
    for (int i = gid; i < Cout * bp; i += gsz) {                      // Wd[c][j], Wu[j][c]
        const int j = i % bp, c = i / bp;
        p.Wd[i] = j < b ? p.Ws[j * Cout + c] : 0.f;
        p.Wu[j * Cout + c] = j < b ? p.We[c * b + j] : 0.f;
    }
    for (int i = gid; i < bp; i += gsz) {
        p.bd[i] = i < b ? p.bs[i] : 0.f;
        float s = 0.f;
        if (i < b)
            for (int q = 0; q < 3; ++q) s = fmaf(p.m_imp[q], p.bt[q][i], s);
        p.beff[i] = s;
    }
    for (int i = gid; i < 15 * bp * bp; i += gsz) {                   // Weff[tap][ci][co]
        const int co = i % bp, ci = (i / bp) % bp, tap = i / (bp * bp);
        float s = 0.f;
        if (co < b && ci < b) {
            const int kt[3] = {3, 9, 15};
            for (int q = 0; q < 3; ++q) {
                const int kk = tap - (15 - kt[q]) / 2;
                if (kk >= 0 && kk < kt[q]) s = fmaf(p.m_imp[q], p.Wt[q][(co * b + ci) * kt[q] + kk], s);
            }
        }
        p.Weff[i] = s;
    }
    if (p.Wres) {
        for (int i = gid; i < Cin * Cout; i += gsz) {                 // Wr[ci][co] = Wres[co][ci]
            const int ci = i % Cin, co = i / Cin;
            p.Wr[ci * Cout + co] = p.Wres[i];
        }
        for (int i = gid; i < V * Cout; i += gsz) p.btr[i] = p.bres[i % Cout];
    }
}

struct PrepBwdP {
    const float *dvals, *dWc, *dbt, *dWd, *dbd, *dWeff, *dbeff, *dWu, *dWr, *dbtr;
    const float *bias, *colsum, *A[3], *Wt[3], *bt[3], *m_imp;
    const int *inv_idx, *id_kw;
    float *dW, *dbias, *dimp[3], *dWs, *dbs, *dWt[3], *dbtc[3], *dWe, *dm_imp, *dWres, *dbres;
    int K, V, Cin, Cout, b, bp, nnz, ns;
};

__global__ void __launch_bounds__(256) block_prep_bwd_kernel(PrepBwdP p) {
    const int gsz = gridDim.x * blockDim.x, gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int K = p.K, V = p.V, Cin = p.Cin, Cout = p.Cout, b = p.b, bp = p.bp;
    // importances: d imp_i[k][v][w] = A_i * (dvals[id] + sum_c bias[k][c] dbt[w][c]) on the pattern, 0 elsewhere
    // (one WARP per stack element: the lanes share the dot product over the channels)
    const int lane = threadIdx.x & 31, gwarp = gid >> 5, nwarp = gsz >> 5;
    for (int i = gwarp; i < K * V * V; i += nwarp) {
        const int id = p.inv_idx[i];
        float gsum = 0.f;
        if (id >= 0) {
            if (p.bias) {
                const int kw = p.id_kw[id], k = kw / V, w = kw - k * V;
                float q = 0.f;
                for (int c = lane; c < Cout; c += 32) q = fmaf(p.bias[k * Cout + c], p.dbt[w * Cout + c], q);
                gsum = warp_sum(q);
            }
            gsum += p.dvals[id];
        }
        if (lane < p.ns && p.dimp[lane]) p.dimp[lane][i] = id >= 0 ? p.A[lane][i] * gsum : 0.f;
    }
    for (int i = gid; i < K * Cout * Cin; i += gsz) {                 // dW[k*Cout+c][ci] = dWc[k*Cin+ci][c]
        const int c = i % Cout, ci = (i / Cout) % Cin, k = i / (Cin * Cout);     // coalesced reads, strided writes
        p.dW[((long long)k * Cout + c) * Cin + ci] = p.dWc[i];
    }
    if (p.dbias)
        for (int i = gid; i < K * Cout; i += gsz) {                   // dbias[k][c] = sum_w colsum[k][w] dbt[w][c]
            const int c = i % Cout, k = i / Cout;
            float s = 0.f;
            for (int w = 0; w < V; ++w) s = fmaf(p.colsum[k * V + w], p.dbt[w * Cout + c], s);
            p.dbias[i] = s;
        }
    for (int i = gid; i < b * Cout; i += gsz) {                       // dWs[j][c] = dWd[c][j]; dWe[c][j] = dWu[j][c]
        const int c = i % Cout, j = i / Cout;
        p.dWs[i] = p.dWd[c * bp + j];
        p.dWe[c * b + j] = p.dWu[j * Cout + c];
    }
    for (int i = gid; i < b; i += gsz) {
        p.dbs[i] = p.dbd[i];
        for (int q = 0; q < 3; ++q) p.dbtc[q][i] = p.m_imp[q] * p.dbeff[i];
    }
    const int kt[3] = {3, 9, 15};
    for (int q = 0; q < 3; ++q)
        for (int i = gid; i < b * b * kt[q]; i += gsz) {              // dW_q[co][ci][kk] = m_q dWeff[off+kk][ci][co]
            const int kk = i % kt[q], ci = (i / kt[q]) % b, co = i / (kt[q] * b);
            p.dWt[q][i] = p.m_imp[q] * p.dWeff[((kk + (15 - kt[q]) / 2) * bp + ci) * bp + co];
        }
    if (p.dWres) {
        for (int i = gid; i < Cout * Cin; i += gsz) {                 // dWres[co][ci] = dWr[ci][co]
            const int co = i % Cout, ci = i / Cout;
            p.dWres[co * Cin + ci] = p.dWr[i];
        }
        for (int i = gid; i < Cout; i += gsz) {
            float s = 0.f;
            for (int v = 0; v < V; ++v) s += p.dbtr[v * Cout + i];
            p.dbres[i] = s;
        }
    }
    // d m_imp[q] = <W_q, dWeff window> + <b_q, dbeff>: blocks 0..2 take one branch each
    if (blockIdx.x < 3) {
        __shared__ float red[8];
        const int q = blockIdx.x;
        float s = 0.f;
        for (int i = threadIdx.x; i < b * b * kt[q]; i += blockDim.x) {
            const int kk = i % kt[q], ci = (i / kt[q]) % b, co = i / (kt[q] * b);
            s = fmaf(p.Wt[q][i], p.dWeff[((kk + (15 - kt[q]) / 2) * bp + ci) * bp + co], s);
        }
        for (int i = threadIdx.x; i < b; i += blockDim.x) s = fmaf(p.bt[q][i], p.dbeff[i], s);
        s = warp_sum(s);
        if (lane == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int wq = 0; wq < (int)(blockDim.x >> 5); ++wq) t += red[wq];
            p.dm_imp[q] = t;
        }
    }
}

}  // namespace istgcn

using namespace istgcn;

ISTGCN_API int istgcn_block_prep_fwd(const float* W, const float* bias, const float* A1, const float* imp1,
                                     const float* A2, const float* imp2, const float* A3, const float* imp3,
                                     const long long* flat_idx, const int* dst_ptr, const int* dst_id, int nnz,
                                     float* vals, float* colsum, float* Wc, float* biasterm,
                                     const float* Ws, const float* bs, const float* W1, const float* b1,
                                     const float* W2, const float* b2, const float* W3, const float* b3,
                                     const float* We, const float* m_imp, float* Wd, float* bd, float* Weff,
                                     float* beff, float* Wu, const float* Wres, const float* bres, float* Wr,
                                     float* btr, int K, int V, int Cin, int Cout, int b, int bp,
                                     istgcn_stream_t s) {
    ISTGCN_REQUIRE(W && A1 && flat_idx && dst_ptr && dst_id && vals && colsum && Wc && biasterm && Ws && bs &&
                       W1 && b1 && W2 && b2 && W3 && b3 && We && m_imp && Wd && bd && Weff && beff && Wu,
                   ISTGCN_E_ARG, "block_prep_fwd: null pointer");
    ISTGCN_REQUIRE((A2 == nullptr) == (A3 == nullptr), ISTGCN_E_ARG, "block_prep_fwd: A2 / A3 come together");
    ISTGCN_REQUIRE((Wres == nullptr) == (Wr == nullptr) && (Wres == nullptr) == (bres == nullptr) &&
                       (Wres == nullptr) == (btr == nullptr), ISTGCN_E_ARG, "block_prep_fwd: residual pointers");
    ISTGCN_REQUIRE(b >= 1 && b <= bp && (bp == 8 || bp == 16), ISTGCN_E_SHAPE, "block_prep_fwd: b=%d bp=%d", b, bp);
    PrepP p{W, bias, {A1, A2, A3}, {imp1, imp2, imp3}, flat_idx, dst_ptr, dst_id, vals, colsum, Wc, biasterm,
            Ws, bs, {W1, W2, W3}, {b1, b2, b3}, We, m_imp, Wd, bd, Weff, beff, Wu, Wres, bres, Wr, btr,
            K, V, Cin, Cout, b, bp, nnz, A2 ? 3 : 1};
    const int work = K * Cin * Cout;
    int blocks = (work + 255) / 256;
    if (blocks > num_sms() * 2) blocks = num_sms() * 2;
    if (blocks < 4) blocks = 4;
    block_prep_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(p);
    return finish_launch("block_prep_fwd");
}

ISTGCN_API int istgcn_block_prep_bwd(const float* dvals, const float* dWc, const float* dbt, const float* dWd,
                                     const float* dbd, const float* dWeff, const float* dbeff, const float* dWu,
                                     const float* dWr, const float* dbtr, const float* bias, const float* colsum,
                                     const float* A1, const float* A2, const float* A3, const int* inv_idx,
                                     const int* id_kw, const float* W1, const float* b1, const float* W2,
                                     const float* b2, const float* W3, const float* b3, const float* m_imp,
                                     float* dW, float* dbias, float* dimp1, float* dimp2, float* dimp3,
                                     float* dWs, float* dbs, float* dW1, float* db1, float* dW2, float* db2,
                                     float* dW3, float* db3, float* dWe, float* dm_imp, float* dWres,
                                     float* dbres, int K, int V, int Cin, int Cout, int b, int bp,
                                     istgcn_stream_t s) {
    ISTGCN_REQUIRE(dvals && dWc && dbt && dWd && dbd && dWeff && dbeff && dWu && colsum && A1 && inv_idx &&
                       id_kw && W1 && b1 && W2 && b2 && W3 && b3 && m_imp && dW && dWs && dbs && dW1 && db1 &&
                       dW2 && db2 && dW3 && db3 && dWe && dm_imp,
                   ISTGCN_E_ARG, "block_prep_bwd: null pointer");
    ISTGCN_REQUIRE((dWres == nullptr) == (dWr == nullptr) && (dWres == nullptr) == (dbtr == nullptr) &&
                       (dWres == nullptr) == (dbres == nullptr), ISTGCN_E_ARG, "block_prep_bwd: residual pointers");
    ISTGCN_REQUIRE((bias == nullptr) == (dbias == nullptr), ISTGCN_E_ARG, "block_prep_bwd: bias pointers");
    PrepBwdP p{dvals, dWc, dbt, dWd, dbd, dWeff, dbeff, dWu, dWr, dbtr, bias, colsum, {A1, A2, A3},
               {W1, W2, W3}, {b1, b2, b3}, m_imp, inv_idx, id_kw, dW, dbias, {dimp1, dimp2, dimp3}, dWs, dbs,
               {dW1, dW2, dW3}, {db1, db2, db3}, dWe, dm_imp, dWres, dbres, K, V, Cin, Cout, b, bp, 0,
               A2 ? 3 : 1};
    const int work = K * Cin * Cout;
    int blocks = (work + 255) / 256;
    if (blocks > num_sms() * 2) blocks = num_sms() * 2;
    if (blocks < 4) blocks = 4;
    block_prep_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(p);
    return finish_launch("block_prep_bwd");
}

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
