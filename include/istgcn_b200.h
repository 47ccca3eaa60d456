/* istgcn_b200 -- C ABI of the B200-native IST-GCN hot path (libistgcn_b200.so).
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a CUDA stream and
 * returns 0 on success (a cudaError_t value, or a negative ISTGCN_E_* code for bad arguments;
 * istgcn_last_error() holds the text).  Nothing is allocated or freed inside: the caller owns
 * every buffer.  All activations are fp32, channels-last: a tensor the reference holds as
 * (N*M, C, T, V) (net/st_gcnold.py:80) lives here as rows[(n*T + t)*V + v][C].
 *
 * The reference is pure Python/PyTorch, so there is no FFI to mirror; each function below
 * names the reference lines whose PyTorch ops it replaces.  The Python binding that the
 * reference-side drop-in (ist-gcn_b200/net/*.py) uses is ist-gcn_b200/istgcn/_lib.py (ctypes);
 * INTEGRATION.md shows the stub.
 *
 * Sparse adjacency ("SpA" arguments).  A_eff = A*imp (+ A2*imp2 + A3*imp3) has a static
 * non-zero pattern.  The caller passes the values of the nnz entries in
 * canonical order (sorted by k, v, w) in `vals[nnz]` and four index arrays built once:
 *   dst_ptr[K*V+1], dst_src[nnz], dst_id[nnz]   entries grouped by destination (k, w):
 *                                               source joint v and canonical id
 *   src_ptr[V+1],   src_kw[nnz],  src_id[nnz]   entries grouped by source v: k*V+w and id
 */
#ifndef ISTGCN_B200_H
#define ISTGCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* istgcn_stream_t; /* cudaStream_t */

#define ISTGCN_E_SHAPE (-1)   /* unsupported shape / channel count            */
#define ISTGCN_E_ARG   (-2)   /* null pointer, bad flag                       */
#define ISTGCN_E_ARCH  (-3)   /* device is not sm_100                         */

/* math mode of the tensor-core GEMMs */
#define ISTGCN_MATH_TF32   0  /* one TF32 pass (fast mode, <=2e-2 budget)     */
#define ISTGCN_MATH_3XTF32 1  /* error-compensated split, fp32-grade (<=1e-4) */

const char* istgcn_last_error(void);
int istgcn_version(void);
/* 0 when the current device can run the library (compute capability 10.x) */
int istgcn_check_device(void);

/* ---- data_bn: nn.BatchNorm1d(V*C) on (N*M, V*C, T) + the two permutes ------------------
 * reference: net/st_gcnold.py:74-80.  x is the model input (N, C, T, V, M) contiguous.
 * stats: per channel ch = v*C + c over (N, M, T), written as double sum[VC], sumsq[VC]
 * (must be zeroed by the caller).  apply: y rows[((n*M+m)*T + t)*V + v][C].               */
int istgcn_data_bn_stats(const float* x, double* sum, double* sumsq,
                         int N, int C, int T, int V, int M, istgcn_stream_t s);
int istgcn_data_bn_apply(const float* x, const float* mean, const float* scale,
                         const float* beta, float* y, int N, int C, int T, int V, int M,
                         istgcn_stream_t s);
/* backward of the affine parameters only (the network input needs no gradient):
 * dgamma[ch] = sum g*xhat, dbeta[ch] = sum g, with g rows[...][C] channels-last.           */
int istgcn_data_bn_bwd(const float* x, const float* g, const float* mean, const float* rstd,
                       double* dgamma, double* dbeta, int N, int C, int T, int V, int M,
                       istgcn_stream_t s);

/* ---- BatchNorm bookkeeping (nn.BatchNorm2d train/eval; st_gcnold.py:165,174) -----------
 * Every kernel applies BatchNorm in the cancellation-free form y = (x - mean)*scale + beta and
 * its backward as dx = p*((g - m1) - c*(x - mean)).
 * finalize: from double sum/sumsq over `count` elements -> mean, rstd (biased variance),
 * scale = gamma*rstd; running stats updated with momentum (unbiased variance) when
 * running_mean != NULL.                                                                    */
int istgcn_bn_finalize(const double* sum, const double* sumsq, double count,
                       const float* gamma, float* running_mean, float* running_var,
                       float momentum, float eps, float* scale, float* mean, float* rstd, int C,
                       istgcn_stream_t s);
/* eval mode: scale = gamma / sqrt(running_var + eps) (mean = running_mean is passed as is) */
int istgcn_bn_eval_coeffs(const float* gamma, const float* running_var, float eps, float* scale,
                          int C, istgcn_stream_t s);
/* backward coefficients: with sg = sum g, sgx = sum g*xhat (double) over `count` elements:
 * p = gamma*rstd, m1 = sg/count, c = rstd*sgx/count; also dgamma = sgx, dbeta = sg.        */
int istgcn_bn_bwd_coeffs(const double* sg, const double* sgx, double count, const float* gamma,
                         const float* rstd, float* p, float* m1, float* c, float* dgamma,
                         float* dbeta, int C, istgcn_stream_t s);

/* ---- fused graph convolution ------------------------------------------------------------
 * reference: net/utils/tgcn.py:76-89 (conv1x1 to K*Cout, einsum 'nkctv,kvw->nctw') and
 * net/utils/inceptionv2_gcn.py:64-89 (three einsums with A, A2, A3 == one with A_eff).
 *   z[(f,w)][c] = sum_k sum_ci Wc[k*Cin+ci][c] * (sum_v A_eff[k][v][w] x[(f,v)][ci])
 *                 + biasterm[w][c]
 * Wc[K*Cin][Cout] is the conv weight regrouped (Wc[k*Cin+ci][c] = weight[k*Cout+c][ci]);
 * biasterm[V][Cout] = sum_k bias[k*Cout+c] * colsum(A_eff[k])[w].  `frames` = N*M*T.
 * stat_sum/stat_sumsq (double[Cout], caller-zeroed, may be NULL) receive the BatchNorm
 * statistics of z.
 * Temporal map: with t_out > 0, output frame f = n*t_out + to reads input frame
 * n*t_in + to*t_stride + t_offset (zeros when that frame is outside [0, t_in): the temporal
 * padding), and gin / add_in of the backward use the same mapping.  With K = 1 and an identity
 * adjacency this is the strided 1x1 convolution of the residual branch
 * (net/st_gcnold.py:186-193) and, with t_offset = tap - pad, one tap of the (kt x 1) temporal
 * convolution (net/st_gcnold.py:165-171): the taps are summed through add_rows (partial sums of
 * the previous taps, may alias z; NULL for the first).  biasterm may be NULL.  t_out = 0 means
 * "same frames".                                                                             */
int istgcn_gcn_fwd(const float* x, const float* Wc, const float* biasterm, const float* vals,
                   const int* dst_ptr, const int* dst_src, const int* dst_id, int nnz,
                   const float* add_rows, float* z, double* stat_sum, double* stat_sumsq,
                   int frames, int V, int K, int Cin, int Cout,
                   int t_in, int t_out, int t_stride, int t_offset, int math, istgcn_stream_t s);
/* input gradient + adjacency gradient.  dz is formed on the fly from the BatchNorm-backward
 * coefficients: dz = bn_p*((g - bn_m1) - bn_c*(z - bn_mu)) per channel (bn_p = NULL: dz = g).
 *   gin[(f,v)][ci] = sum_k sum_w A_eff[k][v][w] (dz Wc_k^T)[(f,w)][ci]  (+ add_in if not NULL)
 *   dvals[id] += sum_{f,ci} x[(f,v)][ci] * (dz Wc_k^T)[(f,w)][ci]       (caller-zeroed)     */
int istgcn_gcn_bwd_x(const float* g, const float* z, const float* bn_p, const float* bn_m1,
                     const float* bn_c, const float* bn_mu,
                     const float* x, const float* Wc, const float* vals,
                     const int* src_ptr, const int* src_kw, const int* src_id, int nnz,
                     const float* add_in, float* gin, float* dvals,
                     int frames, int V, int K, int Cin, int Cout,
                     int t_in, int t_out, int t_stride, int t_offset, int math, istgcn_stream_t s);
/* weight gradient: dWc[K*Cin][Cout] += X'^T dz, dbiasterm[V][Cout] += sum_f dz (both
 * caller-zeroed fp32).                                                                     */
int istgcn_gcn_bwd_w(const float* g, const float* z, const float* bn_p, const float* bn_m1,
                     const float* bn_c, const float* bn_mu,
                     const float* x, const float* vals,
                     const int* dst_ptr, const int* dst_src, const int* dst_id, int nnz,
                     float* dWc, float* dbiasterm,
                     int frames, int V, int K, int Cin, int Cout,
                     int t_in, int t_out, int t_stride, int t_offset, int math, istgcn_stream_t s);

/* tcgen05 / TMA / TMEM engine of the same graph convolution (forward form, or input-gradient
 * form when the lists are grouped by (k, source joint) and the weight is Wc):
 *   out[(f,w)][n] = sum_{k,ci} (sum_v A[k][v][w] in'[(f,v)][ci]) * w_rows[k*Cout + n][ci]
 *                   + sum_k colsum[k][w] * bias_k[k][n] + add_rows[(f,w)][n]
 * with in' = in, or in' = bn_p*((in - bn_m1) - bn_c*(in2 - bn_mu)) when bn_p != NULL.
 * w_rows is [K*Cout][CinPad] (CinPad = Cin rounded up to 32, padding columns zero), 16-byte
 * aligned; lptr[K*V+1] / lsrc / lid group the non-zeros by (k, destination joint).
 * in_out (may be NULL) receives a copy of in' ([rows][Cin]): the input-gradient call uses it to
 * materialise dz for the two kernels below.  bias_k [K][Cout] is the conv bias and colsum [K][V]
 * the column sums of A[k] (the bias passes through the aggregation; both NULL: no bias).
 * Whole output tiles leave through TMA tile stores; when add_rows == out the kernel accumulates
 * in place with TMA reduce-add.  map_side selects where the temporal stride of the
 * residual branch applies (0: none, 1: input rows are read from frame n*t_in + to*t_stride,
 * 2: output / add_rows rows are written there).  TF32 inputs, fp32 accumulation in TMEM.
 * Dispatch: plain calls (bn_p == NULL, in_out == NULL, map_side == 0, Cin and Cout multiples of
 * 32, Cout <= 256, add_rows NULL or == out) run the second-generation kernel csrc/gcn_tc2.cu --
 * aggregation AND channel mix on the tensor core, adjacency and aggregated operand in tensor
 * memory, input frames by TMA; everything else the first-generation kernel csrc/gcn_tc.cu.
 * The environment variable ISTGCN_GCN_TC_V1 forces the first generation everywhere.          */
int istgcn_gcn_tc(const float* in, const float* in2, const float* bn_p, const float* bn_m1,
                  const float* bn_c, const float* bn_mu, const float* w_rows, const float* vals,
                  const int* lptr, const int* lsrc, const int* lid, int nnz,
                  const float* bias_k, const float* colsum, const float* add_rows, float* out,
                  float* in_out,
                  double* stat_sum, double* stat_sumsq, int frames, int V, int K, int Cin,
                  int CinPad, int Cout, int t_in, int t_out, int t_stride, int t_offset, int map_side,
                  istgcn_stream_t s);

/* weight AND adjacency gradient from one pass over (dz, x) (csrc/gcn_pair_tc.cu; replaces the two calls
 * above in the fast mode): P[(v,w)][ci][c] = sum_f x[(f,v)][ci] dz[(f,w)][c] for every joint pair of the
 * non-zero pattern on the tensor core (both operands by TMA as they lie in HBM), then
 *   dWc[k*Cin+ci][c] += sum_(v,w) vals[(k,v,w)] P[(v,w)][ci][c],   dvals[(k,v,w)] += <Wc[k], P[(v,w)]>.
 * A work item is a BLOCK of the joint-pair pattern: ns source joints x nd destination joints x ncw
 * output channels of each (the kernel is bound by the L2 -> shared-memory operand stream: rows + columns
 * per K-tile for rows x columns MACs, so blocks should be as square as tensor memory allows):
 * items[nitems][8] = {d0, nd, s0, ns, first output column, ncw, owned-cell mask (bit vi*nd + j), 0} with
 * d0 / s0 indexing joints[]; ceil(ns*Cin/128) * nd*ncw <= 512, nd*ncw <= 256, ncw % 32 == 0, ns*nd <= 32;
 * every pattern pair x column range must be owned by exactly one item.  pair_of[V*V]: (v*V + w) -> pair
 * index or -1.  ctas[nctas][4] = {item, first 64-frame K-tile, K-tile stride, 0}: one thread block each
 * (the caller gives an item thread blocks in proportion to its cost so that all blocks finish together);
 * entry_pair[nnz]: pair index of every canonical entry; k_ptr[K+1]: entries of partition k (canonical
 * order is sorted by k).  P_ws [npairs][Cin][Cout] caller-zeroed scratch; Cin, Cout multiples of 32.   */
int istgcn_gcn_pair_grads(const float* dz, const float* x, const float* vals, const float* Wc,
                          const int* items, int nitems, const int* ctas, int nctas, const int* joints,
                          const int* pair_of, int npairs, const int* entry_pair, const int* k_ptr, int nnz,
                          float* P_ws, float* dWc, float* dvals, int frames, int V, int K, int Cin,
                          int Cout, istgcn_stream_t s);

/* ---- first block (in_channels <= 4: net/st_gcnold.py:46 `st_gcn(in_channels, 64, ...)`, the
 * graph convolution of tgcn.py:76-89 on a 3-channel input) on CUDA cores in full fp32
 * (csrc/gcn_small.cu): 12 B in / 256 B out per row, so no tensor-core slice padding.
 *   gcn_small_fwd: out[(f,w)][n] = sum_{k,c} X'_k[(f,w)][c] * Wc[k*Cin+c][n] + biasterm[w][n]
 *                  (+ BatchNorm sums).  Wc [K*Cin][Cout], Cout = 64 or 128; lists by (k, dest w).
 *   gcn_small_bwd: the whole backward behind dz = p*((g1 - m1) - c*(z - mu)) in ONE kernel:
 *                  dx [frames*V][Cin] (written); dvals, dWc [K*Cin][64], dbt [V][64] accumulated
 *                  (caller-zeroed; dvals / dbt may be NULL).  (tptr, tsrc, tid) = lists grouped by
 *                  (k, source v) with tsrc = destination w.  Cout = 64.                        */
int istgcn_gcn_small_fwd(const float* x, const float* Wc, const float* biasterm, const float* vals,
                         const int* lptr, const int* lsrc, const int* lid, int nnz, float* out,
                         double* stat_sum, double* stat_sumsq, float* xagg_out, float* zsum_out,
                         int frames, int V, int K, int Cin, int Cout, istgcn_stream_t s);
/* The same backward with its heavy part on the tensor core (fast mode).  The forward optionally stores
 * xagg_out [frames*V][16] = the aggregated input X'[(f,w)][k*4 + c] (TF32-rounded) and zsum_out [V][Cout]
 * += sum_f out[(f,w)][n] (Cout = 64).  Backward = istgcn_tcn2_bwd_up(go = g1, u = z, BN1-backward
 * coefficients, h2 = X', Wu = Wc padded to [16][64] rows k*4 + c, C = 64, bp = 16): its dh2 is
 * G[(f,w)][k*4+c] = sum_n dz[(f,w)][n] Wc[k*Cin+c][n], its dWu the weight gradient; then
 * istgcn_joint_colsum(g1) and istgcn_gcn_small_bwd_post: dx (written), dvals += ..., and
 * dbt[w][n] += p[n]((sg1[w][n] - F m1[n]) - c[n](sz[w][n] - F mu[n])) -- dz is affine in (g1, z).   */
int istgcn_joint_colsum(const float* a, float* sums, int frames, int V, int C, istgcn_stream_t s);
int istgcn_gcn_small_bwd_post(const float* G, const float* x, const float* vals, const int* lptr,
                              const int* lsrc, const int* lid, const int* tptr, const int* tsrc,
                              const int* tid, int nnz, float* dx, float* dvals, const float* sg1,
                              const float* sz, const float* bn_p, const float* bn_m1, const float* bn_c,
                              const float* bn_mu, float* dbt, int frames, int V, int K, int Cin, int Cout,
                              istgcn_stream_t s);
int istgcn_gcn_small_bwd(const float* g1, const float* z, const float* bn_p, const float* bn_m1,
                         const float* bn_c, const float* bn_mu, const float* x, const float* Wc,
                         const float* vals, const int* lptr, const int* lsrc, const int* lid,
                         const int* tptr, const int* tsrc, const int* tid, int nnz, float* dx,
                         float* dvals, float* dWc, float* dbt, int frames, int V, int K, int Cin,
                         int Cout, istgcn_stream_t s);

/* adjacency gradient on the tcgen05 engine (both operands fed by TMA):
 *   dvals[id] += sum_{f,ci} x[(f,v)][ci] * (dz Wc_k^T)[(f,w)][ci]   over the non-zeros (k,v,w)
 * dz [frames*V][Cout] (gradient w.r.t. the graph-conv output), Wc [K*Cin][Cout]; lists grouped
 * by (k, destination w) with lsrc = source joint v, lid = canonical id.  Cout % 32 == 0.
 * Cin % 32 == 0 and Cout <= 128: csrc/gcn_tc_da2.cu (second MMA with the A operand from tensor
 * memory, accumulators over all frame tiles in TMEM, one read-out); else csrc/gcn_tc_da.cu.  */
int istgcn_gcn_tc_dvals(const float* dz, const float* x, const float* Wc, const int* lptr,
                        const int* lsrc, const int* lid, int nnz, float* dvals, int frames, int V,
                        int K, int Cin, int Cout, istgcn_stream_t s);

/* weight gradient on the tcgen05 engine: dWc[K*Cin][Cout] += X'^T dz with both operands read
 * MN-major (contraction over the rows of the frame tile), accumulators resident in tensor memory
 * for the whole kernel; dbiasterm[V][Cout] += sum over frames of dz (may be NULL).  Outputs are
 * caller-zeroed; lists grouped by (k, destination w); Cout % 32 == 0.
 * Plain frame maps (t_out == 0), Cin % 32 == 0, Cout <= 256 and at most eight CTA groups:
 * csrc/gcn_tc_dw2.cu (aggregation on the tensor core with the adjacency in tensor memory, converter
 * warps TMEM -> MN-major operand atoms); else csrc/gcn_tc_dw.cu (aggregation on CUDA cores).   */
int istgcn_gcn_tc_dw(const float* dz, const float* x, const float* vals, const int* lptr,
                     const int* lsrc, const int* lid, int nnz, float* dWc, float* dbiasterm,
                     int frames, int V, int K, int Cin, int Cout,
                     int t_in, int t_out, int t_stride, int t_offset, istgcn_stream_t s);

/* ---- Inception TCN with 1x1 bottlenecks (net/st_gcn_mstcn_1x1.py:250-266) ---------------
 *   a  = relu((z - mean1)*scale1 + beta1)              (tcn_start: BN + ReLU)
 *   h1 = a Wd + bd                                     (conv_1x1_start, C -> b)
 *   h2[to] = sum_tap Weff[tap] h1[to*stride + tap - 7] + beff   (tcn_1/2/3 merged: 15 taps,
 *            Weff = imp0*W3x1 (+6) + imp1*W9x1 (+3) + imp2*W15x1, zero padding on h1)
 *   u  = h2 Wu + bu                                    (conv_1x1_end, b -> C)
 * Wd[C][bp], Weff[15][bp(in)][bp(out)], Wu[bp][C]; bp = b rounded up to a multiple of 8
 * (padding columns/rows zero).  h1[(n,t,v)][bp] and h2[(n,to,v)][bp] are saved for backward.
 * stats of u optional.                                                                     */
int istgcn_tcn_fwd(const float* z, const float* mean1, const float* scale1, const float* beta1,
                   const float* Wd, const float* bd, const float* Weff, const float* beff, const float* Wu,
                   const float* bu, float* h1, float* h2, float* u, double* stat_sum,
                   double* stat_sumsq, int NM, int T, int V, int C, int bp, int stride,
                   int math, istgcn_stream_t s);
/* backward.  du = p2*((gy - m12) - c2*(u - mean2)), gy = go * keep(dropout); writes g1 = d(loss)/d(bn1
 * output) masked by the ReLU, accumulates sum g1 and sum g1*zhat (double[C], caller-zeroed)
 * and all weight gradients (caller-zeroed fp32).  dh2_ws[(n,to,v)][bp] and dh1_ws[(n,t,v)][bp]
 * are caller-provided scratch (fully overwritten).                                         */
int istgcn_tcn_bwd(const float* go, const float* u, const float* p2, const float* m12,
                   const float* c2, const float* mean2, const float* z, const float* scale1,
                   const float* beta1, const float* mean1, const float* rstd1,
                   const float* h1, const float* h2,
                   const float* Wd, const float* Weff, const float* Wu,
                   float* dh2_ws, float* dh1_ws, float* g1, double* sg1, double* sg1x, float* dWd, float* dbd, float* dWeff,
                   float* dbeff, float* dWu, float* dbu, int NM, int T, int V, int C, int bp,
                   int stride, float drop_p, uint64_t drop_seed,
                   const unsigned long long* drop_step, int math, istgcn_stream_t s);

/* The same chain in the 'tf32' math mode as six streaming kernels, one entry point per kernel
 * (csrc/tcn2.cu): activation rows go from global memory straight into mma.sync fragments, every big
 * tensor is touched once, only the bp-wide intermediates make round trips.  C in {64, 128, 256},
 * bp in {8, 16}; rows_in = NM*T*V, rows_out = NM*Tout*V, Tout = (T-1)/stride + 1.  Same operands,
 * layouts and accumulation conventions (caller-zeroed fp32 / double accumulators) as above.
 *   tcn2_down      h1 = relu((z - mean1)*scale1 + beta1) Wd + bd          [rows_in][bp]
 *   tcn2_conv      h2[to] = sum_tap Weff[tap] h1[to*stride + tap - 7] + beff   [rows_out][bp]
 *   tcn2_up        u = h2 Wu + bu; stat_sum / stat_sumsq (double[C], may both be NULL)
 *   tcn2_bwd_up    du = p2*((gy - m12) - c2*(u - mean2)); dh2 = du Wu^T (written);
 *                  dWu += h2^T du, dbu += sum du, dbeff += sum dh2
 *   tcn2_bwd_conv  dh1 = transposed temporal conv of dh2 (written); dWeff += ..., dbd += sum dh1
 *                  (two independent kernels: dh1 == NULL -> dWeff only, dWeff == NULL -> dh1 + dbd only)
 *   tcn2_bwd_down  g1 = (dh1 Wd^T) where relu(BN1(z)) > 0 (written); sg1 += sum g1,
 *                  sg1x += sum g1*zhat (double[C]); dWd += a^T dh1                              */
int istgcn_tcn2_down(const float* z, const float* mean1, const float* scale1, const float* beta1,
                     const float* Wd, const float* bd, float* h1, long long rows_in, int C, int bp,
                     istgcn_stream_t s);
int istgcn_tcn2_conv(const float* h1, const float* Weff, const float* beff, float* h2, int NM, int T,
                     int V, int bp, int stride, istgcn_stream_t s);
int istgcn_tcn2_up(const float* h2, const float* Wu, const float* bu, float* u, double* stat_sum,
                   double* stat_sumsq, long long rows_out, int C, int bp, istgcn_stream_t s);
int istgcn_tcn2_bwd_up(const float* go, const float* u, const float* p2, const float* m12,
                       const float* c2, const float* mean2, const float* h2, const float* Wu,
                       float* dh2, float* dWu, float* dbu, float* dbeff, long long rows_out, int C,
                       int bp, float drop_p, uint64_t drop_seed, const unsigned long long* drop_step,
                       istgcn_stream_t s);
int istgcn_tcn2_bwd_conv(const float* dh2, const float* h1, const float* Weff, float* dh1,
                         float* dWeff, float* dbd, int NM, int T, int V, int bp, int stride,
                         istgcn_stream_t s);
int istgcn_tcn2_bwd_down(const float* dh1, const float* z, const float* mean1, const float* scale1,
                         const float* beta1, const float* rstd1, const float* Wd, float* g1,
                         float* dWd, double* sg1, double* sg1x, long long rows_in, int C, int bp,
                         istgcn_stream_t s);

/* ---- the same consumers with the BatchNorm bookkeeping folded in (training mode).  The one-block
 * istgcn_bn_finalize / istgcn_bn_bwd_coeffs launches in front of them cost ~6 us each with their gaps
 * (~75 per training step); here every thread block of the consumer derives the per-channel coefficients
 * it needs from the raw double sums, and the first block also stores them (the arrays named as outputs
 * below keep their meaning for the backward pass and for later consumers), updates the running
 * statistics (bn_finalize's formula; may be NULL) and writes dgamma / dbeta (may be NULL).
 * rstd = 1/sqrt(var + eps) by rsqrtf + one Newton step (float-exact); count = rows behind the sums.  */
int istgcn_tcn2_down_bn(const float* z, const double* stat_sum, const double* stat_sumsq, double count,
                        const float* gamma1, const float* beta1, float* running_mean,
                        float* running_var, float momentum, float eps, float* mean1, float* scale1,
                        float* rstd1, const float* Wd, const float* bd, float* h1, long long rows_in,
                        int C, int bp, istgcn_stream_t s);
int istgcn_tcn2_bwd_up_bn(const float* go, const float* u, const double* sg, const double* sgx,
                          double count, const float* gamma2, const float* rstd2, float* p2, float* m12,
                          float* c2, float* dgamma2, float* dbeta2, const float* mean2, const float* h2,
                          const float* Wu, float* dh2, float* dWu, float* dbu, float* dbeff,
                          long long rows_out, int C, int bp, float drop_p, uint64_t drop_seed,
                          const unsigned long long* drop_step, istgcn_stream_t s);
int istgcn_bn_back_colsum_bn(const float* g, const float* z, const double* sg, const double* sgx,
                             double count, const float* gamma, const float* rstd, float* p, float* m1,
                             float* c, float* dgamma, float* dbeta, const float* mean, float* dz,
                             float* colsum, int frames, int V, int C, istgcn_stream_t s);
int istgcn_block_tail_fwd_bn(const float* u, const double* sum2, const double* sumsq2, double count,
                             const float* gamma2, const float* beta2, float* rmean2, float* rvar2,
                             float momentum2, float eps2, float* mean2, float* scale2, float* rstd2,
                             const float* res, const double* sum_r, const double* sumsq_r,
                             const float* gamma_r, const float* beta_r, float* rmean_r, float* rvar_r,
                             float momentum_r, float eps_r, float* mean_r, float* scale_r, float* rstd_r,
                             float* out, long long rows, int C, float drop_p, uint64_t drop_seed,
                             const unsigned long long* drop_step, istgcn_stream_t s);

/* ---- full-width temporal convolution (net/st_gcnold.py:160-174: BN -> ReLU -> Conv2d(C, C,
 * (kt,1), (stride,1), (pad,0)) -> BN -> Dropout; net/st_gcn_mstcn.py: three such convs of 3/9/15
 * taps scaled by mstcn_importance = one 15-tap conv).  The convolution itself is the sum over
 * taps of the shifted, strided 1x1 engine above (istgcn_gcn_tc / istgcn_gcn_fwd with K = 1, an
 * identity adjacency, t_offset = tap - pad and add_rows = the partial sums); these three
 * element-wise stages sit around it.  rows = N*M*T*V, channels-last.
 *   bn_relu_apply: a = max((z - mean)*scale + beta, 0)
 *   bn_back_apply: du = p*((g' - m1) - c*(u - mean)),  g' = go with the dropout mask of
 *                  (drop_p, drop_seed, drop_step) applied (the tail's convention)
 *   relu_bn_bwd:   g1 = da where a > 0 else 0;  sg[c] += sum g1,  sgx[c] += sum g1*(z-mean1)*rstd1
 *                  (caller-zeroed doubles, the inputs of istgcn_bn_bwd_coeffs)                 */
/* The same convolution as ONE tcgen05 implicit GEMM (csrc/tconv_tc.cu): both operands by TMA, tap
 * `tap` = the activation box shifted by dir*(tap - pad) frames, out-of-range frames read as zeros.
 *   out[(n,to,v)][co] = sum_tap sum_ci in[(n, to*stride + dir*(tap-pad), v)][ci] * w_rows[tap*Cout+co][ci]
 *                       + bias[co]
 * dir = +1: forward; dir = -1 (stride 1 only) with the transposed weights: input gradient.
 * Needs Cin, Cout multiples of 32, Tout a multiple of floor(128/V); stat_* as in istgcn_gcn_tc. */
int istgcn_tconv_tc(const float* in, const float* w_rows, const float* bias, float* out,
                    double* stat_sum, double* stat_sumsq, int NM, int T, int Tout, int V, int Cin,
                    int Cout, int kt, int stride, int dir, istgcn_stream_t s);
/* Weight gradient of the same convolution on the tcgen05 engine (csrc/tconv_dw_tc.cu):
 *   dW[tap*Cin + ci][co] += sum_{n,to,v} a[(n, to*stride + tap - pad, v)][ci] * du[(n,to,v)][co]
 *   dbias_vc[v][co]      += sum_{n,to} du[(n,to,v)][co]                       (may be NULL)
 * a [NM][T][V][Cin], du [NM][Tout][V][Cout]; outputs caller-zeroed.  Cin, Cout multiples of 32,
 * Cout <= 128 or a multiple of 128.                                                           */
int istgcn_tconv_dw_tc(const float* a, const float* du, float* dW, float* dbias_vc, int NM, int T,
                       int Tout, int V, int Cin, int Cout, int kt, int stride, istgcn_stream_t s);
int istgcn_bn_relu_apply(const float* z, const float* mean, const float* scale, const float* beta,
                         float* a, long long rows, int C, istgcn_stream_t s);
int istgcn_bn_back_apply(const float* go, const float* u, const float* p, const float* m1,
                         const float* c, const float* mean, float* du, long long rows, int C,
                         float drop_p, uint64_t drop_seed, const unsigned long long* drop_step,
                         istgcn_stream_t s);
/* dz = p*((g - m1) - c*(z - mean)) per channel, written to dz [frames*V][C], and in the same pass
 * colsum[V][C] += sum over frames of dz (caller-zeroed): the input of the tensor-core graph-conv
 * gradient kernels and the gradient of the graph convolution's bias term (tgcn.py:79-86).    */
int istgcn_bn_back_colsum(const float* g, const float* z, const float* p, const float* m1,
                          const float* c, const float* mean, float* dz, float* colsum, int frames,
                          int V, int C, istgcn_stream_t s);
int istgcn_relu_bn_bwd(const float* da, const float* a, const float* z, const float* mean1,
                       const float* rstd1, float* g1, double* sg, double* sgx, long long rows, int C,
                       istgcn_stream_t s);

/* ---- block tail: BN2 -> dropout -> + residual -> ReLU (st_gcn_mstcn_1x1.py:262-266) -----
 * out = relu(BN2(u)*keep/(1-p) + res) where res = NULL (0), the block input (identity,
 * scale_r = NULL) or BNr(res) (strided conv + BN).  Rows = NM*T*V, C channels.             */
int istgcn_block_tail_fwd(const float* u, const float* mean2, const float* scale2,
                          const float* beta2, const float* res, const float* mean_r,
                          const float* scale_r, const float* beta_r, float* out, long long rows,
                          int C, float drop_p, uint64_t drop_seed,
                          const unsigned long long* drop_step, istgcn_stream_t s);
/* backward reduction pass: go = gout * (out > 0) written to `go` (may alias gout); BN2 sums
 * sum gy, sum gy*uhat with gy = go*keep/(1-p); if rres != NULL also sum go, sum go*rhat.    */
int istgcn_block_tail_bwd(const float* gout, const float* out, const float* u,
                          const float* mean2, const float* rstd2, const float* rres,
                          const float* mean_r, const float* rstd_r, float* go, double* sg2,
                          double* sg2x, double* sgr, double* sgrx, long long rows, int C,
                          float drop_p, uint64_t drop_seed, const unsigned long long* drop_step,
                          istgcn_stream_t s);

/* keep-mask of the counter-based dropout the two functions above use (element index =
 * row*C + c); exported so that parity tests can inject the same mask into the oracle.
 * drop_step (may be NULL) points at a device-resident step counter that is mixed into the
 * seed, so that a captured CUDA graph draws a fresh mask on every replay.                   */
int istgcn_dropout_mask(unsigned char* mask, long long n, float p, uint64_t seed,
                        const unsigned long long* drop_step, istgcn_stream_t s);

/* ---- head: global average pool over (T, V), mean over M, 1x1 conv (st_gcnold.py:89-94) --*/
int istgcn_pool_fwd(const float* x, float* pooled, int N, int M, int TV, int C,
                    istgcn_stream_t s);
int istgcn_pool_bwd(const float* gpooled, float* gx, int N, int M, int TV, int C,
                    istgcn_stream_t s);

/* ---- optimiser step (processor/recognition.py:152-159: optim.SGD(momentum=0.9, nesterov,
 * weight_decay); :287-289 optimizer.step()) over flat fp32 buffers of n elements:
 *   g' = g*grad_scale + weight_decay*p;  buf = momentum*buf + g';
 *   p -= lr * (nesterov ? g' + momentum*buf : buf)
 * lr is read from DEVICE memory (the step schedule of recognition.py:168-176 changes it without
 * re-capturing a CUDA graph); grad_scale = 1/world folds the data-parallel average in.
 * buf zero-initialised by the caller before the first step.  16-byte aligned buffers.        */
int istgcn_sgd_step(float* p, const float* g, float* buf, long long n, const float* lr,
                    float momentum, float weight_decay, int nesterov, float grad_scale,
                    istgcn_stream_t s);

/* ---- parameter regrouping of one IST-GCN block, one kernel each way ------------------------
 * reference-layout parameters (tgcn.py:51-58 / inceptionv2_gcn.py:22-29 conv weight [K*Cout][Cin] + bias,
 * A / A2 / A3 buffers and edge_importance* [K][V][V] (st_gcn_msgcn.py:112-117; A2 = A3 = NULL: single
 * adjacency; imp_i = NULL: importance 1), conv_1x1_start [b][C], tcn_1/2/3 [b][b][3|9|15] with
 * mstcn_importance[3], conv_1x1_end [C][b] (st_gcn_mstcn_1x1.py:190-224), residual conv [Cout][Cin])
 * -> the kernel operands documented above: vals[nnz], colsum[K][V], Wc[K*Cin][Cout], biasterm[V][Cout],
 * Wd[C][bp], bd[bp], Weff[15][bp][bp], beff[bp], Wu[bp][C], Wr[Cin][Cout], btr[V][Cout].
 * flat_idx[nnz] (index of every pattern entry in the [K][V][V] stack), dst_ptr / dst_id as above.
 * bwd: gradients of those operands -> gradients of the reference-layout parameters (written, not
 * accumulated); inv_idx[K*V*V] = id of the pattern entry or -1, id_kw[nnz] = k*V + w of entry id.    */
int istgcn_block_prep_fwd(const float* W, const float* bias, const float* A1, const float* imp1,
                          const float* A2, const float* imp2, const float* A3, const float* imp3,
                          const long long* flat_idx, const int* dst_ptr, const int* dst_id, int nnz,
                          float* vals, float* colsum, float* Wc, float* biasterm, const float* Ws,
                          const float* bs, const float* W1, const float* b1, const float* W2, const float* b2,
                          const float* W3, const float* b3, const float* We, const float* m_imp, float* Wd,
                          float* bd, float* Weff, float* beff, float* Wu, const float* Wres, const float* bres,
                          float* Wr, float* btr, int K, int V, int Cin, int Cout, int b, int bp,
                          istgcn_stream_t s);
int istgcn_block_prep_bwd(const float* dvals, const float* dWc, const float* dbt, const float* dWd,
                          const float* dbd, const float* dWeff, const float* dbeff, const float* dWu,
                          const float* dWr, const float* dbtr, const float* bias, const float* colsum,
                          const float* A1, const float* A2, const float* A3, const int* inv_idx,
                          const int* id_kw, const float* W1, const float* b1, const float* W2, const float* b2,
                          const float* W3, const float* b3, const float* m_imp, float* dW, float* dbias,
                          float* dimp1, float* dimp2, float* dimp3, float* dWs, float* dbs, float* dW1,
                          float* db1, float* dW2, float* db2, float* dW3, float* db3, float* dWe, float* dm_imp,
                          float* dWres, float* dbres, int K, int V, int Cin, int Cout, int b, int bp,
                          istgcn_stream_t s);

/* ---- input pipeline on the device (feeder/feeder.py:70-85, feeder/tools.py:32-102) ---------
 * in (N, C, Tin, V, M) -> out (N, C, Tout, V, M):  out[:, :, t] = in[:, :, t + shift[n]] (zeros outside
 * [0, Tin): random_choose's crop, auto_pading's offset; shift may be NULL) and then random_move on
 * channels 0, 1 of every output frame: (x, y) <- (m0*x - m1*y + m2, m1*x + m0*y + m3) with
 * move[n][t] = {cos(a)*s, sin(a)*s, t_x, t_y} drawn on the host (may be NULL).                */
int istgcn_feeder_augment(const float* in, const int* shift, const float* move, float* out, int N,
                          int C, int Tin, int Tout, int V, int M, istgcn_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* ISTGCN_B200_H */
