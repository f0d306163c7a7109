/* tamgcn.h — C-ABI of the B200-native CTR-GCN / ST-GCN hot path (libtamgcn.so).
 *
 * The reference (Tamnemng/TAM-GCN) is pure Python/PyTorch and has no FFI of its own; the
 * interface each entry point replaces is therefore the ATen call sequence inside the reference
 * module named in its comment (paths relative to the reference checkout).  The Python host side
 * (tam_gcn_b200/_C.py, ctypes) is the only caller; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - plain pointers + sizes, no torch types.  Every pointer is DEVICE memory owned by the caller
 *    (the PyTorch caching allocator); the library never allocates, frees or keeps device memory.
 *  - activations are contiguous planes (T, V) inside (N, C, T, V) tensors, V fastest.  A tensor
 *    argument is (pointer to channel 0 of the slice, sample stride in ELEMENTS); channel stride is
 *    always T*V.  This lets callers pass channel slices of wider tensors without copies.
 *  - `dtype` selects the activation storage type: TAMGCN_F32 or TAMGCN_BF16.  Parameters, BN
 *    coefficients and all gradients of parameters are fp32; BN statistics are fp64 accumulators.
 *    All arithmetic accumulates in fp32.
 *  - everything is enqueued on `stream` (a cudaStream_t), never synchronises, and is CUDA-graph
 *    capturable.  Accumulator outputs (documented "+=") must be zeroed by the caller.
 *  - return 0 on success, <0 on error; tamgcn_last_error() returns a thread-local message.
 */
#ifndef TAMGCN_H
#define TAMGCN_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tamgcn_stream; /* cudaStream_t */

enum { TAMGCN_F32 = 0, TAMGCN_BF16 = 1 };
enum { TAMGCN_RES_NONE = 0, TAMGCN_RES_IDENTITY = 1, TAMGCN_RES_AFFINE = 2 };

int tamgcn_version(void);
const char* tamgcn_last_error(void);
/* Share (percent, 1..100; default 100; process-wide) of the SMs the persistent weight-gradient kernels may occupy;
 * returns the previous value.  A step engine that launches weight gradients on a side stream sets ~50 so that they run
 * next to — not in turns with — the data-gradient chain of the main stream.  tamgcn_set_main_sm_share does the same for
 * the persistent convolution forward / data-gradient kernels (so that, while the side stream holds part of the chip, a
 * main-stream kernel is not split into a running and a waiting wave of CTAs). */
int tamgcn_set_wgrad_sm_share(int percent);
int tamgcn_set_main_sm_share(int percent);
/* number of kernel launches issued by this library in the calling process (bench.py gpu_launches) */
int64_t tamgcn_launch_count(void);

/* A lazily transformed activation operand: value(n,ch,t,v) = f(a[ch]*P + b[ch]*Q + c[ch]),
 * f = max(0,.) when relu != 0.  NULL a -> 1, NULL b (or NULL q) -> no Q term, NULL c -> 0.
 * It is how BatchNorm-apply(+ReLU) (forward), the res - y difference of unit_gcn and the
 * BatchNorm backward formula dY = A*dYhat + B*Y + C are fused into the consumer kernels. */
typedef struct tamgcn_operand {
    const void* p;
    const void* q;
    const float* a;
    const float* b;
    const float* c;
    int64_t p_nstride;
    int64_t q_nstride;
    int32_t relu;
    int32_t reserved;
} tamgcn_operand;

/* (k x 1) convolution geometry: nn.Conv2d(Cin, Cout, (k,1), stride (s,1), padding (pad,0), dilation (dil,1)) */
typedef struct tamgcn_conv_geom {
    int32_t N, Cin, Cout, T, To, V, k, stride, dil, pad;
} tamgcn_conv_geom;

/* ---- (k x 1) convolutions: models/ctrgcn.py:56-62,95-99,114,122,161-164,183-184,212,221;
 *      models/stgcn.py:47-55,79,89 (nn.Conv2d forward / convolution_backward) ---------------------- */
/* Tensor-core (tcgen05) path of the bf16 convolutions: the fp32 weights are first packed into bf16 tiles laid
 * out exactly as the MMA reads them from shared memory (K-major, 128-byte swizzle), one buffer for the forward
 * and one for the data-gradient GEMM.  `tamgcn_conv_pack_bytes` gives the buffer sizes; pass the buffers as
 * `wpack` to tamgcn_conv_fwd / tamgcn_conv_dgrad.  wpack == NULL (or dtype F32) selects the exact-fp32 SIMT path. */
int64_t tamgcn_conv_pack_bytes(int Cout, int Cin, int k, int dgrad);
int tamgcn_conv_pack_weights(const float* W, int Cout, int Cin, int k, void* wpack_fwd, void* wpack_dgrad,
                             tamgcn_stream stream);
/* The same for many weight matrices in ONE launch (a step engine re-packs every convolution of the model right after
 * the optimiser instead of once per layer call).  `table`: DEVICE array of njobs rows of 8 int64:
 * {W pointer, wpack_fwd pointer or 0, wpack_dgrad pointer or 0, Cout, Cin, k, 0, 0}. */
int tamgcn_conv_pack_weights_batched(const int64_t* table, int njobs, tamgcn_stream stream);
/* 1 if the bf16 forward (dgrad = 0) / data-gradient (dgrad = 1) kernel of this shape reads a packed buffer, 0 if it
 * reads the fp32 weights directly (the small-channel temporal convolutions, Cin = Cout in {16, 32, 64}, k >= 2,
 * run on warp-level MMAs with the operand staged once per time block — csrc/tconv_mma.cu). */
int tamgcn_conv_needs_pack(int Cin, int Cout, int k, int stride, int V, int dgrad);
/* y[n,co,to,v] = bias[co] + sum_{ci,j} W[co,ci,j] * X(n,ci,to*s + j*dil - pad, v)   (zero padding)
 * optional epilogue: per-channel sum / sum of squares of y over (n,to,v) for channels >= stat_c0
 * (stat arrays indexed co - stat_c0, "+=").  W is (Cout,Cin,k) fp32. */
int tamgcn_conv_fwd(const tamgcn_conv_geom* g, int dtype, const tamgcn_operand* x, const float* W,
                    const void* wpack, const float* bias, void* y, int64_t y_nstride, double* stat_sum,
                    double* stat_sumsq, int stat_c0, tamgcn_stream stream);
/* dX = conv_transpose(dY) [+ addend] [+ bcast[n,ci,v]*bcast_scale];  if mask != NULL the result is
 * multiplied by [mask.a*mask.P + mask.c > 0] (ReLU backward of the fused forward prologue) and
 * s1[ci] += sum dX, s2[ci] += sum dX*mask.P (BatchNorm-backward reductions, may be NULL). */
int tamgcn_conv_dgrad(const tamgcn_conv_geom* g, int dtype, const tamgcn_operand* dy, const float* W,
                      const void* wpack, void* dx, int64_t dx_nstride, const void* addend, int64_t addend_nstride, const float* bcast,
                      float bcast_scale, const tamgcn_operand* mask, double* s1, double* s2,
                      tamgcn_stream stream);
/* dW[co,ci,j] += sum_{n,to,v} dY(n,co,to,v) * X(n,ci,to*s+j*dil-pad,v);  dbias[co] += sum dY (NULL to skip) */
int tamgcn_conv_wgrad(const tamgcn_conv_geom* g, int dtype, const tamgcn_operand* dy, const tamgcn_operand* x,
                      float* dW, float* dbias, tamgcn_stream stream);

/* ---- CTRGC channel-wise topology refinement: models/ctrgcn.py:172-177 ---------------------------- */
/* m[n,c,v] = mean_t x[n,c,t,v]  (fp32 out, (N,C,V)).  conv1/conv2 followed by .mean(-2) commute with the
 * mean, so x1/x2 are produced by tamgcn_conv_fwd on m viewed as (N,Cin,T=1,V) with the stacked
 * conv1/conv2 weights; their backward is tamgcn_conv_wgrad/dgrad on the same view plus the
 * `bcast` term of tamgcn_conv_dgrad. */
int tamgcn_mean_t(int dtype, const void* x, int64_t x_nstride, int N, int C, int T, int V, float* m,
                  tamgcn_stream stream);
/* y[n,c,t,u] = sum_i sum_v Q_i[n,c,u,v] * x3[n,i*Cout+c,t,v],
 * Q_i[n,c,u,v] = alpha*(sum_r W4[i,c,r]*tanh(x1[n,i,r,u]-x2[n,i,r,v]) + b4[i,c]) + PA[i,u,v].
 * x1,x2: fp32 (K,R,V) blocks per sample with sample stride x12_nstride; W4 (K,Cout,R); b4 (K,Cout);
 * PA (K,V,V); alpha: device pointer to one float.  The (N,C,V,V) topology tensor lives in shared
 * memory only.  Optional BN statistics of y ("+=").  V must be 20 or 25. */
int tamgcn_ctrgc_fwd(int dtype, const void* x3, int64_t x3_nstride, int N, int Cout, int T, int V, int K, int R,
                     const float* x1, const float* x2, int64_t x12_nstride, const float* W4, const float* b4,
                     const float* PA, const float* alpha, void* y, int64_t y_nstride, double* stat_sum,
                     double* stat_sumsq, tamgcn_stream stream);
/* backward of the above for cotangent g (a lazy operand): dx3 (same layout as x3), and "+=" into
 * dx1,dx2 (same layout as x1,x2), dW4 (K,Cout,R), db4 (K,Cout), dPA (K,V,V), dalpha (1). */
int tamgcn_ctrgc_bwd(int dtype, const tamgcn_operand* g, const void* x3, int64_t x3_nstride, int N, int Cout, int T,
                     int V, int K, int R, const float* x1, const float* x2, int64_t x12_nstride, const float* W4,
                     const float* b4, const float* PA, const float* alpha, void* dx3, int64_t dx3_nstride,
                     float* dx1, float* dx2, float* dW4, float* db4, float* dPA, float* dalpha,
                     tamgcn_stream stream);

/* ---- BatchNorm (nn.BatchNorm2d defaults, models/ctrgcn.py:64,100,115,118,123,186,213,222,230) ---- */
typedef struct tamgcn_bn {
    const double* sum;    /* batch statistics accumulated by a producer epilogue (train) */
    const double* sumsq;
    const float* gamma;   /* NULL -> 1 */
    const float* beta;    /* NULL -> 0 */
    float* rmean;         /* running stats: updated in train, read in eval */
    float* rvar;
    int64_t* nbt;         /* num_batches_tracked (+1 in train), may be NULL */
    float* scale;         /* out: gamma*invstd */
    float* shift;         /* out: beta - mean*scale */
    float* mean;          /* out: mean used for normalisation */
    float* invstd;        /* out */
    int32_t C;
    int32_t reserved;
} tamgcn_bn;
/* up to 8 BatchNorms per launch */
int tamgcn_bn_finalize(int n_bn, const tamgcn_bn* bns, double count, float momentum, float eps, int train,
                       tamgcn_stream stream);
typedef struct tamgcn_bn_bwd {
    const double* s1;     /* sum dYhat */
    const double* s2;     /* sum dYhat * Y(raw) */
    const float* gamma;   /* NULL -> 1 */
    const float* mean;
    const float* invstd;
    float* A;             /* out: dY = A*dYhat + B*Y + C */
    float* B;
    float* Cc;
    float* dgamma;        /* out (=, not +=), may be NULL */
    float* dbeta;
    int32_t C;
    int32_t reserved;
} tamgcn_bn_bwd;
int tamgcn_bn_bwd_coef(int n_bn, const tamgcn_bn_bwd* bns, double count, int train, tamgcn_stream stream);
/* coefficients of the lazily formed difference res - y of unit_gcn (models/ctrgcn.py:256-259), per channel:
 * nb = -sb, c = (ha ? ha : 0) - hb */
int tamgcn_coef_diff(int C, const float* sb, const float* ha, const float* hb, float* nb, float* c, tamgcn_stream stream);

/* ---- fused epilogues (BN + tanh/ReLU + residual): models/ctrgcn.py:255-261,145-146,283; stgcn.py:98-99 */
/* out = relu( sg*y0+hg + tanh(so*z+ho) + res );  res = 0 | r | sr*r+hr */
int tamgcn_gcn_epilogue_fwd(int dtype, int N, int C, int TV, const void* y0, const float* sg, const float* hg,
                            const void* z, const float* so, const float* ho, int res_mode, const void* r,
                            int64_t r_nstride, const float* sr, const float* hr, void* out, tamgcn_stream stream);
/* G = g*[out>0];  DZ = G*(1-tanh(so*z+ho)^2);  s1o += sum DZ;  s2o += sum DZ*z */
int tamgcn_gcn_epilogue_bwd(int dtype, int N, int C, int TV, const void* g, const void* out, const void* z,
                            const float* so, const float* ho, void* G, void* DZ, double* s1o, double* s2o,
                            tamgcn_stream stream);
/* DY = G - DD (in place over G);  DR = G + DD (+ extra) (to dr, may be NULL);  s1g += sum DY; s2g += sum DY*y0;
 * s1d += sum DR; s2d += sum DR*r (r, s1d, s2d may be NULL).  `extra` (may be NULL; only without r): a further cotangent
 * of the identity residual — the gradient TCN_GCN_unit's own residual sends to the same input (models/ctrgcn.py:283) —
 * so that no separate add pass is needed. */
int tamgcn_gcn_mid_bwd(int dtype, int N, int C, int TV, void* G, const void* DD, void* dr, int64_t dr_nstride,
                       const void* y0, const void* r, int64_t r_nstride, double* s1g, double* s2g, double* s1d,
                       double* s2d, const void* extra, int64_t extra_nstride, tamgcn_stream stream);
/* out = f( su*u+hu + res ), f = relu if relu else identity */
int tamgcn_tcn_epilogue_fwd(int dtype, int N, int C, int TV, const void* u, int64_t u_nstride, const float* su,
                            const float* hu, int res_mode, const void* r, int64_t r_nstride, const float* sr,
                            const float* hr, int relu, void* out, tamgcn_stream stream);
/* G = relu ? g*[out>0] : g (G may be NULL when relu==0);  s1 += sum G; s2u += sum G*u; s2r += sum G*r */
int tamgcn_tcn_epilogue_bwd(int dtype, int N, int C, int TV, const void* g, const void* out, int relu, const void* u,
                            int64_t u_nstride, const void* r, int64_t r_nstride, void* G, double* s1, double* s2u,
                            double* s2r, tamgcn_stream stream);
/* MaxPool2d((3,1), stride (s,1), padding (1,0)) over the lazy operand x; optional BN stats of y */
int tamgcn_maxpool_fwd(int dtype, int N, int C, int T, int To, int V, int stride, const tamgcn_operand* x, void* y,
                       int64_t y_nstride, double* stat_sum, double* stat_sumsq, tamgcn_stream stream);
/* dh = (sum over windows whose first arg-max is t of dY) * [x.a*x.P + x.c > 0];  s1 += sum dh; s2 += sum dh*x.P */
int tamgcn_maxpool_bwd(int dtype, int N, int C, int T, int To, int V, int stride, const tamgcn_operand* dy,
                       const tamgcn_operand* x, void* dh, int64_t dh_nstride, double* s1, double* s2,
                       tamgcn_stream stream);

/* ---- ST-GCN graph aggregation: models/stgcn.py:60-62  einsum('nkctv,kvw->nctw') -------------------- */
int tamgcn_graph_agg_fwd(int dtype, int N, int K, int C, int T, int V, const void* y, int64_t y_nstride,
                         const float* A, void* out, int64_t out_nstride, double* stat_sum, double* stat_sumsq,
                         tamgcn_stream stream);
/* dy[n,k*C+c,t,v] = sum_w dOut(n,c,t,w) A[k,v,w];  dA[k,v,w] += sum_{n,c,t} y[n,k*C+c,t,v] dOut(n,c,t,w).
 * dy or dA may be NULL (the other half only): dA feeds nothing but the optimiser, so a step engine runs it apart. */
int tamgcn_graph_agg_bwd(int dtype, int N, int K, int C, int T, int V, const tamgcn_operand* dout, const void* y,
                         int64_t y_nstride, const float* A, void* dy, int64_t dy_nstride, float* dA,
                         tamgcn_stream stream);

/* ---- network ends and optimiser -------------------------------------------------- */
/* Model.forward prologue, models/ctrgcn.py:328-332 (fold_m = 0) and models/stgcn.py:174-181 (fold_m = 1):
 *   x.permute(0,4,3,1,2).view(N, M*V*C, T) -> nn.BatchNorm1d -> view/permute -> (N*M, C, T, V) in `dtype`.
 * x is fp32 with arbitrary element strides x_strides[5] for the (n, c, t, v, m) axes (so both the 5-D input and the
 * (N, T, V*C) input of models/ctrgcn.py:325-327 are read in place).  BatchNorm channel = (m*V + v)*C + c, or v*C + c
 * with the persons folded into the batch (fold_m).  Train: batch statistics (fp64 sums inside the CTA), running-stat
 * update with `momentum`, *nbt += 1; eval: running statistics.  save_mean / save_invstd (per channel, may be NULL)
 * are what tamgcn_data_bn_bwd needs. */
int tamgcn_data_bn_fwd(int dtype, const float* x, const int64_t* x_strides, int N, int C, int T, int V, int M, int fold_m,
                       const float* gamma, const float* beta, float* rmean, float* rvar, int64_t* nbt, float momentum,
                       float eps, int train, void* out, float* save_mean, float* save_invstd, tamgcn_stream stream);
/* g: cotangent of the (N*M, C, T, V) output.  dgamma / dbeta "+=" (NULL to skip); dx: contiguous fp32 (N, C, T, V, M)
 * or NULL when the input needs no gradient. */
int tamgcn_data_bn_bwd(int dtype, const void* g, const float* x, const int64_t* x_strides, int N, int C, int T, int V,
                       int M, int fold_m, const float* gamma, const float* mean, const float* invstd, int train,
                       float* dgamma, float* dbeta, float* dx, tamgcn_stream stream);
/* Model.forward head, models/ctrgcn.py:343-348 / models/stgcn.py:187-195: pooled[n,c] = mean over persons and (T*V) of
 * x[(n*M+m), c, :]; logits = (gate .* pooled) W^T + b (W (K,C), b (K) or NULL; W == NULL: pooling only).  pooled (N,C)
 * (always the RAW mean), logits (N,K) fp32.  gate (N,C) or NULL: the channel attention of the cross-modal head,
 * models/resnet_gcn_attention.py:108-118 — f_rgb * att -> AdaptiveAvgPool2d -> classifier == classifier(att .* mean(f_rgb)). */
int tamgcn_pool_fc_fwd(int dtype, const void* x, int N, int M, int C, int TV, int K, const float* gate, const float* W,
                       const float* b, float* pooled, float* logits, tamgcn_stream stream);
/* with d = sum_k dlogits[n,k] W[k,c]:  g[(n*M+m), c, :] = d * gate[n,c] / (M*TV)  (NULL to skip);
 * dW[k,c] += dlogits[n,k]*gate[n,c]*pooled[n,c];  db[k] += dlogits[n,k];  dgate[n,c] = d * pooled[n,c]  (each NULL to
 * skip).  W == NULL (pooling only, K == C): dlogits is the cotangent of pooled. */
int tamgcn_pool_fc_bwd(int dtype, const float* dlogits, const float* pooled, const float* gate, const float* W, int N,
                       int M, int C, int TV, int K, void* g, float* dW, float* db, float* dgate, tamgcn_stream stream);
/* out (C,R) = f(in (R,C))^T, fp32.  mode 0: f = identity; 1: f = sigmoid; 2: f = in * aux * (1 - aux) with aux (R,C) the
 * sigmoid output (backward of mode 1).  Glue of the attention MLP of models/resnet_gcn_attention.py:59-65, which runs on
 * the convolution kernels with the batch as the position axis. */
int tamgcn_transpose_act(const float* in, const float* aux, int R, int C, int mode, float* out, tamgcn_stream stream);
/* nn.CrossEntropyLoss (mean over the samples whose label is in [0,K); other labels, e.g. -100, are ignored),
 * processor/recognition_rgb.py:19,61.  loss: one float; dlogits (N,K) = d loss / d logits (NULL to skip). */
int tamgcn_softmax_ce_fwd(const float* logits, const int64_t* labels, int N, int K, float* loss, float* dlogits,
                          tamgcn_stream stream);
/* dlogits = dl_saved * gloss[0]  (chain rule through the scalar loss) */
int tamgcn_softmax_ce_bwd(const float* dl_saved, const float* gloss, int N, int K, float* dlogits, tamgcn_stream stream);
/* torch.optim.SGD(momentum, nesterov, weight_decay; dampening 0) over flat fp32 buffers of n elements
 * (processor/recognition_rgb.py:21-28):  g = grads*grad_scale + wd*p;  m = momentum*m + g;
 * p -= lr * (nesterov ? g + momentum*m : m).  `lr` is a DEVICE pointer so a captured CUDA graph follows
 * adjust_learning_rate (processor/recognition_rgb.py:43-46).  A zero-initialised momentum buffer reproduces torch's
 * first step (buf = g).  Buffers must be 16-byte aligned. */
int tamgcn_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, const float* lr, float momentum,
                    float weight_decay, int nesterov, float grad_scale, tamgcn_stream stream);

/* ---- GPU-side skeleton feeder: feeder/feeder_nucla_gcn.py:85-130 for a whole batch --------------- */
/* raw: (S, Lmax, V, 3) fp32 padded skeleton sequences resident on the device, length[S] their frame counts.  For batch
 * element b: sample[b] selects the sequence, view[b] = (agx degrees, agy degrees, scale) the random view transform
 * (0, 0, 1 for evaluation), frame_idx[b, T] the (sorted) frames to keep.  mode 0 joint / 1 bone (bone_parent[V]: for
 * joint a the 0-based joint b subtracted from it, -1 -> zero) / 2 motion.  out: (B, 3, T, V, 1) fp32 in [-1, 1]. */
int tamgcn_feeder_nucla(const float* raw, const int32_t* length, const int64_t* sample, const float* view,
                        const int32_t* frame_idx, const int32_t* bone_parent, int B, int Lmax, int V, int T, int mode,
                        float* out, tamgcn_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* TAMGCN_H */
