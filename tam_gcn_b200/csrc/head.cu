// head.cu — the ends of the network and the optimiser: everything of a training step that is not a
// CTR-GCN / ST-GCN block, hand-written so that a whole step is this library's kernels only.
//
//   data_bn_fwd / data_bn_bwd   Model.forward prologue: permute (N,C,T,V,M) -> (N, M*V*C, T), BatchNorm1d, permute back
//                               to (N*M, C, T, V) and cast to the activation dtype — ONE kernel, one CTA per BatchNorm
//                               channel (m, v, c); statistics, running-stat update and normalisation in the same CTA.
//                               reference: models/ctrgcn.py:324-332, models/stgcn.py:174-181
//   pool_fc_fwd / pool_fc_bwd   global average pool over (T*V) and persons + nn.Linear   (models/ctrgcn.py:343-348)
//   softmax_ce_fwd / _bwd       nn.CrossEntropyLoss (mean reduction, ignore_index -100)   (processor/recognition_rgb.py:19,61)
//   sgd_step                    torch.optim.SGD(momentum, nesterov, weight_decay) over ONE flat parameter / gradient /
//                               momentum buffer, learning rate read from device memory so a captured CUDA graph follows
//                               adjust_learning_rate (processor/recognition_rgb.py:21-28,43-46)
//
// All of these are tiny, latency-bound passes (the input batch is 0.8 MB, the parameters 6.8 MB); the design goal is
// ONE launch each and coalesced 16-byte accesses where the tensor is large enough to matter (sgd_step, pool).
#include "common.cuh"

namespace tamgcn {

// block-wide sum of two doubles (result valid in every thread)
__device__ __forceinline__ void block_sum2d(double& a, double& b, double* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    __syncthreads();
    if (lane == 0) { scratch[w] = a; scratch[32 + w] = b; }
    __syncthreads();
    double x = 0.0, y = 0.0;
    for (int i = 0; i < nw; ++i) { x += scratch[i]; y += scratch[32 + i]; }
    a = x; b = y;
}

// ---- data_bn ----------------------------------------------------------------------------------------------------
struct DataBnP {
    int N, C, T, V, M;
    int fold_m;                          // 1: persons belong to the batch axis (channels = V*C, ST-GCN); 0: channels = M*V*C
    long long sn, sc, st, sv, sm;        // element strides of the (n, c, t, v, m) axes of the fp32 input
    float momentum, eps;
    int train;
};

// channel ch = (m*V + v)*C + c  (the order x.permute(0,4,3,1,2).view(N, M*V*C, T) gives, models/ctrgcn.py:329);
// with fold_m the channel is v*C + c and the statistics also run over the persons (view(N*M, V*C, T), models/stgcn.py:176).
// Element i of a channel: (n, mm, t) = (i / (ML*T), (i / T) % ML, i % T), person m = m0 + mm.
struct DataBnIdx { int n, m, t; };
__device__ __forceinline__ DataBnIdx data_bn_idx(int i, int T, int ML, int m0) {
    DataBnIdx r;
    const int nm = i / T;
    r.t = i - nm * T;
    r.n = nm / ML;
    r.m = m0 + (nm - r.n * ML);
    return r;
}
template <typename T>
__global__ void __launch_bounds__(256)
data_bn_fwd_kernel(DataBnP p, const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* rmean, float* rvar, long long* nbt, T* __restrict__ out, float* save_mean, float* save_invstd) {
    __shared__ double scratch[64];
    const int ch = blockIdx.x;
    const int c = ch % p.C, v = (ch / p.C) % p.V, m0 = p.fold_m ? 0 : ch / (p.C * p.V);
    const int ML = p.fold_m ? p.M : 1;
    const float* xc = x + c * p.sc + v * p.sv;
    const int NT = p.N * ML * p.T;
    double mean, var;
    if (p.train) {
        double s = 0.0, q = 0.0;
        for (int i = threadIdx.x; i < NT; i += blockDim.x) {
            const DataBnIdx k = data_bn_idx(i, p.T, ML, m0);
            const double val = (double)__ldg(xc + k.n * p.sn + k.t * p.st + k.m * p.sm);
            s += val; q += val * val;
        }
        block_sum2d(s, q, scratch);
        mean = s / NT;
        var = q / NT - mean * mean;
        if (var < 0.0) var = 0.0;
        if (threadIdx.x == 0) {
            if (rmean) {
                const double unb = NT > 1 ? var * NT / (NT - 1.0) : var;
                rmean[ch] = (float)((1.0 - (double)p.momentum) * (double)rmean[ch] + (double)p.momentum * mean);
                rvar[ch] = (float)((1.0 - (double)p.momentum) * (double)rvar[ch] + (double)p.momentum * unb);
            }
            if (nbt && ch == 0) *nbt += 1;
        }
    } else {
        mean = (double)rmean[ch];
        var = (double)rvar[ch];
    }
    const float invstd = (float)(1.0 / sqrt(var + (double)p.eps));
    const float scale = (gamma ? gamma[ch] : 1.f) * invstd;
    const float shift = (beta ? beta[ch] : 0.f) - (float)mean * scale;
    if (threadIdx.x == 0) {
        if (save_mean) save_mean[ch] = (float)mean;
        if (save_invstd) save_invstd[ch] = invstd;
    }
    T* oc = out + (long long)c * p.T * p.V + v;
    for (int i = threadIdx.x; i < NT; i += blockDim.x) {
        const DataBnIdx k = data_bn_idx(i, p.T, ML, m0);
        const float val = __ldg(xc + k.n * p.sn + k.t * p.st + k.m * p.sm);
        stf<T>(oc + ((long long)(k.n * p.M + k.m) * p.C * p.T + k.t) * p.V, fmaf(val, scale, shift));
    }
}

// dgamma = sum g*xhat, dbeta = sum g ("+="), dx = gamma*invstd*(g - dbeta/NT - xhat*dgamma/NT)  (train)
//                                             dx = gamma*invstd*g                                 (eval)
template <typename T>
__global__ void __launch_bounds__(256)
data_bn_bwd_kernel(DataBnP p, const T* __restrict__ g, const float* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ mean, const float* __restrict__ invstd, float* dgamma, float* dbeta,
                   float* __restrict__ dx) {
    __shared__ double scratch[64];
    const int ch = blockIdx.x;
    const int c = ch % p.C, v = (ch / p.C) % p.V, m0 = p.fold_m ? 0 : ch / (p.C * p.V);
    const int ML = p.fold_m ? p.M : 1;
    const float* xc = x + c * p.sc + v * p.sv;
    const T* gc = g + (long long)c * p.T * p.V + v;
    const int NT = p.N * ML * p.T;
    const float mu = mean[ch], is = invstd[ch];
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < NT; i += blockDim.x) {
        const DataBnIdx k = data_bn_idx(i, p.T, ML, m0);
        const float gv = ldf<T>(gc + ((long long)(k.n * p.M + k.m) * p.C * p.T + k.t) * p.V);
        const float xh = (__ldg(xc + k.n * p.sn + k.t * p.st + k.m * p.sm) - mu) * is;
        s1 += (double)gv; s2 += (double)gv * (double)xh;
    }
    block_sum2d(s1, s2, scratch);
    if (threadIdx.x == 0) {
        if (dgamma) dgamma[ch] += (float)s2;
        if (dbeta) dbeta[ch] += (float)s1;
    }
    if (!dx) return;
    const float a = (gamma ? gamma[ch] : 1.f) * is;
    const float k1 = p.train ? (float)(s1 / NT) : 0.f, k2 = p.train ? (float)(s2 / NT) : 0.f;
    // dx is contiguous (N, C, T, V, M)
    float* dc = dx + ((long long)c * p.T * p.V + v) * p.M;
    for (int i = threadIdx.x; i < NT; i += blockDim.x) {
        const DataBnIdx k = data_bn_idx(i, p.T, ML, m0);
        const float gv = ldf<T>(gc + ((long long)(k.n * p.M + k.m) * p.C * p.T + k.t) * p.V);
        const float xh = (__ldg(xc + k.n * p.sn + k.t * p.st + k.m * p.sm) - mu) * is;
        dc[((long long)k.n * p.C * p.T + k.t) * p.V * p.M + k.m] = a * (gv - k1 - xh * k2);
    }
}

// ---- global average pool + linear ----------------------------------------------------------------------------------
// pooled[n,c] = mean over (m, t, v) of x[n*M+m, c, :, :];  logits[n,k] = b[k] + sum_c W[k,c]*gate[n,c]*pooled[n,c]
// (x.view(N,M,C,-1).mean(3).mean(1) == mean over all M*TV elements since every person has the same TV)
// phase 0: one CTA per sample does both (small planes: one launch).  Large planes with few samples (ST-GCN: 16 samples
// x 2 persons x 75 x 25 positions) would leave most SMs idle, so the work is split: phase 1 pools a slice of CS channels
// per CTA (grid N x C/CS), phase 2 (grid N) reads the pooled vector back and applies the classifier.
template <typename T>
__global__ void __launch_bounds__(256)
pool_fc_fwd_kernel(int N, int M, int C, int TV, int K, int phase, int CS, const T* __restrict__ x,
                   const float* __restrict__ gate, const float* __restrict__ W, const float* __restrict__ b,
                   float* __restrict__ pooled, float* __restrict__ logits) {
    extern __shared__ float sp[];                       // C floats
    const int n = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float inv = 1.f / (float)(M * TV);
    const int c_lo = phase == 1 ? blockIdx.y * CS : 0, c_hi = phase == 1 ? min(C, c_lo + CS) : C;
    if (phase == 2) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const float s = pooled[(long long)n * C + c];
            sp[c] = gate ? s * __ldg(gate + (long long)n * C + c) : s;
        }
    } else
    for (int c = c_lo + w; c < c_hi; c += nw) {
        float s = 0.f;
        for (int m = 0; m < M; ++m) {
            const T* px = x + ((long long)(n * M + m) * C + c) * TV;
            for (int e = lane; e < TV; e += 32) s += ldf<T>(px + e);
        }
        s = warp_sum(s) * inv;
        if (lane == 0) {
            pooled[(long long)n * C + c] = s;                       // the raw mean is what the backward needs
            sp[c] = gate ? s * __ldg(gate + (long long)n * C + c) : s;
        }
    }
    __syncthreads();
    if (!W || phase == 1) return;
    for (int k = w; k < K; k += nw) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(__ldg(W + (long long)k * C + c), sp[c], s);
        s = warp_sum(s);
        if (lane == 0) logits[(long long)n * K + k] = s + (b ? __ldg(b + k) : 0.f);
    }
}

// dpooled[n,c] = sum_k dl[n,k] W[k,c];  g[n*M+m, c, :, :] = dpooled[n,c] / (M*TV);
// dW[k,c] += dl[n,k]*pooled[n,c];  db[k] += dl[n,k]
template <typename T>
__global__ void __launch_bounds__(256)
pool_fc_bwd_kernel(int N, int M, int C, int TV, int K, const float* __restrict__ dl, const float* __restrict__ pooled,
                   const float* __restrict__ gate, const float* __restrict__ W, T* __restrict__ g, float* dW, float* db,
                   float* __restrict__ dgate, int CS) {
    extern __shared__ float sm[];                       // K floats dl, C floats dpooled
    float* sdl = sm;
    float* sdp = sm + K;
    const int n = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int c_lo = blockIdx.y * CS, c_hi = min(C, c_lo + CS);      // this CTA's channel slice (grid N x C/CS)
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float v = __ldg(dl + (long long)n * K + k);
        sdl[k] = v;
        if (db && blockIdx.y == 0) atomicAdd(db + k, v);
    }
    __syncthreads();
    const float inv = 1.f / (float)(M * TV);
    for (int c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x) {
        float s = 0.f;
        const float gt = gate ? __ldg(gate + (long long)n * C + c) : 1.f;
        if (W) {
            const float mean = __ldg(pooled + (long long)n * C + c);
            const float pc = mean * gt;
            for (int k = 0; k < K; ++k) {
                s = fmaf(sdl[k], __ldg(W + (long long)k * C + c), s);
                if (dW) atomicAdd(dW + (long long)k * C + c, sdl[k] * pc);
            }
            if (dgate) dgate[(long long)n * C + c] = s * mean;
        } else {
            s = sdl[c];                                 // pooling only: dlogits IS the cotangent of pooled (K == C)
        }
        sdp[c] = s * gt * inv;
    }
    __syncthreads();
    if (!g) return;
    const int cs = c_hi - c_lo;
    for (int mc = w; mc < M * cs; mc += nw) {
        const int m = mc / cs, c = c_lo + (mc - m * cs);
        const float v = sdp[c];
        T* pg = g + ((long long)(n * M + m) * C + c) * TV;
        for (int e = lane; e < TV; e += 32) stf<T>(pg + e, v);
    }
}

// ---- softmax cross-entropy (mean over the non-ignored samples) -------------------------------------------------------
// single CTA (N*K is a few hundred values).  dl[n,k] = (softmax - onehot)/count is saved for the backward.
__global__ void __launch_bounds__(256)
softmax_ce_fwd_kernel(int N, int K, const float* __restrict__ logits, const long long* __restrict__ labels,
                      float* __restrict__ loss, float* __restrict__ dl) {
    __shared__ float s_loss[256];
    __shared__ int s_cnt[256];
    float my = 0.f;
    int cnt = 0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const long long y = labels[n];
        if (y >= 0 && y < K) ++cnt;
    }
    s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o];
        __syncthreads();
    }
    const int count = s_cnt[0];
    const float invc = count > 0 ? 1.f / (float)count : 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float* l = logits + (long long)n * K;
        const long long y = labels[n];
        const bool valid = (y >= 0 && y < K);
        float mx = -INFINITY;
        for (int k = 0; k < K; ++k) mx = fmaxf(mx, l[k]);
        float se = 0.f;
        for (int k = 0; k < K; ++k) se += expf(l[k] - mx);
        const float lse = mx + logf(se);
        if (valid) my += lse - l[y];
        if (dl) {
            for (int k = 0; k < K; ++k)
                dl[(long long)n * K + k] = valid ? (expf(l[k] - lse) - (k == (int)y ? 1.f : 0.f)) * invc : 0.f;
        }
    }
    // deterministic tree sum
    __syncthreads();
    s_loss[threadIdx.x] = my;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s_loss[threadIdx.x] += s_loss[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = count > 0 ? s_loss[0] * invc : __int_as_float(0x7fc00000);
}

__global__ void __launch_bounds__(256)
scale_by_scalar_kernel(long long n, const float* __restrict__ a, const float* __restrict__ s, float* __restrict__ out) {
    const float k = __ldg(s);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = a[i] * k;
}

// ---- (R, C) -> (C, R) transpose of a small fp32 matrix with an optional activation ------------------------------------
// mode 0: out = in^T;  1: out = sigmoid(in)^T;  2: out = (in * aux * (1 - aux))^T  (aux = the sigmoid output, (R, C):
// backward of mode 1 for a cotangent `in` of the TRANSPOSED result, i.e. in and aux are both (R, C))
__global__ void __launch_bounds__(256)
transpose_act_kernel(int R, int C, int mode, const float* __restrict__ in, const float* __restrict__ aux,
                     float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int r = r0 + j, c = c0 + tx;
        float v = 0.f;
        if (r < R && c < C) {
            v = __ldg(in + (long long)r * C + c);
            if (mode == 1) v = 1.f / (1.f + expf(-v));
            else if (mode == 2) { const float a = __ldg(aux + (long long)r * C + c); v = v * a * (1.f - a); }
        }
        tile[j][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, r = r0 + tx;
        if (r < R && c < C) out[(long long)c * R + r] = tile[tx][j];
    }
}

// ---- SGD (momentum, nesterov, weight decay) over flat buffers -------------------------------------------------------
// g = G*grad_scale + wd*p;  m = mu*m + g;  p -= lr * (nesterov ? g + mu*m : m)      [torch.optim.SGD, dampening 0;
// a zero-initialised momentum buffer reproduces torch's "first step: buf = g"]
__global__ void __launch_bounds__(256)
sgd_step_kernel(long long n4, long long n, float4* __restrict__ P, const float4* __restrict__ G, float4* __restrict__ Mo,
                const float* __restrict__ lr_p, float mu, float wd, int nesterov, float gs) {
    const float lr = __ldg(lr_p);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 p = P[i], g = __ldg(G + i), m = Mo[i];
#define TG_SGD1(c)                                                  \
        {                                                           \
            const float gg = fmaf(g.c, gs, wd * p.c);               \
            m.c = fmaf(mu, m.c, gg);                                \
            p.c -= lr * (nesterov ? fmaf(mu, m.c, gg) : m.c);       \
        }
        TG_SGD1(x) TG_SGD1(y) TG_SGD1(z) TG_SGD1(w)
        P[i] = p; Mo[i] = m;
    }
    // tail (n not a multiple of 4)
    float* Ps = (float*)P; const float* Gs = (const float*)G; float* Ms = (float*)Mo;
    for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gg = fmaf(Gs[i], gs, wd * Ps[i]);
        const float mm = fmaf(mu, Ms[i], gg);
        Ms[i] = mm;
        Ps[i] -= lr * (nesterov ? fmaf(mu, mm, gg) : mm);
    }
#undef TG_SGD1
}

}  // namespace tamgcn

using namespace tamgcn;

static int data_bn_params(DataBnP& p, const char* who, int N, int C, int T, int V, int M, int fold_m,
                          const int64_t* strides, float momentum, float eps, int train) {
    TG_REQUIRE(N > 0 && C > 0 && T > 0 && V > 0 && M > 0, "%s: empty shape N=%d C=%d T=%d V=%d M=%d", who, N, C, T, V, M);
    TG_REQUIRE(strides, "%s: strides missing", who);
    p.N = N; p.C = C; p.T = T; p.V = V; p.M = M; p.fold_m = fold_m ? 1 : 0;
    p.sn = strides[0]; p.sc = strides[1]; p.st = strides[2]; p.sv = strides[3]; p.sm = strides[4];
    p.momentum = momentum; p.eps = eps; p.train = train;
    return 0;
}

extern "C" int tamgcn_data_bn_fwd(int dtype, const float* x, const int64_t* x_strides, int N, int C, int T, int V, int M,
                                  int fold_m, const float* gamma, const float* beta, float* rmean, float* rvar, int64_t* nbt,
                                  float momentum, float eps, int train, void* out, float* save_mean, float* save_invstd,
                                  tamgcn_stream stream) {
    DataBnP p;
    if (int rc = data_bn_params(p, "data_bn_fwd", N, C, T, V, M, fold_m, x_strides, momentum, eps, train)) return rc;
    TG_REQUIRE(x && out, "data_bn_fwd: null tensor");
    TG_REQUIRE(train || (rmean && rvar), "data_bn_fwd: eval mode needs running statistics");
    TG_REQUIRE(dtype == TAMGCN_F32 || dtype == TAMGCN_BF16, "data_bn_fwd: bad dtype %d", dtype);
    const int grid = (fold_m ? 1 : M) * V * C;
    if (dtype == TAMGCN_F32)
        data_bn_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(p, x, gamma, beta, rmean, rvar, (long long*)nbt,
                                                                          (float*)out, save_mean, save_invstd);
    else
        data_bn_fwd_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(p, x, gamma, beta, rmean, rvar, (long long*)nbt,
                                                                         (bf16*)out, save_mean, save_invstd);
    count_launch();
    return check_launch("data_bn_fwd");
}

extern "C" int tamgcn_data_bn_bwd(int dtype, const void* g, const float* x, const int64_t* x_strides, int N, int C, int T,
                                  int V, int M, int fold_m, const float* gamma, const float* mean, const float* invstd, int train,
                                  float* dgamma, float* dbeta, float* dx, tamgcn_stream stream) {
    DataBnP p;
    if (int rc = data_bn_params(p, "data_bn_bwd", N, C, T, V, M, fold_m, x_strides, 0.f, 0.f, train)) return rc;
    TG_REQUIRE(g && x && mean && invstd, "data_bn_bwd: null tensor");
    TG_REQUIRE(dtype == TAMGCN_F32 || dtype == TAMGCN_BF16, "data_bn_bwd: bad dtype %d", dtype);
    const int grid = (fold_m ? 1 : M) * V * C;
    if (dtype == TAMGCN_F32)
        data_bn_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(p, (const float*)g, x, gamma, mean, invstd, dgamma,
                                                                          dbeta, dx);
    else
        data_bn_bwd_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(p, (const bf16*)g, x, gamma, mean, invstd, dgamma,
                                                                         dbeta, dx);
    count_launch();
    return check_launch("data_bn_bwd");
}

extern "C" int tamgcn_pool_fc_fwd(int dtype, const void* x, int N, int M, int C, int TV, int K, const float* gate,
                                  const float* W, const float* b, float* pooled, float* logits, tamgcn_stream stream) {
    TG_REQUIRE(N > 0 && M > 0 && C > 0 && TV > 0, "pool_fc_fwd: empty shape N=%d M=%d C=%d TV=%d", N, M, C, TV);
    TG_REQUIRE(x && pooled, "pool_fc_fwd: null tensor");
    TG_REQUIRE(!W || (K > 0 && logits), "pool_fc_fwd: linear layer needs K > 0 and a logits buffer");
    TG_REQUIRE(C * sizeof(float) <= 48 * 1024, "pool_fc_fwd: C=%d too large", C);
    TG_REQUIRE(dtype == TAMGCN_F32 || dtype == TAMGCN_BF16, "pool_fc_fwd: bad dtype %d", dtype);
    const size_t sm = C * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    // few samples with large planes: split into a channel-sliced pooling launch and a classifier launch
    const bool split = (long long)N * 2 < num_sms() && (long long)M * TV >= 1024 && C >= 32;
    const int CS = 16;
    auto launch = [&](dim3 grid, int phase) {
        if (dtype == TAMGCN_F32)
            pool_fc_fwd_kernel<float><<<grid, 256, sm, st>>>(N, M, C, TV, K, phase, CS, (const float*)x, gate, W, b, pooled, logits);
        else
            pool_fc_fwd_kernel<bf16><<<grid, 256, sm, st>>>(N, M, C, TV, K, phase, CS, (const bf16*)x, gate, W, b, pooled, logits);
        count_launch();
    };
    if (!split) {
        launch(dim3(N), 0);
    } else {
        launch(dim3(N, (C + CS - 1) / CS), 1);
        if (W) launch(dim3(N), 2);
    }
    return check_launch("pool_fc_fwd");
}

extern "C" int tamgcn_pool_fc_bwd(int dtype, const float* dlogits, const float* pooled, const float* gate, const float* W,
                                  int N, int M, int C, int TV, int K, void* g, float* dW, float* db, float* dgate,
                                  tamgcn_stream stream) {
    TG_REQUIRE(N > 0 && M > 0 && C > 0 && TV > 0 && K > 0, "pool_fc_bwd: empty shape");
    TG_REQUIRE(dlogits && (W ? pooled != nullptr : (K == C && !dW && !db && !dgate)), "pool_fc_bwd: null tensor / pooling-only needs K == C");
    TG_REQUIRE((size_t)(C + K) * sizeof(float) <= 48 * 1024, "pool_fc_bwd: C+K too large");
    TG_REQUIRE(dtype == TAMGCN_F32 || dtype == TAMGCN_BF16, "pool_fc_bwd: bad dtype %d", dtype);
    const size_t sm = (size_t)(C + K) * sizeof(float);
    // channel slices per CTA: enough CTAs to fill the chip when there are few samples
    int CS = C;
    while (CS > 16 && (long long)N * ((C + CS - 1) / CS) < 2LL * num_sms()) CS = (CS + 1) / 2;
    dim3 grid(N, (C + CS - 1) / CS);
    if (dtype == TAMGCN_F32)
        pool_fc_bwd_kernel<float><<<grid, 256, sm, (cudaStream_t)stream>>>(N, M, C, TV, K, dlogits, pooled, gate, W, (float*)g, dW, db, dgate, CS);
    else
        pool_fc_bwd_kernel<bf16><<<grid, 256, sm, (cudaStream_t)stream>>>(N, M, C, TV, K, dlogits, pooled, gate, W, (bf16*)g, dW, db, dgate, CS);
    count_launch();
    return check_launch("pool_fc_bwd");
}

extern "C" int tamgcn_softmax_ce_fwd(const float* logits, const int64_t* labels, int N, int K, float* loss, float* dlogits,
                                     tamgcn_stream stream) {
    TG_REQUIRE(N > 0 && K > 0 && logits && labels && loss, "softmax_ce_fwd: bad arguments N=%d K=%d", N, K);
    softmax_ce_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(N, K, logits, (const long long*)labels, loss, dlogits);
    count_launch();
    return check_launch("softmax_ce_fwd");
}

extern "C" int tamgcn_softmax_ce_bwd(const float* dl_saved, const float* gloss, int N, int K, float* dlogits,
                                     tamgcn_stream stream) {
    TG_REQUIRE(N > 0 && K > 0 && dl_saved && gloss && dlogits, "softmax_ce_bwd: bad arguments");
    const long long n = (long long)N * K;
    int grid = (int)((n + 255) / 256);
    if (grid > 1024) grid = 1024;
    scale_by_scalar_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, dl_saved, gloss, dlogits);
    count_launch();
    return check_launch("softmax_ce_bwd");
}

extern "C" int tamgcn_transpose_act(const float* in, const float* aux, int R, int C, int mode, float* out,
                                    tamgcn_stream stream) {
    TG_REQUIRE(R > 0 && C > 0 && in && out && mode >= 0 && mode <= 2 && (mode != 2 || aux), "transpose_act: bad arguments");
    dim3 grid((C + 31) / 32, (R + 31) / 32);
    TG_REQUIRE(grid.y <= 65535, "transpose_act: too many rows");
    transpose_act_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(R, C, mode, in, aux, out);
    count_launch();
    return check_launch("transpose_act");
}

extern "C" int tamgcn_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, const float* lr,
                               float momentum, float weight_decay, int nesterov, float grad_scale, tamgcn_stream stream) {
    TG_REQUIRE(n > 0 && params && grads && momentum_buf && lr, "sgd_step: bad arguments");
    TG_REQUIRE(((uintptr_t)params & 15) == 0 && ((uintptr_t)grads & 15) == 0 && ((uintptr_t)momentum_buf & 15) == 0,
               "sgd_step: buffers must be 16-byte aligned");
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    sgd_step_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(n4, n, (float4*)params, (const float4*)grads,
                                                                  (float4*)momentum_buf, lr, momentum, weight_decay,
                                                                  nesterov, grad_scale);
    count_launch();
    return check_launch("sgd_step");
}
