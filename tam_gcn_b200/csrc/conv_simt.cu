// conv_simt.cu — (k x 1) convolutions as implicit GEMMs on the fp32 SIMT pipes.
//
// This is the exact-fp32 path (parity mode) and the fallback shape-general path; the bf16 tensor-core
// (tcgen05) path for the same entry points lives in conv_tc.cu.  One kernel each for forward,
// data-gradient and weight-gradient; all three read their activation operands through the lazy
// `Opnd` transform (BatchNorm-apply + ReLU forward, BatchNorm-backward affine, res - y difference),
// and carry the BatchNorm statistics reductions in their epilogues.
//
// Tiling: 64 x 64 output tile per CTA, K step 16, 256 threads, 4 x 4 register micro-tile.
#include "common.cuh"

namespace tamgcn {

struct ConvP {
    int N, Cin, Cout, T, To, V, k, s, d, p;
};

#define TM 64
#define TN 64
#define TK 16
#define TPAD 4

// ------------------------------------------------------------------------------------------------
// forward:  M = Cout, N = output positions (to, v) of one sample, K = Cin * k
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
conv_fwd_kernel(ConvP g, Opnd x, const float* __restrict__ W, const float* __restrict__ bias, T* __restrict__ y,
                long long yns, double* ssum, double* ssq, int stat_c0) {
    __shared__ __align__(16) float Ws[TK][TM + TPAD];
    __shared__ __align__(16) float Xs[TK][TN + TPAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n = blockIdx.z, co0 = blockIdx.y * TM, pos0 = blockIdx.x * TN;
    const int TV = g.T * g.V, Lo = g.To * g.V, Kt = g.Cin * g.k;

    const int nx = tid & 63, kx0 = tid >> 6;
    const int pos = pos0 + nx;
    const bool pvalid = pos < Lo;
    const int to = pos / g.V, v = pos - to * g.V;
    const int tbase = to * g.s - g.p;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < Kt; k0 += TK) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int e = tid + r * 256, m = e >> 4, kk = e & 15;
            float w = 0.f;
            if (co0 + m < g.Cout && k0 + kk < Kt) w = __ldg(W + (long long)(co0 + m) * Kt + k0 + kk);
            Ws[kk][m] = w;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int kx = kx0 + 4 * r, kk = k0 + kx;
            float xv = 0.f;
            if (pvalid && kk < Kt) {
                const int ci = kk / g.k, j = kk - ci * g.k;
                const int t = tbase + j * g.d;
                if (t >= 0 && t < g.T) {
                    const OpCoef cf = opnd_coef(x, ci);
                    xv = opnd_val<T>(x, cf, n, (long long)ci * TV + t * g.V + v);
                }
            }
            Xs[kx][nx] = xv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&Ws[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        const bool cvalid = co < g.Cout;
        const float bv = (cvalid && bias) ? __ldg(bias + co) : 0.f;
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p2 = pos0 + tx * 4 + j;
            if (cvalid && p2 < Lo) {
                const float val = rnd<T>(acc[i][j] + bv);
                stf<T>(y + (long long)n * yns + (long long)co * Lo + p2, val);
                s += val;
                q = fmaf(val, val, q);
            }
        }
        if (ssum) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, o);
                q += __shfl_xor_sync(0xffffffffu, q, o);
            }
            if (tx == 0 && cvalid && co >= stat_c0) {
                atomicAdd(ssum + (co - stat_c0), (double)s);
                atomicAdd(ssq + (co - stat_c0), (double)q);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// data gradient:  M = Cin, N = input positions (t, v) of one sample, K = Cout * k
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
conv_dgrad_kernel(ConvP g, Opnd dy, const float* __restrict__ W, T* __restrict__ dx, long long dxns,
                  const T* __restrict__ addend, long long addns, const float* __restrict__ bcast, float bscale,
                  Opnd mask, int has_mask, double* s1, double* s2) {
    __shared__ __align__(16) float Ws[TK][TM + TPAD];
    __shared__ __align__(16) float Xs[TK][TN + TPAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n = blockIdx.z, ci0 = blockIdx.y * TM, pos0 = blockIdx.x * TN;
    const int TV = g.T * g.V, Lo = g.To * g.V, Kt = g.Cout * g.k, CK = g.Cin * g.k;

    const int nx = tid & 63, kx0 = tid >> 6;
    const int pos = pos0 + nx;
    const bool pvalid = pos < TV;
    const int t = pos / g.V, v = pos - t * g.V;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < Kt; k0 += TK) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int kx = kx0 + 4 * r, kk = k0 + kx;  // m = nx (ci), K index = kk = co*k + j
            float w = 0.f;
            if (ci0 + nx < g.Cin && kk < Kt) {
                const int co = kk / g.k, j = kk - co * g.k;
                w = __ldg(W + (long long)co * CK + (ci0 + nx) * g.k + j);
            }
            Ws[kx][nx] = w;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int kx = kx0 + 4 * r, kk = k0 + kx;
            float xv = 0.f;
            if (pvalid && kk < Kt) {
                const int co = kk / g.k, j = kk - co * g.k;
                const int num = t + g.p - j * g.d;
                if (num >= 0) {
                    const int to = num / g.s;
                    if (to * g.s == num && to < g.To) {
                        const OpCoef cf = opnd_coef(dy, co);
                        xv = opnd_val<T>(dy, cf, n, (long long)co * Lo + to * g.V + v);
                    }
                }
            }
            Xs[kx][nx] = xv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&Ws[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ci = ci0 + ty * 4 + i;
        const bool cvalid = ci < g.Cin;
        OpCoef mc;
        mc.a = 1.f; mc.b = 0.f; mc.c = 0.f;
        if (has_mask && cvalid) mc = opnd_coef(mask, ci);
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p2 = pos0 + tx * 4 + j;
            if (cvalid && p2 < TV) {
                const long long off = (long long)ci * TV + p2;
                float val = acc[i][j];
                if (addend) val += ldf<T>(addend + (long long)n * addns + off);
                if (bcast) val = fmaf(__ldg(bcast + ((long long)n * g.Cin + ci) * g.V + (p2 % g.V)), bscale, val);
                if (has_mask) {
                    const float pv = ldf<T>((const T*)mask.p + (long long)n * mask.pns + off);
                    if (!(fmaf(mc.a, pv, mc.c) > 0.f)) val = 0.f;
                    val = rnd<T>(val);
                    s += val;
                    q = fmaf(val, pv, q);
                }
                stf<T>(dx + (long long)n * dxns + off, val);
            }
        }
        if (has_mask && s1) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, o);
                q += __shfl_xor_sync(0xffffffffu, q, o);
            }
            if (tx == 0 && cvalid) {
                atomicAdd(s1 + ci, (double)s);
                atomicAdd(s2 + ci, (double)q);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// weight gradient:  M = Cout, N = (ci, j), K = (n, to, v) split over blockIdx.z; fp32 atomics into dW
// ------------------------------------------------------------------------------------------------
#define WG_BLK 256  // positions per work unit
template <typename T>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(ConvP g, Opnd dy, Opnd x, float* __restrict__ dW, float* __restrict__ dbias, int nblk) {
    __shared__ __align__(16) float Ds[TK][TM + TPAD];
    __shared__ __align__(16) float Xs[TK][TN + TPAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int co0 = blockIdx.y * TM, col0 = blockIdx.x * TN;
    const int TV = g.T * g.V, Lo = g.To * g.V, CK = g.Cin * g.k;
    const int kk = tid & 15, mm0 = tid >> 4;  // loader: position kk of the chunk, rows mm0 + 16 r

    int lci[4], lj[4];
    OpCoef xcf[4], dcf[4];
    bool lvalid[4], dvalid[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int col = col0 + mm0 + 16 * r;
        lvalid[r] = col < CK;
        lci[r] = lvalid[r] ? col / g.k : 0;
        lj[r] = col - lci[r] * g.k;
        xcf[r] = opnd_coef(x, lci[r]);
        const int co = co0 + mm0 + 16 * r;
        dvalid[r] = co < g.Cout;
        dcf[r] = opnd_coef(dy, dvalid[r] ? co : 0);
    }
    float acc[4][4];
    float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int units = g.N * nblk;
    for (int u = blockIdx.z; u < units; u += gridDim.z) {
        const int n = u / nblk, pbase = (u - n * nblk) * WG_BLK;
        const int pend = min(Lo, pbase + WG_BLK);
        for (int p0 = pbase; p0 < pend; p0 += TK) {
            const int pos = p0 + kk;
            const bool pvalid = pos < pend;
            const int to = pos / g.V, v = pos - to * g.V;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float dv = 0.f;
                if (pvalid && dvalid[r])
                    dv = opnd_val<T>(dy, dcf[r], n, (long long)(co0 + mm0 + 16 * r) * Lo + pos);
                Ds[kk][mm0 + 16 * r] = dv;
                dbacc[r] += dv;
                float xv = 0.f;
                if (pvalid && lvalid[r]) {
                    const int t = to * g.s + lj[r] * g.d - g.p;
                    if (t >= 0 && t < g.T) xv = opnd_val<T>(x, xcf[r], n, (long long)lci[r] * TV + t * g.V + v);
                }
                Xs[kk][mm0 + 16 * r] = xv;
            }
            __syncthreads();
#pragma unroll
            for (int k2 = 0; k2 < TK; ++k2) {
                const float4 a = *reinterpret_cast<const float4*>(&Ds[k2][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Xs[k2][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        if (co >= g.Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = col0 + tx * 4 + j;
            if (col < CK) atomicAdd(dW + (long long)co * CK + col, acc[i][j]);
        }
    }
    if (dbias && blockIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float s = dbacc[r];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (kk == 0 && dvalid[r]) atomicAdd(dbias + co0 + mm0 + 16 * r, s);
        }
    }
}

static int check_geom(const tamgcn_conv_geom* g) {
    TG_REQUIRE(g != nullptr, "conv: null geometry");
    TG_REQUIRE(g->N > 0 && g->Cin > 0 && g->Cout > 0 && g->T > 0 && g->To > 0 && g->V > 0, "conv: empty dimension");
    TG_REQUIRE(g->k >= 1 && g->stride >= 1 && g->dil >= 1 && g->pad >= 0, "conv: bad kernel/stride/dilation/pad");
    const int to = (g->T + 2 * g->pad - g->dil * (g->k - 1) - 1) / g->stride + 1;
    TG_REQUIRE(to == g->To, "conv: To=%d inconsistent with T=%d k=%d s=%d d=%d p=%d (expect %d)", g->To, g->T, g->k,
               g->stride, g->dil, g->pad, to);
    TG_REQUIRE(g->N <= 65535, "conv: N=%d exceeds grid.z limit", g->N);
    return 0;
}
static ConvP to_p(const tamgcn_conv_geom* g) {
    ConvP p = {g->N, g->Cin, g->Cout, g->T, g->To, g->V, g->k, g->stride, g->dil, g->pad};
    return p;
}

}  // namespace tamgcn

using namespace tamgcn;

// tensor-core (tcgen05) paths, conv_tc.cu: return 1 if they handled the call, 0 to fall through, <0 on error
namespace tamgcn {
int conv_fwd_tc(const tamgcn_conv_geom* g, const Opnd& x, const void* wpack, const float* bias, void* y, long long yns,
                double* ssum, double* ssq, int stat_c0, cudaStream_t st);
int conv_dgrad_tc(const tamgcn_conv_geom* g, const Opnd& dy, const void* wpack, void* dx, long long dxns,
                  const void* addend, long long addns, const float* bcast, float bscale, const Opnd* mask, double* s1,
                  double* s2, cudaStream_t st);
int conv_wgrad_tc(const tamgcn_conv_geom* g, const Opnd& dy, const Opnd& x, float* dW, float* dbias, cudaStream_t st);
int conv_wgrad_tc2(const tamgcn_conv_geom* g, const Opnd& dy, const Opnd& x, float* dW, float* dbias, cudaStream_t st);
// warp-level MMA kernels for the small-channel temporal convolutions, tconv_mma.cu (kind: 0 forward, 1 data gradient)
size_t conv_pack_t9_offset(int Cout, int Cin, int k, int dgrad);
int tconv9_launch(int mode, int N, int Cin, int Cout, int T, int To, int V, int k, int stride, int dil, int pad, const Opnd& in,
                  const void* wpack9, const float* bias, void* out, long long ons, const Opnd* mask, double* s1, double* s2,
                  int stat_c0, cudaStream_t st);
int tconv9_wgrad_launch(int N, int Cin, int Cout, int T, int To, int V, int k, int stride, int dil, int pad, const Opnd& dy,
                        const Opnd& x, float* dW, float* db, cudaStream_t st);
int tconv_mma_fwd_dgrad(const tamgcn_conv_geom* g, int kind, const Opnd& in, const float* W, const float* bias, void* out,
                        long long ons, const Opnd* mask, double* s1, double* s2, int stat_c0, cudaStream_t st);
int tconv_mma_wgrad(const tamgcn_conv_geom* g, const Opnd& dy, const Opnd& x, float* dW, float* db, cudaStream_t st);
}

// ------------------------------------------------------------------------------------------------
// 1x1 convolutions over a handful of positions per sample (the conv1 / conv2 branches of CTRGC act on the
// T-mean of x: L = V positions, models/ctrgcn.py:173): fp32, plain operands.  The 64 x 64 tiles above would be
// two thirds empty and walk K in 16-wide steps with two barriers each; here the sample's whole input sits in shared
// memory and every thread owns a few outputs.
// ------------------------------------------------------------------------------------------------
#define SL_OT 24      // output channels per CTA (forward)
#define SL_CT 32      // input channels per CTA (data gradient)

// y[n,o,l] = sum_c W[o,c] x[n,c,l] + b[o]
__global__ void __launch_bounds__(256)
conv1x1_smallL_fwd_kernel(int Cin, int Cout, int L, const float* __restrict__ x, long long xns, const float* __restrict__ W,
                          const float* __restrict__ bias, float* __restrict__ y, long long yns) {
    extern __shared__ __align__(16) float sl_xs[];             // [Cin][L]
    const int n = blockIdx.x, o0 = blockIdx.y * SL_OT;
    const float* xn = x + (long long)n * xns;
    for (int i = threadIdx.x; i < Cin * L; i += blockDim.x) sl_xs[i] = __ldg(xn + i);
    __syncthreads();
    const int no = min(SL_OT, Cout - o0);
    for (int idx = threadIdx.x; idx < no * L; idx += blockDim.x) {
        const int o = idx / L, l = idx - o * L;
        const float* w = W + (long long)(o0 + o) * Cin;
        float acc = bias ? __ldg(bias + o0 + o) : 0.f;
        int c = 0;
        if ((Cin & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
            for (; c < Cin; c += 4) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(w + c));
                acc = fmaf(wv.x, sl_xs[c * L + l], acc);
                acc = fmaf(wv.y, sl_xs[(c + 1) * L + l], acc);
                acc = fmaf(wv.z, sl_xs[(c + 2) * L + l], acc);
                acc = fmaf(wv.w, sl_xs[(c + 3) * L + l], acc);
            }
        }
        for (; c < Cin; ++c) acc = fmaf(__ldg(w + c), sl_xs[c * L + l], acc);
        y[(long long)n * yns + (long long)(o0 + o) * L + l] = acc;
    }
}

// dx[n,c,l] = sum_o W[o,c] dy[n,o,l]
__global__ void __launch_bounds__(256)
conv1x1_smallL_dgrad_kernel(int Cin, int Cout, int L, const float* __restrict__ dy, long long dyns,
                            const float* __restrict__ W, float* __restrict__ dx, long long dxns) {
    extern __shared__ __align__(16) float sl_ds[];             // [Cout][L]
    const int n = blockIdx.x, c0 = blockIdx.y * SL_CT;
    const float* dn = dy + (long long)n * dyns;
    for (int i = threadIdx.x; i < Cout * L; i += blockDim.x) sl_ds[i] = __ldg(dn + i);
    __syncthreads();
    const int nc = min(SL_CT, Cin - c0);
    for (int idx = threadIdx.x; idx < nc * L; idx += blockDim.x) {
        const int l = idx / nc, c = idx - l * nc;             // consecutive threads: consecutive c (coalesced rows of W)
        const float* w = W + c0 + c;
        float acc = 0.f;
#pragma unroll 4
        for (int o = 0; o < Cout; ++o) acc = fmaf(__ldg(w + (long long)o * Cin), sl_ds[o * L + l], acc);
        dx[(long long)n * dxns + (long long)(c0 + c) * L + l] = acc;
    }
}

static inline bool opnd_is_plain(const Opnd& o) { return !o.a && !o.b && !o.c && !o.q && !o.relu; }
static inline bool small_l_geom(const ConvP& p) {
    return p.k == 1 && p.s == 1 && p.p == 0 && p.T == p.To && p.T * p.V <= 32 && p.N <= 65535;
}

extern "C" int tamgcn_conv_fwd(const tamgcn_conv_geom* g, int dtype, const tamgcn_operand* x, const float* W,
                               const void* wpack, const float* bias, void* y, int64_t y_nstride, double* stat_sum,
                               double* stat_sumsq, int stat_c0, tamgcn_stream stream) {
    if (check_geom(g)) return -1;
    TG_REQUIRE(x && x->p && W && y, "conv_fwd: null pointer");
    TG_REQUIRE((stat_sum == nullptr) == (stat_sumsq == nullptr), "conv_fwd: stat_sum/stat_sumsq must both be set");
    const ConvP p = to_p(g);
    const Opnd xo = make_opnd(x);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TAMGCN_BF16) {
        int rc = tconv9_launch(0, p.N, p.Cin, p.Cout, p.T, p.To, p.V, p.k, p.s, p.d, p.p, xo,
                               wpack ? (const unsigned char*)wpack + conv_pack_t9_offset(p.Cout, p.Cin, p.k, 0) : nullptr, bias, y,
                               y_nstride, nullptr, stat_sum, stat_sumsq, stat_c0, st);
        if (rc != 0) return rc < 0 ? rc : 0;
        rc = tconv_mma_fwd_dgrad(g, 0, xo, W, bias, y, y_nstride, nullptr, stat_sum, stat_sumsq, stat_c0, st);
        if (rc != 0) return rc < 0 ? rc : 0;
        rc = conv_fwd_tc(g, xo, wpack, bias, y, y_nstride, stat_sum, stat_sumsq, stat_c0, st);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    if (dtype == TAMGCN_F32 && small_l_geom(p) && opnd_is_plain(xo) && !stat_sum && (size_t)p.Cin * p.T * p.V * 4 <= 48 * 1024) {
        const int L = p.T * p.V;
        conv1x1_smallL_fwd_kernel<<<dim3(p.N, cdiv(p.Cout, SL_OT)), 256, (size_t)p.Cin * L * 4, st>>>(
            p.Cin, p.Cout, L, (const float*)xo.p, xo.pns, W, bias, (float*)y, y_nstride);
        count_launch();
        return check_launch("conv_fwd(small L)");
    }
    dim3 grid(cdiv((long long)p.To * p.V, TN), cdiv(p.Cout, TM), p.N);
    if (dtype == TAMGCN_F32) {
        conv_fwd_kernel<float><<<grid, 256, 0, st>>>(p, xo, W, bias, (float*)y, y_nstride, stat_sum, stat_sumsq, stat_c0);
    } else if (dtype == TAMGCN_BF16) {
        conv_fwd_kernel<bf16><<<grid, 256, 0, st>>>(p, xo, W, bias, (bf16*)y, y_nstride, stat_sum, stat_sumsq, stat_c0);
    } else {
        return set_error("conv_fwd: bad dtype %d", dtype);
    }
    count_launch();
    return check_launch("conv_fwd");
}

extern "C" int tamgcn_conv_dgrad(const tamgcn_conv_geom* g, int dtype, const tamgcn_operand* dy, const float* W,
                                 const void* wpack, void* dx, int64_t dx_nstride, const void* addend, int64_t addend_nstride,
                                 const float* bcast, float bcast_scale, const tamgcn_operand* mask, double* s1,
                                 double* s2, tamgcn_stream stream) {
    if (check_geom(g)) return -1;
    TG_REQUIRE(dy && dy->p && W && dx, "conv_dgrad: null pointer");
    TG_REQUIRE((s1 == nullptr) == (s2 == nullptr), "conv_dgrad: s1/s2 must both be set");
    TG_REQUIRE(!(s1 && !mask), "conv_dgrad: statistics need a mask operand");
    TG_REQUIRE(!mask || mask->p, "conv_dgrad: mask operand without tensor");
    const ConvP p = to_p(g);
    const Opnd dyo = make_opnd(dy);
    Opnd mo = plain_opnd(nullptr, 0);
    if (mask) mo = make_opnd(mask);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TAMGCN_BF16) {
        int rc = (!addend && !bcast)
                     ? tconv9_launch(1, p.N, p.Cin, p.Cout, p.T, p.To, p.V, p.k, p.s, p.d, p.p, dyo,
                                     wpack ? (const unsigned char*)wpack + conv_pack_t9_offset(p.Cout, p.Cin, p.k, 1) : nullptr,
                                     nullptr, dx, dx_nstride, mask ? &mo : nullptr, s1, s2, 0, st)
                     : 0;
        if (rc != 0) return rc < 0 ? rc : 0;
        rc = (!addend && !bcast) ? tconv_mma_fwd_dgrad(g, 1, dyo, W, nullptr, dx, dx_nstride, mask ? &mo : nullptr, s1, s2, 0, st) : 0;
        if (rc != 0) return rc < 0 ? rc : 0;
        rc = conv_dgrad_tc(g, dyo, wpack, dx, dx_nstride, addend, addend_nstride, bcast, bcast_scale,
                           mask ? &mo : nullptr, s1, s2, st);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    if (dtype == TAMGCN_F32 && small_l_geom(p) && opnd_is_plain(dyo) && !addend && !bcast && !mask && !s1 &&
        (size_t)p.Cout * p.T * p.V * 4 <= 48 * 1024) {
        const int L = p.T * p.V;
        conv1x1_smallL_dgrad_kernel<<<dim3(p.N, cdiv(p.Cin, SL_CT)), 256, (size_t)p.Cout * L * 4, st>>>(
            p.Cin, p.Cout, L, (const float*)dyo.p, dyo.pns, W, (float*)dx, dx_nstride);
        count_launch();
        return check_launch("conv_dgrad(small L)");
    }
    dim3 grid(cdiv((long long)p.T * p.V, TN), cdiv(p.Cin, TM), p.N);
    if (dtype == TAMGCN_F32) {
        conv_dgrad_kernel<float><<<grid, 256, 0, st>>>(p, dyo, W, (float*)dx, dx_nstride, (const float*)addend,
                                                        addend_nstride, bcast, bcast_scale, mo, mask != nullptr, s1, s2);
    } else if (dtype == TAMGCN_BF16) {
        conv_dgrad_kernel<bf16><<<grid, 256, 0, st>>>(p, dyo, W, (bf16*)dx, dx_nstride, (const bf16*)addend,
                                                       addend_nstride, bcast, bcast_scale, mo, mask != nullptr, s1, s2);
    } else {
        return set_error("conv_dgrad: bad dtype %d", dtype);
    }
    count_launch();
    return check_launch("conv_dgrad");
}

extern "C" int tamgcn_conv_wgrad(const tamgcn_conv_geom* g, int dtype, const tamgcn_operand* dy,
                                 const tamgcn_operand* x, float* dW, float* dbias, tamgcn_stream stream) {
    if (check_geom(g)) return -1;
    TG_REQUIRE(dy && dy->p && x && x->p && dW, "conv_wgrad: null pointer");
    const ConvP p = to_p(g);
    const Opnd dyo = make_opnd(dy), xo = make_opnd(x);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TAMGCN_BF16) {
        int rc = tconv9_wgrad_launch(p.N, p.Cin, p.Cout, p.T, p.To, p.V, p.k, p.s, p.d, p.p, dyo, xo, dW, dbias, st);   // tconv9.cu
        if (rc != 0) return rc < 0 ? rc : 0;
        rc = tconv_mma_wgrad(g, dyo, xo, dW, dbias, st);           // small-channel temporal convolutions (tconv_mma.cu)
        if (rc != 0) return rc < 0 ? rc : 0;
        rc = conv_wgrad_tc2(g, dyo, xo, dW, dbias, st);            // vector-access kernel (conv_wg2.cu)
        if (rc != 0) return rc < 0 ? rc : 0;
        rc = conv_wgrad_tc(g, dyo, xo, dW, dbias, st);             // element-granular fallback (conv_tc.cu)
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    const int gx = cdiv((long long)p.Cin * p.k, TN), gy = cdiv(p.Cout, TM);
    const int nblk = cdiv((long long)p.To * p.V, WG_BLK);
    const long long units = (long long)p.N * nblk;
    long long S = (148LL * 4 + gx * gy - 1) / (gx * gy);
    if (S > units) S = units;
    if (S < 1) S = 1;
    if (S > 65535) S = 65535;
    dim3 grid(gx, gy, (unsigned)S);
    if (dtype == TAMGCN_F32) {
        conv_wgrad_kernel<float><<<grid, 256, 0, st>>>(p, dyo, xo, dW, dbias, nblk);
    } else if (dtype == TAMGCN_BF16) {
        conv_wgrad_kernel<bf16><<<grid, 256, 0, st>>>(p, dyo, xo, dW, dbias, nblk);
    } else {
        return set_error("conv_wgrad: bad dtype %d", dtype);
    }
    count_launch();
    return check_launch("conv_wgrad");
}
