// ctrgc.cu — fused CTRGC channel-wise topology refinement (reference models/ctrgcn.py:172-177).
//
//   D_i[n,r,u,v] = tanh(x1_i[n,r,u] - x2_i[n,r,v])
//   Q_i[n,c,u,v] = alpha * (sum_r W4_i[c,r] D_i[n,r,u,v] + b4_i[c]) + PA_i[u,v]
//   y[n,c,t,u]   = sum_i sum_v Q_i[n,c,u,v] * x3_i[n,c,t,v]
//
// One CTA owns one sample n and a tile of CT output channels.  D and Q (the N x C x V x V topology
// tensor of the reference) exist only in shared memory; the CTA then streams the K*CT contiguous
// (T x V) planes of x3 once and writes the CT planes of y once.  The sum over the K subsets and the
// BatchNorm statistics of y are accumulated in the same pass.  The backward kernel recomputes D and
// Q, streams g and x3 once, writes dx3 once and reduces dQ in shared memory into dPA, dalpha,
// dW4, db4 and (through tanh') dx1, dx2.
//
// SIMT fp32 math for both storage types.  V is a template parameter (20 = NW-UCLA, 25 = NTU).
#include "common.cuh"
#include <cstdio>
#include <type_traits>
#include "rows.cuh"
#include <atomic>
#include <cstdlib>

namespace tamgcn {

// ---- mean over T (conv1/conv2 commute with it: conv(x).mean(-2) == conv(x.mean(-2))) --------------
template <typename T>
__global__ void __launch_bounds__(256)
mean_t_kernel(const T* __restrict__ x, long long xns, int C, int Tn, int V, float* __restrict__ m) {
    const int n = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 8 + w;
    if (c >= C) return;
    const T* px = x + (long long)n * xns + (long long)c * Tn * V;
    const float inv = 1.f / (float)Tn;
    for (int v = lane; v < V; v += 32) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int t = 0;
        for (; t + 3 < Tn; t += 4) {
            s0 += ldf<T>(px + (t + 0) * V + v);
            s1 += ldf<T>(px + (t + 1) * V + v);
            s2 += ldf<T>(px + (t + 2) * V + v);
            s3 += ldf<T>(px + (t + 3) * V + v);
        }
        for (; t < Tn; ++t) s0 += ldf<T>(px + t * V + v);
        m[((long long)n * C + c) * V + v] = ((s0 + s1) + (s2 + s3)) * inv;
    }
}

struct CtrgcP {
    int N, Cout, T, K, R, CT;
    long long x3ns, x12ns, yns;
};

// D_i -> shared (pitch DP); also used by the backward kernel.  FAST (bf16 activations): hardware tanh.approx (2^-11
// relative, below the bf16 rounding of everything downstream) instead of the ~30-instruction tanhf — every CTA of a
// sample rebuilds the same R*V*V table, which made this the whole cost of the kernel at R = 32.
template <int V, int DP, bool FAST>
__device__ __forceinline__ void build_D(float* Ds, float* xs, const float* __restrict__ x1, const float* __restrict__ x2, int R) {
    // x1 / x2 of the plane -> shared once; then one (r, u) row of V values per thread: no divisions or global loads per
    // element (every CTA of a sample rebuilds this table, so its cost is multiplied by Cout / CT)
    for (int idx = threadIdx.x; idx < R * V; idx += blockDim.x) {
        xs[idx] = __ldg(x1 + idx);
        xs[R * V + idx] = __ldg(x2 + idx);
    }
    __syncthreads();
    for (int task = threadIdx.x; task < R * V; task += blockDim.x) {
        const int r = task / V;
        const float a = xs[task];
        const float* b = xs + R * V + r * V;
        float* d = Ds + (size_t)task * DP;
#pragma unroll 5
        for (int v = 0; v < V; ++v) {
            const float z = a - b[v];
            float t;
            if (FAST) asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(z));
            else t = tanhf(z);
            d[v] = t;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256)
ctrgc_fwd_kernel(CtrgcP g, const T* __restrict__ x3, const float* __restrict__ x1, const float* __restrict__ x2,
                 const float* __restrict__ W4, const float* __restrict__ b4, const float* __restrict__ PA,
                 const float* __restrict__ alpha_p, T* __restrict__ y, double* ssum, double* ssq) {
    constexpr int VP = VPad<V>::VP, DP = VPad<V>::DP;
    extern __shared__ __align__(16) float smem[];
    const int CT = g.CT, R = g.R, K = g.K;
    float* Qs = smem;                          // [K][CT][V][VP]
    float* Ds = Qs + K * CT * V * VP;          // [R][V][DP]
    float* W4s = Ds + R * V * DP;              // [CT][R]
    float* st = W4s + CT * R;                  // [CT][2]
    float* xs = st + CT * 2;                   // [2][R][V]  x1 / x2 of the current plane
    const int n = blockIdx.y, c0 = blockIdx.x * CT;
    const int nc = min(CT, g.Cout - c0);
    const float alpha = __ldg(alpha_p);

    for (int idx = threadIdx.x; idx < K * CT * V * VP; idx += blockDim.x) Qs[idx] = 0.f;
    for (int idx = threadIdx.x; idx < CT * 2; idx += blockDim.x) st[idx] = 0.f;

    for (int i = 0; i < K; ++i) {
        __syncthreads();
        build_D<V, DP, !std::is_same<T, float>::value>(Ds, xs, x1 + (long long)n * g.x12ns + i * R * V, x2 + (long long)n * g.x12ns + i * R * V, R);
        if ((CT & 3) == 0 && (R & 3) == 0) {
            // W4 of the CTA's channels transposed to [r][CT]: a thread owns one (u, v) and FOUR channels, so every tanh value
            // it reads feeds four FMAs and the four weights arrive as one 16-byte shared-memory load (the scalar loop below
            // spent two loads per FMA and was 55 % of the kernel's instructions at R = 32)
            for (int idx = threadIdx.x; idx < CT * R; idx += blockDim.x) {
                const int r = idx / CT, c = idx - r * CT;
                W4s[idx] = c < nc ? __ldg(W4 + ((long long)i * g.Cout + c0 + c) * R + r) : 0.f;
            }
            __syncthreads();
            const int ncq = (nc + 3) >> 2;
            for (int idx = threadIdx.x; idx < ncq * V * V; idx += blockDim.x) {
                const int cq = idx / (V * V), rem = idx - cq * V * V, u = rem / V, v = rem - u * V;
                float acc[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] = (4 * cq + k < nc) ? __ldg(b4 + i * g.Cout + c0 + 4 * cq + k) : 0.f;
                const float* d = Ds + u * DP + v;
                const float* w = W4s + 4 * cq;
#pragma unroll 4
                for (int r = 0; r < R; ++r) {
                    const float dv = d[r * V * DP];
                    const float4 wv = *reinterpret_cast<const float4*>(w + r * CT);
                    acc[0] = fmaf(wv.x, dv, acc[0]); acc[1] = fmaf(wv.y, dv, acc[1]);
                    acc[2] = fmaf(wv.z, dv, acc[2]); acc[3] = fmaf(wv.w, dv, acc[3]);
                }
                const float pa = __ldg(PA + (i * V + u) * V + v);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (4 * cq + k < nc) Qs[((i * CT + 4 * cq + k) * V + u) * VP + v] = fmaf(alpha, acc[k], pa);
            }
        } else {
            for (int idx = threadIdx.x; idx < nc * R; idx += blockDim.x)
                W4s[idx] = __ldg(W4 + ((long long)i * g.Cout + c0) * R + idx);
            __syncthreads();
            for (int idx = threadIdx.x; idx < nc * V * V; idx += blockDim.x) {
                const int c = idx / (V * V), rem = idx - c * V * V, u = rem / V, v = rem - u * V;
                float acc = __ldg(b4 + i * g.Cout + c0 + c);
                const float* d = Ds + u * DP + v;
                const float* w = W4s + c * R;
                for (int r = 0; r < R; ++r) acc = fmaf(w[r], d[r * V * DP], acc);
                Qs[((i * CT + c) * V + u) * VP + v] = fmaf(alpha, acc, __ldg(PA + (i * V + u) * V + v));
            }
        }
    }
    __syncthreads();

    const int Tn = g.T;
    for (int row = threadIdx.x; row < nc * Tn; row += blockDim.x) {
        const int c = row / Tn, t = row - c * Tn;
        float acc[V];
#pragma unroll
        for (int u = 0; u < V; ++u) acc[u] = 0.f;
        for (int i = 0; i < K; ++i) {
            float xr[VP];
            load_row<T, V, VP>(x3 + (long long)n * g.x3ns + (((long long)i * g.Cout + c0 + c) * Tn + t) * V, xr);
            const float* q = Qs + (i * CT + c) * V * VP;
#pragma unroll
            for (int u = 0; u < V; ++u) {
                float s = acc[u];
#pragma unroll
                for (int v4 = 0; v4 < VP / 4; ++v4) {
                    const float4 qq = *reinterpret_cast<const float4*>(q + u * VP + 4 * v4);
                    s = fmaf(qq.x, xr[4 * v4], s);
                    s = fmaf(qq.y, xr[4 * v4 + 1], s);
                    s = fmaf(qq.z, xr[4 * v4 + 2], s);
                    s = fmaf(qq.w, xr[4 * v4 + 3], s);
                }
                acc[u] = s;
            }
        }
        float s = 0.f, q2 = 0.f;
#pragma unroll
        for (int u = 0; u < V; ++u) {
            acc[u] = rnd<T>(acc[u]);
            s += acc[u];
            q2 = fmaf(acc[u], acc[u], q2);
        }
        store_row<T, V>(y + (long long)n * g.yns + ((long long)(c0 + c) * Tn + t) * V, acc);
        if (ssum) {
            atomicAdd(&st[2 * c], s);
            atomicAdd(&st[2 * c + 1], q2);
        }
    }
    if (ssum) {
        __syncthreads();
        for (int c = threadIdx.x; c < nc; c += blockDim.x) {
            atomicAdd(ssum + c0 + c, (double)st[2 * c]);
            atomicAdd(ssq + c0 + c, (double)st[2 * c + 1]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256)
ctrgc_bwd_kernel(CtrgcP g, Opnd go, const T* __restrict__ x3, const float* __restrict__ x1,
                 const float* __restrict__ x2, const float* __restrict__ W4, const float* __restrict__ b4,
                 const float* __restrict__ PA, const float* __restrict__ alpha_p, T* __restrict__ dx3,
                 long long dx3ns, float* dx1, float* dx2, float* dW4, float* db4, float* dPA, float* dalpha) {
    constexpr int VP = VPad<V>::VP, DP = VPad<V>::DP;
    extern __shared__ __align__(16) float smem[];
    const int CT = g.CT, R = g.R, K = g.K, Tn = g.T;
    float* Qs = smem;                  // [CT][V][VP]
    float* P4s = Qs + CT * V * VP;     // [CT][V][DP]   W4.D + b4
    float* dQs = P4s + CT * V * DP;    // [CT][V][DP]
    float* Ds = dQs + CT * V * DP;     // [R][V][DP]
    float* W4s = Ds + R * V * DP;      // [CT][R]
    float* red = W4s + CT * R;         // [64]
    float* xs = red + 64;              // [2][R][V]  x1 / x2 of the current plane
    const int n = blockIdx.y, c0 = blockIdx.x * CT;
    const int nc = min(CT, g.Cout - c0);
    const float alpha = __ldg(alpha_p);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int TV = Tn * V;

    for (int idx = threadIdx.x; idx < CT * V * VP; idx += blockDim.x) Qs[idx] = 0.f;
    float dalpha_acc = 0.f;

    for (int i = 0; i < K; ++i) {
        __syncthreads();
        build_D<V, DP, !std::is_same<T, float>::value>(Ds, xs, x1 + (long long)n * g.x12ns + i * R * V, x2 + (long long)n * g.x12ns + i * R * V, R);
        for (int idx = threadIdx.x; idx < nc * R; idx += blockDim.x)
            W4s[idx] = __ldg(W4 + ((long long)i * g.Cout + c0) * R + idx);
        __syncthreads();
        for (int idx = threadIdx.x; idx < nc * V * V; idx += blockDim.x) {
            const int c = idx / (V * V), rem = idx - c * V * V, u = rem / V, v = rem - u * V;
            float acc = __ldg(b4 + i * g.Cout + c0 + c);
            const float* d = Ds + u * DP + v;
            const float* w = W4s + c * R;
            for (int r = 0; r < R; ++r) acc = fmaf(w[r], d[r * V * DP], acc);
            P4s[(c * V + u) * DP + v] = acc;
            Qs[(c * V + u) * VP + v] = fmaf(alpha, acc, __ldg(PA + (i * V + u) * V + v));
        }
        __syncthreads();

        // (a) dx3_i[c,t,v] = sum_u Q[c,u,v] g[c,t,u]
        for (int row = threadIdx.x; row < nc * Tn; row += blockDim.x) {
            const int c = row / Tn, t = row - c * Tn;
            const OpCoef cf = opnd_coef(go, c0 + c);
            const long long off = (long long)(c0 + c) * TV + t * V;
            float acc[VP];
#pragma unroll
            for (int v = 0; v < VP; ++v) acc[v] = 0.f;
            const float* q = Qs + c * V * VP;
#pragma unroll
            for (int u = 0; u < V; ++u) {
                const float gu = opnd_val<T>(go, cf, n, off + u);
#pragma unroll
                for (int v4 = 0; v4 < VP / 4; ++v4) {
                    const float4 qq = *reinterpret_cast<const float4*>(q + u * VP + 4 * v4);
                    acc[4 * v4] = fmaf(qq.x, gu, acc[4 * v4]);
                    acc[4 * v4 + 1] = fmaf(qq.y, gu, acc[4 * v4 + 1]);
                    acc[4 * v4 + 2] = fmaf(qq.z, gu, acc[4 * v4 + 2]);
                    acc[4 * v4 + 3] = fmaf(qq.w, gu, acc[4 * v4 + 3]);
                }
            }
            store_row<T, V>(dx3 + (long long)n * dx3ns + (((long long)i * g.Cout + c0 + c) * Tn + t) * V, acc);
        }
        // (b) dQ[c,u,v] = sum_t g[c,t,u] x3_i[c,t,v]     one thread per (c,u)
        for (int task = threadIdx.x; task < nc * V; task += blockDim.x) {
            const int c = task / V, u = task - c * V;
            const OpCoef cf = opnd_coef(go, c0 + c);
            const long long goff = (long long)(c0 + c) * TV + u;
            const T* px = x3 + (long long)n * g.x3ns + ((long long)i * g.Cout + c0 + c) * TV;
            float acc[VP];
#pragma unroll
            for (int v = 0; v < VP; ++v) acc[v] = 0.f;
            for (int t = 0; t < Tn; ++t) {
                const float gu = opnd_val<T>(go, cf, n, goff + t * V);
                float xr[VP];
                load_row<T, V, VP>(px + t * V, xr);
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] = fmaf(gu, xr[v], acc[v]);
            }
#pragma unroll
            for (int v = 0; v < V; ++v) dQs[(c * V + u) * DP + v] = acc[v];
        }
        __syncthreads();

        // dPA_i[u,v] += sum_c dQ ;  dalpha += sum dQ * P4
        for (int idx = threadIdx.x; idx < V * V; idx += blockDim.x) {
            const int u = idx / V, v = idx - u * V;
            float s = 0.f;
            for (int c = 0; c < nc; ++c) {
                const float dq = dQs[(c * V + u) * DP + v];
                s += dq;
                dalpha_acc = fmaf(dq, P4s[(c * V + u) * DP + v], dalpha_acc);
            }
            atomicAdd(dPA + (i * V + u) * V + v, s);
        }
        // db4_i[c] += alpha * sum_uv dQ ;  dW4_i[c,r] += alpha * sum_uv dQ[c,u,v] D[r,u,v]   (warp per task)
        for (int task = warp; task < nc * (R + 1); task += nwarp) {
            const int c = task / (R + 1), r = task - c * (R + 1);
            float s = 0.f;
            for (int idx = lane; idx < V * V; idx += 32) {
                const int u = idx / V, v = idx - u * V;
                const float dq = dQs[(c * V + u) * DP + v];
                s += (r < R) ? dq * Ds[(r * V + u) * DP + v] : dq;
            }
            s = warp_sum(s);
            if (lane == 0) {
                if (r < R) atomicAdd(dW4 + ((long long)i * g.Cout + c0 + c) * R + r, alpha * s);
                else atomicAdd(db4 + i * g.Cout + c0 + c, alpha * s);
            }
        }
        // dD[r,u,v] = alpha sum_c W4[c,r] dQ[c,u,v];  dS = dD (1 - D^2);  dx1[r,u] += sum_v dS;  dx2[r,v] -= sum_u dS
        for (int task = threadIdx.x; task < 2 * R * V; task += blockDim.x) {
            const int which = task / (R * V), rem = task - which * R * V, r = rem / V, w = rem - r * V;
            float s = 0.f;
            for (int o = 0; o < V; ++o) {
                const int u = which ? o : w, v = which ? w : o;
                float dd = 0.f;
                for (int c = 0; c < nc; ++c) dd = fmaf(W4s[c * R + r], dQs[(c * V + u) * DP + v], dd);
                const float dv = Ds[(r * V + u) * DP + v];
                s = fmaf(dd, 1.f - dv * dv, s);
            }
            s *= alpha;
            float* dst = (which ? dx2 : dx1) + (long long)n * g.x12ns + (i * R + r) * V + w;
            atomicAdd(dst, which ? -s : s);
        }
    }
    // dalpha
    float dv[1] = {dalpha_acc};
    block_sum<1>(dv, red);
    if (threadIdx.x == 0) atomicAdd(dalpha, dv[0]);
}

// ------------------------------------------------------------------------------------------------
// backward, bf16 storage: the two contractions over (t) and (u) run on the tensor cores (warp-level mma.sync, fp32
// accumulation) with every operand fragment loaded straight from global memory in its natural row layout and
// re-distributed inside the warp with movmatrix; one warp owns one channel.  The topology tensors D, Q, dQ still live
// in shared memory only and the parameter reductions are unchanged.  (The forward kernel, ctrgc_tc.cu, is the
// tcgen05/TMEM one; this kernel is bound by the fp32 SIMT reductions after the two GEMMs.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movm_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// two consecutive bf16 at element offset off (4-byte access when aligned), zero when !ok / second element when !ok1
__device__ __forceinline__ uint32_t ld_pair(const bf16* __restrict__ p, long long off, bool ok, bool ok1) {
    if (!ok) return 0u;
    if (ok1 && ((off & 1) == 0) && ((reinterpret_cast<uintptr_t>(p) & 3) == 0)) return __ldg(reinterpret_cast<const unsigned*>(p + off));
    const unsigned lo = __ldg(reinterpret_cast<const unsigned short*>(p + off));
    const unsigned hi = ok1 ? __ldg(reinterpret_cast<const unsigned short*>(p + off + 1)) : 0u;
    return lo | (hi << 16);
}
__device__ __forceinline__ void st_pair(bf16* __restrict__ p, long long off, uint32_t w, bool ok, bool ok1) {
    if (!ok) return;
    if (ok1 && ((off & 1) == 0) && ((reinterpret_cast<uintptr_t>(p) & 3) == 0)) { *reinterpret_cast<unsigned*>(p + off) = w; return; }
    reinterpret_cast<unsigned short*>(p)[off] = (unsigned short)(w & 0xffffu);
    if (ok1) reinterpret_cast<unsigned short*>(p)[off + 1] = (unsigned short)(w >> 16);
}

__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];"
                 : "=r"(r0), "=r"(r1) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#define CBM_CT 16            // channels per CTA (one warp each)
#define CBM_THREADS 512

// Every contraction of the backward runs on warp-level tensor-core MMAs (m16n8k16, bf16 in, fp32 accumulate):
//   P4[c,uv]   = sum_r W4[c,r] D[r,uv]          (-> Q = alpha (P4 + b4) + PA, stored transposed as the B operand of)
//   dx3[t,v]   = sum_u g[t,u] Q[u,v]            per channel, fragments straight from global memory
//   dQ[u,v]    = sum_t g[t,u] x3[t,v]           per channel, operands re-distributed with movmatrix
//   raw[c,r]   = sum_uv dQ[c,uv] D[r,uv]        (row R of D is all ones: raw[c,R] = sum_uv dQ)  -> dW4, db4, dalpha
//   dD[r,uv]   = sum_c W4[c,r] dQ[c,uv]         -> dS = alpha dD (1 - D^2) -> dx1, dx2
// LEAN (large R: the full tables do not fit 227 KB): the fp32 tanh table is not kept — the last phase recomputes
// tanh from x1/x2 (same instruction, same value) and stages dS for 16 rows of r at a time.
__device__ int g_cbm_dbg = 0;        // build with -DTAMGCN_CBM_PHASES, run with TAMGCN_CBM_DBG=1: block (0,0) prints its cycles per phase

// CT channels per CTA, one warp each: 16 (512 threads) or 8 (256 threads: half the shared memory and registers per
// CTA, so two or three CTAs share an SM and one CTA's barrier / load latencies hide behind the others' work)
template <int V, bool LEAN, int CT>
__global__ void __launch_bounds__(CT * 32, CT == 8 ? 2 : 1)
ctrgc_bwd_mma_kernel(CtrgcP g, Opnd go, const bf16* __restrict__ x3, const float* __restrict__ x1,
                     const float* __restrict__ x2, const float* __restrict__ W4, const float* __restrict__ b4,
                     const float* __restrict__ PA, const float* __restrict__ alpha_p, bf16* __restrict__ dx3,
                     long long dx3ns, float* dx1, float* dx2, float* dW4, float* db4, float* dPA, float* dalpha) {
    constexpr int THR = CT * 32;
    constexpr int UV = V * V, UVp = (UV + 15) & ~15, UP = UVp + 8;      // uv = u*V + v; bf16 row pitch (conflict-free)
    constexpr int NTn = (V + 7) / 8;          // 8-wide v tiles (20 -> 3, 25 -> 4)
    constexpr int QP = 40;                    // pitch (bf16) of a Qt row: 32 u + 8 padding
    extern __shared__ __align__(16) float smem[];
    const int R = g.R, K = g.K, Tn = g.T;
    const int Rp = (R + 15) & ~15, RW = Rp + 8;          // K extent of the r contraction, pitch of W4b rows
    const int NR = R + 1, NRt = (NR + 7) / 8, DR = NRt * 8 > Rp ? NRt * 8 : Rp;   // rows of Db (R data rows, ones row, zero rows)
    float* Df = smem;                                   // [R][UVp]     tanh table, later dS   (LEAN: [16][UVp], dS only)
    float* dQs = Df + (size_t)(LEAN ? 16 : R) * UVp;    // [CT][UV]     fp32 dQ (for dPA)
    float* b4s = dQs + CT * UV;                         // [CT]
    float* x12s = b4s + CT;                             // [2][R*V]
    float* red = x12s + 2 * R * V;                      // [64]
    bf16* Db = reinterpret_cast<bf16*>(red + 64);       // [DR][UP]     D (bf16), ones row, zero rows
    bf16* dQc = Db + (size_t)DR * UP;                   // [16][UP]     dQ (bf16)
    bf16* W4b = dQc + 16 * UP;                          // [16][RW]     W4[c][r]
    bf16* W4T = W4b + 16 * RW;                          // [Rp][24]     W4[c][r] transposed
    bf16* Qt = W4T + Rp * 24;                           // [CT][8*NTn][QP]   Qt[c][v][u] = Q[c][u][v]
    const int n = blockIdx.y, c0 = blockIdx.x * CT;
    const int nc = min(CT, g.Cout - c0);
    const float alpha = __ldg(alpha_p);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const long long TV = (long long)Tn * V;
    const bf16 zero = __float2bfloat16_rn(0.f), one = __float2bfloat16_rn(1.f);

    for (int idx = tid; idx < CT * 8 * NTn * QP; idx += THR) Qt[idx] = zero;
    // Db: zero everywhere (16-byte stores; DR * UP * 2 bytes is a multiple of 16), then the ones row
    for (int idx = tid; idx < DR * UP / 8; idx += THR) reinterpret_cast<uint4*>(Db)[idx] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    for (int idx = tid; idx < UV; idx += THR) Db[R * UP + idx] = one;
    float dalpha_acc = 0.f;
#ifdef TAMGCN_CBM_PHASES
    long long tph[6] = {0, 0, 0, 0, 0, 0}, tmark = clock64();
#define CBM_MARK(ix) do { if (g_cbm_dbg) { const long long t_ = clock64(); tph[ix] += t_ - tmark; tmark = t_; } } while (0)
#else
#define CBM_MARK(ix) do { } while (0)
#endif

    for (int i = 0; i < K; ++i) {
        __syncthreads();
        // ---- parameters of this subset and the tanh table ----
        for (int idx = tid; idx < R * V; idx += THR) {
            x12s[idx] = __ldg(x1 + (long long)n * g.x12ns + i * R * V + idx);
            x12s[R * V + idx] = __ldg(x2 + (long long)n * g.x12ns + i * R * V + idx);
        }
        for (int idx = tid; idx < 16 * RW; idx += THR) {
            const int c = idx / RW, r = idx - c * RW;
            W4b[idx] = (c < nc && r < R) ? __float2bfloat16_rn(__ldg(W4 + ((long long)i * g.Cout + c0 + c) * R + r)) : zero;
        }
        for (int idx = tid; idx < Rp * 24; idx += THR) {
            const int r = idx / 24, c = idx - r * 24;
            W4T[idx] = (c < nc && r < R) ? __float2bfloat16_rn(__ldg(W4 + ((long long)i * g.Cout + c0 + c) * R + r)) : zero;
        }
        for (int idx = tid; idx < CT; idx += THR) b4s[idx] = idx < nc ? __ldg(b4 + i * g.Cout + c0 + idx) : 0.f;
        for (int idx = tid; idx < 16 * UP; idx += THR) dQc[idx] = zero;
        __syncthreads();
        if (V % 4 == 0) {
            // four consecutive v of one (r, u) per step: one x1 value, one 16-byte x2 load, vector stores
            constexpr int QR = UVp / 4, QU = V / 4;
            for (int idx = tid; idx < R * QR; idx += THR) {
                const int r = idx / QR, q = idx - r * QR, u = q / QU, v = 4 * (q - u * QU);
                const float a = x12s[r * V + u];
                const float4 b = *reinterpret_cast<const float4*>(x12s + R * V + r * V + v);
                const float4 d = make_float4(tanh_fast(a - b.x), tanh_fast(a - b.y), tanh_fast(a - b.z), tanh_fast(a - b.w));
                if (!LEAN) *reinterpret_cast<float4*>(Df + (size_t)r * UVp + 4 * q) = d;
                *reinterpret_cast<uint2*>(Db + (size_t)r * UP + 4 * q) = make_uint2(pack2_bf16(d.x, d.y), pack2_bf16(d.z, d.w));
            }
        } else {
            for (int idx = tid; idx < R * UVp; idx += THR) {
                const int r = idx / UVp, uv = idx - r * UVp;
                float d = 0.f;
                if (uv < UV) {
                    const int u = uv / V, v = uv - u * V;
                    d = tanh_fast(x12s[r * V + u] - x12s[R * V + r * V + v]);
                }
                if (!LEAN) Df[idx] = d;
                Db[r * UP + uv] = __float2bfloat16_rn(d);
            }
        }
        __syncthreads();
        CBM_MARK(0);
        // ---- Q = alpha (W4 . D + b4) + PA, as Qt[c][v][u] (bf16) ----
        for (int nt = warp; nt < UVp / 8; nt += THR / 32) {
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            // PA[i][uv], PA[i][uv+1] of this lane's accumulator columns: requested before the MMAs so that the global
            // load latency is hidden behind them (it used to stall the scatter below: 10 % of the kernel's samples)
            const int uv0 = nt * 8 + 2 * tig;
            const float pa0 = uv0 < UV ? __ldg(PA + i * UV + uv0) : 0.f;
            const float pa1 = uv0 + 1 < UV ? __ldg(PA + i * UV + uv0 + 1) : 0.f;
            for (int ks = 0; ks < Rp / 16; ++ks) {
                const bf16* wa = W4b + gid * RW + ks * 16 + 2 * tig;
                const uint32_t a[4] = {*reinterpret_cast<const uint32_t*>(wa), *reinterpret_cast<const uint32_t*>(wa + 8 * RW),
                                       *reinterpret_cast<const uint32_t*>(wa + 8), *reinterpret_cast<const uint32_t*>(wa + 8 * RW + 8)};
                uint32_t b0, b1;
                ldsm_x2_trans(b0, b1, Db + (size_t)(ks * 16 + (lane & 15)) * UP + nt * 8);
                mma_bf16_16816(d, a, b0, b1);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = gid + 8 * (e >> 1), uv = nt * 8 + 2 * tig + (e & 1);
                if (c < nc && uv < UV) {
                    const int u = uv / V, v = uv - u * V;
                    Qt[(c * 8 * NTn + v) * QP + u] = __float2bfloat16_rn(fmaf(alpha, d[e] + b4s[c], (e & 1) ? pa1 : pa0));
                }
            }
        }
        __syncthreads();
        CBM_MARK(1);

        // ---- per channel (one warp each): dx3 and dQ ----
        if (warp < nc) {
            const int c = warp;
            const OpCoef cf = opnd_coef(go, c0 + c);
            const bf16* gp = (const bf16*)go.p + (long long)n * go.pns + (long long)(c0 + c) * TV;
            const bf16* gq = go.q ? (const bf16*)go.q + (long long)n * go.qns + (long long)(c0 + c) * TV : nullptr;
            const bf16* xp = x3 + (long long)n * g.x3ns + ((long long)i * g.Cout + c0 + c) * TV;
            bf16* dxp = dx3 + (long long)n * dx3ns + ((long long)i * g.Cout + c0 + c) * TV;
            uint32_t bq[2][NTn][2];
            const bf16* qt = Qt + (size_t)c * 8 * NTn * QP;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int nt = 0; nt < NTn; ++nt) {
                    const bf16* q0 = qt + (nt * 8 + gid) * QP + ks * 16 + 2 * tig;
                    bq[ks][nt][0] = *reinterpret_cast<const uint32_t*>(q0);
                    bq[ks][nt][1] = *reinterpret_cast<const uint32_t*>(q0 + 8);
                }
            float dq[2][NTn][4];
#pragma unroll
            for (int mu = 0; mu < 2; ++mu)
#pragma unroll
                for (int nt = 0; nt < NTn; ++nt) dq[mu][nt][0] = dq[mu][nt][1] = dq[mu][nt][2] = dq[mu][nt][3] = 0.f;

            // 4-byte aligned planes: every fragment word is one aligned 32-bit load (even V), or — odd V with an even number
            // of rows, where odd rows start 2 bytes into a word — two aligned loads and a funnel shift whose amount is a
            // per-lane constant (row parity = parity of gid); the last row is odd, so no word reaches past the plane
            const bool al4 = ((reinterpret_cast<uintptr_t>(gp) | reinterpret_cast<uintptr_t>(xp) |
                               (gq ? reinterpret_cast<uintptr_t>(gq) : 0)) & 3) == 0;
            const bool oddld = (V % 2 == 1) && al4 && (Tn % 2 == 0);
            const bool fastld = ((V % 2 == 0) && al4) || oddld;
            const int odd = oddld ? (gid & 1) : 0, fsh = odd * 16;
            for (int t0 = 0; t0 < Tn; t0 += 16) {
                uint32_t ga[2][4];                       // cotangent, [row half][u block], lazy operand applied
                uint32_t xa[2][NTn];                     // x3 rows, [row half][v block]
                if (fastld) {
                    // even V, 4-byte aligned rows: every fragment word is one aligned 32-bit load.  All loads of the time
                    // block are issued before the first use (the element-wise path below serialises load -> use).
                    uint32_t pw[2][4], qw[2][4];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int t = t0 + gid + 8 * h;
                        const bool tok = t < Tn;
                        const long long ro = (long long)t * V + 2 * tig;
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const bool ok = tok && 8 * b + 2 * tig < V, ok1 = odd && 8 * b + 2 * tig + 1 < V;   // second word: only if its element is real
                            const unsigned* wp = reinterpret_cast<const unsigned*>(gp + ro + 8 * b - odd);
                            const unsigned w0 = ok ? __ldg(wp) : 0u, w1 = (ok && ok1) ? __ldg(wp + 1) : 0u;
                            pw[h][b] = (V % 2) ? __funnelshift_r(w0, w1, fsh) : w0;
                            if (gq) {
                                const unsigned* wq = reinterpret_cast<const unsigned*>(gq + ro + 8 * b - odd);
                                const unsigned q0 = ok ? __ldg(wq) : 0u, q1 = (ok && ok1) ? __ldg(wq + 1) : 0u;
                                qw[h][b] = (V % 2) ? __funnelshift_r(q0, q1, fsh) : q0;
                            } else {
                                qw[h][b] = 0u;
                            }
                        }
#pragma unroll
                        for (int b = 0; b < NTn; ++b) {
                            const bool ok = tok && 8 * b + 2 * tig < V, ok1 = odd && 8 * b + 2 * tig + 1 < V;
                            const unsigned* wx = reinterpret_cast<const unsigned*>(xp + ro + 8 * b - odd);
                            const unsigned x0 = ok ? __ldg(wx) : 0u, x1 = (ok && ok1) ? __ldg(wx + 1) : 0u;
                            unsigned xw = (V % 2) ? __funnelshift_r(x0, x1, fsh) : x0;
                            if ((V % 2) && 8 * b + 2 * tig + 1 >= V) xw &= 0xffffu;
                            xa[h][b] = xw;
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const bool ok = (t0 + gid + 8 * h < Tn) && 8 * b + 2 * tig < V;
                            float lo = fmaf(cf.a, __uint_as_float(pw[h][b] << 16), cf.c), hi = fmaf(cf.a, __uint_as_float(pw[h][b] & 0xffff0000u), cf.c);
                            if (gq) {
                                lo = fmaf(cf.b, __uint_as_float(qw[h][b] << 16), lo);
                                hi = fmaf(cf.b, __uint_as_float(qw[h][b] & 0xffff0000u), hi);
                            }
                            if (go.relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
                            if ((V % 2) && 8 * b + 2 * tig + 1 >= V) hi = 0.f;
                            ga[h][b] = ok ? pack2_bf16(lo, hi) : 0u;
                        }
                } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int t = t0 + gid + 8 * h;
                    const bool tok = t < Tn;
                    const long long ro = (long long)t * V;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int u = 8 * b + 2 * tig;
                        const bool ok = tok && u < V, ok1 = u + 1 < V;
                        uint32_t w = 0u;
                        if (ok) {
                            const uint32_t pw = ld_pair(gp, ro + u, true, ok1);
                            float lo = fmaf(cf.a, __uint_as_float(pw << 16), cf.c), hi = fmaf(cf.a, __uint_as_float(pw & 0xffff0000u), cf.c);
                            if (gq) {
                                const uint32_t qw = ld_pair(gq, ro + u, true, ok1);
                                lo = fmaf(cf.b, __uint_as_float(qw << 16), lo);
                                hi = fmaf(cf.b, __uint_as_float(qw & 0xffff0000u), hi);
                            }
                            if (go.relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
                            w = pack2_bf16(lo, ok1 ? hi : 0.f);
                        }
                        ga[h][b] = w;
                    }
#pragma unroll
                    for (int b = 0; b < NTn; ++b) {
                        const int v = 8 * b + 2 * tig;
                        xa[h][b] = ld_pair(xp, ro + v, tok && v < V, v + 1 < V);
                    }
                }
                }
#pragma unroll
                for (int nt = 0; nt < NTn; ++nt) {
                    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint32_t a[4] = {ga[0][2 * ks], ga[1][2 * ks], ga[0][2 * ks + 1], ga[1][2 * ks + 1]};
                        mma_bf16_16816(d, a, bq[ks][nt][0], bq[ks][nt][1]);
                    }
                    const int v = 8 * nt + 2 * tig;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int t = t0 + gid + 8 * h;
                        st_pair(dxp, (long long)t * V + v, pack2_bf16(d[2 * h], d[2 * h + 1]), t < Tn && v < V, v + 1 < V);
                    }
                }
                uint32_t bx[NTn][2];
#pragma unroll
                for (int nt = 0; nt < NTn; ++nt) { bx[nt][0] = movm_trans(xa[0][nt]); bx[nt][1] = movm_trans(xa[1][nt]); }
#pragma unroll
                for (int mu = 0; mu < 2; ++mu) {
                    const uint32_t a[4] = {movm_trans(ga[0][2 * mu]), movm_trans(ga[0][2 * mu + 1]), movm_trans(ga[1][2 * mu]),
                                           movm_trans(ga[1][2 * mu + 1])};
#pragma unroll
                    for (int nt = 0; nt < NTn; ++nt) mma_bf16_16816(dq[mu][nt], a, bx[nt][0], bx[nt][1]);
                }
            }
#pragma unroll
            for (int mu = 0; mu < 2; ++mu)
#pragma unroll
                for (int nt = 0; nt < NTn; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int u = 16 * mu + gid + 8 * (e >> 1), v = 8 * nt + 2 * tig + (e & 1);
                        if (u < V && v < V) {
                            dQs[c * UV + u * V + v] = dq[mu][nt][e];
                            dQc[c * UP + u * V + v] = __float2bfloat16_rn(dq[mu][nt][e]);
                        }
                    }
        }
        __syncthreads();
        CBM_MARK(2);

        // ---- dPA_i[u,v] += sum_c dQ ----
        for (int uv = tid; uv < UV; uv += THR) {
            float s = 0.f;
            for (int c = 0; c < nc; ++c) s += dQs[c * UV + uv];
            atomicAdd(dPA + i * UV + uv, s);
        }
        // ---- raw[c,r] = sum_uv dQ[c,uv] D[r,uv]  (r = R: ones row): one warp per 8 values of r ----
        if (warp < NRt) {
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            for (int ks = 0; ks < UVp / 16; ++ks) {
                const bf16* qa = dQc + gid * UP + ks * 16 + 2 * tig;
                const uint32_t a[4] = {*reinterpret_cast<const uint32_t*>(qa), *reinterpret_cast<const uint32_t*>(qa + 8 * UP),
                                       *reinterpret_cast<const uint32_t*>(qa + 8), *reinterpret_cast<const uint32_t*>(qa + 8 * UP + 8)};
                const bf16* db = Db + (size_t)(warp * 8 + gid) * UP + ks * 16 + 2 * tig;
                mma_bf16_16816(d, a, *reinterpret_cast<const uint32_t*>(db), *reinterpret_cast<const uint32_t*>(db + 8));
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = gid + 8 * (e >> 1), r = warp * 8 + 2 * tig + (e & 1);
                if (c < nc && r < R) {
                    atomicAdd(dW4 + ((long long)i * g.Cout + c0 + c) * R + r, alpha * d[e]);
                    dalpha_acc = fmaf(__bfloat162float(W4b[c * RW + r]), d[e], dalpha_acc);
                } else if (c < nc && r == R) {
                    atomicAdd(db4 + i * g.Cout + c0 + c, alpha * d[e]);
                    dalpha_acc = fmaf(b4s[c], d[e], dalpha_acc);
                }
            }
        }
        CBM_MARK(3);
        // ---- dD[r,uv] = sum_c W4[c,r] dQ[c,uv];  dS = alpha dD (1 - D^2) overwrites the tanh table ----
        // ---- dx1[r,u] += sum_v dS[r,u,v];  dx2[r,v] -= sum_u dS[r,u,v] ----
        const int MT = Rp / 16;
        for (int pass = 0; pass < (LEAN ? MT : 1); ++pass) {
            const int mt_lo = LEAN ? pass : 0, mt_hi = LEAN ? pass + 1 : MT;
            const int r_lo = mt_lo * 16, r_n = min(R, mt_hi * 16) - r_lo;       // rows staged in Df this pass
            for (int nt = warp; nt < UVp / 8; nt += THR / 32) {
                uint32_t b0, b1;
                ldsm_x2_trans(b0, b1, dQc + (size_t)(lane & 15) * UP + nt * 8);
                const int uv = nt * 8 + 2 * tig;
                int u0 = 0, v0 = 0, u1 = 0, v1 = 0;
                if (LEAN) { u0 = uv / V; v0 = uv - u0 * V; u1 = (uv + 1) / V; v1 = uv + 1 - u1 * V; }
                for (int mt = mt_lo; mt < mt_hi; ++mt) {
                    const bf16* wa = W4T + (mt * 16 + gid) * 24 + 2 * tig;
                    const uint32_t a[4] = {*reinterpret_cast<const uint32_t*>(wa), *reinterpret_cast<const uint32_t*>(wa + 8 * 24),
                                           *reinterpret_cast<const uint32_t*>(wa + 8), *reinterpret_cast<const uint32_t*>(wa + 8 * 24 + 8)};
                    float d[4] = {0.f, 0.f, 0.f, 0.f};
                    mma_bf16_16816(d, a, b0, b1);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = mt * 16 + gid + 8 * h;
                        if (r < R) {
                            float2* pd = reinterpret_cast<float2*>(Df + (size_t)(r - r_lo) * UVp + uv);
                            float2 dv;
                            if (LEAN) {
                                dv.x = uv < UV ? tanh_fast(x12s[r * V + u0] - x12s[R * V + r * V + v0]) : 0.f;
                                dv.y = uv + 1 < UV ? tanh_fast(x12s[r * V + u1] - x12s[R * V + r * V + v1]) : 0.f;
                            } else {
                                dv = *pd;
                            }
                            *pd = make_float2(alpha * d[2 * h] * (1.f - dv.x * dv.x), alpha * d[2 * h + 1] * (1.f - dv.y * dv.y));
                        }
                    }
                }
            }
            __syncthreads();
            CBM_MARK(4);
            if (V % 4 == 0) {
                // 16-byte shared-memory loads: a dx1 task sums one row of V values, a dx2 task four adjacent columns
                constexpr int QU = V / 4;
                for (int task = tid; task < r_n * V; task += THR) {
                    const int rl = task / V, u = task - rl * V, r = r_lo + rl;
                    const float4* ds = reinterpret_cast<const float4*>(Df + (size_t)rl * UVp + u * V);
                    float s = 0.f;
#pragma unroll
                    for (int o = 0; o < QU; ++o) { const float4 d = ds[o]; s += (d.x + d.y) + (d.z + d.w); }
                    atomicAdd(dx1 + (long long)n * g.x12ns + (i * R + r) * V + u, s);
                }
                for (int task = tid; task < r_n * QU; task += THR) {
                    const int rl = task / QU, vq = task - rl * QU, r = r_lo + rl;
                    const float* ds = Df + (size_t)rl * UVp + 4 * vq;
                    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
                    for (int o = 0; o < V; ++o) {
                        const float4 d = *reinterpret_cast<const float4*>(ds + o * V);
                        s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
                    }
                    float* dst = dx2 + (long long)n * g.x12ns + (i * R + r) * V + 4 * vq;
                    atomicAdd(dst, -s.x); atomicAdd(dst + 1, -s.y); atomicAdd(dst + 2, -s.z); atomicAdd(dst + 3, -s.w);
                }
            } else {
                for (int task = tid; task < 2 * r_n * V; task += THR) {
                    const int which = task / (r_n * V), rem = task - which * r_n * V, rl = rem / V, w = rem - rl * V;
                    const int r = r_lo + rl;
                    const float* ds = Df + (size_t)rl * UVp + (which ? w : w * V);
                    const int step = which ? V : 1;
                    float s = 0.f;
#pragma unroll 5
                    for (int o = 0; o < V; ++o) s += ds[o * step];
                    float* dst = (which ? dx2 : dx1) + (long long)n * g.x12ns + (i * R + r) * V + w;
                    atomicAdd(dst, which ? -s : s);
                }
            }
            if (LEAN) __syncthreads();
            CBM_MARK(5);
        }
    }
#ifdef TAMGCN_CBM_PHASES
    if (g_cbm_dbg && blockIdx.x == 0 && blockIdx.y == 0 && (tid == 0 || tid == 300))
        printf("ctrgc_bwd_mma tid %d: table %lld  Q %lld  channel loop %lld  dPA+raw %lld  dD/dS %lld  dx1/dx2 %lld\n", tid, tph[0], tph[1],
               tph[2], tph[3], tph[4], tph[5]);
#endif
#undef CBM_MARK
    float dv[1] = {dalpha_acc};
    block_sum<1>(dv, red);
    if (tid == 0) atomicAdd(dalpha, dv[0]);
}


// ------------------------------------------------------------------------------------------------
// forward on warp-level tensor-core MMAs (bf16): the shapes the tcgen05 kernels (ctrgc_tc*.cu) do not cover — large R,
// where their fp16 tanh table no longer fits next to the pipeline stages (NTU l6-l10: V = 25, R = 16 / 32).
//   Q_i[c,uv]  = alpha (sum_r W4_i[c,r] D_i[r,uv] + b4_i[c]) + PA_i[uv]     m16n8k16, M = 16 channels of a tile
//   y[t,u]     = sum_i sum_v x3_i[t,v] Q_i[u,v]                             per channel (one warp each)
// A CTA owns one sample and every S-th 16-channel tile: the tanh tables of all K subsets are built ONCE per CTA
// (bf16, shared memory) and reused by all its tiles.  Steps j = (tile, subset): the x3 rows of step j+1 stream into a
// per-warp staging buffer with cp.async while step j computes; with two Q buffers (NQ = 2) the Q build of step j+1
// and the contraction of step j share one barrier interval.
// ------------------------------------------------------------------------------------------------
struct CfmP {
    int S;            // channel-tile stride (gridDim.x)
    int NQ;           // Q buffers (1 or 2)
    int cpw;          // cp.async width in bytes for x3 (16 / 8 / 4)
    int yw;           // copy-out width in bytes for y (16 / 8 / 4 / 2)
    int xs_bytes;     // staging bytes per warp and buffer
};

__device__ __forceinline__ void cfm_cp_async(uint32_t dst, const void* src, int w) {
    if (w == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else if (w == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

#define CFM_KS 2      // R <= 32

template <int V, int TB>
__global__ void __launch_bounds__(CBM_THREADS, 1)
ctrgc_fwd_mma_kernel(CtrgcP g, CfmP f, const bf16* __restrict__ x3, const float* __restrict__ x1,
                     const float* __restrict__ x2, const float* __restrict__ W4, const float* __restrict__ b4,
                     const float* __restrict__ PA, const float* __restrict__ alpha_p, bf16* __restrict__ y, double* ssum,
                     double* ssq) {
    constexpr int CT = CBM_CT;
    constexpr int UV = V * V, UVp = (UV + 15) & ~15, UP = UVp + 8;
    constexpr int NTn = (V + 7) / 8;          // 8-wide u / v tiles (20 -> 3, 25 -> 4)
    constexpr int QP = 40;                    // pitch (bf16) of a Q row: 32 v + 8 padding (conflict-free fragment reads)
    constexpr int CQ = 8 * NTn * QP + 8;      // channel stride of a Q buffer (+8: the build's stores of 8 channels hit 8 bank groups)
    constexpr int QSZ = CT * CQ;              // one Q buffer (bf16 elements)
    extern __shared__ __align__(16) float smem[];
    const int R = g.R, K = g.K, Tn = g.T;
    const int Rp = (R + 15) & ~15, KS = Rp / 16;
    bf16* Db = reinterpret_cast<bf16*>(smem);                               // [K][Rp][UP]  tanh tables
    bf16* Qn = Db + (size_t)K * Rp * UP;                                    // [NQ][CT][CQ]  Q[c][u][v]
    unsigned char* xs = reinterpret_cast<unsigned char*>(Qn + (size_t)f.NQ * QSZ);   // [2][16 warps][xs_bytes]
    float* x12s = reinterpret_cast<float*>(xs + 2 * 16 * (size_t)f.xs_bytes);        // [2][R*V]  (table build only)
    const int n = blockIdx.y;
    const float alpha = __ldg(alpha_p);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const long long TV = (long long)Tn * V;
    const int nTiles = (g.Cout + CT - 1) / CT;
    const int myTiles = (nTiles - (int)blockIdx.x + f.S - 1) / f.S;
    const int J = myTiles * K;

    // padding rows / columns of Q and D stay zero: only valid entries are ever rewritten
    for (int idx = tid; idx < f.NQ * QSZ / 8; idx += CBM_THREADS) reinterpret_cast<uint4*>(Qn)[idx] = make_uint4(0u, 0u, 0u, 0u);
    for (int idx = tid; idx < K * Rp * UP / 8; idx += CBM_THREADS) reinterpret_cast<uint4*>(Db)[idx] = make_uint4(0u, 0u, 0u, 0u);

    // x3 rows of step j (this warp's channel) -> staging buffer j & 1
    const uint32_t xs_s = (uint32_t)__cvta_generic_to_shared(xs);
    const int row_bytes = (int)TV * 2;
    int sj_tile = (int)blockIdx.x, sj_i = 0;             // (tile, subset) of the next step to stage
    auto stage = [&](int j) {
        if (j < J) {
            const int c = sj_tile * CT + warp;
            if (c < g.Cout) {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(x3 + (long long)n * g.x3ns + ((long long)sj_i * g.Cout + c) * TV);
                const uint32_t dst = xs_s + (uint32_t)((j & 1) * 16 + warp) * (uint32_t)f.xs_bytes;
                for (int o = lane * f.cpw; o < row_bytes; o += 32 * f.cpw) cfm_cp_async(dst + o, src + o, f.cpw);
            }
            if (++sj_i == K) { sj_i = 0; sj_tile += f.S; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0);

    // ---- tanh tables of all subsets ----
    for (int i = 0; i < K; ++i) {
        __syncthreads();
        for (int idx = tid; idx < R * V; idx += CBM_THREADS) {
            x12s[idx] = __ldg(x1 + (long long)n * g.x12ns + i * R * V + idx);
            x12s[R * V + idx] = __ldg(x2 + (long long)n * g.x12ns + i * R * V + idx);
        }
        __syncthreads();
        bf16* Di = Db + (size_t)i * Rp * UP;
        for (int idx = tid; idx < R * (UVp / 2); idx += CBM_THREADS) {
            const int r = idx / (UVp / 2), uv = 2 * (idx - r * (UVp / 2));
            float d0 = 0.f, d1 = 0.f;
            if (uv < UV) { const int u = uv / V, v = uv - u * V; d0 = tanh_fast(x12s[r * V + u] - x12s[R * V + r * V + v]); }
            if (uv + 1 < UV) { const int u = (uv + 1) / V, v = uv + 1 - u * V; d1 = tanh_fast(x12s[r * V + u] - x12s[R * V + r * V + v]); }
            *reinterpret_cast<uint32_t*>(Di + (size_t)r * UP + uv) = pack2_bf16(d0, d1);
        }
    }
    __syncthreads();

    // A fragments of the Q build (W4 rows of a tile's 16 channels, bf16) straight from global memory, requested one
    // phase before their use
    uint32_t afr[CFM_KS][4];
    float b4a = 0.f, b4b = 0.f;
    int bq_tile = (int)blockIdx.x, bq_i = 0;             // (tile, subset) of the next Q build
    auto load_w4 = [&](int j) {
        if (j >= J) return;
        const int c0 = bq_tile * CT, ca = c0 + gid, cb = c0 + gid + 8;
        const bool oka = ca < g.Cout, okb = cb < g.Cout;
        const float* wa = W4 + ((long long)bq_i * g.Cout + ca) * R;
        const float* wb = W4 + ((long long)bq_i * g.Cout + cb) * R;
        b4a = oka ? __ldg(b4 + bq_i * g.Cout + ca) : 0.f;
        b4b = okb ? __ldg(b4 + bq_i * g.Cout + cb) : 0.f;
#pragma unroll
        for (int ks = 0; ks < CFM_KS; ++ks) {
            const int r0 = ks * 16 + 2 * tig;
            float w[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = r0 + (q & 1) + 8 * (q >> 1);
                w[q] = (oka && r < R) ? __ldg(wa + r) : 0.f;
                w[4 + q] = (okb && r < R) ? __ldg(wb + r) : 0.f;
            }
            afr[ks][0] = pack2_bf16(w[0], w[1]);
            afr[ks][1] = pack2_bf16(w[4], w[5]);
            afr[ks][2] = pack2_bf16(w[2], w[3]);
            afr[ks][3] = pack2_bf16(w[6], w[7]);
        }
    };
    // Q of the step whose fragments load_w4 fetched last, into buffer Qb
    auto build_q = [&](bf16* Qb) {
        const int i = bq_i, c0 = bq_tile * CT;
        if (++bq_i == K) { bq_i = 0; bq_tile += f.S; }
        const bf16* Di = Db + (size_t)i * Rp * UP;
        const bool oka = c0 + gid < g.Cout, okb = c0 + gid + 8 < g.Cout;
        bf16* qa = Qb + gid * CQ;
        bf16* qb = qa + 8 * CQ;
        const float* pa = PA + i * UV;
        // (u, v) of this lane's first accumulator column, advanced by 16 tiles = 128 columns per iteration
        int uv = warp * 8 + 2 * tig;
        int u = uv / V, v = uv - u * V;
        constexpr int DU = 128 / V, DV = 128 - DU * V;
        for (int nt = warp; nt < UVp / 8; nt += CBM_THREADS / 32) {
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            const float pa0 = uv < UV ? __ldg(pa + uv) : 0.f;
            const float pa1 = uv + 1 < UV ? __ldg(pa + uv + 1) : 0.f;
#pragma unroll
            for (int ks = 0; ks < CFM_KS; ++ks) {
                if (ks < KS) {
                    uint32_t b0, b1;
                    ldsm_x2_trans(b0, b1, Di + (size_t)(ks * 16 + (lane & 15)) * UP + nt * 8);
                    mma_bf16_16816(d, afr[ks], b0, b1);
                }
            }
            const int o0 = u * QP + v;
            int o1 = o0 + 1;
            if (v + 1 == V) o1 = o0 + QP - (V - 1);
            if (uv < UV) {
                if (oka) qa[o0] = __float2bfloat16_rn(fmaf(alpha, d[0] + b4a, pa0));
                if (okb) qb[o0] = __float2bfloat16_rn(fmaf(alpha, d[2] + b4b, pa0));
            }
            if (uv + 1 < UV) {
                if (oka) qa[o1] = __float2bfloat16_rn(fmaf(alpha, d[1] + b4a, pa1));
                if (okb) qb[o1] = __float2bfloat16_rn(fmaf(alpha, d[3] + b4b, pa1));
            }
            uv += 128; u += DU; v += DV;
            if (v >= V) { v -= V; ++u; }
        }
    };

    float acc[TB][NTn][4];
    float s1 = 0.f, s2 = 0.f;
    // element e = t*V + v of this lane's first fragment word (t = gid, v = 2*tig); every other word of the lane is an
    // even number of elements further, so the word offset and the half-word shift are per-lane constants
    const int e0 = gid * V + 2 * tig;
    const int sh = (V % 2) ? (e0 & 1) * 16 : 0;
    int ct_tile = (int)blockIdx.x, ct_i = 0;             // (tile, subset) of the next contraction
    // contraction of step j: x3 fragments from the staging buffer, Q fragments from Qb
    auto contract = [&](int j, const bf16* Qb) {
        const int i = ct_i, c = ct_tile * CT + warp;
        if (++ct_i == K) { ct_i = 0; ct_tile += f.S; }
        if (i == 0) {
#pragma unroll
            for (int tb = 0; tb < TB; ++tb)
#pragma unroll
                for (int nt = 0; nt < NTn; ++nt) acc[tb][nt][0] = acc[tb][nt][1] = acc[tb][nt][2] = acc[tb][nt][3] = 0.f;
        }
        if (c >= g.Cout) return;
        unsigned char* xb = xs + (size_t)((j & 1) * 16 + warp) * f.xs_bytes;
        const uint32_t* xw = reinterpret_cast<const uint32_t*>(xb) + (e0 >> 1);
        uint32_t xa[TB][2][4];
#pragma unroll
        for (int tb = 0; tb < TB; ++tb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool tok = tb * 16 + gid + 8 * h < Tn;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    constexpr int dummy = 0; (void)dummy;
                    const int wo = ((tb * 16 + 8 * h) * V + 8 * b) / 2;         // compile-time word offset
                    uint32_t w = 0u;
                    if (8 * b + 2 * tig < V && tok) {
                        if (V % 2 == 0) w = xw[wo];
                        else w = __funnelshift_r(xw[wo], xw[wo + 1], sh);
                        if (8 * b + 2 * tig + 1 >= V) w &= 0xffffu;
                    }
                    xa[tb][h][b] = w;
                }
            }
        const bf16* qt = Qb + (size_t)warp * CQ;
#pragma unroll
        for (int nt = 0; nt < NTn; ++nt) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const bf16* q0 = qt + (nt * 8 + gid) * QP + ks * 16 + 2 * tig;
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(q0), b1 = *reinterpret_cast<const uint32_t*>(q0 + 8);
#pragma unroll
                for (int tb = 0; tb < TB; ++tb) {
                    const uint32_t a[4] = {xa[tb][0][2 * ks], xa[tb][1][2 * ks], xa[tb][0][2 * ks + 1], xa[tb][1][2 * ks + 1]};
                    mma_bf16_16816(acc[tb][nt], a, b0, b1);
                }
            }
        }
        if (i == K - 1) {
            // the finished (channel, all t) block has the layout of the staging buffer: written there, copied out in
            // wide coalesced words
            __syncwarp();
            s1 = 0.f; s2 = 0.f;
            unsigned short* xh = reinterpret_cast<unsigned short*>(xb);
#pragma unroll
            for (int tb = 0; tb < TB; ++tb)
#pragma unroll
                for (int nt = 0; nt < NTn; ++nt) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const bool ok = (tb * 16 + gid + 8 * h < Tn) && 8 * nt + 2 * tig < V, ok1 = 8 * nt + 2 * tig + 1 < V;
                        const uint32_t w = pack2_bf16(acc[tb][nt][2 * h], acc[tb][nt][2 * h + 1]);
                        if (ok) {
                            const float lo = __uint_as_float(w << 16), hi = ok1 ? __uint_as_float(w & 0xffff0000u) : 0.f;
                            s1 += lo + hi;
                            s2 = fmaf(lo, lo, fmaf(hi, hi, s2));
                            const int e = e0 + (tb * 16 + 8 * h) * V + 8 * nt;
                            xh[e] = (unsigned short)(w & 0xffffu);
                            if (ok1) xh[e + 1] = (unsigned short)(w >> 16);
                        }
                    }
                }
            __syncwarp();
            unsigned char* yp = reinterpret_cast<unsigned char*>(y + (long long)n * g.yns + (long long)c * TV);
            if (f.yw == 16) {
                for (int o = lane * 16; o < row_bytes; o += 512) *reinterpret_cast<uint4*>(yp + o) = *reinterpret_cast<const uint4*>(xb + o);
            } else if (f.yw == 8) {
                for (int o = lane * 8; o < row_bytes; o += 256) *reinterpret_cast<uint2*>(yp + o) = *reinterpret_cast<const uint2*>(xb + o);
            } else if (f.yw == 4) {
                for (int o = lane * 4; o < row_bytes; o += 128) *reinterpret_cast<uint32_t*>(yp + o) = *reinterpret_cast<const uint32_t*>(xb + o);
            } else {
                for (int o = lane * 2; o < row_bytes; o += 64) *reinterpret_cast<unsigned short*>(yp + o) = *reinterpret_cast<const unsigned short*>(xb + o);
            }
            if (ssum) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
                if (lane == 0) { atomicAdd(ssum + c, (double)s1); atomicAdd(ssq + c, (double)s2); }
            }
        }
    };

    if (f.NQ == 2) {
        load_w4(0);
        build_q(Qn);
        __syncthreads();
        for (int j = 0; j < J; ++j) {
            stage(j + 1);
            load_w4(j + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            contract(j, Qn + (size_t)(j & 1) * QSZ);
            if (j + 1 < J) build_q(Qn + (size_t)((j + 1) & 1) * QSZ);
            __syncthreads();
        }
    } else {
        load_w4(0);
        for (int j = 0; j < J; ++j) {
            stage(j + 1);
            build_q(Qn);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncthreads();
            load_w4(j + 1);
            contract(j, Qn);
            __syncthreads();
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// returns 1 if launched, 0 if the shape does not fit (caller falls back to the SIMT kernel), <0 on error
static int launch_fwd_mma(const CtrgcP& g0, int V, const void* x3, const float* x1, const float* x2, const float* W4,
                          const float* b4, const float* PA, const float* alpha, void* y, double* ssum, double* ssq,
                          cudaStream_t st) {
    static const bool off = [] { const char* e = getenv("TAMGCN_DISABLE_FWD_MMA"); return e && e[0] == '1'; }();
    if (off) return 0;
    CtrgcP g = g0;
    g.CT = CBM_CT;
    if (g.T > 64 || g.R > 16 * CFM_KS) return 0;                       // accumulators of all T live in registers
    const long long TVb = (long long)g.T * V * 2;
    if (TVb % 4) return 0;
    const uintptr_t xa = (uintptr_t)x3;
    CfmP f;
    f.cpw = (TVb % 16 == 0 && (g.x3ns * 2) % 16 == 0 && (xa & 15) == 0) ? 16
          : ((TVb % 8 == 0 && (g.x3ns * 2) % 8 == 0 && (xa & 7) == 0) ? 8 : (((g.x3ns * 2) % 4 == 0 && (xa & 3) == 0) ? 4 : 0));
    if (f.cpw == 0) return 0;
    const uintptr_t ya = (uintptr_t)y;
    f.yw = (TVb % 16 == 0 && (g.yns * 2) % 16 == 0 && (ya & 15) == 0) ? 16
         : ((TVb % 8 == 0 && (g.yns * 2) % 8 == 0 && (ya & 7) == 0) ? 8 : (((g.yns * 2) % 4 == 0 && (ya & 3) == 0) ? 4 : 2));
    f.xs_bytes = (int)((TVb + 15) & ~15LL) + 16;
    const int UV = V * V, UVp = (UV + 15) & ~15, UP = UVp + 8, NTn = (V + 7) / 8;
    const int Rp = (g.R + 15) & ~15;
    const size_t szD = sizeof(bf16) * (size_t)g.K * Rp * UP, szQ = sizeof(bf16) * (size_t)CBM_CT * (8 * NTn * 40 + 8);
    const size_t szX = 2 * 16 * (size_t)f.xs_bytes, szP = sizeof(float) * 2 * (size_t)g.R * V;
    f.NQ = (szD + 2 * szQ + szX + szP + 16 <= 227 * 1024) ? 2 : 1;
    const size_t sm = szD + f.NQ * szQ + szX + szP + 16;
    if (sm > 227 * 1024) return 0;
    // channel-tile stride S: waves(S) * (table cost + tiles per CTA * tile cost), table : tile ~ R : 16
    const int nTiles = cdiv(g.Cout, CBM_CT);
    int S = 1;
    {
        double best = 1e30;
        for (int c = 1; c <= nTiles && c <= 16; ++c) {
            const double waves = (double)cdiv((long long)g.N * c, num_sms());
            const double cost = waves * ((double)g.R / 16.0 + (double)cdiv(nTiles, c));
            if (cost < best - 1e-9) { best = cost; S = c; }
        }
    }
    f.S = S;
    dim3 grid(S, g.N);
#define CFM_LAUNCH(VV, TT)                                                                                                  \
    do {                                                                                                                    \
        static SmemLimit lim;                                                                                               \
        ensure_smem(ctrgc_fwd_mma_kernel<VV, TT>, lim, sm);                                                                 \
        ctrgc_fwd_mma_kernel<VV, TT><<<grid, CBM_THREADS, sm, st>>>(g, f, (const bf16*)x3, x1, x2, W4, b4, PA, alpha,       \
                                                                    (bf16*)y, ssum, ssq);                                   \
    } while (0)
    const int TB = g.T <= 16 ? 1 : (g.T <= 32 ? 2 : 4);
    if (V == 20) { if (TB == 1) CFM_LAUNCH(20, 1); else if (TB == 2) CFM_LAUNCH(20, 2); else CFM_LAUNCH(20, 4); }
    else         { if (TB == 1) CFM_LAUNCH(25, 1); else if (TB == 2) CFM_LAUNCH(25, 2); else CFM_LAUNCH(25, 4); }
#undef CFM_LAUNCH
    count_launch();
    const int rc = check_launch("ctrgc_fwd(mma)");
    return rc < 0 ? rc : 1;
}

template <typename T>
static int launch_fwd(const CtrgcP& g, int V, const void* x3, const float* x1, const float* x2, const float* W4,
                      const float* b4, const float* PA, const float* alpha, void* y, double* ssum, double* ssq,
                      cudaStream_t st) {
    dim3 grid(cdiv(g.Cout, g.CT), g.N);
    if (V == 20) {
        constexpr int VP = VPad<20>::VP, DP = VPad<20>::DP;
        const size_t sm = sizeof(float) * ((size_t)g.K * g.CT * 20 * VP + (size_t)g.R * 20 * DP + g.CT * g.R + g.CT * 2 + 2 * g.R * 20);
        TG_REQUIRE(sm <= 227 * 1024, "ctrgc_fwd: shared memory %zu too large (R=%d)", sm, g.R);
        static SmemLimit lim;
        ensure_smem(ctrgc_fwd_kernel<T, 20>, lim, sm);
        ctrgc_fwd_kernel<T, 20><<<grid, 256, sm, st>>>(g, (const T*)x3, x1, x2, W4, b4, PA, alpha, (T*)y, ssum, ssq);
    } else {
        constexpr int VP = VPad<25>::VP, DP = VPad<25>::DP;
        const size_t sm = sizeof(float) * ((size_t)g.K * g.CT * 25 * VP + (size_t)g.R * 25 * DP + g.CT * g.R + g.CT * 2 + 2 * g.R * 25);
        TG_REQUIRE(sm <= 227 * 1024, "ctrgc_fwd: shared memory %zu too large (R=%d)", sm, g.R);
        static SmemLimit lim;
        ensure_smem(ctrgc_fwd_kernel<T, 25>, lim, sm);
        ctrgc_fwd_kernel<T, 25><<<grid, 256, sm, st>>>(g, (const T*)x3, x1, x2, W4, b4, PA, alpha, (T*)y, ssum, ssq);
    }
    count_launch();
    return check_launch("ctrgc_fwd");
}

static size_t ctrgc_bwd_mma_smem(int V, int R, bool lean, int CT = CBM_CT) {
    const int UV = V * V, UVp = (UV + 15) & ~15, UP = UVp + 8, NTn = (V + 7) / 8;
    const int Rp = (R + 15) & ~15, RW = Rp + 8, NRt = (R + 1 + 7) / 8, DR = NRt * 8 > Rp ? NRt * 8 : Rp;
    return sizeof(float) * ((size_t)(lean ? 16 : R) * UVp + (size_t)CT * UV + CT + 2 * (size_t)R * V + 64) +
           sizeof(bf16) * ((size_t)DR * UP + 16 * (size_t)UP + 16 * RW + (size_t)Rp * 24 + (size_t)CT * 8 * NTn * 40) + 16;
}

// returns 1 if launched, 0 if the shape does not fit (caller falls back to the SIMT kernel), <0 on error
static int launch_bwd_mma(const CtrgcP& g0, int V, const Opnd& go, const void* x3, const float* x1, const float* x2,
                          const float* W4, const float* b4, const float* PA, const float* alpha, void* dx3, long long dx3ns,
                          float* dx1, float* dx2, float* dW4, float* db4, float* dPA, float* dalpha, cudaStream_t st) {
    CtrgcP g = g0;
    g.CT = CBM_CT;
    static const int lean_env = [] { const char* e = getenv("TAMGCN_CBM_LEAN"); return e ? atoi(e) : 0; }();
    static const int dbg_env = [] {
        const char* e = getenv("TAMGCN_CBM_DBG");
        const int v = e ? atoi(e) : 0;
        if (v) cudaMemcpyToSymbol(g_cbm_dbg, &v, sizeof(int));
        return v;
    }();
    (void)dbg_env;   // 1: force
    bool lean = lean_env == 1 && g.R > 16;
    size_t sm = ctrgc_bwd_mma_smem(V, g.R, lean);
    if (sm > 227 * 1024 && g.R > 16) { lean = true; sm = ctrgc_bwd_mma_smem(V, g.R, true); }
    if (sm > 227 * 1024 || g.R > 128 || (g.R & 1)) return 0;
    // 8-channel CTAs, two per SM: measured slower than one 16-channel CTA (2.23 vs 2.13 ms at N'=2048; the tables are
    // rebuilt twice as often) -> opt-in, TAMGCN_CBM_CT=8
    static const int ct_env = [] { const char* e = getenv("TAMGCN_CBM_CT"); return e ? atoi(e) : 0; }();
    const size_t sm8 = ctrgc_bwd_mma_smem(V, g.R, lean, 8);
    const bool ct8 = ct_env == 8 && !lean && g.R <= 55 && 2 * (sm8 + 1024) <= 227 * 1024;
    if (ct8) { g.CT = 8; sm = sm8; }
    dim3 grid(cdiv(g.Cout, g.CT), g.N);
#define CBM_LAUNCH(VV, LL, CC)                                                                                              \
    do {                                                                                                                    \
        static SmemLimit lim;                                                                                               \
        ensure_smem(ctrgc_bwd_mma_kernel<VV, LL, CC>, lim, sm);                                                             \
        ctrgc_bwd_mma_kernel<VV, LL, CC><<<grid, CC * 32, sm, st>>>(g, go, (const bf16*)x3, x1, x2, W4, b4, PA, alpha,      \
                                                                    (bf16*)dx3, dx3ns, dx1, dx2, dW4, db4, dPA, dalpha);   \
    } while (0)
    if (V == 20) { if (lean) CBM_LAUNCH(20, true, 16); else if (ct8) CBM_LAUNCH(20, false, 8); else CBM_LAUNCH(20, false, 16); }
    else         { if (lean) CBM_LAUNCH(25, true, 16); else if (ct8) CBM_LAUNCH(25, false, 8); else CBM_LAUNCH(25, false, 16); }
#undef CBM_LAUNCH
    count_launch();
    const int rc = check_launch("ctrgc_bwd(mma)");
    return rc < 0 ? rc : 1;
}

template <typename T>
static int launch_bwd(const CtrgcP& g, int V, const Opnd& go, const void* x3, const float* x1, const float* x2,
                      const float* W4, const float* b4, const float* PA, const float* alpha, void* dx3,
                      long long dx3ns, float* dx1, float* dx2, float* dW4, float* db4, float* dPA, float* dalpha,
                      cudaStream_t st) {
    dim3 grid(cdiv(g.Cout, g.CT), g.N);
    if (V == 20) {
        constexpr int VP = VPad<20>::VP, DP = VPad<20>::DP;
        const size_t sm = sizeof(float) * ((size_t)g.CT * 20 * VP + 2 * (size_t)g.CT * 20 * DP + (size_t)g.R * 20 * DP +
                                           g.CT * g.R + 64 + 2 * g.R * 20);
        TG_REQUIRE(sm <= 227 * 1024, "ctrgc_bwd: shared memory %zu too large (R=%d)", sm, g.R);
        static SmemLimit lim;
        ensure_smem(ctrgc_bwd_kernel<T, 20>, lim, sm);
        ctrgc_bwd_kernel<T, 20><<<grid, 256, sm, st>>>(g, go, (const T*)x3, x1, x2, W4, b4, PA, alpha, (T*)dx3, dx3ns,
                                                       dx1, dx2, dW4, db4, dPA, dalpha);
    } else {
        constexpr int VP = VPad<25>::VP, DP = VPad<25>::DP;
        const size_t sm = sizeof(float) * ((size_t)g.CT * 25 * VP + 2 * (size_t)g.CT * 25 * DP + (size_t)g.R * 25 * DP +
                                           g.CT * g.R + 64 + 2 * g.R * 25);
        TG_REQUIRE(sm <= 227 * 1024, "ctrgc_bwd: shared memory %zu too large (R=%d)", sm, g.R);
        static SmemLimit lim;
        ensure_smem(ctrgc_bwd_kernel<T, 25>, lim, sm);
        ctrgc_bwd_kernel<T, 25><<<grid, 256, sm, st>>>(g, go, (const T*)x3, x1, x2, W4, b4, PA, alpha, (T*)dx3, dx3ns,
                                                       dx1, dx2, dW4, db4, dPA, dalpha);
    }
    count_launch();
    return check_launch("ctrgc_bwd");
}

}  // namespace tamgcn

namespace tamgcn {
int ctrgc_fwd_tc(const void* x3, long long x3ns, int N, int Cout, int T, int V, int K, int R, const float* x1,
                 const float* x2, long long x12ns, const float* W4, const float* b4, const float* PA, const float* alpha,
                 void* y, long long yns, double* ssum, double* ssq, cudaStream_t st);
}

using namespace tamgcn;

extern "C" int tamgcn_mean_t(int dtype, const void* x, int64_t x_nstride, int N, int C, int T, int V, float* m,
                             tamgcn_stream stream) {
    TG_REQUIRE(x && m && N > 0 && C > 0 && T > 0 && V > 0, "mean_t: bad arguments");
    TG_REQUIRE(N <= 65535, "mean_t: N too large");
    dim3 grid(cdiv(C, 8), N);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TAMGCN_F32) mean_t_kernel<float><<<grid, 256, 0, st>>>((const float*)x, x_nstride, C, T, V, m);
    else if (dtype == TAMGCN_BF16) mean_t_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, x_nstride, C, T, V, m);
    else return set_error("mean_t: bad dtype %d", dtype);
    count_launch();
    return check_launch("mean_t");
}

static int ctrgc_common(const char* who, int N, int Cout, int T, int V, int K, int R) {
    TG_REQUIRE(N > 0 && Cout > 0 && T > 0 && K > 0 && R > 0, "%s: empty dimension", who);
    TG_REQUIRE(V == 20 || V == 25, "%s: V=%d not supported (20 = NW-UCLA, 25 = NTU RGB+D)", who, V);
    TG_REQUIRE(N <= 65535, "%s: N too large for one launch", who);
    return 0;
}

extern "C" int tamgcn_ctrgc_fwd(int dtype, const void* x3, int64_t x3_nstride, int N, int Cout, int T, int V, int K,
                                int R, const float* x1, const float* x2, int64_t x12_nstride, const float* W4,
                                const float* b4, const float* PA, const float* alpha, void* y, int64_t y_nstride,
                                double* stat_sum, double* stat_sumsq, tamgcn_stream stream) {
    if (ctrgc_common("ctrgc_fwd", N, Cout, T, V, K, R)) return -1;
    TG_REQUIRE(x3 && x1 && x2 && W4 && b4 && PA && alpha && y, "ctrgc_fwd: null pointer");
    TG_REQUIRE((stat_sum == nullptr) == (stat_sumsq == nullptr), "ctrgc_fwd: stats must both be set");
    CtrgcP g;
    g.N = N; g.Cout = Cout; g.T = T; g.K = K; g.R = R;
    g.CT = (V == 20) ? 16 : 8;
    if (V == 25) {
        // as many channels per CTA as fit ~200 KB of shared memory: every CTA of a sample rebuilds the tanh table, so the
        // table cost scales with Cout / CT
        const int s = dtype == TAMGCN_BF16 ? 4 : 4;          // the SIMT kernel keeps Q and D in fp32
        while (g.CT > 1 && (size_t)s * ((size_t)K * g.CT * 25 * 28 + (size_t)R * 25 * 25 + g.CT * R + g.CT * 2 + 2 * R * 25) > 200 * 1024)
            g.CT >>= 1;
    }
    g.x3ns = x3_nstride; g.x12ns = x12_nstride; g.yns = y_nstride;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TAMGCN_BF16) {     // tensor-core path (ctrgc_tc.cu); 0 = shape not covered, use the SIMT kernel
        const int rc = ctrgc_fwd_tc(x3, x3_nstride, N, Cout, T, V, K, R, x1, x2, x12_nstride, W4, b4, PA, alpha, y, y_nstride,
                                    stat_sum, stat_sumsq, st);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    if (dtype == TAMGCN_F32) return launch_fwd<float>(g, V, x3, x1, x2, W4, b4, PA, alpha, y, stat_sum, stat_sumsq, st);
    if (dtype == TAMGCN_BF16) {     // warp-MMA kernel for what the tcgen05 kernels leave (large R)
        const int rc = launch_fwd_mma(g, V, x3, x1, x2, W4, b4, PA, alpha, y, stat_sum, stat_sumsq, st);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    if (dtype == TAMGCN_BF16) return launch_fwd<bf16>(g, V, x3, x1, x2, W4, b4, PA, alpha, y, stat_sum, stat_sumsq, st);
    return set_error("ctrgc_fwd: bad dtype %d", dtype);
}

extern "C" int tamgcn_ctrgc_bwd(int dtype, const tamgcn_operand* gop, const void* x3, int64_t x3_nstride, int N,
                                int Cout, int T, int V, int K, int R, const float* x1, const float* x2,
                                int64_t x12_nstride, const float* W4, const float* b4, const float* PA,
                                const float* alpha, void* dx3, int64_t dx3_nstride, float* dx1, float* dx2,
                                float* dW4, float* db4, float* dPA, float* dalpha, tamgcn_stream stream) {
    if (ctrgc_common("ctrgc_bwd", N, Cout, T, V, K, R)) return -1;
    TG_REQUIRE(gop && gop->p && x3 && x1 && x2 && W4 && b4 && PA && alpha && dx3 && dx1 && dx2 && dW4 && db4 && dPA &&
                   dalpha, "ctrgc_bwd: null pointer");
    CtrgcP g;
    g.N = N; g.Cout = Cout; g.T = T; g.K = K; g.R = R;
    g.CT = 8;
    g.x3ns = x3_nstride; g.x12ns = x12_nstride; g.yns = 0;
    const Opnd go = make_opnd(gop);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TAMGCN_F32)
        return launch_bwd<float>(g, V, go, x3, x1, x2, W4, b4, PA, alpha, dx3, dx3_nstride, dx1, dx2, dW4, db4, dPA, dalpha, st);
    if (dtype == TAMGCN_BF16) {
        const char* e = getenv("TAMGCN_DISABLE_TC");
        if (!(e && e[0] == '1')) {
            const int rc = launch_bwd_mma(g, V, go, x3, x1, x2, W4, b4, PA, alpha, dx3, dx3_nstride, dx1, dx2, dW4, db4, dPA, dalpha, st);
            if (rc != 0) return rc < 0 ? rc : 0;
        }
        return launch_bwd<bf16>(g, V, go, x3, x1, x2, W4, b4, PA, alpha, dx3, dx3_nstride, dx1, dx2, dW4, db4, dPA, dalpha, st);
    }
    return set_error("ctrgc_bwd: bad dtype %d", dtype);
}
