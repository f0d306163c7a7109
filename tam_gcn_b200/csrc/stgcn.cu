// stgcn.cu — ST-GCN graph aggregation  out[n,c,t,w] = sum_{k,v} y[n,k*C+c,t,v] * A[k,v,w]
// (reference models/stgcn.py:60-62, einsum 'nkctv,kvw->nctw'), forward and backward.
//
// A (K x V x V, shared by all samples) sits in shared memory; one thread owns one (n,c,t) row of V
// joints in registers, so y is read once and out written once (HBM-bound streaming pass).  A CTA owns
// one output channel and strides over samples, which keeps the BatchNorm statistics of `out` (its
// consumer is st_gcn.tcn[0]) in registers.  dA is reduced per CTA in shared memory, then fp32 atomics.
#include "common.cuh"
#include "rows.cuh"

namespace tamgcn {

static inline dim3 agg_grid(int N, int C) {
    int ng = (148 * 8 + C - 1) / C;
    if (ng > N) ng = N;
    if (ng < 1) ng = 1;
    return dim3(C, ng);
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
graph_agg_fwd_kernel(int N, int K, int C, int Tn, const T* __restrict__ y, long long yns, const float* __restrict__ A,
                     T* __restrict__ out, long long ons, double* ssum, double* ssq) {
    constexpr int VP = VPad<V>::VP;
    extern __shared__ __align__(16) float As[];  // [K][V][VP]
    __shared__ float scratch[64];
    for (int idx = threadIdx.x; idx < K * V * VP; idx += blockDim.x) {
        const int w = idx % VP, kv = idx / VP;
        As[idx] = (w < V) ? __ldg(A + kv * V + w) : 0.f;
    }
    __syncthreads();
    const int c = blockIdx.x;
    const int ns = (N - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;  // samples of this CTA
    float st[2] = {0.f, 0.f};
    for (int row = threadIdx.x; row < ns * Tn; row += blockDim.x) {
        const int n = blockIdx.y + (row / Tn) * gridDim.y, t = row % Tn;
        float acc[VP];
#pragma unroll
        for (int w = 0; w < VP; ++w) acc[w] = 0.f;
        for (int k = 0; k < K; ++k) {
            float xr[VP];
            load_row<T, V, VP>(y + (long long)n * yns + (((long long)k * C + c) * Tn + t) * V, xr);
            const float* a = As + k * V * VP;
#pragma unroll
            for (int v = 0; v < V; ++v) {
#pragma unroll
                for (int w4 = 0; w4 < VP / 4; ++w4) {
                    const float4 aa = *reinterpret_cast<const float4*>(a + v * VP + 4 * w4);
                    acc[4 * w4] = fmaf(aa.x, xr[v], acc[4 * w4]);
                    acc[4 * w4 + 1] = fmaf(aa.y, xr[v], acc[4 * w4 + 1]);
                    acc[4 * w4 + 2] = fmaf(aa.z, xr[v], acc[4 * w4 + 2]);
                    acc[4 * w4 + 3] = fmaf(aa.w, xr[v], acc[4 * w4 + 3]);
                }
            }
        }
#pragma unroll
        for (int w = 0; w < V; ++w) {
            acc[w] = rnd<T>(acc[w]);
            st[0] += acc[w];
            st[1] = fmaf(acc[w], acc[w], st[1]);
        }
        store_row<T, V>(out + (long long)n * ons + ((long long)c * Tn + t) * V, acc);
    }
    if (ssum) {
        block_sum<2>(st, scratch);
        if (threadIdx.x == 0) {
            atomicAdd(ssum + c, (double)st[0]);
            atomicAdd(ssq + c, (double)st[1]);
        }
    }
}

// dy[n,k*C+c,t,v] = sum_w g(n,c,t,w) A[k,v,w]
template <typename T, int V>
__global__ void __launch_bounds__(256)
graph_agg_dy_kernel(int N, int K, int C, int Tn, Opnd go, const float* __restrict__ A, T* __restrict__ dy,
                    long long dyns) {
    constexpr int VP = VPad<V>::VP;
    extern __shared__ __align__(16) float As[];  // [K][V][VP]  (w fastest)
    for (int idx = threadIdx.x; idx < K * V * VP; idx += blockDim.x) {
        const int w = idx % VP, kv = idx / VP;
        As[idx] = (w < V) ? __ldg(A + kv * V + w) : 0.f;
    }
    __syncthreads();
    const int c = blockIdx.x;
    const OpCoef cf = opnd_coef(go, c);
    const int ns = (N - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;
    for (int row = threadIdx.x; row < ns * Tn; row += blockDim.x) {
        const int n = blockIdx.y + (row / Tn) * gridDim.y, t = row % Tn;
        float gr[VP];
#pragma unroll
        for (int w = 0; w < VP; ++w) gr[w] = (w < V) ? opnd_val<T>(go, cf, n, ((long long)c * Tn + t) * V + w) : 0.f;
        for (int k = 0; k < K; ++k) {
            float o[V];
            const float* a = As + k * V * VP;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                float s = 0.f;
#pragma unroll
                for (int w4 = 0; w4 < VP / 4; ++w4) {
                    const float4 aa = *reinterpret_cast<const float4*>(a + v * VP + 4 * w4);
                    s = fmaf(aa.x, gr[4 * w4], s);
                    s = fmaf(aa.y, gr[4 * w4 + 1], s);
                    s = fmaf(aa.z, gr[4 * w4 + 2], s);
                    s = fmaf(aa.w, gr[4 * w4 + 3], s);
                }
                o[v] = s;
            }
            store_row<T, V>(dy + (long long)n * dyns + (((long long)k * C + c) * Tn + t) * V, o);
        }
    }
}

// dA[k,v,w] += sum_{n,c,t} y[n,k*C+c,t,v] g(n,c,t,w):  rows staged in shared memory in chunks of RB,
// each thread owns a strided subset of the K*V*V accumulators in registers.
#define AGG_RB 32
#define AGG_MAXACC 8
template <typename T>
__global__ void __launch_bounds__(256)
graph_agg_dA_kernel(int N, int K, int C, int Tn, int V, Opnd go, const T* __restrict__ y, long long yns,
                    float* __restrict__ dA) {
    extern __shared__ __align__(16) float sm[];
    float* ys = sm;                       // [RB][K*V]
    float* gs = ys + AGG_RB * K * V;      // [RB][V]
    const int c = blockIdx.x;
    const OpCoef cf = opnd_coef(go, c);
    const int ns = (N - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;
    const int rows = ns * Tn, KV = K * V, nacc = K * V * V;
    float acc[AGG_MAXACC];
#pragma unroll
    for (int i = 0; i < AGG_MAXACC; ++i) acc[i] = 0.f;
    for (int r0 = 0; r0 < rows; r0 += AGG_RB) {
        const int nr = min(AGG_RB, rows - r0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < nr * (KV + V); idx += blockDim.x) {
            const int rr = idx / (KV + V), j = idx - rr * (KV + V);
            const int row = r0 + rr, n = blockIdx.y + (row / Tn) * gridDim.y, t = row % Tn;
            if (j < KV) {
                const int k = j / V, v = j - k * V;
                ys[rr * KV + j] = ldf<T>(y + (long long)n * yns + (((long long)k * C + c) * Tn + t) * V + v);
            } else {
                gs[rr * V + (j - KV)] = opnd_val<T>(go, cf, n, ((long long)c * Tn + t) * V + (j - KV));
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < AGG_MAXACC; ++i) {
            const int e = threadIdx.x + i * 256;
            if (e < nacc) {
                const int kv = e / V, w = e - kv * V;
                float s = acc[i];
                for (int rr = 0; rr < nr; ++rr) s = fmaf(ys[rr * KV + kv], gs[rr * V + w], s);
                acc[i] = s;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < AGG_MAXACC; ++i) {
        const int e = threadIdx.x + i * 256;
        if (e < nacc) atomicAdd(dA + e, acc[i]);
    }
}

// bf16 tensor-core versions (stgcn_mma.cu): 1 = launched, 0 = not covered, < 0 = error
int graph_agg_fwd_mma(int N, int K, int C, int T, int V, const void* y, long long yns, const float* A, void* out,
                      long long ons, double* ssum, double* ssq, cudaStream_t st);
int graph_agg_bwd_mma(int N, int K, int C, int T, int V, const Opnd& go, const void* y, long long yns, const float* A,
                      void* dy, long long dyns, float* dA, cudaStream_t st);

}  // namespace tamgcn

using namespace tamgcn;

extern "C" int tamgcn_graph_agg_fwd(int dtype, int N, int K, int C, int T, int V, const void* y, int64_t y_nstride,
                                    const float* A, void* out, int64_t out_nstride, double* stat_sum,
                                    double* stat_sumsq, tamgcn_stream stream) {
    TG_REQUIRE(N > 0 && K > 0 && C > 0 && T > 0 && y && A && out, "graph_agg_fwd: bad arguments");
    TG_REQUIRE(V == 20 || V == 25, "graph_agg_fwd: V=%d not supported (20 or 25)", V);
    TG_REQUIRE((stat_sum == nullptr) == (stat_sumsq == nullptr), "graph_agg_fwd: stats must both be set");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TAMGCN_BF16) {
        const int rc = graph_agg_fwd_mma(N, K, C, T, V, y, y_nstride, A, out, out_nstride, stat_sum, stat_sumsq, st);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    const dim3 grid = agg_grid(N, C);
    const size_t sm = sizeof(float) * K * V * ((V + 3) & ~3);
#define AGG_FWD(T_, V_)                                                                                        \
    graph_agg_fwd_kernel<T_, V_><<<grid, 256, sm, st>>>(N, K, C, T, (const T_*)y, y_nstride, A, (T_*)out, out_nstride, \
                                                         stat_sum, stat_sumsq)
    if (dtype == TAMGCN_F32) { if (V == 20) AGG_FWD(float, 20); else AGG_FWD(float, 25); }
    else if (dtype == TAMGCN_BF16) { if (V == 20) AGG_FWD(bf16, 20); else AGG_FWD(bf16, 25); }
    else return set_error("graph_agg_fwd: bad dtype %d", dtype);
#undef AGG_FWD
    count_launch();
    return check_launch("graph_agg_fwd");
}

extern "C" int tamgcn_graph_agg_bwd(int dtype, int N, int K, int C, int T, int V, const tamgcn_operand* dout,
                                    const void* y, int64_t y_nstride, const float* A, void* dy, int64_t dy_nstride,
                                    float* dA, tamgcn_stream stream) {
    TG_REQUIRE(N > 0 && K > 0 && C > 0 && T > 0 && dout && dout->p && A && (dy || dA), "graph_agg_bwd: bad arguments");
    TG_REQUIRE(V == 20 || V == 25, "graph_agg_bwd: V=%d not supported (20 or 25)", V);
    TG_REQUIRE(!dA || y, "graph_agg_bwd: dA needs y");
    TG_REQUIRE(K * V * V <= AGG_MAXACC * 256, "graph_agg_bwd: K*V*V=%d too large", K * V * V);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid = agg_grid(N, C);
    const Opnd go = make_opnd(dout);
    if (dtype == TAMGCN_BF16) {
        const int rc = graph_agg_bwd_mma(N, K, C, T, V, go, y, y_nstride, A, dy, dy_nstride, dA, st);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    const size_t sm = sizeof(float) * K * V * ((V + 3) & ~3);
#define AGG_DY(T_, V_) graph_agg_dy_kernel<T_, V_><<<grid, 256, sm, st>>>(N, K, C, T, go, A, (T_*)dy, dy_nstride)
    if (dtype != TAMGCN_F32 && dtype != TAMGCN_BF16) return set_error("graph_agg_bwd: bad dtype %d", dtype);
    if (dy) {
        if (dtype == TAMGCN_F32) { if (V == 20) AGG_DY(float, 20); else AGG_DY(float, 25); }
        else { if (V == 20) AGG_DY(bf16, 20); else AGG_DY(bf16, 25); }
        count_launch();
        if (check_launch("graph_agg_bwd(dy)")) return -2;
    }
#undef AGG_DY
    if (dA) {
        const size_t sm2 = sizeof(float) * AGG_RB * (K * V + V);
        if (dtype == TAMGCN_F32)
            graph_agg_dA_kernel<float><<<grid, 256, sm2, st>>>(N, K, C, T, V, go, (const float*)y, y_nstride, dA);
        else
            graph_agg_dA_kernel<bf16><<<grid, 256, sm2, st>>>(N, K, C, T, V, go, (const bf16*)y, y_nstride, dA);
        count_launch();
        return check_launch("graph_agg_bwd(dA)");
    }
    return 0;
}
