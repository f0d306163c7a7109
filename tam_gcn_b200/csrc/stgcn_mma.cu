// stgcn_mma.cu — ST-GCN graph aggregation for bf16 activations on warp-level tensor-core MMAs.
//
//   out[n,c,t,w]     = sum_k sum_v y[n,kC+c,t,v] A[k,v,w]        (reference models/stgcn.py:60-62)
//   dy[n,kC+c,t,v]   = sum_w g[n,c,t,w] A[k,v,w]
//   dA[k,v,w]       += sum_{n,c,t} y[n,kC+c,t,v] g[n,c,t,w]
//
// All three are streaming passes over (sample, channel) planes whose rows are V joints (40 / 50 bytes): HBM-bound, but the
// 2*K*V FMAs per element kept the SIMT kernels (stgcn.cu: one thread per row, A in shared memory) at 5 % of the
// bandwidth.  Here a warp owns one plane at a time and walks it in blocks of 16 rows:
//   * the 16*V*2 contiguous bytes of a block (of each of the K input planes) are fetched with 16-byte cp.async into a
//     per-warp staging buffer — from the 16-byte-aligned address below the block, so any plane alignment works —
//     double buffered across blocks and planes;
//   * m16n8k16 fragments are read from the staging buffer as aligned words + funnel shift (the odd-row misalignment of
//     V = 25 is a per-lane constant), the lazy operand f(a*P + b*Q + c) of the backward is applied on the fly;
//   * results go back through the staging buffer and leave as 16-byte coalesced stores.
// A (K x V x V) lives in registers as B fragments (fwd, dy); dA is accumulated in registers per warp (operands
// re-distributed with movmatrix), reduced per CTA in shared memory, then one fp32 atomic per element and CTA.
#include "common.cuh"
#include <cstdlib>

namespace tamgcn {

#define SG_WARPS 8
#define SG_D 2                                   // input blocks in flight per warp (ring depth; 3 measured equal: the kernels are issue-bound)
#define SG_THREADS (SG_WARPS * 32)

struct SgP {
    int N, K, C, T;
    long long yns, ons;     // sample strides (elements) of y / of out (fwd) or dy (bwd)
    int units;              // N * C
    int nblk;               // ceil(T / 16)
};

template <int V> struct SgCfg {
    static const int NTn = (V + 7) / 8;
    static const int BLK = 16 * V * 2;                         // bytes of a full 16-row block
    static const int SGB = ((BLK + 15 + 15) & ~15) + 16;       // staging bytes per plane block (misalignment + one word of slack)
};

__device__ __forceinline__ void sg_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t sg_movm(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}
__device__ __forceinline__ uint32_t sg_pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// [src, src + bytes) -> shared memory, as the 16-byte chunks of [src - mis, ...) (mis = src & 15): dst corresponds to
// src - mis.  Chunks that would touch memory outside [beg, end) are copied element by element.
__device__ __forceinline__ void sg_stage(uint32_t dst_s, unsigned char* dst, const unsigned char* src, int bytes,
                                         const unsigned char* beg, const unsigned char* end, int lane) {
    const int mis = (int)(reinterpret_cast<uintptr_t>(src) & 15);
    const unsigned char* base = src - mis;
    const int total = (mis + bytes + 15) & ~15;
    for (int o = lane * 16; o < total; o += 32 * 16) {
        const unsigned char* s = base + o;
        if (s >= beg && s + 16 <= end) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_s + o), "l"(s) : "memory");
        } else {
#pragma unroll
            for (int b = 0; b < 16; b += 2)
                if (s + b >= beg && s + b + 2 <= end)
                    *reinterpret_cast<unsigned short*>(dst + o + b) = __ldg(reinterpret_cast<const unsigned short*>(s + b));
        }
    }
}
__device__ __forceinline__ void sg_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void sg_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// staged block -> [src-aligned] global, bytes [dst, dst + bytes): whole 16-byte chunks as vector stores, the ragged
// head / tail element by element.  buf corresponds to dst - (dst & 15).
__device__ __forceinline__ void sg_copy_out(unsigned char* dst, const unsigned char* buf, int bytes, int lane) {
    const int mis = (int)(reinterpret_cast<uintptr_t>(dst) & 15);
    unsigned char* base = dst - mis;
    const int total = (mis + bytes + 15) & ~15;
    for (int o = lane * 16; o < total; o += 32 * 16) {
        if (o >= mis && o + 16 <= mis + bytes) {
            *reinterpret_cast<uint4*>(base + o) = *reinterpret_cast<const uint4*>(buf + o);
        } else {
#pragma unroll
            for (int b = 0; b < 16; b += 2)
                if (o + b >= mis && o + b < mis + bytes)
                    *reinterpret_cast<unsigned short*>(base + o + b) = *reinterpret_cast<const unsigned short*>(buf + o + b);
        }
    }
}

// the 2 x 4 A-fragment words of a 16-row block (rows gid, gid + 8; columns 8 b + 2 tig, +1) from a staged block.
// E0 = element index of (row gid, column 2 tig) in the buffer, including the buffer's misalignment shift.
template <int V>
__device__ __forceinline__ void sg_frags(const unsigned char* buf, int E0, int rows, int gid, int tig, uint32_t (&xa)[2][4]) {
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(buf) + (E0 >> 1);
    const int sh = (E0 & 1) * 16;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const bool tok = gid + 8 * h < rows;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int wo = (8 * h * V + 8 * b) / 2;              // compile-time word offset (8 h V + 8 b is even)
            uint32_t w = 0u;
            if (tok && 8 * b + 2 * tig < V) {
                w = __funnelshift_r(xw[wo], xw[wo + 1], sh);
                if (8 * b + 2 * tig + 1 >= V) w &= 0xffffu;
            }
            xa[h][b] = w;
        }
    }
}
// the same for a lazy operand f(a P + b Q + c)
template <int V>
__device__ __forceinline__ void sg_frags_lazy(const unsigned char* bufp, const unsigned char* bufq, int E0p, int E0q, int rows,
                                              int gid, int tig, const OpCoef& cf, bool has_q, bool relu, bool plain,
                                              uint32_t (&ga)[2][4]) {
    sg_frags<V>(bufp, E0p, rows, gid, tig, ga);
    if (plain) return;
    uint32_t qa[2][4];
    if (has_q) sg_frags<V>(bufq, E0q, rows, gid, tig, qa);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const bool tok = gid + 8 * h < rows;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const bool ok = tok && 8 * b + 2 * tig < V, ok1 = 8 * b + 2 * tig + 1 < V;
            float lo = fmaf(cf.a, __uint_as_float(ga[h][b] << 16), cf.c), hi = fmaf(cf.a, __uint_as_float(ga[h][b] & 0xffff0000u), cf.c);
            if (has_q) {
                lo = fmaf(cf.b, __uint_as_float(qa[h][b] << 16), lo);
                hi = fmaf(cf.b, __uint_as_float(qa[h][b] & 0xffff0000u), hi);
            }
            if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
            ga[h][b] = ok ? sg_pack(lo, ok1 ? hi : 0.f) : 0u;
        }
    }
}

// accumulator tile rows (gid, gid + 8) x columns (8 nt + 2 tig, +1) -> staged output block (bf16), returns nothing;
// E0 as above for the output buffer.  Optionally accumulates sum / sum of squares of the rounded values.
template <int V, bool STATS>
__device__ __forceinline__ void sg_store_acc(unsigned char* buf, int E0, int rows, int gid, int tig,
                                             const float (&acc)[SgCfg<V>::NTn][4], float& s1, float& s2) {
    unsigned short* xh = reinterpret_cast<unsigned short*>(buf);
#pragma unroll
    for (int nt = 0; nt < SgCfg<V>::NTn; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const bool ok = gid + 8 * h < rows && 8 * nt + 2 * tig < V, ok1 = 8 * nt + 2 * tig + 1 < V;
            if (ok) {
                const uint32_t w = sg_pack(acc[nt][2 * h], acc[nt][2 * h + 1]);
                if (STATS) {
                    const float lo = __uint_as_float(w << 16), hi = ok1 ? __uint_as_float(w & 0xffff0000u) : 0.f;
                    s1 += lo + hi;
                    s2 = fmaf(lo, lo, fmaf(hi, hi, s2));
                }
                const int e = E0 + 8 * h * V + 8 * nt;
                xh[e] = (unsigned short)(w & 0xffffu);
                if (ok1) xh[e + 1] = (unsigned short)(w >> 16);
            }
        }
}

// ---- forward ------------------------------------------------------------------------------------------------
template <int V, int K>
__global__ void __launch_bounds__(SG_THREADS)
graph_agg_fwd_mma_kernel(SgP p, const bf16* __restrict__ y, const float* __restrict__ A, bf16* __restrict__ out,
                         double* ssum, double* ssq) {
    constexpr int NTn = SgCfg<V>::NTn, SGB = SgCfg<V>::SGB;
    extern __shared__ __align__(16) unsigned char sg_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    unsigned char* my = sg_smem + (size_t)warp * (SG_D * K + 1) * SGB;       // [SG_D][K] input blocks + 1 output block
    const uint32_t my_s = (uint32_t)__cvta_generic_to_shared(my);
    unsigned char* obuf = my + SG_D * K * SGB;
    // B fragments: B[k = v][n = w] = A_k[v][w], pairs along v
    uint32_t bq[K][2][NTn][2];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int nt = 0; nt < NTn; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int v = ks * 16 + 2 * tig + 8 * j, w = nt * 8 + gid;
                    const float lo = (v < V && w < V) ? __ldg(A + (k * V + v) * V + w) : 0.f;
                    const float hi = (v + 1 < V && w < V) ? __ldg(A + (k * V + v + 1) * V + w) : 0.f;
                    bq[k][ks][nt][j] = sg_pack(lo, hi);
                }
    const long long TV = (long long)p.T * V;
    const unsigned char* ybeg = reinterpret_cast<const unsigned char*>(y);
    const unsigned char* yend = reinterpret_cast<const unsigned char*>(y + (long long)(p.N - 1) * p.yns + (long long)p.K * p.C * TV);
    const int wg = blockIdx.x * SG_WARPS + warp, wstride = gridDim.x * SG_WARPS;
    const long long total = (long long)p.units * p.nblk;                     // steps: (unit, block)
    const int my_units = wg < p.units ? (p.units - wg + wstride - 1) / wstride : 0;
    const long long J = (long long)my_units * p.nblk;

    auto src_of = [&](int unit, int tb, int k) {
        const int n = unit / p.C, c = unit - n * p.C;
        return reinterpret_cast<const unsigned char*>(y + (long long)n * p.yns + ((long long)k * p.C + c) * TV + (long long)tb * 16 * V);
    };
    int s_unit = wg, s_tb = 0;
    auto stage = [&](long long j) {
        if (j < J) {
            const int rows = min(16, p.T - s_tb * 16);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const unsigned char* src = src_of(s_unit, s_tb, k);
                const int o = ((int)(j % SG_D) * K + k) * SGB;
                sg_stage(my_s + o, my + o, src, rows * V * 2, ybeg, yend, lane);
            }
            if (++s_tb == p.nblk) { s_tb = 0; s_unit += wstride; }
        }
        sg_commit();
    };
    (void)total;
    for (int d = 0; d < SG_D - 1; ++d) stage(d);
    int unit = wg, tb = 0;
    float s1 = 0.f, s2 = 0.f;
    for (long long j = 0; j < J; ++j) {
        stage(j + SG_D - 1);
        sg_wait<SG_D - 1>();
        __syncwarp();
        const int rows = min(16, p.T - tb * 16);
        const int n = unit / p.C, c = unit - n * p.C;
        float acc[NTn][4];
#pragma unroll
        for (int nt = 0; nt < NTn; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const unsigned char* src = src_of(unit, tb, k);
            const int es = (int)(reinterpret_cast<uintptr_t>(src) & 15) >> 1;
            uint32_t xa[2][4];
            sg_frags<V>(my + ((int)(j % SG_D) * K + k) * SGB, es + gid * V + 2 * tig, rows, gid, tig, xa);
#pragma unroll
            for (int nt = 0; nt < NTn; ++nt)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t a[4] = {xa[0][2 * ks], xa[1][2 * ks], xa[0][2 * ks + 1], xa[1][2 * ks + 1]};
                    sg_mma(acc[nt], a, bq[k][ks][nt][0], bq[k][ks][nt][1]);
                }
        }
        bf16* dst = out + (long long)n * p.ons + (long long)c * TV + (long long)tb * 16 * V;
        const int eo = (int)(reinterpret_cast<uintptr_t>(dst) & 15) >> 1;
        sg_store_acc<V, true>(obuf, eo + gid * V + 2 * tig, rows, gid, tig, acc, s1, s2);
        __syncwarp();
        sg_copy_out(reinterpret_cast<unsigned char*>(dst), obuf, rows * V * 2, lane);
        __syncwarp();
        if (++tb == p.nblk) {
            if (ssum) {
                s1 = warp_sum(s1); s2 = warp_sum(s2);
                if (lane == 0) { atomicAdd(ssum + c, (double)s1); atomicAdd(ssq + c, (double)s2); }
            }
            s1 = 0.f; s2 = 0.f;
            tb = 0; unit += wstride;
        }
    }
    sg_wait<0>();
}

// ---- backward: dy -------------------------------------------------------------------------------------------
template <int V, int K>
__global__ void __launch_bounds__(SG_THREADS)
graph_agg_dy_mma_kernel(SgP p, Opnd go, const float* __restrict__ A, bf16* __restrict__ dy) {
    constexpr int NTn = SgCfg<V>::NTn, SGB = SgCfg<V>::SGB;
    extern __shared__ __align__(16) unsigned char sg_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    unsigned char* my = sg_smem + (size_t)warp * (2 * SG_D + 1) * SGB;       // [SG_D][P, Q] input blocks + 1 output block
    const uint32_t my_s = (uint32_t)__cvta_generic_to_shared(my);
    unsigned char* obuf = my + 2 * SG_D * SGB;
    // B fragments: B[k = w][n = v] = A_k[v][w], pairs along w
    uint32_t bq[K][2][NTn][2];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int nt = 0; nt < NTn; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int w = ks * 16 + 2 * tig + 8 * j, v = nt * 8 + gid;
                    const float lo = (v < V && w < V) ? __ldg(A + (k * V + v) * V + w) : 0.f;
                    const float hi = (v < V && w + 1 < V) ? __ldg(A + (k * V + v) * V + w + 1) : 0.f;
                    bq[k][ks][nt][j] = sg_pack(lo, hi);
                }
    const long long TV = (long long)p.T * V;
    const bool has_q = go.q != nullptr, plain = !go.a && !go.c && !has_q && !go.relu;
    const unsigned char* pbeg = reinterpret_cast<const unsigned char*>(go.p);
    const unsigned char* pend = reinterpret_cast<const unsigned char*>((const bf16*)go.p + (long long)(p.N - 1) * go.pns + (long long)p.C * TV);
    const unsigned char* qbeg = reinterpret_cast<const unsigned char*>(go.q);
    const unsigned char* qend = has_q ? reinterpret_cast<const unsigned char*>((const bf16*)go.q + (long long)(p.N - 1) * go.qns + (long long)p.C * TV) : nullptr;
    const int wg = blockIdx.x * SG_WARPS + warp, wstride = gridDim.x * SG_WARPS;
    const int my_units = wg < p.units ? (p.units - wg + wstride - 1) / wstride : 0;
    const long long J = (long long)my_units * p.nblk;

    auto src_of = [&](const void* base, long long ns, int unit, int tb) {
        const int n = unit / p.C, c = unit - n * p.C;
        return reinterpret_cast<const unsigned char*>((const bf16*)base + (long long)n * ns + (long long)c * TV + (long long)tb * 16 * V);
    };
    int s_unit = wg, s_tb = 0;
    auto stage = [&](long long j) {
        if (j < J) {
            const int rows = min(16, p.T - s_tb * 16);
            const int o = (int)(j % SG_D) * 2 * SGB;
            sg_stage(my_s + o, my + o, src_of(go.p, go.pns, s_unit, s_tb), rows * V * 2, pbeg, pend, lane);
            if (has_q) sg_stage(my_s + o + SGB, my + o + SGB, src_of(go.q, go.qns, s_unit, s_tb), rows * V * 2, qbeg, qend, lane);
            if (++s_tb == p.nblk) { s_tb = 0; s_unit += wstride; }
        }
        sg_commit();
    };
    for (int d = 0; d < SG_D - 1; ++d) stage(d);
    int unit = wg, tb = 0;
    OpCoef cf = {1.f, 0.f, 0.f};
    for (long long j = 0; j < J; ++j) {
        stage(j + SG_D - 1);
        sg_wait<SG_D - 1>();
        __syncwarp();
        const int rows = min(16, p.T - tb * 16);
        const int n = unit / p.C, c = unit - n * p.C;
        if (tb == 0) cf = opnd_coef(go, c);
        const unsigned char* sp = src_of(go.p, go.pns, unit, tb);
        const int esp = (int)(reinterpret_cast<uintptr_t>(sp) & 15) >> 1;
        int esq = 0;
        if (has_q) esq = (int)(reinterpret_cast<uintptr_t>(src_of(go.q, go.qns, unit, tb)) & 15) >> 1;
        const unsigned char* bp = my + (int)(j % SG_D) * 2 * SGB;
        uint32_t ga[2][4];
        sg_frags_lazy<V>(bp, bp + SGB, esp + gid * V + 2 * tig, esq + gid * V + 2 * tig, rows, gid, tig, cf, has_q, go.relu != 0, plain, ga);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float acc[NTn][4];
#pragma unroll
            for (int nt = 0; nt < NTn; ++nt) {
                acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t a[4] = {ga[0][2 * ks], ga[1][2 * ks], ga[0][2 * ks + 1], ga[1][2 * ks + 1]};
                    sg_mma(acc[nt], a, bq[k][ks][nt][0], bq[k][ks][nt][1]);
                }
            }
            bf16* dst = dy + (long long)n * p.ons + ((long long)k * p.C + c) * TV + (long long)tb * 16 * V;
            const int eo = (int)(reinterpret_cast<uintptr_t>(dst) & 15) >> 1;
            float d1 = 0.f, d2 = 0.f;
            sg_store_acc<V, false>(obuf, eo + gid * V + 2 * tig, rows, gid, tig, acc, d1, d2);
            __syncwarp();
            sg_copy_out(reinterpret_cast<unsigned char*>(dst), obuf, rows * V * 2, lane);
            __syncwarp();
        }
        if (++tb == p.nblk) { tb = 0; unit += wstride; }
    }
    sg_wait<0>();
}

// ---- backward: dA (blockIdx.y = k) --------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(SG_THREADS)
graph_agg_dA_mma_kernel(SgP p, Opnd go, const bf16* __restrict__ y, float* __restrict__ dA) {
    constexpr int NTn = SgCfg<V>::NTn, SGB = SgCfg<V>::SGB;
    extern __shared__ __align__(16) unsigned char sg_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    float* dAs = reinterpret_cast<float*>(sg_smem);                          // [V*V] CTA accumulator
    unsigned char* my = sg_smem + ((V * V * 4 + 15) & ~15) + (size_t)warp * 3 * SG_D * SGB;      // [SG_D][y, P, Q]
    const uint32_t my_s = (uint32_t)__cvta_generic_to_shared(my);
    for (int i = threadIdx.x; i < V * V; i += SG_THREADS) dAs[i] = 0.f;
    __syncthreads();
    const int k = blockIdx.y;
    const long long TV = (long long)p.T * V;
    const bool has_q = go.q != nullptr, plain = !go.a && !go.c && !has_q && !go.relu;
    const unsigned char* ybeg = reinterpret_cast<const unsigned char*>(y);
    const unsigned char* yend = reinterpret_cast<const unsigned char*>(y + (long long)(p.N - 1) * p.yns + (long long)p.K * p.C * TV);
    const unsigned char* pbeg = reinterpret_cast<const unsigned char*>(go.p);
    const unsigned char* pend = reinterpret_cast<const unsigned char*>((const bf16*)go.p + (long long)(p.N - 1) * go.pns + (long long)p.C * TV);
    const unsigned char* qbeg = reinterpret_cast<const unsigned char*>(go.q);
    const unsigned char* qend = has_q ? reinterpret_cast<const unsigned char*>((const bf16*)go.q + (long long)(p.N - 1) * go.qns + (long long)p.C * TV) : nullptr;
    const int wg = blockIdx.x * SG_WARPS + warp, wstride = gridDim.x * SG_WARPS;
    const int my_units = wg < p.units ? (p.units - wg + wstride - 1) / wstride : 0;
    const long long J = (long long)my_units * p.nblk;

    auto src_of = [&](const void* base, long long ns, long long plane, int unit, int tb) {
        const int n = unit / p.C, c = unit - n * p.C;
        return reinterpret_cast<const unsigned char*>((const bf16*)base + (long long)n * ns + (plane + c) * TV + (long long)tb * 16 * V);
    };
    int s_unit = wg, s_tb = 0;
    auto stage = [&](long long j) {
        if (j < J) {
            const int rows = min(16, p.T - s_tb * 16);
            const int o = (int)(j % SG_D) * 3 * SGB;
            sg_stage(my_s + o, my + o, src_of(y, p.yns, (long long)k * p.C, s_unit, s_tb), rows * V * 2, ybeg, yend, lane);
            sg_stage(my_s + o + SGB, my + o + SGB, src_of(go.p, go.pns, 0, s_unit, s_tb), rows * V * 2, pbeg, pend, lane);
            if (has_q) sg_stage(my_s + o + 2 * SGB, my + o + 2 * SGB, src_of(go.q, go.qns, 0, s_unit, s_tb), rows * V * 2, qbeg, qend, lane);
            if (++s_tb == p.nblk) { s_tb = 0; s_unit += wstride; }
        }
        sg_commit();
    };
    for (int d = 0; d < SG_D - 1; ++d) stage(d);
    int unit = wg, tb = 0;
    OpCoef cf = {1.f, 0.f, 0.f};
    float acc[2][NTn][4];                                // dA_k[v = 16 mu + ...][w = 8 nt + ...]
#pragma unroll
    for (int mu = 0; mu < 2; ++mu)
#pragma unroll
        for (int nt = 0; nt < NTn; ++nt) acc[mu][nt][0] = acc[mu][nt][1] = acc[mu][nt][2] = acc[mu][nt][3] = 0.f;
    for (long long j = 0; j < J; ++j) {
        stage(j + SG_D - 1);
        sg_wait<SG_D - 1>();
        __syncwarp();
        const int rows = min(16, p.T - tb * 16);
        const int c = unit % p.C;
        if (tb == 0) cf = opnd_coef(go, c);
        const unsigned char* bb = my + (int)(j % SG_D) * 3 * SGB;
        const int esy = (int)(reinterpret_cast<uintptr_t>(src_of(y, p.yns, (long long)k * p.C, unit, tb)) & 15) >> 1;
        const int esp = (int)(reinterpret_cast<uintptr_t>(src_of(go.p, go.pns, 0, unit, tb)) & 15) >> 1;
        int esq = 0;
        if (has_q) esq = (int)(reinterpret_cast<uintptr_t>(src_of(go.q, go.qns, 0, unit, tb)) & 15) >> 1;
        uint32_t ya[2][4], ga[2][4];
        sg_frags<V>(bb, esy + gid * V + 2 * tig, rows, gid, tig, ya);
        sg_frags_lazy<V>(bb + SGB, bb + 2 * SGB, esp + gid * V + 2 * tig, esq + gid * V + 2 * tig, rows, gid, tig, cf, has_q,
                         go.relu != 0, plain, ga);
        // dA[v,w] += sum_t y[t,v] g[t,w]: A operand = y^T, B operand = g, both re-distributed with movmatrix
        uint32_t bx[NTn][2];
#pragma unroll
        for (int nt = 0; nt < NTn; ++nt) { bx[nt][0] = sg_movm(ga[0][nt]); bx[nt][1] = sg_movm(ga[1][nt]); }
#pragma unroll
        for (int mu = 0; mu < 2; ++mu) {
            const uint32_t a[4] = {sg_movm(ya[0][2 * mu]), sg_movm(ya[0][2 * mu + 1]), sg_movm(ya[1][2 * mu]), sg_movm(ya[1][2 * mu + 1])};
#pragma unroll
            for (int nt = 0; nt < NTn; ++nt) sg_mma(acc[mu][nt], a, bx[nt][0], bx[nt][1]);
        }
        if (++tb == p.nblk) { tb = 0; unit += wstride; }
    }
    sg_wait<0>();
#pragma unroll
    for (int mu = 0; mu < 2; ++mu)
#pragma unroll
        for (int nt = 0; nt < NTn; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int v = 16 * mu + gid + 8 * (e >> 1), w = 8 * nt + 2 * tig + (e & 1);
                if (v < V && w < V) atomicAdd(dAs + v * V + w, acc[mu][nt][e]);
            }
    __syncthreads();
    for (int i = threadIdx.x; i < V * V; i += SG_THREADS) atomicAdd(dA + k * V * V + i, dAs[i]);
}

static bool sg_disabled() {
    static const bool off = [] { const char* e = getenv("TAMGCN_DISABLE_AGG_MMA"); return e && e[0] == '1'; }();
    return off;
}
static int sg_grid(const SgP& p, size_t smem_per_cta) {
    int per_sm = (int)((227 * 1024) / (smem_per_cta + 1024));
    if (per_sm > 2) per_sm = 2;                          // registers (B fragments of A live in registers) allow two
    if (per_sm < 1) per_sm = 1;
    long long ctas = (long long)num_sms() * per_sm;
    const long long need = ((long long)p.units + SG_WARPS - 1) / SG_WARPS;
    if (ctas > need) ctas = need;
    return (int)(ctas < 1 ? 1 : ctas);
}

// returns 1 if launched, 0 if not covered (caller uses the SIMT kernels), <0 on error
int graph_agg_fwd_mma(int N, int K, int C, int T, int V, const void* y, long long yns, const float* A, void* out,
                      long long ons, double* ssum, double* ssq, cudaStream_t st) {
    if (sg_disabled() || K < 1 || K > 3 || (V != 20 && V != 25)) return 0;
    if ((reinterpret_cast<uintptr_t>(y) & 1) || (reinterpret_cast<uintptr_t>(out) & 1)) return 0;
    SgP p = {N, K, C, T, yns, ons, N * C, (T + 15) / 16};
    const size_t sgb = V == 20 ? SgCfg<20>::SGB : SgCfg<25>::SGB;
    const size_t sm = (size_t)SG_WARPS * (SG_D * K + 1) * sgb;
    const int grid = sg_grid(p, sm);
#define SG_FWD(VV, KK)                                                                                                 \
    do {                                                                                                                \
        static SmemLimit lim;                                                                                           \
        ensure_smem(graph_agg_fwd_mma_kernel<VV, KK>, lim, sm);                                                         \
        graph_agg_fwd_mma_kernel<VV, KK><<<grid, SG_THREADS, sm, st>>>(p, (const bf16*)y, A, (bf16*)out, ssum, ssq);    \
    } while (0)
    if (V == 20) { if (K == 1) SG_FWD(20, 1); else if (K == 2) SG_FWD(20, 2); else SG_FWD(20, 3); }
    else         { if (K == 1) SG_FWD(25, 1); else if (K == 2) SG_FWD(25, 2); else SG_FWD(25, 3); }
#undef SG_FWD
    count_launch();
    const int rc = check_launch("graph_agg_fwd(mma)");
    return rc < 0 ? rc : 1;
}

int graph_agg_bwd_mma(int N, int K, int C, int T, int V, const Opnd& go, const void* y, long long yns, const float* A,
                      void* dy, long long dyns, float* dA, cudaStream_t st) {
    if (sg_disabled() || K < 1 || K > 3 || (V != 20 && V != 25)) return 0;
    if ((reinterpret_cast<uintptr_t>(go.p) & 1) || (reinterpret_cast<uintptr_t>(go.q) & 1) || (reinterpret_cast<uintptr_t>(dy) & 1) ||
        (reinterpret_cast<uintptr_t>(y) & 1))
        return 0;
    SgP p = {N, K, C, T, yns, dyns, N * C, (T + 15) / 16};
    const size_t sgb = V == 20 ? SgCfg<20>::SGB : SgCfg<25>::SGB;
    if (dy) {
        const size_t sm = (size_t)SG_WARPS * (2 * SG_D + 1) * sgb;
        const int grid = sg_grid(p, sm);
#define SG_DY(VV, KK)                                                                                                  \
    do {                                                                                                                \
        static SmemLimit lim;                                                                                           \
        ensure_smem(graph_agg_dy_mma_kernel<VV, KK>, lim, sm);                                                          \
        graph_agg_dy_mma_kernel<VV, KK><<<grid, SG_THREADS, sm, st>>>(p, go, A, (bf16*)dy);                             \
    } while (0)
        if (V == 20) { if (K == 1) SG_DY(20, 1); else if (K == 2) SG_DY(20, 2); else SG_DY(20, 3); }
        else         { if (K == 1) SG_DY(25, 1); else if (K == 2) SG_DY(25, 2); else SG_DY(25, 3); }
#undef SG_DY
        count_launch();
        if (check_launch("graph_agg_bwd(dy, mma)") < 0) return -2;
    }
    if (dA) {
        const size_t sm = (size_t)((V * V * 4 + 15) & ~15) + (size_t)SG_WARPS * 3 * SG_D * sgb;
        int gx = sg_grid(p, sm) / K;
        if (gx < 1) gx = 1;
        dim3 grid(gx, K);
        if (V == 20) {
            static SmemLimit lim;
            ensure_smem(graph_agg_dA_mma_kernel<20>, lim, sm);
            graph_agg_dA_mma_kernel<20><<<grid, SG_THREADS, sm, st>>>(p, go, (const bf16*)y, dA);
        } else {
            static SmemLimit lim;
            ensure_smem(graph_agg_dA_mma_kernel<25>, lim, sm);
            graph_agg_dA_mma_kernel<25><<<grid, SG_THREADS, sm, st>>>(p, go, (const bf16*)y, dA);
        }
        count_launch();
        if (check_launch("graph_agg_bwd(dA, mma)") < 0) return -2;
    }
    return 1;
}

}  // namespace tamgcn
