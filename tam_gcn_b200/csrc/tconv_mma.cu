// tconv_mma.cu — small-channel (k x 1) temporal convolutions (the MS-TCN branch convolutions of
// models/ctrgcn.py:52-69,95-110: Cin = Cout = 16 / 32 / 64, k = 5, dilation 1|2, stride 1|2) on warp-level tensor-core
// MMAs (mma.sync m16n8k16, bf16 in, fp32 accumulate).  bf16 storage only.
//
// Why not the tcgen05 kernel (conv_tc2.cu): with 16..64 channels that kernel is producer-bound — every tap re-stages
// (and re-transforms) the activation operand, ~75 instructions per output element.  Here a CTA stages the lazy operand
// of one (sample, block of time steps) ONCE into shared memory as [channel][time][V padded to a multiple of 8]; a tap
// is then only a row offset (a multiple of 8 positions, i.e. 16-byte aligned for ldmatrix), and the whole contraction
// for 16..64 channels fits a few hundred MMAs per CTA.
//
//   forward / data gradient (stride 1):  out[oc, to, v] = sum_j sum_ic A_j[oc][ic] * in[ic, to*s + off_j, v]
//        fwd  : off_j = j*d - p,  A_j[oc][ic] = W[oc, ic, j]      (+ bias; BatchNorm sums of the stored values)
//        dgrad: off_j = p - j*d,  A_j[ci][co] = W[co, ci, j]      (ReLU mask from the saved activation + BN-backward sums)
//   weight gradient: dW[oc, ic, j] += sum_{n,to,v} dy[oc, to, v] * x[ic, to*s + j*d - p, v],  db[oc] += sum dy
//        persistent CTAs keep their dW tiles in registers over all their (sample, time block) units: one atomic per
//        weight per CTA at the end.
#include "common.cuh"
#include <cstdlib>
#include <type_traits>

namespace tamgcn {

#define TM_THREADS 256
#define TM_MAXK 9

struct TmP {
    int N, Tin, Tout, V, VP, k, s, TB, tps, RIN, omin;
    int off[TM_MAXK];
    int xpitch, opitch, ypitch;
    int mode;                    // 0 forward, 1 data gradient
    int vec_in, vec_in2, vec_out, stat_c0;
    long long ons;
};

__device__ __forceinline__ void tm_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tm_ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void tm_ldsm_x2(uint32_t& r0, uint32_t& r1, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];"
                 : "=r"(r0), "=r"(r1) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void tm_ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];"
                 : "=r"(r0), "=r"(r1) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ float tm_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float tm_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t tm_pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// tile[ch][pitch] (bf16, shared) <- rows [row0, row0 + nrows) x VP columns of the lazy operand `o`, sample n, channels
// 0..CB-1 of a (T x V) plane; rows outside [0, T) and columns >= V are zero.  coef: [3][CB] (a, b, c) in shared memory.
template <int CB>
__device__ __forceinline__ void tm_load_tile(bf16* __restrict__ tile, int pitch, const Opnd& o, const float* coef, int n,
                                             int T, int V, int VP, int row0, int nrows, int vec) {
    const bf16* pp = (const bf16*)o.p + (long long)n * o.pns;
    const bf16* qq = o.q ? (const bf16*)o.q + (long long)n * o.qns : nullptr;
    const long long TV = (long long)T * V;
    if (vec == 4) {
        const int VQ = VP >> 2, per = nrows * VQ, total = CB * per;
        for (int idx = threadIdx.x; idx < total; idx += TM_THREADS) {
            const int ch = idx / per, rem = idx - ch * per, r = rem / VQ, v = (rem - r * VQ) << 2;
            const int t = row0 + r;
            uint2 w = make_uint2(0u, 0u);
            if (t >= 0 && t < T && v < V) {
                const long long e = (long long)ch * TV + (long long)t * V + v;
                const uint2 x = __ldg(reinterpret_cast<const uint2*>(pp + e));
                const float a = coef[ch], c = coef[2 * CB + ch];
                float f0 = fmaf(a, tm_lo(x.x), c), f1 = fmaf(a, tm_hi(x.x), c), f2 = fmaf(a, tm_lo(x.y), c), f3 = fmaf(a, tm_hi(x.y), c);
                if (qq) {
                    const uint2 y = __ldg(reinterpret_cast<const uint2*>(qq + e));
                    const float b = coef[CB + ch];
                    f0 = fmaf(b, tm_lo(y.x), f0); f1 = fmaf(b, tm_hi(y.x), f1); f2 = fmaf(b, tm_lo(y.y), f2); f3 = fmaf(b, tm_hi(y.y), f3);
                }
                if (o.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); f2 = fmaxf(f2, 0.f); f3 = fmaxf(f3, 0.f); }
                w.x = tm_pack(f0, f1); w.y = tm_pack(f2, f3);
            }
            *reinterpret_cast<uint2*>(tile + (size_t)ch * pitch + r * VP + v) = w;
        }
    } else if (vec == 8) {
        // Rows that are not 8-byte aligned (V = 25: 50-byte rows).  The rows row0 .. row0 + nrows - 1 of a channel are ONE
        // contiguous element range in global memory, whatever V is: it is read in 16-byte pieces aligned in the GLOBAL
        // address space (head / tail elements masked), and each element goes to its (row, v) slot of the padded tile with a
        // 2-byte shared-memory store.  One division per 8 elements instead of two per element, 1/8 of the load instructions.
        const int r_lo = row0 < 0 ? -row0 : 0, r_hi = (T - row0) < nrows ? (T - row0) : nrows;   // valid rows [r_lo, r_hi)
        for (int idx = threadIdx.x; idx < CB * ((nrows * VP) >> 3); idx += TM_THREADS) {
            const int ch = idx / ((nrows * VP) >> 3), k = idx - ch * ((nrows * VP) >> 3);
            *reinterpret_cast<uint4*>(tile + (size_t)ch * pitch + 8 * k) = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        if (r_hi > r_lo) {
            const int g_lo = (row0 + r_lo) * V, g_hi = (row0 + r_hi) * V;           // element range inside a channel plane
            const int nck = ((g_hi - g_lo) >> 3) + 2;                               // 16-byte pieces per channel (upper bound)
            for (int idx = threadIdx.x; idx < CB * nck; idx += TM_THREADS) {
                const int ch = idx / nck, k = idx - ch * nck;
                const bf16* cp = pp + (long long)ch * TV;
                const uintptr_t a0 = reinterpret_cast<uintptr_t>(cp + g_lo) & ~(uintptr_t)15;
                const bf16* src = reinterpret_cast<const bf16*>(a0) + 8 * k;
                int g = (int)(src - cp);                                            // element index of the piece's first element
                if (g >= g_hi || g + 8 <= g_lo) continue;
                const uint4 x = __ldg(reinterpret_cast<const uint4*>(src));
                uint4 y = make_uint4(0u, 0u, 0u, 0u);
                if (qq) y = __ldg(reinterpret_cast<const uint4*>(qq + (long long)ch * TV + g));   // same alignment required of q (host check)
                const float a = coef[ch], b = coef[CB + ch], c = coef[2 * CB + ch];
                const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
                int t = g >= 0 ? g / V : -1, v = g - t * V;                          // g >= g_lo - 7 >= -7
                if (g < 0) { t = -1; v = g + V; }
                bf16* dst = tile + (size_t)ch * pitch;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (g + j >= g_lo && g + j < g_hi) {
                        float f = fmaf(a, (j & 1) ? tm_hi(xw[j >> 1]) : tm_lo(xw[j >> 1]), c);
                        if (qq) f = fmaf(b, (j & 1) ? tm_hi(yw[j >> 1]) : tm_lo(yw[j >> 1]), f);
                        if (o.relu) f = fmaxf(f, 0.f);
                        dst[(t - row0) * VP + v] = __float2bfloat16_rn(f);
                    }
                    if (++v == V) { v = 0; ++t; }
                }
            }
        }
    } else {
        const int per = nrows * VP, total = CB * per;
        for (int idx = threadIdx.x; idx < total; idx += TM_THREADS) {
            const int ch = idx / per, rem = idx - ch * per, r = rem / VP, v = rem - r * VP;
            const int t = row0 + r;
            float f = 0.f;
            if (t >= 0 && t < T && v < V) {
                const long long e = (long long)ch * TV + (long long)t * V + v;
                f = fmaf(coef[ch], ldf<bf16>(pp + e), coef[2 * CB + ch]);
                if (qq) f = fmaf(coef[CB + ch], ldf<bf16>(qq + e), f);
                if (o.relu) f = fmaxf(f, 0.f);
            }
            tile[(size_t)ch * pitch + r * VP + v] = __float2bfloat16_rn(f);
        }
    }
}

template <int CB>
__device__ __forceinline__ void tm_load_coef(float* coef, const Opnd& o) {
    for (int i = threadIdx.x; i < CB; i += TM_THREADS) {
        const OpCoef cf = opnd_coef(o, i);
        coef[i] = cf.a; coef[CB + i] = cf.b; coef[2 * CB + i] = cf.c;
    }
}

// ------------------------------------------------------------------------------------------------
// forward / data gradient
// ------------------------------------------------------------------------------------------------
template <int CB, int NTW>
__global__ void __launch_bounds__(TM_THREADS)
tconv_mma_kernel(TmP p, Opnd xo, const float* __restrict__ W, const float* __restrict__ bias, bf16* __restrict__ out,
                 Opnd mo, int has_mask, double* __restrict__ s1, double* __restrict__ s2) {
    constexpr int MT = CB / 16, KC = CB / 16, WP = CB + 8, RPW = CB / 8;
    extern __shared__ __align__(16) unsigned char tm_smem[];
    bf16* Wsm = reinterpret_cast<bf16*>(tm_smem);            // [k][CB][WP]   A_j[oc][ic]
    bf16* X = Wsm + (size_t)p.k * CB * WP;                   // [CB][xpitch]  staged operand
    bf16* O = X + (size_t)CB * p.xpitch;                     // [CB][opitch]  staged output rows
    float* coef = reinterpret_cast<float*>(O + (size_t)CB * p.opitch);   // [3][CB] operand, [CB] bias, [2][CB] mask a / c
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const int V = p.V, VP = p.VP;

    for (int idx = tid; idx < CB * CB * p.k; idx += TM_THREADS) {
        const int co = idx / (CB * p.k), rem = idx - co * CB * p.k, ci = rem / p.k, j = rem - ci * p.k;
        const bf16 w = __float2bfloat16_rn(__ldg(W + idx));
        if (p.mode == 0) Wsm[((size_t)j * CB + co) * WP + ci] = w;
        else Wsm[((size_t)j * CB + ci) * WP + co] = w;
    }
    tm_load_coef<CB>(coef, xo);
    for (int i = tid; i < CB; i += TM_THREADS) {
        coef[3 * CB + i] = (p.mode == 0 && bias) ? __ldg(bias + i) : 0.f;
        coef[4 * CB + i] = (has_mask && mo.a) ? __ldg(mo.a + i) : 1.f;
        coef[5 * CB + i] = (has_mask && mo.c) ? __ldg(mo.c + i) : 0.f;
    }
    float sa[RPW], sb[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) sa[i] = sb[i] = 0.f;

    const int units = p.N * p.tps;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const int n = unit / p.tps, t0 = (unit - n * p.tps) * p.TB;
        const int rows = min(p.TB, p.Tout - t0);
        __syncthreads();
        tm_load_tile<CB>(X, p.xpitch, xo, coef, n, p.Tin, V, VP, t0 * p.s + p.omin, p.RIN, p.vec_in);
        __syncthreads();
        for (int tl = warp; tl < rows; tl += TM_THREADS / 32) {
            float acc[MT][NTW][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTW; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
            for (int j = 0; j < p.k; ++j) {
                const int r = tl * p.s + p.off[j] - p.omin;
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    uint32_t b[NTW][2];
                    const bf16* xb = X + (size_t)(kc * 16 + (lane & 15)) * p.xpitch + r * VP;
#pragma unroll
                    for (int nt = 0; nt < NTW; ++nt) tm_ldsm_x2_trans(b[nt][0], b[nt][1], xb + nt * 8);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        uint32_t a[4];
                        tm_ldsm_x4(a, Wsm + ((size_t)j * CB + mt * 16 + (lane & 15)) * WP + kc * 16 + (lane >> 4) * 8);
#pragma unroll
                        for (int nt = 0; nt < NTW; ++nt) tm_mma(acc[mt][nt], a, b[nt][0], b[nt][1]);
                    }
                }
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int oc = mt * 16 + gid + 8 * (e >> 1), v = nt * 8 + 2 * tig + (e & 1);
                        if (v < V) O[(size_t)oc * p.opitch + tl * V + v] = __float2bfloat16_rn(acc[mt][nt][e] + coef[3 * CB + oc]);
                    }
        }
        __syncthreads();
        // ---- copy-out: one warp per channel row (rows * V contiguous elements); mask and BatchNorm sums on the way
        const int L = rows * V;
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
            const int oc = warp + 8 * i;
            const bf16* src = O + (size_t)oc * p.opitch;
            const long long go = ((long long)oc * p.Tout + t0) * V;
            bf16* dst = out + (long long)n * p.ons + go;
            const bf16* mk = has_mask ? (const bf16*)mo.p + (long long)n * mo.pns + go : nullptr;
            const float ma = coef[4 * CB + oc], mc = coef[5 * CB + oc];
            float a1 = 0.f, a2 = 0.f;
            if (p.vec_out == 4) {
                for (int e = lane * 4; e < L; e += 128) {
                    uint2 w = *reinterpret_cast<const uint2*>(src + e);
                    float f[4] = {tm_lo(w.x), tm_hi(w.x), tm_lo(w.y), tm_hi(w.y)};
                    if (mk) {
                        const uint2 hw = __ldg(reinterpret_cast<const uint2*>(mk + e));
                        const float h[4] = {tm_lo(hw.x), tm_hi(hw.x), tm_lo(hw.y), tm_hi(hw.y)};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (!(fmaf(ma, h[u], mc) > 0.f)) f[u] = 0.f;
                            a1 += f[u]; a2 = fmaf(f[u], h[u], a2);
                        }
                        w.x = tm_pack(f[0], f[1]); w.y = tm_pack(f[2], f[3]);
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; ++u) { a1 += f[u]; a2 = fmaf(f[u], f[u], a2); }
                    }
                    *reinterpret_cast<uint2*>(dst + e) = w;
                }
            } else {
                for (int e = lane; e < L; e += 32) {
                    float f = __bfloat162float(src[e]);
                    if (mk) {
                        const float h = ldf<bf16>(mk + e);
                        if (!(fmaf(ma, h, mc) > 0.f)) f = 0.f;
                        a1 += f; a2 = fmaf(f, h, a2);
                    } else {
                        a1 += f; a2 = fmaf(f, f, a2);
                    }
                    dst[e] = __float2bfloat16_rn(f);
                }
            }
            sa[i] += a1; sb[i] += a2;
        }
    }
    if (s1) {
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
            const float t1 = warp_sum(sa[i]), t2 = warp_sum(sb[i]);
            const int oc = warp + 8 * i;
            if (lane == 0 && oc >= p.stat_c0 && (t1 != 0.f || t2 != 0.f)) {
                atomicAdd(s1 + oc - p.stat_c0, (double)t1);
                atomicAdd(s2 + oc - p.stat_c0, (double)t2);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
template <int CB, int NTW, int KS>
__global__ void __launch_bounds__(TM_THREADS)
tconv_wgrad_mma_kernel(TmP p, Opnd dyo, Opnd xo, float* __restrict__ dW, float* __restrict__ db) {
    constexpr int MT = CB / 16, NT8 = CB / 8, TILES = KS * MT * NT8, TPW = (TILES + 7) / 8, RPW = CB / 8;
    extern __shared__ __align__(16) unsigned char tm_smem[];
    bf16* Y = reinterpret_cast<bf16*>(tm_smem);              // [CB][ypitch]  dy rows of the unit (zero padded)
    bf16* X = Y + (size_t)CB * p.ypitch;                     // [CB][xpitch]
    float* coef = reinterpret_cast<float*>(X + (size_t)CB * p.xpitch);   // [3][CB] dy, [3][CB] x
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const int VP = p.VP;
    tm_load_coef<CB>(coef, dyo);
    tm_load_coef<CB>(coef + 3 * CB, xo);
    float acc[TPW][4];
#pragma unroll
    for (int i = 0; i < TPW; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    float dbacc[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) dbacc[i] = 0.f;
    const int NB = p.TB * NTW, KB = (NB + 1) >> 1;          // 8-position blocks / 16-position K steps of a unit
    const int units = p.N * p.tps;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const int n = unit / p.tps, t0 = (unit - n * p.tps) * p.TB;
        __syncthreads();
        tm_load_tile<CB>(Y, p.ypitch, dyo, coef, n, p.Tout, p.V, VP, t0, p.TB, p.vec_in);
        if (NB & 1)
            for (int idx = tid; idx < CB * 8; idx += TM_THREADS) Y[(size_t)(idx >> 3) * p.ypitch + NB * 8 + (idx & 7)] = __float2bfloat16_rn(0.f);
        tm_load_tile<CB>(X, p.xpitch, xo, coef + 3 * CB, n, p.Tin, p.V, VP, t0 * p.s + p.omin, p.RIN, p.vec_in2);
        __syncthreads();
        if (db) {
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const bf16* yr = Y + (size_t)(warp + 8 * i) * p.ypitch;
                float s = 0.f;
                for (int e = lane * 2; e < NB * 8; e += 64) {
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(yr + e);
                    s += tm_lo(w) + tm_hi(w);
                }
                dbacc[i] += s;
            }
        }
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
            const int tile = warp + 8 * i;
            if (tile < TILES) {
                const int j = tile / (MT * NT8), rem = tile - j * (MT * NT8), mt = rem / NT8, n8 = rem - mt * NT8;
                const bf16* ya = Y + (size_t)(mt * 16 + (lane & 15)) * p.ypitch + (lane >> 4) * 8;
                const bf16* xr = X + (size_t)(n8 * 8 + (lane & 7)) * p.xpitch;
                const int roff = p.off[j] - p.omin;
                for (int kb = 0; kb < KB; ++kb) {
                    uint32_t a[4], b0, b1;
                    tm_ldsm_x4(a, ya + kb * 16);
                    int blk = 2 * kb + ((lane >> 3) & 1);
                    if (blk >= NB) blk = NB - 1;                   // tail block: dy is zero there, any finite x will do
                    const int tl = blk / NTW, c = blk - tl * NTW;
                    tm_ldsm_x2(b0, b1, xr + (tl * p.s + roff) * VP + c * 8);
                    tm_mma(acc[i], a, b0, b1);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
        const int tile = warp + 8 * i;
        if (tile < TILES) {
            const int j = tile / (MT * NT8), rem = tile - j * (MT * NT8), mt = rem / NT8, n8 = rem - mt * NT8;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int oc = mt * 16 + gid + 8 * (e >> 1), ic = n8 * 8 + 2 * tig + (e & 1);
                if (acc[i][e] != 0.f) atomicAdd(dW + ((long long)oc * CB + ic) * KS + j, acc[i][e]);
            }
        }
    }
    if (db) {
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
            const float s = warp_sum(dbacc[i]);
            if (lane == 0 && s != 0.f) atomicAdd(db + warp + 8 * i, s);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static bool tm_disabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TAMGCN_DISABLE_TCONV_MMA");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
static int tm_num_sms() { return num_sms(); }
static bool tm_al8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }
static int tm_vec(const Opnd& o, int V) {
    bool ok = (V % 4 == 0) && (o.pns % 4 == 0) && tm_al8(o.p);
    if (o.q) ok = ok && (o.qns % 4 == 0) && tm_al8(o.q);
    if (ok) return 4;
    // flat 16-byte pieces (tm_load_tile, vec == 8): with a second tensor both must share the 16-byte phase of every
    // channel plane, which holds when both bases are 16-byte aligned and the sample strides are multiples of 8 elements
    static const int flat = [] { const char* e = getenv("TAMGCN_TCONV_FLAT"); return e ? atoi(e) : 1; }();
    bool okf = flat && (o.pns % 8 == 0) && ((reinterpret_cast<uintptr_t>(o.p) & 15) == 0);
    if (o.q) okf = okf && (o.qns % 8 == 0) && ((reinterpret_cast<uintptr_t>(o.q) & 15) == 0);
    return okf ? 8 : 1;
}
static int tm_oddpitch(int elems) {                // multiple of 8 elements with an odd number of 16-byte pieces
    int p = (elems + 7) & ~7;
    if (((p >> 3) & 1) == 0) p += 8;
    return p;
}

// shared geometry: taps, time blocking.  kind: 0 fwd, 1 dgrad, 2 wgrad.  Returns false when the shape is not covered.
static bool tm_setup(TmP& p, const tamgcn_conv_geom* g, int kind, size_t& smem) {
    const int CB = g->Cin;
    if (tm_disabled() || g->Cin != g->Cout || (CB != 16 && CB != 32 && CB != 64)) return false;
    if (g->k < 2 || g->k > TM_MAXK || g->V > 32 || g->V < 1) return false;
    if (kind == 1 && g->stride != 1) return false;
    if (kind == 2 && g->k != 3 && g->k != 5) return false;
    p = TmP{};
    p.N = g->N; p.V = g->V; p.VP = (g->V + 7) & ~7; p.k = g->k; p.s = g->stride;
    p.mode = kind == 1 ? 1 : 0;
    if (kind == 1) { p.Tin = g->To; p.Tout = g->T; } else { p.Tin = g->T; p.Tout = g->To; }
    int omin = 1 << 30, omax = -(1 << 30);
    for (int j = 0; j < g->k; ++j) {
        p.off[j] = kind == 1 ? g->pad - j * g->dil : j * g->dil - g->pad;
        omin = p.off[j] < omin ? p.off[j] : omin;
        omax = p.off[j] > omax ? p.off[j] : omax;
    }
    p.omin = omin;
    const int NTW = p.VP / 8;
    int tps = (256 + p.N - 1) / p.N;
    if (tps < 1) tps = 1;
    if (tps > p.Tout) tps = p.Tout;
    if (kind == 2) {
        // weight gradient: a persistent grid of one CTA per SM walks N * tps units; pick the time blocking with the fewest
        // (rounds x input rows per unit) — longer blocks amortise the tap halo and avoid a second, mostly empty round
        static const int wg_search = [] { const char* e = getenv("TAMGCN_TCONV_WG_SEARCH"); return e ? atoi(e) : 1; }();
        if (wg_search) {
            long long best = -1;
            int best_tps = tps;
            for (int cand = 1; cand <= 16 && cand <= p.Tout; ++cand) {
                const int TB = (p.Tout + cand - 1) / cand, RIN = (TB - 1) * p.s + 1 + (omax - omin);
                const size_t sm_c = (size_t)CB * tm_oddpitch(((TB * NTW + 1) / 2) * 16) * 2 + (size_t)CB * tm_oddpitch(RIN * p.VP) * 2 + 6 * CB * 4;
                if (sm_c > 100 * 1024) continue;
                const long long units = (long long)p.N * ((p.Tout + TB - 1) / TB);
                const long long rounds = (units + wgrad_sms() - 1) / wgrad_sms();
                const long long cost = rounds * (RIN + 4);
                if (best < 0 || cost < best) { best = cost; best_tps = cand; }
            }
            tps = best_tps;
        }
    }
    for (;;) {
        p.TB = (p.Tout + tps - 1) / tps;
        p.RIN = (p.TB - 1) * p.s + 1 + (omax - omin);
        p.xpitch = tm_oddpitch(p.RIN * p.VP);
        p.opitch = ((p.TB * p.V + 3) & ~3) + 4;
        p.ypitch = tm_oddpitch(((p.TB * NTW + 1) / 2) * 16);
        const size_t xs = (size_t)CB * p.xpitch * 2;
        if (kind == 2) smem = (size_t)CB * p.ypitch * 2 + xs + 6 * CB * 4;
        else smem = (size_t)p.k * CB * (CB + 8) * 2 + xs + (size_t)CB * p.opitch * 2 + 6 * CB * 4;
        if (smem <= 100 * 1024 || p.TB == 1) break;
        ++tps;
    }
    if (smem > 200 * 1024) return false;
    p.tps = (p.Tout + p.TB - 1) / p.TB;
    return true;
}


// compile-time (channel count, 8-wide position tiles per row) from the run-time shape
template <typename F>
static void tm_dispatch(int CB, int NTW, F&& f) {
    auto by_ntw = [&](auto cb) {
        switch (NTW) {
            case 1: f(cb, std::integral_constant<int, 1>{}); break;
            case 2: f(cb, std::integral_constant<int, 2>{}); break;
            case 3: f(cb, std::integral_constant<int, 3>{}); break;
            default: f(cb, std::integral_constant<int, 4>{}); break;
        }
    };
    if (CB == 16) by_ntw(std::integral_constant<int, 16>{});
    else if (CB == 32) by_ntw(std::integral_constant<int, 32>{});
    else by_ntw(std::integral_constant<int, 64>{});
}

// return 1 if handled, 0 if the caller should use another kernel, <0 on error
int tconv_mma_fwd_dgrad(const tamgcn_conv_geom* g, int kind, const Opnd& in, const float* W, const float* bias, void* out,
                        long long ons, const Opnd* mask, double* s1, double* s2, int stat_c0, cudaStream_t st) {
    TmP p;
    size_t sm = 0;
    if (!W || !tm_setup(p, g, kind, sm)) return 0;
    if (mask && (mask->q || mask->b)) return 0;
    p.ons = ons;
    p.stat_c0 = stat_c0;
    p.vec_in = tm_vec(in, p.V);
    const long long row = (long long)p.Tout * p.V;
    bool vo = (p.V % 4 == 0) && (ons % 4 == 0) && tm_al8(out) && (row % 4 == 0);
    if (mask) vo = vo && (mask->pns % 4 == 0) && tm_al8(mask->p);
    p.vec_out = vo ? 4 : 1;
    const int CB = g->Cin, NTW = p.VP / 8;
    long long grid = (long long)p.N * p.tps;
    const long long cap = 2LL * tm_num_sms();
    if (grid > cap) grid = cap;
    Opnd mo = mask ? *mask : plain_opnd(nullptr, 0);
    tm_dispatch(CB, NTW, [&](auto cb, auto ntw) {
        constexpr int CB_ = decltype(cb)::value, NTW_ = decltype(ntw)::value;
        static SmemLimit lim;
        ensure_smem(tconv_mma_kernel<CB_, NTW_>, lim, sm);
        tconv_mma_kernel<CB_, NTW_><<<(int)grid, TM_THREADS, sm, st>>>(p, in, W, bias, (bf16*)out, mo, mask ? 1 : 0, s1, s2);
    });
    count_launch();
    const int rc = check_launch(kind == 1 ? "conv_dgrad(mma)" : "conv_fwd(mma)");
    return rc < 0 ? rc : 1;
}

int tconv_mma_wgrad(const tamgcn_conv_geom* g, const Opnd& dy, const Opnd& x, float* dW, float* db, cudaStream_t st) {
    TmP p;
    size_t sm = 0;
    if (!tm_setup(p, g, 2, sm)) return 0;
    p.vec_in = tm_vec(dy, p.V);
    p.vec_in2 = tm_vec(x, p.V);
    const int CB = g->Cin, NTW = p.VP / 8;
    long long grid = (long long)p.N * p.tps;
    if (grid > wgrad_sms()) grid = wgrad_sms();
    tm_dispatch(CB, NTW, [&](auto cb, auto ntw) {
        constexpr int CB_ = decltype(cb)::value, NTW_ = decltype(ntw)::value;
        if (g->k == 5) {
            static SmemLimit lim;
            ensure_smem(tconv_wgrad_mma_kernel<CB_, NTW_, 5>, lim, sm);
            tconv_wgrad_mma_kernel<CB_, NTW_, 5><<<(int)grid, TM_THREADS, sm, st>>>(p, dy, x, dW, db);
        } else {
            static SmemLimit lim;
            ensure_smem(tconv_wgrad_mma_kernel<CB_, NTW_, 3>, lim, sm);
            tconv_wgrad_mma_kernel<CB_, NTW_, 3><<<(int)grid, TM_THREADS, sm, st>>>(p, dy, x, dW, db);
        }
    });
    count_launch();
    const int rc = check_launch("conv_wgrad(mma)");
    return rc < 0 ? rc : 1;
}

}  // namespace tamgcn

namespace tamgcn { bool tconv9_covers(int Cin, int Cout, int k, int stride, int dil, int pad, int V, int dgrad); }

extern "C" int tamgcn_conv_needs_pack(int Cin, int Cout, int k, int stride, int V, int dgrad) {
    // the V-padded tcgen05 kernel (tconv9.cu) takes "same" temporal convolutions of V = 25 first and reads packed tiles
    if ((k & 1) && tamgcn::tconv9_covers(Cin, Cout, k, stride, 1, (k - 1) / 2, V, dgrad)) return 1;
    // mirrors tm_setup: the MMA kernels read the fp32 weights directly, the tcgen05 kernels need the packed tiles
    if (tamgcn::tm_disabled()) return 1;
    const bool small = (Cin == Cout) && (Cin == 16 || Cin == 32 || Cin == 64) && k >= 2 && k <= TM_MAXK && V <= 32;
    if (!small) return 1;
    return (dgrad && stride != 1) ? 1 : 0;
}
