// ctrgc_tc.cu — fused CTRGC forward for bf16 activations on the 5th-generation tensor cores.
//
//   y[n,c,t,u] = sum_i sum_v Q_i[n,c,u,v] * x3_i[n,c,t,v]      (reference models/ctrgcn.py:174-176, :252-254)
//   Q_i[n,c,u,v] = alpha * (sum_r W4_i[c,r] tanh(x1_i[n,r,u] - x2_i[n,r,v]) + b4_i[c]) + PA_i[u,v]
//
// The kernel is bound by how fast x3 can be streamed from HBM (15 FLOP/B), but the 2*K*V FMAs per output
// element do not fit the fp32 SIMT pipe at that rate (74 TFLOP/s fp32 vs ~100 TFLOP/s needed at 6.5 TB/s),
// so the V x V contraction runs as tcgen05.mma with the accumulator in TMEM; everything else is arranged so
// that the CUDA cores only move bytes and build the tiny per-(n,c) topology matrices:
//
//   tile      = one sample n, G output channels, TR time steps (G*TR = 128 = UMMA M)
//   A (smem)  = x3 rows  [(g,t)] x [(i,v)]   bf16, K-major, SWIZZLE_128B — gathered straight from the K planes
//               of x3 with 8-byte cp.async (V=20) / vector loads (V=25); never touches registers for V=20
//   B (smem)  = Q rows   [(g,u)] x [(i,v)]   bf16, K-major, SWIZZLE_128B — built by the CUDA cores from the
//               per-sample tanh table D (fp16, shared memory) and the W4/b4/PA/alpha parameters
//   D (TMEM)  = [(g,t)] x [(g',u)] fp32; only the diagonal blocks g == g' are read back (the off-diagonal
//               products are wasted tensor-pipe work, which is idle anyway)
//
// Persistent, warp-specialised CTA (one per SM), mbarrier pipelines:
//   warps 0-3   epilogue: TMEM -> registers -> bf16 -> staging smem -> coalesced 16-byte stores; BatchNorm
//               sum / sum-of-squares by warp shuffles, per-CTA smem accumulators, one fp64 atomic per channel
//   warp  4     MMA issuer (one lane) + TMEM allocation; accumulators double buffered
//   warps 5-12  two loader groups that alternate tiles (keeps >= 2 tiles of loads in flight per SM)
//   warps 13-22 topology builders: tanh table once per sample, Q tile per (sample, channel group)
// The N x C x V x V topology tensor exists only in shared memory.
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace tamgcn {

template <int V> struct CtcCfg;
template <> struct CtcCfg<20> { static const int VS = 20, VN = 24, VP = 20, NQ = 5; };
template <> struct CtcCfg<25> { static const int VS = 32, VN = 32, VP = 28, NQ = 7; };

#define CTC_SMAX 4
#define CTC_EPI_T 128
#define CTC_MMA_W 4
#define CTC_LD_T0 160            // first loader thread
#define CTC_LDG_T 128            // threads per loader group
#define CTC_Q_T0 416             // first topology-builder thread
#define CTC_Q_T 320
#define CTC_THREADS (CTC_Q_T0 + CTC_Q_T)   // 736

struct CtcP {
    int N, Cout, T, K, R;
    long long x3ns, x12ns, yns;
    int TR, G, nTC, nCG, tps, n_tiles;
    int S, NMMA, Nmma, NKB, tmem_cols;
    int av, ov;
    unsigned kmagic;                 // 65536 / K + 1
    uint32_t a_bytes, b_blk_bytes, stage_bytes;
    uint32_t off_D, off_PA, off_W4, off_out, off_stat, off_hdr, off_x12;
    int w4t_floats;
    int dbg;                         // TAMGCN_CTC_DBG bit mask (profiling aid): 1 skip x3 copies, 2 skip Q math, 4 skip copy-out
};

// (sample, channel group, time chunk) of consecutive tiles without integer divisions
struct CtcTile {
    int n, cg, tc;
    __device__ __forceinline__ void init(const CtcP& p, int tile) {
        n = tile / p.tps;
        const int rem = tile - n * p.tps;
        cg = rem / p.nTC;
        tc = rem - cg * p.nTC;
    }
    __device__ __forceinline__ void next(const CtcP& p) {
        if (++tc == p.nTC) { tc = 0; if (++cg == p.nCG) { cg = 0; ++n; } }
    }
};

struct CtcHdr {
    uint64_t a_full[CTC_SMAX], b_full[CTC_SMAX], empty[CTC_SMAX], tfull[2], tempty[2];
    uint32_t tmem_base;
    volatile uint32_t error;
};

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void st_shared_v2(uint32_t saddr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_shared_u16(uint32_t saddr, unsigned short a) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(a) : "memory");
}
// 32 lanes x 8 consecutive fp32 columns, no wait
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool ctc_wait(CtcHdr* hdr, uint64_t* bar, uint32_t parity) {
    if (hdr->error) return false;
    if (!mbar_wait(bar, parity)) { hdr->error = 1; return false; }
    return true;
}

// ---- topology builder: Q tile of one (sample, channel group) into the B stage -----------------------
template <int V, int G>
__device__ __forceinline__ void ctc_build_q(const CtcP& p, int qt, const __half* __restrict__ Dsm, const float* __restrict__ PAs,
                                            const float* __restrict__ w4t, float alpha, uint32_t sB) {
    typedef CtcCfg<V> Cf;
    const int R = p.R, K = p.K;
    const float* b4t = w4t + K * R * G;
    const int ntask = K * V * Cf::NQ;
    for (int task = qt; task < ntask; task += CTC_Q_T) {
        const int i = task / (V * Cf::NQ), rem = task - i * (V * Cf::NQ), u = rem / Cf::NQ, vq = rem - u * Cf::NQ;
        float acc[G][4];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float b = b4t[i * G + g];
            acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = b;
        }
        const __half* dp = Dsm + ((size_t)(i * R) * V + u) * Cf::VP + 4 * vq;
        const float* wp = w4t + i * R * G;
#pragma unroll 4
        for (int r = 0; r < R; ++r) {
            const uint2 dd = *reinterpret_cast<const uint2*>(dp + (size_t)r * V * Cf::VP);
            const float2 d01 = __half22float2(*reinterpret_cast<const __half2*>(&dd.x));
            const float2 d23 = __half22float2(*reinterpret_cast<const __half2*>(&dd.y));
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float w = wp[r * G + g];
                acc[g][0] = fmaf(w, d01.x, acc[g][0]);
                acc[g][1] = fmaf(w, d01.y, acc[g][1]);
                acc[g][2] = fmaf(w, d23.x, acc[g][2]);
                acc[g][3] = fmaf(w, d23.y, acc[g][3]);
            }
        }
        const float4 pa = *reinterpret_cast<const float4*>(PAs + (i * V + u) * Cf::VP + 4 * vq);
        const int col = i * Cf::VS + 4 * vq, blk = col >> 6, cb = col & 63;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            float q0 = fmaf(alpha, acc[g][0], pa.x), q1 = fmaf(alpha, acc[g][1], pa.y);
            float q2 = fmaf(alpha, acc[g][2], pa.z), q3 = fmaf(alpha, acc[g][3], pa.w);
            if (V == 25 && vq == Cf::NQ - 1) q1 = q2 = q3 = 0.f;          // v = 25..27 are K padding
            const int row = g * Cf::VN + u;
            const uint32_t a = sB + (uint32_t)blk * p.b_blk_bytes + (uint32_t)row * 128u +
                               ((uint32_t)((cb >> 3) ^ (row & 7)) << 4) + (uint32_t)(cb & 7) * 2u;
            st_shared_v2(a, pack_bf16(q0, q1), pack_bf16(q2, q3));
        }
    }
}

template <int V>
__global__ void __launch_bounds__(CTC_THREADS, 1)
ctrgc_fwd_tc_kernel(CtcP p, const bf16* __restrict__ x3, const float* __restrict__ x1, const float* __restrict__ x2,
                    const float* __restrict__ W4, const float* __restrict__ b4, const float* __restrict__ PA,
                    const float* __restrict__ alpha_p, bf16* __restrict__ y, double* ssum, double* ssq) {
    typedef CtcCfg<V> Cf;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
    CtcHdr* hdr = (CtcHdr*)(smem + p.off_hdr);
    __half* Dsm = (__half*)(smem + p.off_D);
    float* PAs = (float*)(smem + p.off_PA);
    float* W4t = (float*)(smem + p.off_W4);
    uint8_t* outs = smem + p.off_out;
    float* stat = (float*)(smem + p.off_stat);       // [2][Cout]
    float* x12s = (float*)(smem + p.off_x12);        // [2][K*R*V]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile_begin = (int)((long long)blockIdx.x * p.n_tiles / gridDim.x);
    const int tile_end = (int)((long long)(blockIdx.x + 1) * p.n_tiles / gridDim.x);
    const int nt = tile_end - tile_begin;
    const int S = p.S;

    // ---- one-time setup ----
    if (warp == CTC_MMA_W) tmem_alloc(&hdr->tmem_base, (uint32_t)p.tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < CTC_SMAX; ++i) {
            mbar_init(&hdr->a_full[i], CTC_LDG_T);
            mbar_init(&hdr->b_full[i], CTC_Q_T);
            mbar_init(&hdr->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&hdr->tfull[i], 1); mbar_init(&hdr->tempty[i], CTC_EPI_T); }
        hdr->error = 0;
        fence_mbar_init();
    }
    // operand stages start as zeros: K padding columns are never written afterwards and must not hold NaN bit patterns
    for (uint32_t i = tid; i < (uint32_t)S * p.stage_bytes / 16; i += CTC_THREADS)
        st_shared_v4(smem_u32(smem) + i * 16, 0u, 0u, 0u, 0u);
    for (int i = tid; i < p.K * V * Cf::VP; i += CTC_THREADS) {
        const int v = i % Cf::VP, iu = i / Cf::VP;
        PAs[i] = (v < V) ? __ldg(PA + iu * V + v) : 0.f;
    }
    for (int i = tid; i < 2 * p.Cout; i += CTC_THREADS) stat[i] = 0.f;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = hdr->tmem_base;
    const uint32_t s0 = smem_u32(smem);

    if (warp < 4) {
        // =============================== epilogue ===============================
        const int TR = p.TR, G = p.G;
        const int row = tid;                                   // TMEM lane == tile row (g, t)
        const int g_row = row / TR, tl = row - g_row * TR;
        const int width = TR < 32 ? TR : 32;                   // lanes of a warp that share a channel
        CtcTile tl_;
        tl_.init(p, tile_begin);
        for (int it = 0; it < nt; ++it, tl_.next(p)) {
            const int n = tl_.n;
            const int c0 = tl_.cg * G, t0 = tl_.tc * TR;
            const int TRv = min(TR, p.T - t0), Gv = min(G, p.Cout - c0);
            const int buf = it & 1;
            ctc_wait(hdr, &hdr->tfull[buf], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            float acc[Cf::VN];
            const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * p.Nmma);
            if (TR >= 32) {
                const uint32_t ta = tbase + (uint32_t)(min(g_row, G - 1) * Cf::VN);
                if (V == 20) {
                    tmem_ld8_nowait(ta, acc); tmem_ld8_nowait(ta + 8, acc + 8); tmem_ld8_nowait(ta + 16, acc + 16);
                    tmem_wait_ld();
                } else {
                    tmem_ld32(ta, *reinterpret_cast<float(*)[32]>(acc));
                }
            } else {
#pragma unroll
                for (int j = 0; j < Cf::VN; ++j) acc[j] = 0.f;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float tmp[Cf::VN];
                    const uint32_t ta = tbase + (uint32_t)(min(warp * 2 + h, G - 1) * Cf::VN);
                    if (V == 20) {
                        tmem_ld8_nowait(ta, tmp); tmem_ld8_nowait(ta + 8, tmp + 8); tmem_ld8_nowait(ta + 16, tmp + 16);
                        tmem_wait_ld();
                    } else {
                        tmem_ld32(ta, *reinterpret_cast<float(*)[32]>(tmp));
                    }
                    const bool mine = (lane >> 4) == h;
#pragma unroll
                    for (int j = 0; j < Cf::VN; ++j) acc[j] = mine ? tmp[j] : acc[j];
                }
            }
            tc_fence_before();
            mbar_arrive(&hdr->tempty[buf]);                    // the MMA warp may refill this accumulator

            const bool valid = (g_row < Gv) && (tl < TRv);
            float s = 0.f, q = 0.f;
            uint8_t* ob = outs + (size_t)buf * (128 * V * 2);
            if (valid) {
                const uint32_t oa = smem_u32(ob) + (uint32_t)((g_row * TRv + tl) * V * 2);
                if (V == 20) {
#pragma unroll
                    for (int j = 0; j < 20; j += 4) {
                        const uint32_t w0 = pack_bf16(acc[j], acc[j + 1]), w1 = pack_bf16(acc[j + 2], acc[j + 3]);
                        const float f0 = __uint_as_float(w0 << 16), f1 = __uint_as_float(w0 & 0xffff0000u);
                        const float f2 = __uint_as_float(w1 << 16), f3 = __uint_as_float(w1 & 0xffff0000u);
                        s += (f0 + f1) + (f2 + f3);
                        q = fmaf(f0, f0, q); q = fmaf(f1, f1, q); q = fmaf(f2, f2, q); q = fmaf(f3, f3, q);
                        st_shared_v2(oa + j * 2, w0, w1);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 25; ++j) {
                        const __nv_bfloat16 hb = __float2bfloat16_rn(acc[j]);
                        const float f = __bfloat162float(hb);
                        s += f;
                        q = fmaf(f, f, q);
                        st_shared_u16(oa + j * 2, *reinterpret_cast<const unsigned short*>(&hb));
                    }
                }
            }
            if (ssum) {
                for (int o = width >> 1; o > 0; o >>= 1) {
                    s += __shfl_xor_sync(0xffffffffu, s, o);
                    q += __shfl_xor_sync(0xffffffffu, q, o);
                }
                if ((lane & (width - 1)) == 0 && g_row < Gv) {
                    atomicAdd(&stat[c0 + g_row], s);
                    atomicAdd(&stat[p.Cout + c0 + g_row], q);
                }
            }
            bar_sync(1, CTC_EPI_T);
            // coalesced copy-out of the staged tile
            const int ov = (p.dbg & 4) ? 0 : p.ov;
            if (ov == 0) {
            } else if (p.nTC == 1) {
                const int bytes = Gv * TRv * V * 2;
                uint8_t* dst = (uint8_t*)(y + (long long)n * p.yns + (long long)c0 * p.T * V);
                if (ov == 16) {
                    for (int j = tid; j < bytes / 16; j += CTC_EPI_T) reinterpret_cast<uint4*>(dst)[j] = reinterpret_cast<const uint4*>(ob)[j];
                } else if (ov == 8) {
                    for (int j = tid; j < bytes / 8; j += CTC_EPI_T) reinterpret_cast<uint2*>(dst)[j] = reinterpret_cast<const uint2*>(ob)[j];
                } else {
                    for (int j = tid; j < bytes / 2; j += CTC_EPI_T) reinterpret_cast<unsigned short*>(dst)[j] = reinterpret_cast<const unsigned short*>(ob)[j];
                }
            } else {
                const int bytes = TRv * V * 2;
                for (int g = 0; g < Gv; ++g) {
                    uint8_t* dst = (uint8_t*)(y + (long long)n * p.yns + ((long long)(c0 + g) * p.T + t0) * V);
                    const uint8_t* src = ob + (size_t)g * bytes;
                    if (ov == 16) {
                        for (int j = tid; j < bytes / 16; j += CTC_EPI_T) reinterpret_cast<uint4*>(dst)[j] = reinterpret_cast<const uint4*>(src)[j];
                    } else if (ov == 8) {
                        for (int j = tid; j < bytes / 8; j += CTC_EPI_T) reinterpret_cast<uint2*>(dst)[j] = reinterpret_cast<const uint2*>(src)[j];
                    } else {
                        for (int j = tid; j < bytes / 2; j += CTC_EPI_T) reinterpret_cast<unsigned short*>(dst)[j] = reinterpret_cast<const unsigned short*>(src)[j];
                    }
                }
            }
        }
        if (ssum) {
            bar_sync(1, CTC_EPI_T);
            for (int c = tid; c < p.Cout; c += CTC_EPI_T) {
                const float a = stat[c], b = stat[p.Cout + c];
                if (a != 0.f || b != 0.f) {
                    atomicAdd(ssum + c, (double)a);
                    atomicAdd(ssq + c, (double)b);
                }
            }
        }
    } else if (warp == CTC_MMA_W) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)p.Nmma);
            for (int it = 0; it < nt; ++it) {
                const int s = it % S, ph = (it / S) & 1, buf = it & 1;
                ctc_wait(hdr, &hdr->tempty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
                ctc_wait(hdr, &hdr->a_full[s], (uint32_t)ph);
                ctc_wait(hdr, &hdr->b_full[s], (uint32_t)ph);
                tc_fence_after();
                const uint32_t sA = s0 + (uint32_t)s * p.stage_bytes, sB = sA + p.a_bytes;
                const uint32_t td = tmem + (uint32_t)(buf * p.Nmma);
                for (int j = 0; j < p.NMMA; ++j) {
                    const uint32_t blk = (uint32_t)j >> 2, kk = (uint32_t)j & 3;
                    umma_bf16(td, umma_desc_sw128(sA + blk * 16384u + kk * 32u), umma_desc_sw128(sB + blk * p.b_blk_bytes + kk * 32u),
                              idesc, j > 0 ? 1u : 0u);
                }
                umma_commit(&hdr->empty[s]);
                umma_commit(&hdr->tfull[buf]);
            }
        }
    } else if (tid < CTC_Q_T0) {
        // =============================== x3 loaders (two groups alternate tiles) ===============================
        const int lg = (tid - CTC_LD_T0) / CTC_LDG_T, lt = (tid - CTC_LD_T0) % CTC_LDG_T;
        const int TR = p.TR, G = p.G, K = p.K;
        const bool lagged = (V == 20) && (S >= 3);
        int prev_s = -1;
        CtcTile tl_;
        tl_.init(p, tile_begin + lg);
        for (int it = lg; it < nt; it += 2, tl_.next(p), tl_.next(p)) {
            const int n = tl_.n;
            const int c0 = tl_.cg * G, t0 = tl_.tc * TR;
            const int TRv = min(TR, p.T - t0), Gv = min(G, p.Cout - c0);
            const int s = it % S, ph = (it / S) & 1;
            const uint32_t sA = s0 + (uint32_t)s * p.stage_bytes;
            const bf16* xn = x3 + (long long)n * p.x3ns;
            if (V == 20) {
                const unsigned nu = (unsigned)TRv * 5u, total = (unsigned)(Gv * K) * nu;
                const unsigned magic = 0xFFFFFFFFu / nu + 1u;
                ctc_wait(hdr, &hdr->empty[s], (uint32_t)(ph ^ 1));
                for (unsigned idx = lt; idx < ((p.dbg & 1) ? 0u : total); idx += CTC_LDG_T) {
                    const unsigned gi = __umulhi(idx, magic), j = idx - gi * nu;
                    const unsigned g = (gi * p.kmagic) >> 16, i = gi - g * K;
                    const unsigned t = (j * 52429u) >> 18, q = j - 5u * t;
                    const bf16* src = xn + ((long long)(i * p.Cout + c0 + g) * p.T + t0) * 20 + j * 4;
                    const unsigned row = g * TR + t, byte = i * 40u + 8u * q;
                    cp_async8(sA + row * 128u + (((byte >> 4) ^ (row & 7u)) << 4) + (byte & 15u), src);
                }
                cp_async_commit();
                if (lagged) {
                    if (prev_s >= 0) {
                        cp_async_wait<1>();
                        fence_proxy_async_smem();
                        mbar_arrive(&hdr->a_full[prev_s]);
                    }
                    prev_s = s;
                } else {
                    cp_async_wait<0>();
                    fence_proxy_async_smem();
                    mbar_arrive(&hdr->a_full[s]);
                }
            } else {
                const int av = p.av;
                const unsigned nvec = (unsigned)(TRv * 25 / av), total = (unsigned)(Gv * K) * nvec;
                const unsigned magic = 0xFFFFFFFFu / nvec + 1u;
                ctc_wait(hdr, &hdr->empty[s], (uint32_t)(ph ^ 1));
#pragma unroll 2
                for (unsigned idx = lt; idx < total; idx += CTC_LDG_T) {
                    const unsigned gi = __umulhi(idx, magic), jv = idx - gi * nvec;
                    const unsigned g = (gi * p.kmagic) >> 16, i = gi - g * K;
                    const unsigned e0 = jv * av;
                    const bf16* src = xn + ((long long)(i * p.Cout + c0 + g) * p.T + t0) * 25 + e0;
                    unsigned short e[8];
                    if (av == 8) {
                        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src));
                        e[0] = u.x & 0xffff; e[1] = u.x >> 16; e[2] = u.y & 0xffff; e[3] = u.y >> 16;
                        e[4] = u.z & 0xffff; e[5] = u.z >> 16; e[6] = u.w & 0xffff; e[7] = u.w >> 16;
                    } else if (av == 4) {
                        const uint2 u = __ldg(reinterpret_cast<const uint2*>(src));
                        e[0] = u.x & 0xffff; e[1] = u.x >> 16; e[2] = u.y & 0xffff; e[3] = u.y >> 16;
                    } else if (av == 2) {
                        const unsigned u = __ldg(reinterpret_cast<const unsigned*>(src));
                        e[0] = u & 0xffff; e[1] = u >> 16;
                    } else {
                        e[0] = __ldg(reinterpret_cast<const unsigned short*>(src));
                    }
                    unsigned t = e0 / 25u, v = e0 - 25u * t;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (k < av) {
                            const unsigned row = g * TR + t, col = i * 32u + v, blk = col >> 6, cb = col & 63u;
                            st_shared_u16(sA + blk * 16384u + row * 128u + (((cb >> 3) ^ (row & 7u)) << 4) + (cb & 7u) * 2u, e[k]);
                            if (++v == 25u) { v = 0; ++t; }
                        }
                    }
                }
                fence_proxy_async_smem();
                mbar_arrive(&hdr->a_full[s]);
            }
        }
        if (lagged && prev_s >= 0) {
            cp_async_wait<0>();
            fence_proxy_async_smem();
            mbar_arrive(&hdr->a_full[prev_s]);
        }
    } else {
        // =============================== topology builders ===============================
        const int qt = tid - CTC_Q_T0;
        const int G = p.G, K = p.K, R = p.R;
        const float alpha = __ldg(alpha_p);
        const int nw4 = K * R * G + K * G;                     // per-tile parameter slice: W4 [K][R][G] then b4 [K][G]
        // software prefetch of the next tile's slice (global -> registers during the current tile's math).
        // The (i, r, g) decomposition of a thread's slots does not depend on the tile: do it once.
        const float* wsrc[3];
        int wstep[3], wg[3];                                   // element step per channel, g of the slot (-1: unused)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int idx = qt + k * CTC_Q_T;
            wsrc[k] = W4; wstep[k] = 0; wg[k] = -1;
            if (idx < K * R * G) {
                const int i = idx / (R * G), r2 = idx - i * (R * G), r = r2 / G, g = r2 - r * G;
                wsrc[k] = W4 + ((long long)i * p.Cout + g) * R + r; wstep[k] = R; wg[k] = g;
            } else if (idx < nw4) {
                const int j = idx - K * R * G, i = j / G, g = j - i * G;
                wsrc[k] = b4 + i * p.Cout + g; wstep[k] = 1; wg[k] = g;
            }
        }
        float wpre[3];
        auto w4_fetch = [&](int c0) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                wpre[k] = (wg[k] >= 0 && c0 + wg[k] < p.Cout) ? __ldg(wsrc[k] + (long long)c0 * wstep[k]) : 0.f;
        };
        CtcTile tl_, nx_;
        tl_.init(p, tile_begin);
        nx_ = tl_;
        if (nt > 0) w4_fetch(tl_.cg * G);
        int cur_n = -1;
        for (int it = 0; it < nt; ++it, tl_.next(p)) {
            const int n = tl_.n;
            const int s = it % S, ph = (it / S) & 1;
            float* w4t = W4t + (it & 1) * p.w4t_floats;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (qt + k * CTC_Q_T < nw4) w4t[qt + k * CTC_Q_T] = wpre[k];
            nx_.next(p);
            if (it + 1 < nt) w4_fetch(nx_.cg * G);
            if (n != cur_n) {
                // stage x1/x2 of the sample, then the tanh table D[i][r][u][v] (fp16)
                const float* x1n = x1 + (long long)n * p.x12ns;
                const float* x2n = x2 + (long long)n * p.x12ns;
                if (cur_n >= 0) bar_sync(2, CTC_Q_T);          // everybody is done reading the previous sample's table
                for (int idx = qt; idx < K * R * V; idx += CTC_Q_T) {
                    x12s[idx] = __ldg(x1n + idx);
                    x12s[K * R * V + idx] = __ldg(x2n + idx);
                }
                bar_sync(2, CTC_Q_T);
                for (int idx = qt; idx < K * R * V * Cf::VP; idx += CTC_Q_T) {
                    const int v = idx % Cf::VP, t1 = idx / Cf::VP, u = t1 % V, ir = t1 / V;
                    float d = 0.f;
                    if (v < V) d = tanhf(x12s[ir * V + u] - x12s[K * R * V + ir * V + v]);
                    Dsm[idx] = __float2half_rn(d);
                }
                cur_n = n;
            }
            bar_sync(2, CTC_Q_T);
            ctc_wait(hdr, &hdr->empty[s], (uint32_t)(ph ^ 1));
            const uint32_t sB = s0 + (uint32_t)s * p.stage_bytes + p.a_bytes;
            if (!(p.dbg & 2)) {
                if (G == 2) ctc_build_q<V, 2>(p, qt, Dsm, PAs, w4t, alpha, sB);
                else if (G == 4) ctc_build_q<V, 4>(p, qt, Dsm, PAs, w4t, alpha, sB);
                else if (G == 8) ctc_build_q<V, 8>(p, qt, Dsm, PAs, w4t, alpha, sB);
                else ctc_build_q<V, 1>(p, qt, Dsm, PAs, w4t, alpha, sB);
            }
            fence_proxy_async_smem();
            mbar_arrive(&hdr->b_full[s]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == CTC_MMA_W) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
    if (tid == 0 && hdr->error) printf("tamgcn: ctrgc_fwd(tcgen05) pipeline timeout in block %d\n", blockIdx.x);
}

static int ctc_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static bool ctc_disabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TAMGCN_DISABLE_TC");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static uint32_t al16(uint32_t x) { return (x + 15u) & ~15u; }

// returns 1 if launched, 0 if the caller must use the SIMT kernel, <0 on error
int ctrgc_fwd_tc(const void* x3, long long x3ns, int N, int Cout, int T, int V, int K, int R, const float* x1,
                 const float* x2, long long x12ns, const float* W4, const float* b4, const float* PA, const float* alpha,
                 void* y, long long yns, double* ssum, double* ssq, cudaStream_t st) {
    if (ctc_disabled()) return 0;
    if (V != 20 && V != 25) return 0;
    const int VS = V == 20 ? 20 : 32, VN = V == 20 ? 24 : 32, VP = V == 20 ? 20 : 28;
    if (K * VS > 128 || K > 8 || R > 64 || Cout > 2048) return 0;
    if (K * R * 8 + K * 8 > 3 * CTC_Q_T) return 0;          // per-tile parameter slice is prefetched 3 values per thread
    CtcP p = {};
    p.N = N; p.Cout = Cout; p.T = T; p.K = K; p.R = R;
    p.x3ns = x3ns; p.x12ns = x12ns; p.yns = yns;
    p.TR = T > 32 ? 64 : (T > 16 ? 32 : 16);
    p.G = 128 / p.TR;
    if (V == 25 && p.G > 4) p.G = 4;
    while (p.G > 1 && p.G / 2 >= Cout) p.G /= 2;
    p.nTC = (T + p.TR - 1) / p.TR;
    p.nCG = (Cout + p.G - 1) / p.G;
    p.tps = p.nCG * p.nTC;
    const long long tiles = (long long)N * p.tps;
    if (tiles > 0x7fffffffLL) return 0;
    p.n_tiles = (int)tiles;
    p.NMMA = (K * VS + 15) / 16;
    p.NKB = (p.NMMA + 3) / 4;
    p.Nmma = p.G * VN;
    if (p.Nmma < 16) p.Nmma = 16;
    p.Nmma = (p.Nmma + 15) & ~15;
    p.tmem_cols = (int)tmem_cols_pow2((uint32_t)(2 * p.Nmma));
    p.kmagic = 65536u / (unsigned)K + 1u;
    { const char* e = getenv("TAMGCN_CTC_DBG"); p.dbg = e ? atoi(e) : 0; }
    p.a_bytes = (uint32_t)p.NKB * 16384u;
    p.b_blk_bytes = (uint32_t)p.Nmma * 128u;
    p.stage_bytes = p.a_bytes + (uint32_t)p.NKB * p.b_blk_bytes;
    p.w4t_floats = (K * R * p.G + K * p.G + 3) & ~3;
    const uint32_t szD = al16((uint32_t)(K * R * V * VP * 2)), szPA = al16((uint32_t)(K * V * VP * 4));
    const uint32_t szW = al16((uint32_t)(2 * p.w4t_floats * 4)), szO = al16((uint32_t)(2 * 128 * V * 2));
    const uint32_t szS = al16((uint32_t)(2 * Cout * 4)), szH = al16((uint32_t)sizeof(CtcHdr));
    const uint32_t szX = al16((uint32_t)(2 * K * R * V * 4));
    const uint32_t fixed = szD + szPA + szW + szO + szS + szH + szX;
    const uint32_t budget = 227u * 1024u - 1024u;
    if (fixed + 2 * p.stage_bytes > budget) return 0;
    p.S = (int)((budget - fixed) / p.stage_bytes);
    if (p.S > CTC_SMAX) p.S = CTC_SMAX;
    uint32_t off = (uint32_t)p.S * p.stage_bytes;
    p.off_D = off; off += szD;
    p.off_PA = off; off += szPA;
    p.off_W4 = off; off += szW;
    p.off_out = off; off += szO;
    p.off_stat = off; off += szS;
    p.off_hdr = off; off += szH;
    p.off_x12 = off; off += szX;
    const size_t sm = (size_t)off + 1024;
    // alignment-dependent vector widths
    const uintptr_t xa = (uintptr_t)x3, ya = (uintptr_t)y;
    if (V == 20) {
        if ((xa & 7) || (x3ns & 3)) return 0;
        p.av = 4;
    } else {
        p.av = 1;
        for (int a = 8; a > 1; a >>= 1)
            if ((T * 25) % a == 0 && (p.TR * 25) % a == 0 && x3ns % a == 0 && (xa & (uintptr_t)(2 * a - 1)) == 0) { p.av = a; break; }
    }
    p.ov = 2;
    for (int o = 16; o > 2; o >>= 1)
        if ((T * V * 2) % o == 0 && (p.TR * V * 2) % o == 0 && (yns * 2) % o == 0 && (ya & (uintptr_t)(o - 1)) == 0) { p.ov = o; break; }
    int grid = ctc_num_sms();
    if (grid > p.n_tiles) grid = p.n_tiles;
    if (V == 20) {
        static int cur = 48 * 1024;
        if ((int)sm > cur) { cudaFuncSetAttribute(ctrgc_fwd_tc_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); cur = (int)sm; }
        ctrgc_fwd_tc_kernel<20><<<grid, CTC_THREADS, sm, st>>>(p, (const bf16*)x3, x1, x2, W4, b4, PA, alpha, (bf16*)y, ssum, ssq);
    } else {
        static int cur = 48 * 1024;
        if ((int)sm > cur) { cudaFuncSetAttribute(ctrgc_fwd_tc_kernel<25>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); cur = (int)sm; }
        ctrgc_fwd_tc_kernel<25><<<grid, CTC_THREADS, sm, st>>>(p, (const bf16*)x3, x1, x2, W4, b4, PA, alpha, (bf16*)y, ssum, ssq);
    }
    count_launch();
    const int rc = check_launch("ctrgc_fwd(tcgen05)");
    return rc < 0 ? rc : 1;
}

}  // namespace tamgcn
