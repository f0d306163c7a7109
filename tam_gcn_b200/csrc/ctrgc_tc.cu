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
#define CTC_EPI_T 256            // 8 epilogue warps: warp w drains TMEM lane quarter (w & 3) of sub-tile (w >> 2)
#define CTC_MMA_W 8
#define CTC_LD_T0 288            // first loader thread
#define CTC_LDG_T 128            // loader threads
#define CTC_Q_T0 416             // first topology-builder thread
#define CTC_Q_T 320
#define CTC_THREADS (CTC_Q_T0 + CTC_Q_T)   // 736
#define CTC_MAXU 14              // 8-byte units of a full sub-tile per loader thread (register table)

struct CtcP {
    int N, Cout, T, K, R;
    long long x3ns, x12ns, yns;
    int TR, G, P, nTC, nCG, tps, n_tiles;   // tile = P sub-tiles of G channels x TR rows; nCG = groups of P*G channels
    int S, NMMA, Nmma, NKB, tmem_cols;
    int av, ov;
    unsigned kmagic;                 // 65536 / K + 1
    uint32_t a_bytes, b_blk_bytes, b_bytes, stage_bytes;   // per sub-tile A / B block / B; whole stage
    uint32_t off_D, off_PA, off_W4, off_out, off_stat, off_hdr, off_x12;
    int w4t_floats;
    int lg2G;
    int dbg;                         // TAMGCN_CTC_DBG bit mask (profiling aid): 1 skip x3 copies, 2 skip Q math, 4 skip copy-out
    int ra25;                        // V = 25: rows fetched as aligned 16-byte words and realigned in registers
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
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool ctc_wait(CtcHdr* hdr, uint64_t* bar, uint32_t parity) {
    if (hdr->error) return false;
    if (!mbar_wait(bar, parity)) { hdr->error = 1; return false; }
    return true;
}
// profiling aid (TAMGCN_CTC_DBG & 8): cycles a role spends blocked on each of its barriers
#define CTC_TWAIT(acc, call) do { const long long t0_ = clock64(); call; acc += clock64() - t0_; } while (0)

// ---- topology builder: the Q tiles of one (sample, CT = P*G channels) into the B stage --------------
// One task = 4 consecutive v of one (i, u) for ALL CT channels of the tile: the fp16 tanh values are loaded once and
// feed CT independent packed-half FMA chains (HFMA2: two v per instruction).  The R-term sums (|terms| <= ~1) are
// accumulated in fp16, then alpha / b4 / PA are applied in fp32 and the result is rounded to bf16 — the rounding
// of the bf16 store (2^-9) dominates the fp16 accumulation error (2^-11 per term).
template <int V, int CT>
__device__ __forceinline__ void ctc_build_q(const CtcP& p, int qt, const __half* __restrict__ Dsm, const float* __restrict__ PAs,
                                            const uint32_t* __restrict__ w4h, float alpha, uint32_t sB, int lg2G) {
    typedef CtcCfg<V> Cf;
    const int R = p.R, K = p.K;
    const float* b4t = reinterpret_cast<const float*>(w4h + K * R * CT);
    const int ntask = K * V * Cf::NQ;
    for (int task = qt; task < ntask; task += CTC_Q_T) {
        const int i = task / (V * Cf::NQ), rem = task - i * (V * Cf::NQ), u = rem / Cf::NQ, vq = rem - u * Cf::NQ;
        __half2 acc[CT][2];
#pragma unroll
        for (int c = 0; c < CT; ++c) acc[c][0] = acc[c][1] = __float2half2_rn(0.f);
        const __half* dp = Dsm + ((size_t)(i * R) * V + u) * Cf::VP + 4 * vq;
        const uint32_t* wp = w4h + i * R * CT;
#pragma unroll 4
        for (int r = 0; r < R; ++r) {
            const uint2 dd = *reinterpret_cast<const uint2*>(dp + (size_t)r * V * Cf::VP);
            const __half2 d01 = *reinterpret_cast<const __half2*>(&dd.x), d23 = *reinterpret_cast<const __half2*>(&dd.y);
            uint32_t w[CT];
            if (CT >= 4) {
#pragma unroll
                for (int c = 0; c < CT; c += 4) {
                    const uint4 ww = *reinterpret_cast<const uint4*>(wp + r * CT + c);
                    w[c] = ww.x; w[c + 1] = ww.y; w[c + 2] = ww.z; w[c + 3] = ww.w;
                }
            } else if (CT == 2) {
                const uint2 ww = *reinterpret_cast<const uint2*>(wp + r * CT);
                w[0] = ww.x; w[1] = ww.y;
            } else {
                w[0] = wp[r * CT];
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const __half2 wc = *reinterpret_cast<const __half2*>(&w[c]);
                acc[c][0] = __hfma2(wc, d01, acc[c][0]);
                acc[c][1] = __hfma2(wc, d23, acc[c][1]);
            }
        }
        const float4 pa = *reinterpret_cast<const float4*>(PAs + (i * V + u) * Cf::VP + 4 * vq);
        const int col = i * Cf::VS + 4 * vq, blk = col >> 6, cb = col & 63;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const float b = b4t[i * CT + c];
            const float2 a01 = __half22float2(acc[c][0]), a23 = __half22float2(acc[c][1]);
            float q0 = fmaf(alpha, a01.x + b, pa.x), q1 = fmaf(alpha, a01.y + b, pa.y);
            float q2 = fmaf(alpha, a23.x + b, pa.z), q3 = fmaf(alpha, a23.y + b, pa.w);
            if (V == 25 && vq == Cf::NQ - 1) q1 = q2 = q3 = 0.f;          // v = 25..27 are K padding
            const int sp = c >> lg2G, g = c & ((1 << lg2G) - 1);
            const int row = g * Cf::VN + u;
            const uint32_t a = sB + (uint32_t)sp * p.b_bytes + (uint32_t)blk * p.b_blk_bytes + (uint32_t)row * 128u +
                               ((uint32_t)((cb >> 3) ^ (row & 7)) << 4) + (uint32_t)(cb & 7) * 2u;
            st_shared_v2(a, pack_bf16(q0, q1), pack_bf16(q2, q3));
        }
    }
}

// When there is at most one task per builder thread (V = 20: 300 tasks on 320 threads) the task never changes from
// tile to tile: its table pointer, PA values and the CT swizzled destination offsets are computed once.
template <int V, int CT>
struct CtcQPre {
    const __half* dp;
    float4 pa;
    uint32_t off[CT];          // byte offset of the 8-byte store of channel c inside a B stage
    int i;
    bool active, pad_tail;
    __device__ __forceinline__ void init(const CtcP& p, int qt, const __half* Dsm, const float* PAs, int lg2G) {
        typedef CtcCfg<V> Cf;
        active = qt < p.K * V * Cf::NQ;
        const int task = active ? qt : 0;
        i = task / (V * Cf::NQ);
        const int rem = task - i * (V * Cf::NQ), u = rem / Cf::NQ, vq = rem - u * Cf::NQ;
        dp = Dsm + ((size_t)(i * p.R) * V + u) * Cf::VP + 4 * vq;
        pa = *reinterpret_cast<const float4*>(PAs + (i * V + u) * Cf::VP + 4 * vq);
        pad_tail = (V == 25 && vq == Cf::NQ - 1);
        const int col = i * Cf::VS + 4 * vq, blk = col >> 6, cb = col & 63;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const int sp = c >> lg2G, g = c & ((1 << lg2G) - 1), row = g * Cf::VN + u;
            off[c] = (uint32_t)sp * p.b_bytes + (uint32_t)blk * p.b_blk_bytes + (uint32_t)row * 128u +
                     ((uint32_t)((cb >> 3) ^ (row & 7)) << 4) + (uint32_t)(cb & 7) * 2u;
        }
    }
    __device__ __forceinline__ void run(const CtcP& p, const uint32_t* __restrict__ w4h, float alpha, uint32_t sB) const {
        typedef CtcCfg<V> Cf;
        if (!active) return;
        const int R = p.R;
        const float* b4t = reinterpret_cast<const float*>(w4h + p.K * R * CT) + i * CT;
        const uint32_t* wp = w4h + i * R * CT;
        __half2 acc[CT][2];
#pragma unroll
        for (int c = 0; c < CT; ++c) acc[c][0] = acc[c][1] = __float2half2_rn(0.f);
#pragma unroll 4
        for (int r = 0; r < R; ++r) {
            const uint2 dd = *reinterpret_cast<const uint2*>(dp + (size_t)r * V * Cf::VP);
            const __half2 d01 = *reinterpret_cast<const __half2*>(&dd.x), d23 = *reinterpret_cast<const __half2*>(&dd.y);
            uint32_t w[CT];
            if (CT >= 4) {
#pragma unroll
                for (int c = 0; c < CT; c += 4) {
                    const uint4 ww = *reinterpret_cast<const uint4*>(wp + r * CT + c);
                    w[c] = ww.x; w[c + 1] = ww.y; w[c + 2] = ww.z; w[c + 3] = ww.w;
                }
            } else if (CT == 2) {
                const uint2 ww = *reinterpret_cast<const uint2*>(wp + r * CT);
                w[0] = ww.x; w[1] = ww.y;
            } else {
                w[0] = wp[r * CT];
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const __half2 wc = *reinterpret_cast<const __half2*>(&w[c]);
                acc[c][0] = __hfma2(wc, d01, acc[c][0]);
                acc[c][1] = __hfma2(wc, d23, acc[c][1]);
            }
        }
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const float b = b4t[c];
            const float2 a01 = __half22float2(acc[c][0]), a23 = __half22float2(acc[c][1]);
            float q0 = fmaf(alpha, a01.x + b, pa.x), q1 = fmaf(alpha, a01.y + b, pa.y);
            float q2 = fmaf(alpha, a23.x + b, pa.z), q3 = fmaf(alpha, a23.y + b, pa.w);
            if (pad_tail) q1 = q2 = q3 = 0.f;
            st_shared_v2(sB + off[c], pack_bf16(q0, q1), pack_bf16(q2, q3));
        }
    }
};

template <int V, int CT>
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
    float* stat = (float*)(smem + p.off_stat);       // [8 epilogue warps][2][Cout]
    float* x12s = (float*)(smem + p.off_x12);        // [2][K*R*V]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile_begin = (int)((long long)blockIdx.x * p.n_tiles / gridDim.x);
    const int tile_end = (int)((long long)(blockIdx.x + 1) * p.n_tiles / gridDim.x);
    const int nt = tile_end - tile_begin;
    const int S = p.S, P = p.P, G = p.G, TR = p.TR;        // CT == P * G

    // ---- one-time setup ----
    if (warp == CTC_MMA_W) tmem_alloc(&hdr->tmem_base, (uint32_t)p.tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < CTC_SMAX; ++i) {
            mbar_init(&hdr->a_full[i], V == 20 ? CTC_LDG_T : CTC_LDG_T / 2);
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
    for (int i = tid; i < 16 * p.Cout; i += CTC_THREADS) stat[i] = 0.f;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = hdr->tmem_base;
    const uint32_t s0 = smem_u32(smem);
    long long tw0 = 0, tw1 = 0, tw2 = 0;
    const long long t_begin = clock64();

    if (warp < 8) {
        // =============================== epilogue ===============================
        const int row = tid & 127;                             // TMEM lane == sub-tile row (g, t)
        const int qw = warp & 3, my_sp = warp >> 2;            // TMEM lane quarter, sub-tile of this warp
        const int g_row = row / TR, tl = row - g_row * TR;
        const int width = TR < 32 ? TR : 32;                   // lanes of a warp that share a channel
        float* mystat = stat + warp * 2 * p.Cout;              // this warp's private accumulators: no atomics
        CtcTile tl_;
        tl_.init(p, tile_begin);
        for (int it = 0; it < nt; ++it, tl_.next(p)) {
            const int n = tl_.n;
            const int c0 = tl_.cg * CT, t0 = tl_.tc * TR;
            const int TRv = min(TR, p.T - t0), CTv = min(CT, p.Cout - c0);
            const int buf = it & 1;
            uint8_t* ob = outs + (size_t)buf * (P * 128 * V * 2);
            // one warp watches the mbarrier; the others park on a hardware barrier (a parked warp issues nothing,
            // a polling one costs issue slots the builders need)
            if (warp == 0) CTC_TWAIT(tw0, ctc_wait(hdr, &hdr->tfull[buf], (uint32_t)((it >> 1) & 1)));
            bar_sync(3, CTC_EPI_T);
            tc_fence_after();
            for (int sp = my_sp; sp < P; sp += 2) {
                float acc[Cf::VN];
                const uint32_t tbase = tmem + ((uint32_t)(qw * 32) << 16) + (uint32_t)((buf * P + sp) * p.Nmma);
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
                        const uint32_t ta = tbase + (uint32_t)(min(qw * 2 + h, G - 1) * Cf::VN);
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
                if (sp + 2 >= P) {
                    tc_fence_before();
                    mbar_arrive(&hdr->tempty[buf]);            // the MMA warp may refill this accumulator buffer
                }
                const int gc = sp * G + g_row;                 // channel of this row inside the tile
                const bool valid = (g_row < G) && (gc < CTv) && (tl < TRv);
                float s = 0.f, q = 0.f;
                if (valid) {
                    const uint32_t oa = smem_u32(ob) + (uint32_t)((gc * TRv + tl) * V * 2);
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
                    if ((lane & (width - 1)) == 0 && g_row < G && gc < CTv) {
                        mystat[c0 + gc] += s;
                        mystat[p.Cout + c0 + gc] += q;
                    }
                }
            }
            if (my_sp >= P) mbar_arrive(&hdr->tempty[buf]);
            CTC_TWAIT(tw1, bar_sync(1, CTC_EPI_T));
            // coalesced copy-out of the staged tile
            const int ov = (p.dbg & 4) ? 0 : p.ov;
            if (ov == 0) {
            } else if (p.nTC == 1) {
                const int bytes = CTv * TRv * V * 2;
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
                for (int g = 0; g < CTv; ++g) {
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
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) { a += stat[2 * w * p.Cout + c]; b += stat[(2 * w + 1) * p.Cout + c]; }
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
                CTC_TWAIT(tw0, ctc_wait(hdr, &hdr->tempty[buf], (uint32_t)(((it >> 1) & 1) ^ 1)));
                CTC_TWAIT(tw1, ctc_wait(hdr, &hdr->a_full[s], (uint32_t)ph));
                CTC_TWAIT(tw2, ctc_wait(hdr, &hdr->b_full[s], (uint32_t)ph));
                tc_fence_after();
                const uint32_t sA = s0 + (uint32_t)s * p.stage_bytes, sB = sA + (uint32_t)P * p.a_bytes;
                for (int sp = 0; sp < P; ++sp) {
                    const uint32_t td = tmem + (uint32_t)((buf * P + sp) * p.Nmma);
                    const uint32_t sa = sA + (uint32_t)sp * p.a_bytes, sb = sB + (uint32_t)sp * p.b_bytes;
                    for (int j = 0; j < p.NMMA; ++j) {
                        const uint32_t blk = (uint32_t)j >> 2, kk = (uint32_t)j & 3;
                        umma_bf16(td, umma_desc_sw128(sa + blk * 16384u + kk * 32u), umma_desc_sw128(sb + blk * p.b_blk_bytes + kk * 32u),
                                  idesc, j > 0 ? 1u : 0u);
                    }
                }
                umma_commit(&hdr->empty[s]);
                umma_commit(&hdr->tfull[buf]);
            }
        }
    } else if (tid < CTC_Q_T0) {
        // =============================== x3 loaders (two groups alternate tiles) ===============================
        // V = 20: one group of 128 threads, asynchronous copies.  V = 25: the copies go through registers (2-byte
        // granularity), so two groups of 64 threads alternate tiles to overlap each other's load latency.
        const int NLG = (V == 20) ? 1 : 2, LGT = CTC_LDG_T / NLG;
        const int lg = (tid - CTC_LD_T0) / LGT, lt = (tid - CTC_LD_T0) % LGT;
        const int K = p.K;
        // cp.async tiles are signalled `lag` tiles after they were issued, so up to lag + 1 tiles of loads are in flight
        const int lag = (V == 20) ? (S >= 4 ? 2 : S - 2) : 0;
        // V = 20: the (source offset, swizzled destination offset) of a thread's 8-byte units are the same for every full
        // sub-tile (G channels x TRf rows): computed once, kept in registers.
        const int TRf = p.nTC == 1 ? p.T : TR;
        int soff[CTC_MAXU];
        uint32_t doff[CTC_MAXU];
        const unsigned nuf = (unsigned)TRf * 5u, totf = (unsigned)(G * K) * nuf;
        const bool tab_ok = (V == 20) && totf <= CTC_MAXU * CTC_LDG_T;
        if (tab_ok) {
#pragma unroll
            for (int k = 0; k < CTC_MAXU; ++k) {
                const unsigned idx = lt + k * CTC_LDG_T;
                soff[k] = -1; doff[k] = 0;
                if (idx < totf) {
                    const unsigned gi = idx / nuf, j = idx - gi * nuf, g = gi / K, i = gi - g * K, t = j / 5u, q = j - 5u * t;
                    const unsigned row = g * TR + t, byte = i * 40u + 8u * q;
                    soff[k] = (int)((i * p.Cout + g) * p.T * 20 + j * 4);
                    doff[k] = row * 128u + (((byte >> 4) ^ (row & 7u)) << 4) + (byte & 15u);
                }
            }
        }
        CtcTile tl_;
        tl_.init(p, tile_begin + lg);
        for (int it = lg; it < nt; it += NLG, tl_.next(p)) {
            if (NLG == 2 && it > lg) tl_.next(p);
            const int n = tl_.n;
            const int c0 = tl_.cg * CT, t0 = tl_.tc * TR;
            const int TRv = min(TR, p.T - t0);
            const int s = it % S, ph = (it / S) & 1;
            const uint32_t sA = s0 + (uint32_t)s * p.stage_bytes;
            const bf16* xn = x3 + (long long)n * p.x3ns;
            if (lt < 32) CTC_TWAIT(tw0, ctc_wait(hdr, &hdr->empty[s], (uint32_t)(ph ^ 1)));
            bar_sync(4 + lg, LGT);
            for (int sp = 0; sp < P; ++sp) {
                const int cs = c0 + sp * G, Gv = min(G, p.Cout - cs);
                if (Gv <= 0 || (p.dbg & 1)) break;
                const uint32_t sa = sA + (uint32_t)sp * p.a_bytes;
                if (V == 20) {
                    if (tab_ok && Gv == G && TRv == TRf) {
                        const bf16* base = xn + ((long long)cs * p.T + t0) * 20;
#pragma unroll
                        for (int k = 0; k < CTC_MAXU; ++k)
                            if (soff[k] >= 0) cp_async8(sa + doff[k], base + soff[k]);
                    } else {
                        const unsigned nu = (unsigned)TRv * 5u, total = (unsigned)(Gv * K) * nu;
                        const unsigned magic = 0xFFFFFFFFu / nu + 1u;
                        for (unsigned idx = lt; idx < total; idx += CTC_LDG_T) {
                            const unsigned gi = __umulhi(idx, magic), j = idx - gi * nu;
                            const unsigned g = (gi * p.kmagic) >> 16, i = gi - g * K;
                            const unsigned t = (j * 52429u) >> 18, q = j - 5u * t;
                            const bf16* src = xn + ((long long)(i * p.Cout + cs + g) * p.T + t0) * 20 + j * 4;
                            const unsigned row = g * TR + t, byte = i * 40u + 8u * q;
                            cp_async8(sa + row * 128u + (((byte >> 4) ^ (row & 7u)) << 4) + (byte & 15u), src);
                        }
                    }
                } else if (p.ra25) {
                    // V = 25: a row is 50 bytes on a 2-byte boundary.  Its first 24 elements are fetched as aligned 16-byte
                    // words and realigned in registers into the three 16-byte chunks of the operand row, element 24 follows
                    // alone (the rest of its chunk is K padding = 0): 4 stores per row instead of 25.
                    const unsigned rows = (unsigned)(Gv * K * TRv);
                    const unsigned rmagic = 0xFFFFFFFFu / (unsigned)TRv + 1u;
#pragma unroll 2
                    for (unsigned idx = lt; idx < rows; idx += LGT) {
                        const unsigned gi = __umulhi(idx, rmagic), t = idx - gi * (unsigned)TRv;
                        const unsigned g = (gi * p.kmagic) >> 16, i = gi - g * K;
                        const bf16* src = xn + ((long long)(i * p.Cout + cs + g) * p.T + t0 + t) * 25;
                        const unsigned row = g * TR + t, c0 = (i & 1u) * 4u, sw = row & 7u;
                        const uint32_t rb = sa + (i >> 1) * 16384u + row * 128u;
                        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
                        const uint32_t sft = (uint32_t)(a & 15);
                        const uint4* q = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
                        const uint4 w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
                        uint4 k0 = w0, k1 = w1, k2 = w2;
                        if (sft) {
                            const uint4 w3 = __ldg(q + 3);
                            k0 = tc_realign16(w0, w1, sft); k1 = tc_realign16(w1, w2, sft); k2 = tc_realign16(w2, w3, sft);
                        }
                        const uint32_t e24 = __ldg(reinterpret_cast<const unsigned short*>(src + 24));
                        st_shared_v4(rb + (((c0 + 0u) ^ sw) << 4), k0.x, k0.y, k0.z, k0.w);
                        st_shared_v4(rb + (((c0 + 1u) ^ sw) << 4), k1.x, k1.y, k1.z, k1.w);
                        st_shared_v4(rb + (((c0 + 2u) ^ sw) << 4), k2.x, k2.y, k2.z, k2.w);
                        st_shared_v4(rb + (((c0 + 3u) ^ sw) << 4), e24, 0u, 0u, 0u);
                    }
                } else {
                    const int av = p.av;
                    const unsigned nvec = (unsigned)(TRv * 25 / av), total = (unsigned)(Gv * K) * nvec;
                    const unsigned magic = 0xFFFFFFFFu / nvec + 1u;
#pragma unroll 2
                    for (unsigned idx = lt; idx < total; idx += LGT) {
                        const unsigned gi = __umulhi(idx, magic), jv = idx - gi * nvec;
                        const unsigned g = (gi * p.kmagic) >> 16, i = gi - g * K;
                        const unsigned e0 = jv * av;
                        const bf16* src = xn + ((long long)(i * p.Cout + cs + g) * p.T + t0) * 25 + e0;
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
                                st_shared_u16(sa + blk * 16384u + row * 128u + (((cb >> 3) ^ (row & 7u)) << 4) + (cb & 7u) * 2u, e[k]);
                                if (++v == 25u) { v = 0; ++t; }
                            }
                        }
                    }
                }
            }
            if (V == 20) {
                cp_async_commit();
                if (it >= lag) {
                    if (lag == 2) CTC_TWAIT(tw1, cp_async_wait<2>());
                    else if (lag == 1) cp_async_wait<1>();
                    else cp_async_wait<0>();
                    fence_proxy_async_smem();
                    mbar_arrive(&hdr->a_full[(it - lag) % S]);
                }
            } else {
                fence_proxy_async_smem();
                mbar_arrive(&hdr->a_full[s]);
            }
        }
        if (V == 20) {
            cp_async_wait<0>();
            fence_proxy_async_smem();
            for (int it = max(nt - lag, 0); it < nt; ++it) mbar_arrive(&hdr->a_full[it % S]);
        }
    } else {
        // =============================== topology builders ===============================
        const int qt = tid - CTC_Q_T0;
        const int K = p.K, R = p.R;
        const float alpha = __ldg(alpha_p);
        const int nw4 = K * R * CT + K * CT;                   // per-tile parameter slice: W4 [K][R][CT] then b4 [K][CT]
        // software prefetch of the next tile's slice (global -> registers during the current tile's math).
        // The (i, r, g) decomposition of a thread's slots does not depend on the tile: do it once.
        const float* wsrc[3];
        int wstep[3], wg[3];                                   // element step per channel, g of the slot (-1: unused)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int idx = qt + k * CTC_Q_T;
            wsrc[k] = W4; wstep[k] = 0; wg[k] = -1;
            if (idx < K * R * CT) {
                const int i = idx / (R * CT), r2 = idx - i * (R * CT), r = r2 / CT, g = r2 - r * CT;
                wsrc[k] = W4 + ((long long)i * p.Cout + g) * R + r; wstep[k] = R; wg[k] = g;
            } else if (idx < nw4) {
                const int j = idx - K * R * CT, i = j / CT, g = j - i * CT;
                wsrc[k] = b4 + i * p.Cout + g; wstep[k] = 1; wg[k] = g;
            }
        }
        float wpre[3];
        bool wok[3];
        auto w4_fetch = [&](int c0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                wok[k] = wg[k] >= 0 && c0 + wg[k] < p.Cout;
                wpre[k] = __ldg(wok[k] ? wsrc[k] + (long long)c0 * wstep[k] : W4);   // consumed one tile later
            }
        };
        constexpr bool one_task = 3 * V * Cf::NQ <= CTC_Q_T;   // K <= 3 is checked at run time below
        CtcQPre<V, CT> pre;
        const bool hoisted = one_task && K <= 3;
        if (hoisted) pre.init(p, qt, Dsm, PAs, p.lg2G);
        CtcTile tl_, nx_;
        tl_.init(p, tile_begin);
        nx_ = tl_;
        if (nt > 0) w4_fetch(tl_.cg * CT);
        int cur_n = -1;
        for (int it = 0; it < nt; ++it, tl_.next(p)) {
            const int n = tl_.n;
            const int s = it % S, ph = (it / S) & 1;
            float* w4t = W4t + (it & 1) * p.w4t_floats;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (qt + k * CTC_Q_T < nw4) {
                    const float w = wok[k] ? wpre[k] : 0.f;
                    if (qt + k * CTC_Q_T < K * R * CT) {
                        const __half2 h2 = __float2half2_rn(w);                 // (w, w): broadcast operand of HFMA2
                        reinterpret_cast<uint32_t*>(w4t)[qt + k * CTC_Q_T] = *reinterpret_cast<const uint32_t*>(&h2);
                    } else {
                        w4t[qt + k * CTC_Q_T] = w;
                    }
                }
            nx_.next(p);
            if (it + 1 < nt) w4_fetch(nx_.cg * CT);
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
                // one item = 4 consecutive v of one (i, r, u) row: hardware tanh (MUFU, 2^-11 relative) -> fp16 (2^-11)
                for (int idx = qt; idx < K * R * V * Cf::NQ; idx += CTC_Q_T) {
                    const int vq = idx % Cf::NQ, iru = idx / Cf::NQ, ir = iru / V;
                    const float a = x12s[iru];
                    const float* xb = x12s + K * R * V + ir * V + 4 * vq;
                    float d[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        d[e] = 0.f;
                        if (V % 4 == 0 || 4 * vq + e < V) d[e] = tanh_approx(a - xb[e]);
                    }
                    const __half2 h01 = __floats2half2_rn(d[0], d[1]), h23 = __floats2half2_rn(d[2], d[3]);
                    *reinterpret_cast<uint2*>(Dsm + (size_t)iru * Cf::VP + 4 * vq) =
                        make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
                }
                cur_n = n;
            }
            if (qt < 32) CTC_TWAIT(tw0, ctc_wait(hdr, &hdr->empty[s], (uint32_t)(ph ^ 1)));   // one warp polls for the whole role
            CTC_TWAIT(tw1, bar_sync(2, CTC_Q_T));
            const uint32_t sB = s0 + (uint32_t)s * p.stage_bytes + (uint32_t)P * p.a_bytes;
            if (!(p.dbg & 2)) {
                const uint32_t* w4h = reinterpret_cast<const uint32_t*>(w4t);
                if (hoisted) pre.run(p, w4h, alpha, sB);
                else ctc_build_q<V, CT>(p, qt, Dsm, PAs, w4h, alpha, sB, p.lg2G);
            }
            fence_proxy_async_smem();
            mbar_arrive(&hdr->b_full[s]);
        }
    }
    if ((p.dbg & 8) && blockIdx.x == 0 && (tid == 0 || tid == 128 || tid == CTC_MMA_W * 32 || tid == CTC_LD_T0 || tid == CTC_Q_T0))
        printf("ctc role tid %d: tiles %d total %lld clk, blocked %lld / %lld / %lld\n", tid, nt, clock64() - t_begin, tw0, tw1, tw2);
    tc_fence_before();
    __syncthreads();
    if (warp == CTC_MMA_W) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
    if (tid == 0 && hdr->error) printf("tamgcn: ctrgc_fwd(tcgen05) pipeline timeout in block %d\n", blockIdx.x);
}

static int ctc_num_sms() { return num_sms(); }

static bool ctc_disabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TAMGCN_DISABLE_TC");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static uint32_t al16(uint32_t x) { return (x + 15u) & ~15u; }

int ctrgc_fwd_tc3(const void* x3, long long x3ns, int N, int Cout, int T, int V, int K, int R, const float* x1,
                  const float* x2, long long x12ns, const float* W4, const float* b4, const float* PA, const float* alpha,
                  void* y, long long yns, double* ssum, double* ssq, cudaStream_t st);
int ctrgc_fwd_tc4(const void* x3, long long x3ns, int N, int Cout, int T, int V, int K, int R, const float* x1,
                  const float* x2, long long x12ns, const float* W4, const float* b4, const float* PA, const float* alpha,
                  void* y, long long yns, double* ssum, double* ssq, cudaStream_t st);

// returns 1 if launched, 0 if the caller must use the SIMT kernel, <0 on error
int ctrgc_fwd_tc(const void* x3, long long x3ns, int N, int Cout, int T, int V, int K, int R, const float* x1,
                 const float* x2, long long x12ns, const float* W4, const float* b4, const float* PA, const float* alpha,
                 void* y, long long yns, double* ssum, double* ssq, cudaStream_t st) {
    if (ctc_disabled()) return 0;
    {   // the instruction-lean variant (ctrgc_tc3.cu) covers V = 20, R = 8, 32 < T <= 64
        const int rc3 = ctrgc_fwd_tc3(x3, x3ns, N, Cout, T, V, K, R, x1, x2, x12ns, W4, b4, PA, alpha, y, yns, ssum, ssq, st);
        if (rc3 != 0) return rc3;
        // its V = 25 sibling (ctrgc_tc4.cu): V = 25, R = 8, 32 < T <= 64, T a multiple of 8
        const int rc4 = ctrgc_fwd_tc4(x3, x3ns, N, Cout, T, V, K, R, x1, x2, x12ns, W4, b4, PA, alpha, y, yns, ssum, ssq, st);
        if (rc4 != 0) return rc4;
    }
    // TAMGCN_CTC_GENERIC=0: leave the shapes of this (older) kernel to the warp-MMA forward of ctrgc.cu
    static const bool generic_off = [] { const char* e = getenv("TAMGCN_CTC_GENERIC"); return e && e[0] == '0'; }();
    if (generic_off) return 0;
    if (V != 20 && V != 25) return 0;
    const int VS = V == 20 ? 20 : 32, VN = V == 20 ? 24 : 32, VP = V == 20 ? 20 : 28;
    if (K * VS > 128 || K > 8 || R > 64 || Cout > 2048) return 0;
    CtcP p = {};
    p.N = N; p.Cout = Cout; p.T = T; p.K = K; p.R = R;
    p.x3ns = x3ns; p.x12ns = x12ns; p.yns = yns;
    p.TR = T > 32 ? 64 : (T > 16 ? 32 : 16);
    p.G = 128 / p.TR;
    if (V == 25 && p.G > 4) p.G = 4;
    while (p.G > 1 && p.G / 2 >= Cout) p.G /= 2;
    p.nTC = (T + p.TR - 1) / p.TR;
    p.NMMA = (K * VS + 15) / 16;
    p.NKB = (p.NMMA + 3) / 4;
    p.Nmma = p.G * VN;
    if (p.Nmma < 16) p.Nmma = 16;
    p.Nmma = (p.Nmma + 15) & ~15;
    p.kmagic = 65536u / (unsigned)K + 1u;
    { const char* e = getenv("TAMGCN_CTC_DBG"); p.dbg = e ? atoi(e) : 0; }
    { const char* e = getenv("TAMGCN_CTC_RA25"); p.ra25 = (V == 25 && !(e && e[0] == '0')) ? 1 : 0; }
    p.a_bytes = (uint32_t)p.NKB * 16384u;
    p.b_blk_bytes = (uint32_t)p.Nmma * 128u;
    p.b_bytes = (uint32_t)p.NKB * p.b_blk_bytes;
    const uint32_t budget = 227u * 1024u - 1024u;
    const uint32_t szD = al16((uint32_t)(K * R * V * VP * 2)), szPA = al16((uint32_t)(K * V * VP * 4));
    const uint32_t szS = al16((uint32_t)(16 * Cout * 4)), szH = al16((uint32_t)sizeof(CtcHdr));
    const uint32_t szX = al16((uint32_t)(2 * K * R * V * 4));
    uint32_t szW = 0, szO = 0, fixed = 0;
    // P sub-tiles per pipeline step: halves the per-byte cost of the barrier hand-offs when everything still fits
    for (p.P = 2; p.P >= 1; --p.P) {
        const int CT = p.P * p.G;
        p.stage_bytes = (uint32_t)p.P * (p.a_bytes + p.b_bytes);
        p.w4t_floats = (K * R * CT + K * CT + 3) & ~3;
        szW = al16((uint32_t)(2 * p.w4t_floats * 4));
        szO = al16((uint32_t)(2 * p.P * 128 * V * 2));
        fixed = szD + szPA + szW + szO + szS + szH + szX;
        const bool fits = fixed + 3 * p.stage_bytes <= budget && 2 * p.P * p.Nmma <= 512 && K * R * CT + K * CT <= 3 * CTC_Q_T &&
                          CT <= Cout;
        if (fits || p.P == 1) break;
    }
    if (K * R * p.P * p.G + K * p.P * p.G > 3 * CTC_Q_T) return 0;   // per-tile parameter slice is prefetched 3 values per thread
    if (fixed + 2 * p.stage_bytes > budget) return 0;
    p.tmem_cols = (int)tmem_cols_pow2((uint32_t)(2 * p.P * p.Nmma));
    p.lg2G = p.G == 8 ? 3 : (p.G == 4 ? 2 : (p.G == 2 ? 1 : 0));
    if (p.P * p.G > 8) return 0;
    p.nCG = (Cout + p.P * p.G - 1) / (p.P * p.G);
    p.tps = p.nCG * p.nTC;
    const long long tiles = (long long)N * p.tps;
    if (tiles > 0x7fffffffLL) return 0;
    p.n_tiles = (int)tiles;
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
#define CTC_LAUNCH(VV, CC)                                                                                              \
    do {                                                                                                                \
        static SmemLimit lim;                                                                                           \
        ensure_smem(ctrgc_fwd_tc_kernel<VV, CC>, lim, sm);                                                              \
        ctrgc_fwd_tc_kernel<VV, CC><<<grid, CTC_THREADS, sm, st>>>(p, (const bf16*)x3, x1, x2, W4, b4, PA, alpha, (bf16*)y, ssum, ssq); \
    } while (0)
    const int CT = p.P * p.G;
    if (V == 20) {
        if (CT == 8) CTC_LAUNCH(20, 8); else if (CT == 4) CTC_LAUNCH(20, 4); else if (CT == 2) CTC_LAUNCH(20, 2); else CTC_LAUNCH(20, 1);
    } else {
        if (CT == 8) CTC_LAUNCH(25, 8); else if (CT == 4) CTC_LAUNCH(25, 4); else if (CT == 2) CTC_LAUNCH(25, 2); else CTC_LAUNCH(25, 1);
    }
#undef CTC_LAUNCH
    count_launch();
    const int rc = check_launch("ctrgc_fwd(tcgen05)");
    return rc < 0 ? rc : 1;
}

}  // namespace tamgcn
