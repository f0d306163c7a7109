// ctrgc_tc4.cu — fused CTRGC forward, bf16, V = 25 (NTU RGB+D), R = 8: the V = 25 sibling of ctrgc_tc3.cu.
//
//   y[n,c,t,u] = sum_i sum_v Q_i[n,c,u,v] * x3_i[n,c,t,v]      (reference models/ctrgcn.py:174-176, :252-254)
//   Q_i[n,c,u,v] = alpha * (sum_r W4_i[c,r] tanh(x1_i[n,r,u] - x2_i[n,r,v]) + b4_i[c]) + PA_i[u,v]
//
// Same pipeline as ctrgc_tc3.cu (persistent warp-specialised CTA, tcgen05 main contraction with the accumulator in
// TMEM, topology builder on mma.sync, finished tiles leave as one bulk store).  What 25 joints change:
//   * rows of x3 are 50 bytes: no row starts on a 4-byte boundary two rows in a row, so cp.async cannot gather them.
//     The loaders fetch ALIGNED 16-byte words (fully coalesced: four lanes per row cover its 64-byte hull), take the
//     neighbour's word by shuffle and realign in registers.  The shift is 2*(t mod 8) bytes, so loader warp w owns
//     the rows t = w (mod 8) and its realignment is compile-time constant code (no dynamic register indexing).
//   * K = (subset, v) = 3 x 25 does not fit one 128-byte swizzle row: every subset gets its own K-major SWIZZLE_64B
//     block (32 v slots = 64 bytes per row, 25 used), 2 MMA K-steps per block, 6 per sub-tile.
//   * 25 bf16 per output row: the epilogue packs pairs once and selects between the even / odd (16-bit shifted)
//     word streams by row parity, so every store is 4 bytes except one 2-byte edge.
//
// tile      = one sample n, 4 output channels (2 sub-tiles of 2 channels x 64 time rows = UMMA M 128)
// A (smem)  = per subset: x3 rows [(g,t)] x [v 32] bf16, K-major SWIZZLE_64B
// B (smem)  = per subset: Q rows  [(g,u 32)] x [v 32] bf16, K-major SWIZZLE_64B, written by the builder warps
// D (TMEM)  = [(g,t)] x [(g',u 32)] fp32, double buffered; only the diagonal blocks g == g' are read back
// warps 0-7 epilogue | 8-15 x3 loaders (warp = t mod 8) | 16 MMA issue | 17-23 topology builders (one per 4 u)
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace tamgcn {

#define C4_EPI_T 256
#define C4_LD_T0 256
#define C4_LD_T 256
#define C4_MMA_W 16
#define C4_Q_T0 544
#define C4_Q_T 224
#define C4_THREADS 768
#define C4_SMAX 2
#define C4_V 25
#define C4_UP 28                 // u rows of the tanh / PA tables (7 builder warps x 4), rows >= 25 stay zero
#define C4_PAP 26                // v pitch of the PA table (even: 8-byte aligned pair loads)
#define C4_R 8
#define C4_NMMA_N 64             // UMMA N: 2 channels x 32 row slots
#define C4_A_BLK 8192u           // one subset block of A: 128 rows x 64 B
#define C4_B_BLK 4096u           // one subset block of B:  64 rows x 64 B
#define C4_A_BYTES (3u * C4_A_BLK)
#define C4_B_BYTES (3u * C4_B_BLK)
#define C4_STAGE_BYTES (2u * (C4_A_BYTES + C4_B_BYTES))
#define C4_OUT_BYTES 12800u      // staging of one finished tile (4 ch x T<=64 x 25 x 2 B)
#define C4_D_SUB (C4_UP * C4_V * 16)      // bytes of one subset of the tanh table [u 28][v 25][r 8] fp16
#define C4_PA_SUB (C4_UP * C4_PAP * 4)    // bytes of one subset of the PA table [u 28][v 26] fp32

struct C4P {
    int N, Cout, T, K;
    long long x3ns, x12ns, yns;
    int nCG, n_tiles, S;
    int pf;                      // L2 prefetch distance in tiles (0 = off)
    int dbg;                     // TAMGCN_C4_DBG (profiling aid): 1 skip x3 loads, 2 skip Q math, 4 skip output, 8 skip realign
    uint32_t off_D, off_PA, off_W4, off_b4, off_out, off_stat, off_hdr, off_x12;
};

struct C4Hdr {
    uint64_t a_full[C4_SMAX], b_full[C4_SMAX], empty[C4_SMAX], tfull[2], tempty[2];
    uint32_t tmem_base;
    volatile uint32_t error;
};

__device__ __forceinline__ void c4_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void c4_sts32(uint32_t a, uint32_t x) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }
__device__ __forceinline__ void c4_sts16(uint32_t a, uint32_t x) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)x) : "memory");
}
__device__ __forceinline__ uint32_t c4_lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 c4_lds_f2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void c4_tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void c4_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float c4_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// D(16x8, fp32) += A(16x16, fp16, row) * B(16x8, fp16, col)
__device__ __forceinline__ void c4_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void c4_bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ bool c4_wait(C4Hdr* hdr, uint64_t* bar, uint32_t parity) {
    if (mbar_wait_spin(bar, parity)) return true;
    hdr->error = 1;
    return false;
}
// byte offset of 16-byte chunk c (0..3) of row r inside a K-major SWIZZLE_64B block (64-byte rows, Swizzle<2,4,3>)
__device__ __forceinline__ uint32_t c4_sw64(uint32_t r, uint32_t c) { return r * 64u + ((c ^ ((r >> 1) & 3u)) << 4); }

// One loader step: `own` is this lane's aligned 16-byte word of the row's 64-byte hull, the neighbour lane holds the next
// one; the 16 bytes wanted start 2*(t mod 8) bytes into the pair.  The shift is the same for every lane of a warp, so it
// is applied with warp-uniform selects (word offset) and a funnel shift (half-word offset): ONE copy of this code serves
// all loader warps — eight compile-time specialisations thrashed the instruction cache (6 us per tile of pure fetch).
// m0 / m1 mask destination chunk 3 down to its single element v = 24 (the rest is the zero padding of the K block).
__device__ __forceinline__ uint4 c4_realign(const uint4& own, bool wo2, bool wo1, uint32_t sh, uint32_t m0, uint32_t m1) {
    const uint32_t n0 = __shfl_down_sync(0xffffffffu, own.x, 1), n1 = __shfl_down_sync(0xffffffffu, own.y, 1);
    const uint32_t n2 = __shfl_down_sync(0xffffffffu, own.z, 1), n3 = __shfl_down_sync(0xffffffffu, own.w, 1);
    const uint32_t a0 = wo2 ? own.z : own.x, a1 = wo2 ? own.w : own.y, a2 = wo2 ? n0 : own.z, a3 = wo2 ? n1 : own.w;
    const uint32_t a4 = wo2 ? n2 : n0, a5 = wo2 ? n3 : n1;
    const uint32_t v0 = wo1 ? a1 : a0, v1 = wo1 ? a2 : a1, v2 = wo1 ? a3 : a2, v3 = wo1 ? a4 : a3, v4 = wo1 ? a5 : a4;
    return make_uint4(__funnelshift_r(v0, v1, sh) & m0, __funnelshift_r(v1, v2, sh) & m1, __funnelshift_r(v2, v3, sh) & m1,
                      __funnelshift_r(v3, v4, sh) & m1);
}

__global__ void __launch_bounds__(C4_THREADS, 1)
ctrgc_fwd_tc4_kernel(C4P p, const bf16* __restrict__ x3, const float* __restrict__ x1, const float* __restrict__ x2,
                     const float* __restrict__ W4, const float* __restrict__ b4, const float* __restrict__ PA,
                     const float* __restrict__ alpha_p, bf16* __restrict__ y, double* ssum, double* ssq) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    C4Hdr* hdr = (C4Hdr*)(smem + p.off_hdr);
    uint8_t* Dt = smem + p.off_D;                    // tanh table [i][u 28][v 25][r 8] fp16
    float* PAs = (float*)(smem + p.off_PA);          // [i][u 28][v 26]
    __half* W4h = (__half*)(smem + p.off_W4);        // [i][c][r 8]
    float* b4s = (float*)(smem + p.off_b4);          // [i][c]
    uint8_t* outs = smem + p.off_out;                // staged output tile
    float* stat = (float*)(smem + p.off_stat);       // [8 epilogue warps][2][Cout]
    float* x12s = (float*)(smem + p.off_x12);        // [2][K*R*V]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile_begin = (int)((long long)blockIdx.x * p.n_tiles / gridDim.x);
    const int tile_end = (int)((long long)(blockIdx.x + 1) * p.n_tiles / gridDim.x);
    const int nt = tile_end - tile_begin;
    const int S = p.S, K = p.K, T = p.T;

    // ---- one-time setup ----
    if (warp == C4_MMA_W) tmem_alloc(&hdr->tmem_base, 256u);
    if (tid == 0) {
        for (int i = 0; i < C4_SMAX; ++i) {
            mbar_init(&hdr->a_full[i], C4_LD_T);
            mbar_init(&hdr->b_full[i], C4_Q_T);
            mbar_init(&hdr->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&hdr->tfull[i], 1); mbar_init(&hdr->tempty[i], C4_EPI_T); }
        hdr->error = 0;
        fence_mbar_init();
    }
    // operand stages, tanh table and PA table start as zeros: padding rows / columns are never written afterwards
    for (uint32_t i = tid; i < ((uint32_t)S * C4_STAGE_BYTES + 3u * C4_D_SUB + 3u * C4_PA_SUB) / 16; i += C4_THREADS)
        st_shared_v4(smem_u32(smem) + i * 16, 0u, 0u, 0u, 0u);       // off_D / off_PA directly follow the stages
    __syncthreads();
    for (int i = tid; i < K * C4_V * C4_V; i += C4_THREADS) {
        const int sub = i / (C4_V * C4_V), rem = i - sub * (C4_V * C4_V), u = rem / C4_V, v = rem - u * C4_V;
        PAs[(sub * C4_UP + u) * C4_PAP + v] = __ldg(PA + i);
    }
    for (int i = tid; i < K * p.Cout * C4_R; i += C4_THREADS) W4h[i] = __float2half_rn(__ldg(W4 + i));
    for (int i = tid; i < K * p.Cout; i += C4_THREADS) b4s[i] = __ldg(b4 + i);
    for (int i = tid; i < 16 * p.Cout; i += C4_THREADS) stat[i] = 0.f;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = hdr->tmem_base;
    const uint32_t s0 = smem_u32(smem);

    if (warp < 8) {
        // =============================== epilogue ===============================
        const int row = tid & 127;                             // TMEM lane == sub-tile row (g, t)
        const int qw = warp & 3, sp = warp >> 2;               // TMEM lane quarter, sub-tile of this warp
        const int g_row = row >> 6, tl = row & 63;
        const int gc = sp * 2 + g_row;                         // channel of this row inside the tile
        float* mystat = stat + warp * 2 * p.Cout;              // this warp's private accumulators: no atomics
        const bool valid = tl < T;
        const uint32_t tile_bytes = (uint32_t)(4 * T * C4_V * 2);
        const uint32_t ob = smem_u32(outs);
        const int orow = gc * T + tl;                          // output row inside the staged tile (50 bytes each)
        const bool odd = (orow & 1) != 0;                      // odd rows start 2 bytes past a word boundary
        const uint32_t oa = ob + (uint32_t)orow * 50u;
        int n = tile_begin / p.nCG, cg = tile_begin - n * p.nCG;
        for (int it = 0; it < nt; ++it) {
            const int c0 = cg * 4;
            const int buf = it & 1;
            if (warp == 0) {
                c4_wait(hdr, &hdr->tfull[buf], (uint32_t)((it >> 1) & 1));
                // the bulk store of the previous tile must have finished reading the (single) staging buffer
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            c4_bar_sync(3, C4_EPI_T);
            tc_fence_after();
            float acc[32];
            const uint32_t ta = tmem + ((uint32_t)(qw * 32) << 16) + (uint32_t)((buf * 2 + sp) * C4_NMMA_N + g_row * 32);
            c4_tmem_ld8(ta, acc); c4_tmem_ld8(ta + 8, acc + 8); c4_tmem_ld8(ta + 16, acc + 16); c4_tmem_ld8(ta + 24, acc + 24);
            c4_tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(&hdr->tempty[buf]);                    // the MMA warp may refill this accumulator buffer
            float s = 0.f, q = 0.f;
            if (valid && !(p.dbg & 4)) {
                uint32_t pk[13];
#pragma unroll
                for (int j = 0; j < 12; ++j) pk[j] = pack_bf16(acc[2 * j], acc[2 * j + 1]);
                pk[12] = pack_bf16(acc[24], 0.f);
#pragma unroll
                for (int j = 0; j < 25; ++j) { s += acc[j]; q = fmaf(acc[j], acc[j], q); }
                // even row: words (e0,e1) ... (e22,e23) at +0, e24 at +48; odd row: e0 at +0, words (e1,e2) ... (e23,e24) at +2
                const uint32_t wbase = oa + (odd ? 2u : 0u);
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                    const uint32_t sh = __funnelshift_r(pk[j], pk[j + 1], 16);
                    c4_sts32(wbase + 4u * j, odd ? sh : pk[j]);
                }
                c4_sts16(odd ? oa : oa + 48u, odd ? pk[0] : pk[12]);
            }
            if (ssum && !(p.dbg & 32)) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s += __shfl_xor_sync(0xffffffffu, s, o);
                    q += __shfl_xor_sync(0xffffffffu, q, o);
                }
                if (lane == 0) {
                    mystat[c0 + gc] += s;
                    mystat[p.Cout + c0 + gc] += q;
                }
            }
            fence_proxy_async_smem();                          // staged rows -> visible to the bulk-copy engine
            c4_bar_sync(1, C4_EPI_T);
            if (tid == 0 && !(p.dbg & 4)) c4_bulk_s2g(y + (long long)n * p.yns + (long long)c0 * T * C4_V, ob, tile_bytes);
            if (++cg == p.nCG) { cg = 0; ++n; }
        }
        if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (ssum) {
            c4_bar_sync(1, C4_EPI_T);
            for (int c = tid; c < p.Cout; c += C4_EPI_T) {
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) { a += stat[2 * w * p.Cout + c]; b += stat[(2 * w + 1) * p.Cout + c]; }
                if (a != 0.f || b != 0.f) {
                    atomicAdd(ssum + c, (double)a);
                    atomicAdd(ssq + c, (double)b);
                }
            }
        }
    } else if (warp == C4_MMA_W) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, C4_NMMA_N);
            // descriptor low words of stage 0 (start address >> 4 | LBO); a stage / sub-tile / subset / K step adds a constant
            const uint32_t dlo_a = ((s0 >> 4) & 0x3FFFu) | (1u << 16), dlo_b = (((s0 + 2u * C4_A_BYTES) >> 4) & 0x3FFFu) | (1u << 16);
            // K-major SWIZZLE_64B: SBO = 512 B (eight 64-byte rows), descriptor version 1, layout type 4
            const uint32_t dhi = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);
            int s = 0, ph = 0;
            for (int it = 0; it < nt; ++it) {
                const int buf = it & 1;
                c4_wait(hdr, &hdr->tempty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
                c4_wait(hdr, &hdr->a_full[s], (uint32_t)ph);
                c4_wait(hdr, &hdr->b_full[s], (uint32_t)ph);
                tc_fence_after();
                const uint32_t so = (uint32_t)s * (C4_STAGE_BYTES >> 4);
#pragma unroll
                for (int sp = 0; sp < 2; ++sp) {
                    const uint32_t td = tmem + (uint32_t)((buf * 2 + sp) * C4_NMMA_N);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        if (i < K && !(p.dbg & 64)) {
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) {
                                const uint32_t alo = dlo_a + so + (uint32_t)((sp * C4_A_BYTES + i * C4_A_BLK) >> 4) + (uint32_t)(kk * 2);
                                const uint32_t blo = dlo_b + so + (uint32_t)((sp * C4_B_BYTES + i * C4_B_BLK) >> 4) + (uint32_t)(kk * 2);
                                asm volatile(
                                    "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                                    "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                                    "setp.ne.b32 p, %4, 0;\n\t"
                                    "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
                                    ::"r"(td), "r"(alo), "r"(blo), "r"(idesc), "r"((i | kk) > 0 ? 1u : 0u), "r"(dhi)
                                    : "memory");
                            }
                        }
                    }
                }
                umma_commit(&hdr->empty[s]);
                umma_commit(&hdr->tfull[buf]);
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 8 && warp < 16) {
        // =============================== x3 loaders ===============================
        // Warp w owns the rows t = w (mod 8): their shift is 2 w bytes.  Loads of a half tile (one sub-tile: 2 channels x K
        // subsets = 6 words per lane) are issued one half tile ahead of their stores, across tile boundaries, so 6..12
        // 16-byte loads per lane are in flight.
        const int lw = warp - 8;
        const int j = lane >> 2, c = lane & 3, t = lw + 8 * j;
        const bool ok = t < T;
        const int tt = ok ? t : lw;                                        // rows past T re-read row lw and store nothing
        const bool wo2 = (lw & 4) != 0, wo1 = (lw & 2) != 0;
        const uint32_t sh = (lw & 1) ? 16u : 0u;
        const uint32_t m0 = c == 3 ? 0xffffu : 0xffffffffu, m1 = c == 3 ? 0u : 0xffffffffu;
        const uint32_t ps = (uint32_t)T * 50u;                              // bytes of one (channel, subset) block
        const uint32_t sub = (uint32_t)p.Cout * ps;                         // bytes between subsets
        const uint32_t lane_src = (((uint32_t)(50 * tt)) & ~15u) + 16u * (uint32_t)c;   // this lane's aligned word of row t
        const uint32_t dst_r = c4_sw64((uint32_t)tt, (uint32_t)c);
        const bool noload = (p.dbg & 1) != 0;
        int n = tile_begin / p.nCG, cg = tile_begin - n * p.nCG;
        auto tile_ptr = [&](int nn, int cgg) {
            return reinterpret_cast<const uint8_t*>(x3 + (long long)nn * p.x3ns) + (uint32_t)(cgg * 4) * ps + lane_src;
        };
        auto load6 = [&](uint4 (&w)[6], const uint8_t* tb) {
#pragma unroll
            for (int g = 0; g < 2; ++g)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int ii = i < K ? i : 0;                            // absent subsets re-read subset 0 and store nothing
                    w[g * 3 + i] = noload ? make_uint4(0u, 0u, 0u, 0u)
                                          : __ldg(reinterpret_cast<const uint4*>(tb + (uint32_t)g * ps + (uint32_t)ii * sub));
                }
        };
        auto store6 = [&](uint4 (&w)[6], uint32_t sa) {
#pragma unroll
            for (int g = 0; g < 2; ++g)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const uint4 v = c4_realign(w[g * 3 + i], wo2, wo1, sh, m0, m1);
                    if (ok && i < K) st_shared_v4(sa + (uint32_t)i * C4_A_BLK + (uint32_t)g * 4096u + dst_r, v.x, v.y, v.z, v.w);
                }
        };
        // L2 prefetch `pf` tiles ahead (loader warp 0, one lane): the 4 channels of a subset are contiguous (4 * ps bytes)
        auto prefetch_tile = [&](int tile) {
            const int pn = tile / p.nCG, pcg = tile - pn * p.nCG;
            const uint8_t* b = reinterpret_cast<const uint8_t*>(x3 + (long long)pn * p.x3ns) + (uint32_t)(pcg * 4) * ps;
            for (int i = 0; i < K; ++i) bulk_prefetch_l2(b + (uint32_t)i * sub, 4u * ps);
        };
        const bool pf_lane = p.pf > 0 && lw == 0 && lane == 0;
        if (pf_lane)
            for (int k = 1; k < p.pf && k < nt; ++k) prefetch_tile(tile_begin + k);
        uint4 wa[6], wb[6];
        if (nt > 0) load6(wa, tile_ptr(n, cg));
        for (int it = 0; it < nt; ++it) {
            const int s = it % S, ph = (it / S) & 1;
            const uint32_t sA = s0 + (uint32_t)s * C4_STAGE_BYTES;
            if (pf_lane && it + p.pf < nt) prefetch_tile(tile_begin + it + p.pf);
            load6(wb, tile_ptr(n, cg) + 2u * ps);
            c4_wait(hdr, &hdr->empty[s], (uint32_t)(ph ^ 1));               // every lane polls: the warp stays converged
            store6(wa, sA);
            if (++cg == p.nCG) { cg = 0; ++n; }
            if (it + 1 < nt) load6(wa, tile_ptr(n, cg));
            store6(wb, sA + C4_A_BYTES);
            fence_proxy_async_smem();
            mbar_arrive(&hdr->a_full[s]);
        }
    } else {
        // =============================== topology builders ===============================
        const int bw = warp - 17, gid = lane >> 2, tig = lane & 3;
        const int qt = bw * 32 + lane;                         // 0..223; warp bw owns the four u of group bw
        const int sel = gid & 3, cc = gid >> 2;                // MMA row gid = (channel cc, u-select sel); row gid+8 = channel cc+2
        const int u0 = 4 * bw, u = u0 + sel;
        const bool u_ok = u < C4_V;
        const float alpha = __ldg(alpha_p);
        const uint32_t m0 = sel == 0 ? 0xffffffffu : 0u, m1 = sel == 1 ? 0xffffffffu : 0u;
        const uint32_t m2 = sel == 2 ? 0xffffffffu : 0u, m3 = sel == 3 ? 0xffffffffu : 0u;
        const uint32_t dt_base = smem_u32(Dt) + (uint32_t)((u0 * C4_V + gid) * 16 + tig * 4);
        const uint32_t pa_base = smem_u32(PAs) + (uint32_t)((u * C4_PAP + 2 * tig) * 4);
        const uint32_t w_base = smem_u32(W4h) + (uint32_t)((cc * C4_R + 2 * tig) * 2);
        const int qrow = cc * 32 + u;                          // row of (g = cc, u) in its B sub-tile
        uint32_t qoff[4];                                      // swizzled byte offset of this lane's bf16 pair inside a subset block, per v block
#pragma unroll
        for (int vb = 0; vb < 4; ++vb) qoff[vb] = c4_sw64((uint32_t)qrow, (uint32_t)vb) + (uint32_t)tig * 4u;
        const int nx12 = K * C4_R * C4_V;
        // x1 / x2 of a sample are fetched into registers one tile before the sample starts
        float xpre[6];
        auto x12_fetch = [&](int n) {
            const float* x1n = x1 + (long long)n * p.x12ns;
            const float* x2n = x2 + (long long)n * p.x12ns;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int idx = qt + k * C4_Q_T;
                xpre[k] = idx < nx12 ? __ldg(x1n + idx) : 0.f;
                xpre[3 + k] = idx < nx12 ? __ldg(x2n + idx) : 0.f;
            }
        };
        int n = tile_begin / p.nCG, cg = tile_begin - n * p.nCG;
        if (nt > 0) x12_fetch(n);
        int cur_n = -1;
        for (int it = 0; it < nt; ++it) {
            const int c0 = cg * 4;
            const int s = it % S, ph = (it / S) & 1;
            int n_next = n, cg_next = cg + 1;
            if (cg_next == p.nCG) { cg_next = 0; ++n_next; }
            if (n != cur_n) {
                // new sample: x1 / x2 -> shared memory, then the tanh table D[i][u][v][r] (fp16, hardware tanh)
                if (cur_n >= 0) c4_bar_sync(2, C4_Q_T);        // everybody is done reading the previous sample's table
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int idx = qt + k * C4_Q_T;
                    if (idx < nx12) {
                        x12s[idx] = xpre[k];
                        x12s[nx12 + idx] = xpre[3 + k];
                    }
                }
                c4_bar_sync(2, C4_Q_T);
                for (int idx = qt; idx < ((p.dbg & 16) ? 0 : K * C4_V * C4_V); idx += C4_Q_T) {
                    const int i = idx / (C4_V * C4_V), rem = idx - i * (C4_V * C4_V), uu = rem / C4_V, vv = rem - uu * C4_V;
                    const float* xa = x12s + i * C4_R * C4_V + uu;
                    const float* xb = x12s + nx12 + i * C4_R * C4_V + vv;
                    uint32_t h[4];
#pragma unroll
                    for (int r = 0; r < C4_R; r += 2) {
                        const float d0 = c4_tanh(xa[r * C4_V] - xb[r * C4_V]);
                        const float d1 = c4_tanh(xa[(r + 1) * C4_V] - xb[(r + 1) * C4_V]);
                        const __half2 hh = __floats2half2_rn(d0, d1);
                        h[r >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    st_shared_v4(smem_u32(Dt) + (uint32_t)(i * C4_D_SUB + (uu * C4_V + vv) * 16), h[0], h[1], h[2], h[3]);
                }
                c4_bar_sync(2, C4_Q_T);
                cur_n = n;
            }
            if (it + 1 < nt && n_next != n) x12_fetch(n_next);
            // this tile's parameters: W4 pairs (r = 2 tig, 2 tig + 1) and b4 of channels cc and cc + 2, per subset
            uint32_t wlo[3], whi[3];
            float blo[3], bhi[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int ic = (i < K ? i : 0) * p.Cout + c0;
                wlo[i] = c4_lds32(w_base + (uint32_t)(ic * C4_R * 2));
                whi[i] = c4_lds32(w_base + (uint32_t)((ic + 2) * C4_R * 2));
                blo[i] = b4s[ic + cc];
                bhi[i] = b4s[ic + 2 + cc];
            }
            c4_wait(hdr, &hdr->empty[s], (uint32_t)(ph ^ 1));
            const uint32_t sB = s0 + (uint32_t)s * C4_STAGE_BYTES + 2u * C4_A_BYTES;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (i < K && !(p.dbg & 2)) {
                    const uint32_t a00 = wlo[i] & m0, a01 = whi[i] & m0, a02 = wlo[i] & m1, a03 = whi[i] & m1;
                    const uint32_t a10 = wlo[i] & m2, a11 = whi[i] & m2, a12 = wlo[i] & m3, a13 = whi[i] & m3;
#pragma unroll
                    for (int vb = 0; vb < 4; ++vb) {
                        const uint32_t da = dt_base + (uint32_t)(i * C4_D_SUB + vb * 128);
                        const uint32_t b00 = c4_lds32(da), b01 = c4_lds32(da + C4_V * 16);
                        const uint32_t b10 = c4_lds32(da + 2 * C4_V * 16), b11 = c4_lds32(da + 3 * C4_V * 16);
                        float c[4] = {blo[i], blo[i], bhi[i], bhi[i]};
                        c4_mma(c, a00, a01, a02, a03, b00, b01);
                        c4_mma(c, a10, a11, a12, a13, b10, b11);
                        const float2 pa = c4_lds_f2(pa_base + (uint32_t)(i * C4_PA_SUB + vb * 32));
                        uint32_t qlo = pack_bf16(fmaf(alpha, c[0], pa.x), fmaf(alpha, c[1], pa.y));
                        uint32_t qhi = pack_bf16(fmaf(alpha, c[2], pa.x), fmaf(alpha, c[3], pa.y));
                        // v = 8 vb + 2 tig (+1): block 3 only holds v = 24 (tig 0, low half); the high half is K padding
                        if (vb == 3) { qlo &= 0xffffu; qhi &= 0xffffu; }
                        if (u_ok && (vb < 3 || tig == 0)) {
                            c4_sts32(sB + (uint32_t)i * C4_B_BLK + qoff[vb], qlo);
                            c4_sts32(sB + C4_B_BYTES + (uint32_t)i * C4_B_BLK + qoff[vb], qhi);
                        }
                    }
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(&hdr->b_full[s]);
            n = n_next; cg = cg_next;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == C4_MMA_W) tmem_dealloc(tmem, 256u);
    if (tid == 0 && hdr->error) printf("tamgcn: ctrgc_fwd(tc4) pipeline timeout in block %d\n", blockIdx.x);
}

static bool c4_disabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TAMGCN_DISABLE_TC4");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static uint32_t c4_al16(uint32_t x) { return (x + 15u) & ~15u; }

// returns 1 if launched, 0 if the shape is not covered (the caller tries ctrgc_tc.cu next), <0 on error
int ctrgc_fwd_tc4(const void* x3, long long x3ns, int N, int Cout, int T, int V, int K, int R, const float* x1,
                  const float* x2, long long x12ns, const float* W4, const float* b4, const float* PA, const float* alpha,
                  void* y, long long yns, double* ssum, double* ssq, cudaStream_t st) {
    if (c4_disabled()) return 0;
    if (V != C4_V || R != C4_R || K < 1 || K > 3 || T <= 32 || T > 64 || (T & 7) || (Cout & 3) || Cout > 256) return 0;
    // aligned 16-byte words of every (channel, subset) block of x3; 16-byte aligned bulk stores of (4 ch x T x V) blocks of y
    if (((uintptr_t)x3 & 15) || (x3ns & 7) || ((uintptr_t)y & 15) || (yns & 7)) return 0;
    C4P p = {};
    p.N = N; p.Cout = Cout; p.T = T; p.K = K;
    p.x3ns = x3ns; p.x12ns = x12ns; p.yns = yns;
    p.nCG = Cout / 4;
    const long long tiles = (long long)N * p.nCG;
    if (tiles > 0x7fffffffLL) return 0;
    p.n_tiles = (int)tiles;
    const uint32_t budget = 227u * 1024u - 1024u;
    const uint32_t szD = 3u * C4_D_SUB, szPA = 3u * C4_PA_SUB;
    const uint32_t szW = c4_al16((uint32_t)(K * Cout * C4_R * 2)), szB = c4_al16((uint32_t)(K * Cout * 4));
    const uint32_t szO = C4_OUT_BYTES, szS = c4_al16((uint32_t)(16 * Cout * 4)), szH = c4_al16((uint32_t)sizeof(C4Hdr));
    const uint32_t szX = c4_al16((uint32_t)(2 * K * C4_R * C4_V * 4));
    const uint32_t fixed = szD + szPA + szW + szB + szO + szS + szH + szX;
    if (fixed + 2u * C4_STAGE_BYTES > budget) return 0;
    p.S = C4_SMAX;
    { const char* e = getenv("TAMGCN_C4_DBG"); p.dbg = e ? atoi(e) : 0; }
    { static const int pf = [] { const char* e = getenv("TAMGCN_C4_PF"); return e ? atoi(e) : 2; }(); p.pf = pf; }
    uint32_t off = (uint32_t)p.S * C4_STAGE_BYTES;
    p.off_D = off; off += szD;                                  // D and PA must directly follow the stages (zero fill)
    p.off_PA = off; off += szPA;
    p.off_W4 = off; off += szW;
    p.off_b4 = off; off += szB;
    p.off_out = off; off += szO;
    p.off_stat = off; off += szS;
    p.off_hdr = off; off += szH;
    p.off_x12 = off; off += szX;
    const size_t sm = (size_t)off + 1024;
    int grid = num_sms();
    if (grid > p.n_tiles) grid = p.n_tiles;
    static SmemLimit lim;
    ensure_smem(ctrgc_fwd_tc4_kernel, lim, sm);
    ctrgc_fwd_tc4_kernel<<<grid, C4_THREADS, sm, st>>>(p, (const bf16*)x3, x1, x2, W4, b4, PA, alpha, (bf16*)y, ssum, ssq);
    count_launch();
    const int rc = check_launch("ctrgc_fwd(tc4)");
    return rc < 0 ? rc : 1;
}

}  // namespace tamgcn
