// ctrgc_tc3.cu — fused CTRGC forward, bf16, V = 20, R = 8: the instruction-lean variant of ctrgc_tc.cu.
//
//   y[n,c,t,u] = sum_i sum_v Q_i[n,c,u,v] * x3_i[n,c,t,v]      (reference models/ctrgcn.py:174-176, :252-254)
//   Q_i[n,c,u,v] = alpha * (sum_r W4_i[c,r] tanh(x1_i[n,r,u] - x2_i[n,r,v]) + b4_i[c]) + PA_i[u,v]
//
// ctrgc_tc.cu is bound by the instruction issue rate of its CUDA-core roles (ncu, profiles/r01k_ctrgc_tc_full.txt:
// ~8400 warp instructions per 4-channel tile, a third of them the packed-half FMA chains that build Q).  Same tile,
// same operand layouts and the same tcgen05 main contraction here, but
//   * Q is built on the warp-level tensor cores: one mma.sync.m16n8k16 pair produces S[c, u, v] for 4 channels x 4 u x
//     8 v.  Rows are (channel, u-select), K = (u-select', r): the A fragment holds W4[c, r] where select == select'
//     and zeros elsewhere, the B fragment is read straight from the per-sample tanh table (fp16, r contiguous), the
//     accumulator starts at b4[c].  ~20 instructions per 128 Q values instead of ~75; fp32 accumulation.
//   * W4 / b4 of ALL channels are converted to fp16 once per CTA (3 KB at 64 channels): no per-tile parameter
//     staging, prefetch or barrier.
//   * the finished tile leaves shared memory as ONE bulk asynchronous store (cp.async.bulk, the TMA engine): the
//     (4 channels x T x V) block is contiguous in y.
//   * one warp per role polls an mbarrier, the others park on a hardware barrier.
//
// tile      = one sample n, 4 output channels (2 sub-tiles of 2 channels x 64 time rows = UMMA M 128)
// A (smem)  = x3 rows [(g,t)] x [(i,v)] bf16, K-major SWIZZLE_128B, 8-byte cp.async gather from the K planes
// B (smem)  = Q rows  [(g,u)] x [(i,v)] bf16, K-major SWIZZLE_128B, written by the builder warps
// D (TMEM)  = [(g,t)] x [(g',u)] fp32, double buffered; only the diagonal blocks g == g' are read back
// warps 0-7 epilogue | warp 8 MMA issue | warps 9-12 x3 loaders | warps 13-17 topology builders (one per 4 u)
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace tamgcn {

#define C3_EPI_T 256
#define C3_MMA_W 14
#define C3_LD_T0 256
#define C3_LD_T 128
#define C3_Q_T0 416
#define C3_Q_T 160
#define C3_THREADS 576
#define C3_MAXU 15               // 8-byte units of a sub-tile per loader thread (register table): 2 ch x 3 planes x 64 t x 5 / 128
#define C3_SMAX 4
#define C3_V 20
#define C3_VP 24                 // padded v of the tanh / PA tables
#define C3_R 8
#define C3_NMMA_N 48             // UMMA N: 2 channels x 24 row slots
#define C3_A_BYTES 16384u        // 128 rows x 128 B
#define C3_B_BYTES 6144u         // 48 rows x 128 B
#define C3_STAGE_BYTES (2u * (C3_A_BYTES + C3_B_BYTES))
#define C3_OUT_BYTES 10240u      // staging of one finished tile (4 ch x T<=64 x 20 x 2 B)

struct C3P {
    int N, Cout, T, K;
    long long x3ns, x12ns, yns;
    int nCG, n_tiles, S;
    int pf;                      // L2 prefetch distance in tiles (0 = off)
    uint32_t off_D, off_PA, off_W4, off_b4, off_out, off_stat, off_hdr, off_x12;
};

struct C3Hdr {
    uint64_t a_full[C3_SMAX], b_full[C3_SMAX], empty[C3_SMAX], tfull[2], tempty[2];
    uint32_t tmem_base;
    volatile uint32_t error;
};

__device__ __forceinline__ void c3_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void c3_cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void c3_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void c3_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void c3_sts64(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void c3_sts32(uint32_t a, uint32_t x) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }
__device__ __forceinline__ uint32_t c3_lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 c3_lds_f2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void c3_tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void c3_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float c3_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// D(16x8, fp32) += A(16x16, fp16, row) * B(16x8, fp16, col)
__device__ __forceinline__ void c3_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void c3_bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// bounded spin wait (no suspend hint); a timeout flags the CTA so that the host sees a message instead of a hang
__device__ __forceinline__ bool c3_wait(C3Hdr* hdr, uint64_t* bar, uint32_t parity) {
    if (mbar_wait_spin(bar, parity)) return true;
    hdr->error = 1;
    return false;
}

__global__ void __launch_bounds__(C3_THREADS, 1)
ctrgc_fwd_tc3_kernel(C3P p, const bf16* __restrict__ x3, const float* __restrict__ x1, const float* __restrict__ x2,
                     const float* __restrict__ W4, const float* __restrict__ b4, const float* __restrict__ PA,
                     const float* __restrict__ alpha_p, bf16* __restrict__ y, double* ssum, double* ssq) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    C3Hdr* hdr = (C3Hdr*)(smem + p.off_hdr);
    uint8_t* Dt = smem + p.off_D;                    // tanh table [i][u][v 24][r 8] fp16
    float* PAs = (float*)(smem + p.off_PA);          // [i][u][v 24]
    __half* W4h = (__half*)(smem + p.off_W4);        // [i][c][r 8]
    float* b4s = (float*)(smem + p.off_b4);          // [i][c]
    uint8_t* outs = smem + p.off_out;                // 2 staged output tiles
    float* stat = (float*)(smem + p.off_stat);       // [8 epilogue warps][2][Cout]
    float* x12s = (float*)(smem + p.off_x12);        // [2][K*R*V]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile_begin = (int)((long long)blockIdx.x * p.n_tiles / gridDim.x);
    const int tile_end = (int)((long long)(blockIdx.x + 1) * p.n_tiles / gridDim.x);
    const int nt = tile_end - tile_begin;
    const int S = p.S, K = p.K, T = p.T;

    // ---- one-time setup ----
    if (warp == C3_MMA_W) tmem_alloc(&hdr->tmem_base, 256u);
    if (tid == 0) {
        for (int i = 0; i < C3_SMAX; ++i) {
            mbar_init(&hdr->a_full[i], C3_LD_T);
            mbar_init(&hdr->b_full[i], C3_Q_T);
            mbar_init(&hdr->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&hdr->tfull[i], 1); mbar_init(&hdr->tempty[i], C3_EPI_T); }
        hdr->error = 0;
        fence_mbar_init();
    }
    // operand stages and the tanh table start as zeros: padding rows / columns are never written afterwards
    for (uint32_t i = tid; i < ((uint32_t)S * C3_STAGE_BYTES + (uint32_t)(3 * C3_V * C3_VP * 16)) / 16; i += C3_THREADS)
        st_shared_v4(smem_u32(smem) + i * 16, 0u, 0u, 0u, 0u);       // off_D directly follows the stages
    for (int i = tid; i < K * C3_V * C3_VP; i += C3_THREADS) {
        const int v = i % C3_VP, iu = i / C3_VP;
        PAs[i] = (v < C3_V) ? __ldg(PA + iu * C3_V + v) : 0.f;
    }
    for (int i = tid; i < K * p.Cout * C3_R; i += C3_THREADS) W4h[i] = __float2half_rn(__ldg(W4 + i));
    for (int i = tid; i < K * p.Cout; i += C3_THREADS) b4s[i] = __ldg(b4 + i);
    for (int i = tid; i < 16 * p.Cout; i += C3_THREADS) stat[i] = 0.f;
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
        const uint32_t tile_bytes = (uint32_t)(4 * T * C3_V * 2);
        int n = tile_begin / p.nCG, cg = tile_begin - n * p.nCG;
        for (int it = 0; it < nt; ++it) {
            const int c0 = cg * 4;
            const int buf = it & 1;
            const uint32_t ob = smem_u32(outs) + (uint32_t)buf * C3_OUT_BYTES;
            if (warp == 0) {
                c3_wait(hdr, &hdr->tfull[buf], (uint32_t)((it >> 1) & 1));
                // the bulk store issued two tiles ago (same staging buffer) must have finished reading shared memory
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            c3_bar_sync(3, C3_EPI_T);
            tc_fence_after();
            float acc[24];
            const uint32_t ta = tmem + ((uint32_t)(qw * 32) << 16) + (uint32_t)((buf * 2 + sp) * C3_NMMA_N + g_row * 24);
            c3_tmem_ld8(ta, acc); c3_tmem_ld8(ta + 8, acc + 8); c3_tmem_ld8(ta + 16, acc + 16);
            c3_tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(&hdr->tempty[buf]);                    // the MMA warp may refill this accumulator buffer
            float s = 0.f, q = 0.f;
            if (valid) {
                const uint32_t oa = ob + (uint32_t)((gc * T + tl) * C3_V * 2);
#pragma unroll
                for (int j = 0; j < 20; j += 4) {
                    const uint32_t w0 = pack_bf16(acc[j], acc[j + 1]), w1 = pack_bf16(acc[j + 2], acc[j + 3]);
                    s += (acc[j] + acc[j + 1]) + (acc[j + 2] + acc[j + 3]);
                    q = fmaf(acc[j], acc[j], q); q = fmaf(acc[j + 1], acc[j + 1], q); q = fmaf(acc[j + 2], acc[j + 2], q); q = fmaf(acc[j + 3], acc[j + 3], q);
                    c3_sts64(oa + j * 2, w0, w1);
                }
            }
            if (ssum) {
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
            c3_bar_sync(1, C3_EPI_T);
            if (tid == 0) c3_bulk_s2g(y + (long long)n * p.yns + (long long)c0 * T * C3_V, ob, tile_bytes);
            if (++cg == p.nCG) { cg = 0; ++n; }
        }
        if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (ssum) {
            c3_bar_sync(1, C3_EPI_T);
            for (int c = tid; c < p.Cout; c += C3_EPI_T) {
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) { a += stat[2 * w * p.Cout + c]; b += stat[(2 * w + 1) * p.Cout + c]; }
                if (a != 0.f || b != 0.f) {
                    atomicAdd(ssum + c, (double)a);
                    atomicAdd(ssq + c, (double)b);
                }
            }
        }
    } else if (warp == C3_MMA_W) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, C3_NMMA_N);
            // descriptor low words of stage 0 (start address >> 4 | LBO); a stage / sub-tile / K step only adds a constant
            const uint32_t dlo_a = ((s0 >> 4) & 0x3FFFu) | (1u << 16), dlo_b = (((s0 + 2u * C3_A_BYTES) >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t dhi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            int s = 0, ph = 0;
            for (int it = 0; it < nt; ++it) {
                const int buf = it & 1;
                c3_wait(hdr, &hdr->tempty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
                c3_wait(hdr, &hdr->a_full[s], (uint32_t)ph);
                c3_wait(hdr, &hdr->b_full[s], (uint32_t)ph);
                tc_fence_after();
                const uint32_t so = (uint32_t)s * (C3_STAGE_BYTES >> 4);
#pragma unroll
                for (int sp = 0; sp < 2; ++sp) {
                    const uint32_t td = tmem + (uint32_t)((buf * 2 + sp) * C3_NMMA_N);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t alo = dlo_a + so + (uint32_t)(sp * (C3_A_BYTES >> 4) + kk * 2);
                        const uint32_t blo = dlo_b + so + (uint32_t)(sp * (C3_B_BYTES >> 4) + kk * 2);
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                            "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                            "setp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
                            ::"r"(td), "r"(alo), "r"(blo), "r"(idesc), "r"(kk > 0 ? 1u : 0u), "r"(dhi)
                            : "memory");
                    }
                }
                umma_commit(&hdr->empty[s]);
                umma_commit(&hdr->tfull[buf]);
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 8 && warp < 12) {
        // =============================== x3 loaders ===============================
        const int lt = tid - C3_LD_T0;
        const int lag = S >= 4 ? 2 : S - 2;                    // tiles of cp.async in flight beyond the current one
        // (source offset, swizzled destination offset) of this thread's 8-byte units of a sub-tile (2 channels x K
        // planes x T rows): the same for every tile, kept in registers
        int soff[C3_MAXU];
        uint32_t doff[C3_MAXU];
        const unsigned nuf = (unsigned)T * 5u, totf = (unsigned)(2 * K) * nuf;
        const int kfull = (int)(totf / C3_LD_T);
#pragma unroll
        for (int k = 0; k < C3_MAXU; ++k) {
            const unsigned idx = lt + k * C3_LD_T;
            soff[k] = -1; doff[k] = 0;
            if (idx < totf) {
                const unsigned gi = idx / nuf, j = idx - gi * nuf, g = gi / K, i = gi - g * K, t = j / 5u, q = j - 5u * t;
                const unsigned row = g * 64u + t, byte = i * 40u + 8u * q;
                soff[k] = (int)((i * p.Cout + g) * T * 20 + j * 4);
                doff[k] = row * 128u + (((byte >> 4) ^ (row & 7u)) << 4) + (byte & 15u);
            }
        }
        // L2 prefetch `pf` tiles ahead (one lane): the 4 channels of a subset are contiguous (4 * T * 40 bytes)
        const uint32_t pbytes = (uint32_t)(4 * T * C3_V * 2);
        auto prefetch_tile = [&](int tile) {
            const int pn = tile / p.nCG, pcg = tile - pn * p.nCG;
            const bf16* b = x3 + (long long)pn * p.x3ns + (long long)(pcg * 4) * T * C3_V;
            for (int i = 0; i < K; ++i) bulk_prefetch_l2(b + (long long)i * p.Cout * T * C3_V, pbytes);
        };
        const bool pf_lane = p.pf > 0 && lt == 32 && (pbytes & 15u) == 0;   // second loader warp: the first one polls the barrier
        if (pf_lane)
            for (int k = 1; k < p.pf && k < nt; ++k) prefetch_tile(tile_begin + k);
        int n = tile_begin / p.nCG, cg = tile_begin - n * p.nCG;
        for (int it = 0; it < nt; ++it) {
            const int c0 = cg * 4;
            const int s = it % S, ph = (it / S) & 1;
            const uint32_t sA = s0 + (uint32_t)s * C3_STAGE_BYTES;
            const bf16* xn = x3 + (long long)n * p.x3ns;
            if (pf_lane && it + p.pf < nt) prefetch_tile(tile_begin + it + p.pf);
            if (lt < 32) c3_wait(hdr, &hdr->empty[s], (uint32_t)(ph ^ 1));
            c3_bar_sync(4, C3_LD_T);
            // units k < kfull exist for every thread of the role: issued without a predicate (a predicated cp.async costs
            // a compare and two descriptor moves on top of its address arithmetic)
            if (kfull >= 12) {
#pragma unroll
                for (int sp = 0; sp < 2; ++sp) {
                    const uint32_t sa = sA + (uint32_t)sp * C3_A_BYTES;
                    const bf16* base = xn + (long long)(c0 + sp * 2) * T * 20;
#pragma unroll
                    for (int k = 0; k < 12; ++k) c3_cp_async8(sa + doff[k], base + soff[k]);
#pragma unroll
                    for (int k = 12; k < C3_MAXU; ++k)
                        if (soff[k] >= 0) c3_cp_async8(sa + doff[k], base + soff[k]);
                }
            } else {
#pragma unroll
                for (int sp = 0; sp < 2; ++sp) {
                    const uint32_t sa = sA + (uint32_t)sp * C3_A_BYTES;
                    const bf16* base = xn + (long long)(c0 + sp * 2) * T * 20;
#pragma unroll
                    for (int k = 0; k < C3_MAXU; ++k)
                        if (soff[k] >= 0) c3_cp_async8(sa + doff[k], base + soff[k]);
                }
            }
            c3_cp_commit();
            if (it >= lag) {
                if (lag == 2) c3_cp_wait<2>();
                else if (lag == 1) c3_cp_wait<1>();
                else c3_cp_wait<0>();
                fence_proxy_async_smem();
                mbar_arrive(&hdr->a_full[(it - lag) % S]);
            }
            if (++cg == p.nCG) { cg = 0; ++n; }
        }
        c3_cp_wait<0>();
        fence_proxy_async_smem();
        for (int it = max(nt - lag, 0); it < nt; ++it) mbar_arrive(&hdr->a_full[it % S]);
    } else {
        // =============================== topology builders ===============================
        // builder warps 12, 13, 15, 16, 17 (warp 14 issues the MMAs: it shares its scheduler with three warps, not four)
        const int bw = warp < 14 ? warp - 12 : warp - 13, gid = lane >> 2, tig = lane & 3;
        const int qt = bw * 32 + lane;                         // 0..159; warp bw owns the four u of group bw
        const int sel = gid & 3, cc = gid >> 2;                // MMA row gid = (channel cc, u-select sel); row gid+8 = channel cc+2
        const int u0 = 4 * bw, u = u0 + sel;
        const float alpha = __ldg(alpha_p);
        const uint32_t m0 = sel == 0 ? 0xffffffffu : 0u, m1 = sel == 1 ? 0xffffffffu : 0u;
        const uint32_t m2 = sel == 2 ? 0xffffffffu : 0u, m3 = sel == 3 ? 0xffffffffu : 0u;
        const uint32_t dt_base = smem_u32(Dt) + (uint32_t)((u0 * C3_VP + gid) * 16 + tig * 4);
        const uint32_t pa_base = smem_u32(PAs) + (uint32_t)((u * C3_VP + 2 * tig) * 4);
        const uint32_t w_base = smem_u32(W4h) + (uint32_t)((cc * C3_R + 2 * tig) * 2);
        const int qrow = cc * 24 + u;                          // row of (g = cc, u) in its B sub-tile
        uint32_t qoff[9];                                      // swizzled byte offset of this lane's bf16 pair, per (plane, v block)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int vb = 0; vb < 3; ++vb) {
                const int col = i * 20 + vb * 8 + 2 * tig;
                qoff[i * 3 + vb] = (uint32_t)(qrow * 128 + ((((col >> 3) ^ (qrow & 7)) & 7) << 4) + (col & 7) * 2);
            }
        const bool tail_ok = tig < 2;                          // v block 2 covers v = 16 + 2 tig: only v < 20 exists
        // x1 / x2 of a sample are fetched into registers one tile before the sample starts
        float xpre[6];
        auto x12_fetch = [&](int n) {
            const float* x1n = x1 + (long long)n * p.x12ns;
            const float* x2n = x2 + (long long)n * p.x12ns;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                xpre[k] = k < K ? __ldg(x1n + qt + k * C3_Q_T) : 0.f;
                xpre[3 + k] = k < K ? __ldg(x2n + qt + k * C3_Q_T) : 0.f;
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
                if (cur_n >= 0) c3_bar_sync(2, C3_Q_T);        // everybody is done reading the previous sample's table
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (k < K) {
                        x12s[qt + k * C3_Q_T] = xpre[k];
                        x12s[K * C3_R * C3_V + qt + k * C3_Q_T] = xpre[3 + k];
                    }
                c3_bar_sync(2, C3_Q_T);
                for (int idx = qt; idx < K * C3_V * C3_V; idx += C3_Q_T) {
                    const int i = idx / (C3_V * C3_V), rem = idx - i * (C3_V * C3_V), uu = rem / C3_V, vv = rem - uu * C3_V;
                    const float* xa = x12s + i * C3_R * C3_V + uu;
                    const float* xb = x12s + K * C3_R * C3_V + i * C3_R * C3_V + vv;
                    uint32_t h[4];
#pragma unroll
                    for (int r = 0; r < C3_R; r += 2) {
                        const float d0 = c3_tanh(xa[r * C3_V] - xb[r * C3_V]);
                        const float d1 = c3_tanh(xa[(r + 1) * C3_V] - xb[(r + 1) * C3_V]);
                        const __half2 hh = __floats2half2_rn(d0, d1);
                        h[r >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    st_shared_v4(smem_u32(Dt) + (uint32_t)(((i * C3_V + uu) * C3_VP + vv) * 16), h[0], h[1], h[2], h[3]);
                }
                c3_bar_sync(2, C3_Q_T);
                cur_n = n;
            }
            if (it + 1 < nt && n_next != n) x12_fetch(n_next);
            // this tile's parameters: W4 pairs (r = 2 tig, 2 tig + 1) and b4 of channels cc and cc + 2, per plane
            uint32_t wlo[3], whi[3];
            float blo[3], bhi[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int ic = (i < K ? i : 0) * p.Cout + c0;
                wlo[i] = c3_lds32(w_base + (uint32_t)(ic * C3_R * 2));
                whi[i] = c3_lds32(w_base + (uint32_t)((ic + 2) * C3_R * 2));
                blo[i] = b4s[ic + cc];
                bhi[i] = b4s[ic + 2 + cc];
            }
            c3_wait(hdr, &hdr->empty[s], (uint32_t)(ph ^ 1));
            const uint32_t sB = s0 + (uint32_t)s * C3_STAGE_BYTES + 2u * C3_A_BYTES;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (i < K) {
                    const uint32_t a00 = wlo[i] & m0, a01 = whi[i] & m0, a02 = wlo[i] & m1, a03 = whi[i] & m1;
                    const uint32_t a10 = wlo[i] & m2, a11 = whi[i] & m2, a12 = wlo[i] & m3, a13 = whi[i] & m3;
#pragma unroll
                    for (int vb = 0; vb < 3; ++vb) {
                        const uint32_t da = dt_base + (uint32_t)(i * (C3_V * C3_VP * 16) + vb * 128);
                        const uint32_t b00 = c3_lds32(da), b01 = c3_lds32(da + C3_VP * 16);
                        const uint32_t b10 = c3_lds32(da + 2 * C3_VP * 16), b11 = c3_lds32(da + 3 * C3_VP * 16);
                        float c[4] = {blo[i], blo[i], bhi[i], bhi[i]};
                        c3_mma(c, a00, a01, a02, a03, b00, b01);
                        c3_mma(c, a10, a11, a12, a13, b10, b11);
                        const float2 pa = c3_lds_f2(pa_base + (uint32_t)(i * (C3_V * C3_VP * 4) + vb * 32));
                        const uint32_t qlo = pack_bf16(fmaf(alpha, c[0], pa.x), fmaf(alpha, c[1], pa.y));
                        const uint32_t qhi = pack_bf16(fmaf(alpha, c[2], pa.x), fmaf(alpha, c[3], pa.y));
                        if (vb < 2 || tail_ok) {
                            c3_sts32(sB + qoff[i * 3 + vb], qlo);
                            c3_sts32(sB + C3_B_BYTES + qoff[i * 3 + vb], qhi);
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
    if (warp == C3_MMA_W) tmem_dealloc(tmem, 256u);
    if (tid == 0 && hdr->error) printf("tamgcn: ctrgc_fwd(tc3) pipeline timeout in block %d\n", blockIdx.x);
}

static int c3_num_sms() { return num_sms(); }

static bool c3_disabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TAMGCN_DISABLE_TC3");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static uint32_t c3_al16(uint32_t x) { return (x + 15u) & ~15u; }

// returns 1 if launched, 0 if the shape is not covered (the caller tries ctrgc_tc.cu next), <0 on error
int ctrgc_fwd_tc3(const void* x3, long long x3ns, int N, int Cout, int T, int V, int K, int R, const float* x1,
                  const float* x2, long long x12ns, const float* W4, const float* b4, const float* PA, const float* alpha,
                  void* y, long long yns, double* ssum, double* ssq, cudaStream_t st) {
    if (c3_disabled()) return 0;
    if (V != C3_V || R != C3_R || K < 1 || K > 3 || T <= 32 || T > 64 || (Cout & 3) || Cout > 512) return 0;
    // 8-byte gathers of x3 rows, 16-byte aligned bulk stores of (4 channels x T x V) blocks of y
    if (((uintptr_t)x3 & 7) || (x3ns & 3) || ((uintptr_t)y & 15) || (yns & 7) || ((4 * T * C3_V * 2) & 15)) return 0;
    C3P p = {};
    p.N = N; p.Cout = Cout; p.T = T; p.K = K;
    p.x3ns = x3ns; p.x12ns = x12ns; p.yns = yns;
    p.nCG = Cout / 4;
    const long long tiles = (long long)N * p.nCG;
    if (tiles > 0x7fffffffLL) return 0;
    p.n_tiles = (int)tiles;
    const uint32_t budget = 227u * 1024u - 1024u;
    const uint32_t szD = 3u * C3_V * C3_VP * 16u, szPA = c3_al16((uint32_t)(K * C3_V * C3_VP * 4));
    const uint32_t szW = c3_al16((uint32_t)(K * Cout * C3_R * 2)), szB = c3_al16((uint32_t)(K * Cout * 4));
    const uint32_t szO = 2u * C3_OUT_BYTES, szS = c3_al16((uint32_t)(16 * Cout * 4)), szH = c3_al16((uint32_t)sizeof(C3Hdr));
    const uint32_t szX = c3_al16((uint32_t)(2 * K * C3_R * C3_V * 4));
    const uint32_t fixed = szD + szPA + szW + szB + szO + szS + szH + szX;
    if (fixed + 2u * C3_STAGE_BYTES > budget) return 0;
    p.S = (int)((budget - fixed) / C3_STAGE_BYTES);
    if (p.S > C3_SMAX) p.S = C3_SMAX;
    uint32_t off = (uint32_t)p.S * C3_STAGE_BYTES;
    p.off_D = off; off += szD;                                  // must directly follow the stages (zero fill)
    p.off_PA = off; off += szPA;
    p.off_W4 = off; off += szW;
    p.off_b4 = off; off += szB;
    p.off_out = off; off += szO;
    p.off_stat = off; off += szS;
    p.off_hdr = off; off += szH;
    p.off_x12 = off; off += szX;
    const size_t sm = (size_t)off + 1024;
    // x3 rows must be 16-byte aligned blocks for the bulk prefetch (the 8-byte gathers only need 8)
    { static const int pf = [] { const char* e = getenv("TAMGCN_C3_PF"); return e ? atoi(e) : 0; }();
      p.pf = (((uintptr_t)x3 & 15) == 0 && (x3ns & 7) == 0) ? pf : 0; }
    int grid = c3_num_sms();
    if (grid > p.n_tiles) grid = p.n_tiles;
    static SmemLimit lim;
    ensure_smem(ctrgc_fwd_tc3_kernel, lim, sm);
    ctrgc_fwd_tc3_kernel<<<grid, C3_THREADS, sm, st>>>(p, (const bf16*)x3, x1, x2, W4, b4, PA, alpha, (bf16*)y, ssum, ssq);
    count_launch();
    const int rc = check_launch("ctrgc_fwd(tc3)");
    return rc < 0 ? rc : 1;
}

}  // namespace tamgcn
