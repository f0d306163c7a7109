// conv_wg2.cu — (k x 1) convolution weight gradient on tcgen05 (bf16 storage, fp32 accumulation in TMEM).
//
//   dW[co, ci, j] += sum_{n,pos} dY(n, co, pos) * X(n, ci, in(pos, j))        dbias[co] += sum_{n,pos} dY(n, co, pos)
//
// Both operands are K-major in global memory already (K = output positions, which are contiguous), so a CTA streams
// 64-position chunks of   A = dY rows [128 co] x [64 pos]   and, per tap,   B_j = X rows [NT ci] x [64 pos]
// with 16/8-byte cp.async into SWIZZLE_128B tiles, transforms lazy operands (BatchNorm-backward affine of dY,
// BatchNorm-apply + ReLU of X) in place, and a single thread issues 4 x k MMAs (128 x NT x 16) per chunk that
// accumulate into k column groups of TMEM for the whole kernel.  An extra all-ones row of B_0 makes the tensor core
// produce dbias as one more accumulator column.  Split-K over CTAs (persistent, one per SM), fp32 atomics at the end.
#include "tc_common.cuh"
#include <cstdlib>

namespace tamgcn {

struct ConvP {
    int N, Cin, Cout, T, To, V, k, s, d, p;
};

#define W2_PR_T 384           // warps 0-11: producers (and the final TMEM drain)
#define W2_MMA_W 12           // warp 12: MMA issue
#define W2_THREADS 416
#define W2_SMAX 6

struct W2P {
    ConvP g;
    int Lin, Lout, NT, NTp, nchunk, units;
    int gran_a, gran_b, fast, S, lag, tmem_cols, has_bias, flat, rab, raa;
    uint32_t a_bytes, aq_bytes, b_bytes, bq_bytes, stage_bytes, off_hdr, off_coef;
};

struct W2Hdr {
    uint64_t full[W2_SMAX], empty[W2_SMAX], done;
    uint32_t tmem_base;
    volatile uint32_t error;
};

__device__ __forceinline__ void w2_cp16(uint32_t dst, const void* src, uint32_t nbytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void w2_cp8(uint32_t dst, const void* src, uint32_t nbytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void w2_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void w2_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float w2_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float w2_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// byte offset of positions [pu, pu + gran) of row r inside a K-major SW128 tile (rows of 64 positions)
__device__ __forceinline__ uint32_t w2_off(uint32_t r, uint32_t pu) {
    return r * 128u + (((pu >> 3) ^ (r & 7u)) << 4) + (pu & 7u) * 2u;
}

// in-place transform of GR (8 or 4) landed elements: f(a*p + b*q + c)
template <int GR>
__device__ __forceinline__ void w2_xform(uint32_t dst, uint32_t dstq, bool has_q, float a, float b, float c, int relu) {
    uint32_t w[4], q[4] = {0u, 0u, 0u, 0u};
    if (GR == 8) {
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(dst));
        if (has_q) asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(dstq));
    } else {
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(dst));
        if (has_q) asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(q[0]), "=r"(q[1]) : "r"(dstq));
    }
    uint32_t o[GR / 2];
#pragma unroll
    for (int e = 0; e < GR / 2; ++e) {
        float lo = fmaf(a, w2_lo(w[e]), c), hi = fmaf(a, w2_hi(w[e]), c);
        if (has_q) { lo = fmaf(b, w2_lo(q[e]), lo); hi = fmaf(b, w2_hi(q[e]), hi); }
        if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
        o[e] = pack_bf16(lo, hi);
    }
    if (GR == 8) st_shared_v4(dst, o[0], o[1], o[2], o[3]);
    else asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(o[0]), "r"(o[1]) : "memory");
}

__global__ void __launch_bounds__(W2_THREADS, 1)
conv_wg2_kernel(W2P p, Opnd dyo, Opnd xo, float* __restrict__ dW, float* __restrict__ dbias) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    W2Hdr* hdr = (W2Hdr*)(smem + p.off_hdr);
    float* coefA = (float*)(smem + p.off_coef);          // [3][128]
    float* coefB = coefA + 3 * 128;                      // [3][NT]
    const ConvP g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.S, NT = p.NT, NTp = p.NTp, k = g.k, Lin = p.Lin, Lout = p.Lout;
    const int co0 = blockIdx.x * 128, ci0 = blockIdx.y * NT;
    const int nco = min(128, g.Cout - co0), nci = min(NT, g.Cin - ci0);
    const bool do_bias = p.has_bias && blockIdx.y == 0;
    const uint32_t s0 = smem_u32(smem);

    if (warp == W2_MMA_W) tmem_alloc(&hdr->tmem_base, (uint32_t)p.tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < W2_SMAX; ++i) { mbar_init(&hdr->full[i], W2_PR_T); mbar_init(&hdr->empty[i], 1); }
        mbar_init(&hdr->done, 1);
        hdr->error = 0;
        fence_mbar_init();
    }
    for (int i = tid; i < 128; i += W2_THREADS) {
        const OpCoef cf = opnd_coef(dyo, min(co0 + i, g.Cout - 1));
        coefA[i] = cf.a; coefA[128 + i] = cf.b; coefA[256 + i] = cf.c;
    }
    for (int i = tid; i < NT; i += W2_THREADS) {
        const OpCoef cf = opnd_coef(xo, min(ci0 + i, g.Cin - 1));
        coefB[i] = cf.a; coefB[NT + i] = cf.b; coefB[2 * NT + i] = cf.c;
    }
    // rows that are never written must read as zero (co >= nco, ci >= nci, padding rows), the ones row as 1.0
    for (uint32_t i = tid; i < (uint32_t)S * p.stage_bytes / 16; i += W2_THREADS) st_shared_v4(s0 + i * 16, 0u, 0u, 0u, 0u);
    __syncthreads();
    if (do_bias)
        for (int i = tid; i < S * 8; i += W2_THREADS) {
            const uint32_t a = s0 + (uint32_t)(i >> 3) * p.stage_bytes + p.a_bytes + p.aq_bytes + w2_off((uint32_t)NT, (uint32_t)(i & 7) * 8u);
            st_shared_v4(a, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
        }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = hdr->tmem_base;
    const int nmine = (p.units - (int)blockIdx.z + (int)gridDim.z - 1) / (int)gridDim.z;     // chunks of this CTA

    if (warp < W2_MMA_W) {
        // =============================== producers ===============================
        const bf16* ap = (const bf16*)dyo.p; const bf16* aq = (const bf16*)dyo.q;
        const bf16* bp = (const bf16*)xo.p;  const bf16* bq = (const bf16*)xo.q;
        const bool a_q = aq != nullptr, b_q = bq != nullptr;
        const bool a_lazy = dyo.a || dyo.b || dyo.c || a_q || dyo.relu, b_lazy = xo.a || xo.b || xo.c || b_q || xo.relu;
        const int ga = p.gran_a, gb = p.gran_b, upa = 64 / ga, upb = 64 / gb;
        const int lga = ga == 8 ? 3 : 2, lgb = gb == 8 ? 3 : 2;      // log2(units per row): 8 -> 8 units, 4 -> 16 units
        const int sha = 6 - lga, shb = 6 - lgb;                      // idx >> sh = row
        const int lag = p.lag;
        int u0 = 0, u1 = 0, u2 = 0, u3 = 0, u4 = 0;                   // chunk ids in flight, newest first
        // source offset of a B unit or -1
        auto b_src = [&](int j, int pos) -> int {
            if (pos >= Lout) return -1;
            if (p.fast) return pos;
            const int tq = pos / g.V, v = pos - tq * g.V;
            const int t = tq * g.s + j * g.d - g.p;
            return (t >= 0 && t < g.T) ? t * g.V + v : -1;
        };
        // A chunk is 64 consecutive positions.  Per-sample chunking wastes most of the last chunk of every sample when
        // T*V is just above a multiple of 64 (260 -> 5 chunks for 4.06); for 1x1 stride-1 convolutions the chunks tile the
        // flat (sample, position) axis instead and a chunk may continue into the next sample (p.flat; T*V >= 64, T*V a
        // multiple of the copy granularity, so no unit straddles the boundary).
        auto chunk_origin = [&](int u, int& n, int& pos0) {
            if (p.flat) { const long long f0 = (long long)u * 64; n = (int)(f0 / Lout); pos0 = (int)(f0 - (long long)n * Lout); }
            else { n = u / p.nchunk; pos0 = (u - n * p.nchunk) * 64; }
        };
        auto unit_valid = [&](int n, int pos) -> bool { return p.flat ? (n + (pos >= Lout ? 1 : 0) < g.N) : (pos < Lout); };
        auto retire = [&](int age, int newest) {
            int sp = newest - age; if (sp < 0) sp += S;
            const int u = age == 0 ? u0 : (age == 1 ? u1 : (age == 2 ? u2 : (age == 3 ? u3 : u4)));
            int n, pos0;
            chunk_origin(u, n, pos0);
            const uint32_t sA = s0 + (uint32_t)sp * p.stage_bytes, sB = sA + p.a_bytes + p.aq_bytes;
            if (a_lazy && !p.raa) {
                const int tot = nco << sha;
#pragma unroll 1
                for (int idx = tid; idx < tot; idx += W2_PR_T) {
                    const int r = idx >> sha, pu = (idx & (upa - 1)) * ga;
                    if (!unit_valid(n, pos0 + pu)) continue;
                    const uint32_t ud = sA + w2_off((uint32_t)r, (uint32_t)pu);
                    if (ga == 8) w2_xform<8>(ud, ud + p.a_bytes, a_q, coefA[r], coefA[128 + r], coefA[256 + r], dyo.relu);
                    else w2_xform<4>(ud, ud + p.a_bytes, a_q, coefA[r], coefA[128 + r], coefA[256 + r], dyo.relu);
                }
            }
            if (b_lazy && !p.rab) {
                const int per = nci << shb, tot = k * per;
#pragma unroll 1
                for (int idx = tid; idx < tot; idx += W2_PR_T) {
                    const int j = idx / per, i2 = idx - j * per;
                    const int r = i2 >> shb, pu = (i2 & (upb - 1)) * gb;
                    if (p.flat ? !unit_valid(n, pos0 + pu) : (b_src(j, pos0 + pu) < 0)) continue;
                    const uint32_t ud = sB + (uint32_t)(j * NTp) * 128u + w2_off((uint32_t)r, (uint32_t)pu);
                    if (gb == 8) w2_xform<8>(ud, ud + p.b_bytes, b_q, coefB[r], coefB[NT + r], coefB[2 * NT + r], xo.relu);
                    else w2_xform<4>(ud, ud + p.b_bytes, b_q, coefB[r], coefB[NT + r], coefB[2 * NT + r], xo.relu);
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(&hdr->full[sp]);
        };
        int stg = 0, ph = 0, cnt = 0;
        for (int u = blockIdx.z; u < p.units; u += gridDim.z, ++cnt) {
            int n, pos0;
            chunk_origin(u, n, pos0);
            if (!mbar_wait(&hdr->empty[stg], (uint32_t)(ph ^ 1))) hdr->error = 1;
            const uint32_t sA = s0 + (uint32_t)stg * p.stage_bytes, sB = sA + p.a_bytes + p.aq_bytes;
            if (p.raa) {
                // dY rows that are not 8-byte aligned (odd T*V: channel planes start on 2- or 4-byte boundaries): same register
                // path as the tap-shifted operand below, shift 0
                const bf16* pn = ap + (long long)n * dyo.pns + (long long)co0 * Lout;
                const bf16* qn = a_q ? aq + (long long)n * dyo.qns + (long long)co0 * Lout : nullptr;
                const int tot = nco << 3;
#pragma unroll 1
                for (int idx = tid; idx < tot; idx += W2_PR_T) {
                    const int r = idx >> 3, pu = (idx & 7) * 8;
                    const int pos = pos0 + pu;
                    const uint32_t dst = sA + w2_off((uint32_t)r, (uint32_t)pu);
                    if (pos >= Lout) { st_shared_v4(dst, 0u, 0u, 0u, 0u); continue; }
                    const float ca = coefA[r], cb = coefA[128 + r], cc = coefA[256 + r];
                    uint32_t w[4];
                    if (pos + 8 <= Lout) {
                        const long long e = (long long)r * Lout + pos;
                        const uint4 x = tc_ld8_unaligned(pn + e);
                        uint4 y = make_uint4(0u, 0u, 0u, 0u);
                        if (a_q) y = tc_ld8_unaligned(qn + e);
                        const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            float lo = fmaf(ca, w2_lo(xw[h]), cc), hi = fmaf(ca, w2_hi(xw[h]), cc);
                            if (a_q) { lo = fmaf(cb, w2_lo(yw[h]), lo); hi = fmaf(cb, w2_hi(yw[h]), hi); }
                            if (dyo.relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
                            w[h] = pack_bf16(lo, hi);
                        }
                    } else {
                        float v8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float val = 0.f;
                            if (pos + i < Lout) {
                                const long long e = (long long)r * Lout + pos + i;
                                val = fmaf(ca, __bfloat162float(pn[e]), cc);
                                if (a_q) val = fmaf(cb, __bfloat162float(qn[e]), val);
                                if (dyo.relu) val = fmaxf(val, 0.f);
                            }
                            v8[i] = val;
                        }
#pragma unroll
                        for (int h = 0; h < 4; ++h) w[h] = pack_bf16(v8[2 * h], v8[2 * h + 1]);
                    }
                    st_shared_v4(dst, w[0], w[1], w[2], w[3]);
                }
            } else {   // A: dY rows
                const bf16* pn = ap + (long long)n * dyo.pns + (long long)co0 * Lout + pos0;
                const bf16* qn = a_q ? aq + (long long)n * dyo.qns + (long long)co0 * Lout + pos0 : nullptr;
                const int tot = nco << sha;
#pragma unroll 1
                for (int idx = tid; idx < tot; idx += W2_PR_T) {
                    const int r = idx >> sha, pu = (idx & (upa - 1)) * ga;
                    const bool ok = unit_valid(n, pos0 + pu);
                    const bool nxt = p.flat && pos0 + pu >= Lout;           // the unit lies in sample n + 1
                    const long long e = ok ? (long long)r * Lout + pu : 0;
                    const long long ep = (ok && nxt) ? e + dyo.pns - Lout : e, eq = (ok && nxt) ? e + dyo.qns - Lout : e;
                    const uint32_t dst = sA + w2_off((uint32_t)r, (uint32_t)pu);
                    if (ga == 8) { w2_cp16(dst, pn + ep, ok ? 16u : 0u); if (a_q) w2_cp16(dst + p.a_bytes, qn + eq, ok ? 16u : 0u); }
                    else { w2_cp8(dst, pn + ep, ok ? 8u : 0u); if (a_q) w2_cp8(dst + p.a_bytes, qn + eq, ok ? 8u : 0u); }
                }
            }
            if (p.rab) {
                // stride-1 taps whose V-row shift is only 2-byte aligned (V = 25): no cp.async granularity fits.  8 consecutive
                // positions of a row are 8 consecutive source elements: aligned 16-byte words, realigned in registers,
                // transformed here (retire() skips B), stored as one 16-byte chunk.
                const bf16* pn = bp + (long long)n * xo.pns + (long long)ci0 * Lin;
                const bf16* qn = b_q ? bq + (long long)n * xo.qns + (long long)ci0 * Lin : nullptr;
                const int per = nci << 3, tot = k * per;
#pragma unroll 1
                for (int idx = tid; idx < tot; idx += W2_PR_T) {
                    const int j = idx / per, i2 = idx - j * per;
                    const int r = i2 >> 3, pu = (i2 & 7) * 8;
                    const int pos = pos0 + pu, off = pos + (j * g.d - g.p) * g.V;
                    const uint32_t dst = sB + (uint32_t)(j * NTp) * 128u + w2_off((uint32_t)r, (uint32_t)pu);
                    if (pos >= Lout || off + 8 <= 0 || off >= Lin) { st_shared_v4(dst, 0u, 0u, 0u, 0u); continue; }
                    const float ca = coefB[r], cb = coefB[NT + r], cc = coefB[2 * NT + r];
                    uint32_t w[4];
                    if (off >= 0 && off + 8 <= Lin && pos + 8 <= Lout) {
                        const long long e = (long long)r * Lin + off;
                        const uint4 x = tc_ld8_unaligned(pn + e);
                        uint4 y = make_uint4(0u, 0u, 0u, 0u);
                        if (b_q) y = tc_ld8_unaligned(qn + e);
                        const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            float lo = fmaf(ca, w2_lo(xw[h]), cc), hi = fmaf(ca, w2_hi(xw[h]), cc);
                            if (b_q) { lo = fmaf(cb, w2_lo(yw[h]), lo); hi = fmaf(cb, w2_hi(yw[h]), hi); }
                            if (xo.relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
                            w[h] = pack_bf16(lo, hi);
                        }
                    } else {
                        float v8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float val = 0.f;
                            if (off + i >= 0 && off + i < Lin && pos + i < Lout) {
                                const long long e = (long long)r * Lin + off + i;
                                val = fmaf(ca, __bfloat162float(pn[e]), cc);
                                if (b_q) val = fmaf(cb, __bfloat162float(qn[e]), val);
                                if (xo.relu) val = fmaxf(val, 0.f);
                            }
                            v8[i] = val;
                        }
#pragma unroll
                        for (int h = 0; h < 4; ++h) w[h] = pack_bf16(v8[2 * h], v8[2 * h + 1]);
                    }
                    st_shared_v4(dst, w[0], w[1], w[2], w[3]);
                }
            } else {   // B_j: X rows shifted by tap j
                const bf16* pn = bp + (long long)n * xo.pns + (long long)ci0 * Lin;
                const bf16* qn = b_q ? bq + (long long)n * xo.qns + (long long)ci0 * Lin : nullptr;
                const int per = nci << shb, tot = k * per;
#pragma unroll 1
                for (int idx = tid; idx < tot; idx += W2_PR_T) {
                    const int j = idx / per, i2 = idx - j * per;
                    const int r = i2 >> shb, pu = (i2 & (upb - 1)) * gb;
                    const bool nxt = p.flat && pos0 + pu >= Lout;
                    const int off = p.flat ? pos0 + pu : b_src(j, pos0 + pu);
                    const bool ok = p.flat ? unit_valid(n, pos0 + pu) : off >= 0;
                    const long long e = ok ? (long long)r * Lin + off : 0;
                    const long long ep = (ok && nxt) ? e + xo.pns - Lin : e, eq = (ok && nxt) ? e + xo.qns - Lin : e;
                    const uint32_t dst = sB + (uint32_t)(j * NTp) * 128u + w2_off((uint32_t)r, (uint32_t)pu);
                    if (gb == 8) { w2_cp16(dst, pn + ep, ok ? 16u : 0u); if (b_q) w2_cp16(dst + p.b_bytes, qn + eq, ok ? 16u : 0u); }
                    else { w2_cp8(dst, pn + ep, ok ? 8u : 0u); if (b_q) w2_cp8(dst + p.b_bytes, qn + eq, ok ? 8u : 0u); }
                }
            }
            w2_commit();
            u4 = u3; u3 = u2; u2 = u1; u1 = u0; u0 = u;
            if (cnt >= lag) {
                if (lag == 4) w2_wait_group<4>(); else if (lag == 3) w2_wait_group<3>();
                else if (lag == 2) w2_wait_group<2>(); else if (lag == 1) w2_wait_group<1>(); else w2_wait_group<0>();
                retire(lag, stg);
            }
            if (++stg == S) { stg = 0; ph ^= 1; }
        }
        w2_wait_group<0>();
        {
            int newest = stg - 1; if (newest < 0) newest += S;
            for (int b = min(lag, cnt); b >= 1; --b) retire(b - 1, newest);
        }
    } else if (lane == 0) {
        // =============================== MMA issuer ===============================
        const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)NTp);
        int stg = 0, ph = 0;
        for (int it = 0; it < nmine; ++it) {
            if (!mbar_wait(&hdr->full[stg], (uint32_t)ph)) hdr->error = 1;
            tc_fence_after();
            const uint32_t sA = s0 + (uint32_t)stg * p.stage_bytes, sB = sA + p.a_bytes + p.aq_bytes;
            for (int j = 0; j < k; ++j) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tmem + (uint32_t)(j * NTp), umma_desc_sw128(sA + kk * 32u),
                              umma_desc_sw128(sB + (uint32_t)(j * NTp) * 128u + kk * 32u), idesc, (it > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&hdr->empty[stg]);
            if (++stg == S) { stg = 0; ph ^= 1; }
        }
        umma_commit(&hdr->done);
    }
    // =============================== drain: TMEM -> fp32 atomics ===============================
    __syncwarp();
    if (warp < W2_MMA_W && nmine > 0) {
        if (!mbar_wait(&hdr->done, 0u)) hdr->error = 1;
        tc_fence_after();
        const int q = warp & 3, grp = warp >> 2;          // lane quarter, one of three warps sharing it
        const int row = q * 32 + lane;
        const int CK = g.Cin * k, nblk = (k * NTp + 15) >> 4;
        // All split-K CTAs of a (co, ci) tile finish at about the same time and add into the same dW entries: walked in the
        // same order, every address would take gridDim.z back-to-back atomics (measured: the drain was a quarter of the
        // kernel).  Each CTA therefore starts at its own column block and column quarter.
        const int rot_b = (int)(blockIdx.z % (unsigned)nblk), rot_c = (int)((blockIdx.z / (unsigned)nblk) & 3u);
#define W2_DRAIN_COLS(R)                                                                                          \
    _Pragma("unroll") for (int c0 = 0; c0 < 16; ++c0) {                                                          \
        const int c = (c0 + (R)) & 15;                                                                            \
        const int col = b * 16 + c, j = col / NTp, ci = col - j * NTp;                                            \
        if (j < k && ci < nci) atomicAdd(dW + (long long)(co0 + row) * CK + (ci0 + ci) * k + j, acc[c]);          \
        else if (do_bias && j == 0 && ci == NT) atomicAdd(dbias + co0 + row, acc[c]);                             \
    }
        // k = 1: the 16 columns a thread holds are 64 contiguous bytes of its dW row -> 16-byte vector reductions
        // (red.global.add.v4.f32): a quarter of the atomic instructions and L2 atomic transactions
        const bool vec4 = k == 1 && (CK & 3) == 0 && (ci0 & 3) == 0 && ((reinterpret_cast<uintptr_t>(dW) & 15) == 0);
        for (int bb = grp; bb < nblk; bb += 3) {
            int b = bb + rot_b;
            if (b >= nblk) b -= nblk;
            float acc[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 16), acc);
            if (vec4 && row < nco && !hdr->error) {
                float* rowp = dW + (long long)(co0 + row) * CK + ci0;
#define W2_DRAIN_V4(R)                                                                                            \
    _Pragma("unroll") for (int g4 = 0; g4 < 4; ++g4) {                                                            \
        const int c4 = ((g4 + (R)) & 3) * 4;                                                                       \
        const int ci = b * 16 + c4;                                                                                \
        if (ci + 3 < nci) {                                                                                        \
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(rowp + ci), "f"(acc[c4]),           \
                         "f"(acc[c4 + 1]), "f"(acc[c4 + 2]), "f"(acc[c4 + 3]) : "memory");                         \
        } else {                                                                                                   \
            _Pragma("unroll") for (int e = 0; e < 4; ++e) {                                                        \
                if (ci + e < nci) atomicAdd(rowp + ci + e, acc[c4 + e]);                                            \
                else if (do_bias && ci + e == NT) atomicAdd(dbias + co0 + row, acc[c4 + e]);                       \
            }                                                                                                      \
        }                                                                                                          \
    }
                if (rot_c == 0) { W2_DRAIN_V4(0) }
                else if (rot_c == 1) { W2_DRAIN_V4(1) }
                else if (rot_c == 2) { W2_DRAIN_V4(2) }
                else { W2_DRAIN_V4(3) }
#undef W2_DRAIN_V4
            } else if (row < nco && !hdr->error) {
                if (rot_c == 0) { W2_DRAIN_COLS(0) }
                else if (rot_c == 1) { W2_DRAIN_COLS(4) }
                else if (rot_c == 2) { W2_DRAIN_COLS(8) }
                else { W2_DRAIN_COLS(12) }
            }
        }
#undef W2_DRAIN_COLS
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W2_MMA_W) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
    if (tid == 0 && hdr->error) printf("tamgcn: conv_wgrad(tcgen05) pipeline timeout in block (%d,%d,%d)\n", blockIdx.x, blockIdx.y, blockIdx.z);
}

static bool w2_disabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TAMGCN_DISABLE_TC");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
static int w2_num_sms() { return num_sms(); }

// return 1 if handled, 0 if the caller should use another kernel, <0 on error
int conv_wgrad_tc2(const tamgcn_conv_geom* gg, const Opnd& dy, const Opnd& x, float* dW, float* dbias, cudaStream_t st) {
    if (w2_disabled() || gg->k > 16) return 0;
    W2P p = {};
    p.g = {gg->N, gg->Cin, gg->Cout, gg->T, gg->To, gg->V, gg->k, gg->stride, gg->dil, gg->pad};
    const ConvP& g = p.g;
    p.Lin = g.T * g.V; p.Lout = g.To * g.V;
    p.fast = (g.k == 1 && g.s == 1) ? 1 : 0;
    p.has_bias = dbias != nullptr;
    auto al = [](const void* ptr, long long ns, int L, int gr) {
        return ptr == nullptr || ((L % gr == 0) && (ns % gr == 0) && ((((uintptr_t)ptr) & (uintptr_t)(2 * gr - 1)) == 0));
    };
    p.gran_a = (al(dy.p, dy.pns, p.Lout, 8) && al(dy.q, dy.qns, p.Lout, 8)) ? 8 : ((al(dy.p, dy.pns, p.Lout, 4) && al(dy.q, dy.qns, p.Lout, 4)) ? 4 : 0);
    auto okb = [&](int gr) { return al(x.p, x.pns, p.Lin, gr) && al(x.q, x.qns, p.Lin, gr) && (p.fast ? (p.Lout % gr == 0) : (g.V % gr == 0)); };
    p.gran_b = okb(8) ? 8 : (okb(4) ? 4 : 0);
    static const int rab_env = [] { const char* e = getenv("TAMGCN_W2_RAB"); return e ? atoi(e) : 1; }();
    if (p.gran_b == 0 && rab_env && g.s == 1) { p.rab = 1; p.gran_b = 8; }               // register path for the B operand
    if (p.gran_a == 0 && rab_env) { p.raa = 1; p.gran_a = 8; }                             // ... and for the A operand
    if (p.gran_a == 0 || p.gran_b == 0) return 0;
    // ci tile: k column groups of NTp (>= NT + 1 for the ones row) must fit 512 TMEM columns
    int NT = (g.Cin + 15) & ~15;
    while (NT > 16 && (g.k * ((NT + 1 + 15) & ~15) > 512 || ((NT + 1 + 15) & ~15) > 256)) NT -= 16;
    if (g.k * ((NT + 1 + 15) & ~15) > 512) return 0;
    p.NT = NT;
    p.NTp = (NT + 1 + 15) & ~15;
    p.tmem_cols = (int)tmem_cols_pow2((uint32_t)(g.k * p.NTp));
    p.nchunk = (p.Lout + 63) / 64;
    static const int flat_env = [] { const char* e = getenv("TAMGCN_W2_FLAT"); return e ? atoi(e) : 1; }();
    p.flat = (flat_env && p.fast && !p.rab && !p.raa && p.Lout >= 64 && p.Lout % 64 != 0) ? 1 : 0;
    const long long units = p.flat ? ((long long)g.N * p.Lout + 63) / 64 : (long long)g.N * p.nchunk;
    if (units > 0x7fffffffLL) return 0;
    p.units = (int)units;
    p.a_bytes = 16384u;
    p.aq_bytes = dy.q ? 16384u : 0u;
    p.b_bytes = (uint32_t)(g.k * p.NTp) * 128u;
    p.b_bytes = (p.b_bytes + 1023u) & ~1023u;
    p.bq_bytes = x.q ? p.b_bytes : 0u;
    p.stage_bytes = p.a_bytes + p.aq_bytes + p.b_bytes + p.bq_bytes;
    const uint32_t szH = (sizeof(W2Hdr) + 15) & ~15u, szC = (uint32_t)(3 * (128 + NT) * 4 + 15) & ~15u;
    // TAMGCN_W2_OCC=2 (experiment, off by default): two CTAs per SM with half the shared memory each.  Measured on the
    // N-UCLA step: no gain (11.78 vs 11.52 ms) — the extra split-K atomics and prologues cost what the overlap wins.
    static const int occ_env = [] { const char* e = getenv("TAMGCN_W2_OCC"); return e ? atoi(e) : 1; }();
    uint32_t budget = 227u * 1024u - 1024u;
    int occ = 1;
    if (occ_env >= 2 && p.tmem_cols <= 256 && szH + szC + 3 * p.stage_bytes + 1024u <= 113u * 1024u) {
        occ = 2;
        budget = 113u * 1024u - 1024u;
    }
    if (szH + szC + 2 * p.stage_bytes > budget) return 0;
    p.S = (int)((budget - szH - szC) / p.stage_bytes);
    if (p.S > W2_SMAX) p.S = W2_SMAX;
    // chunks of cp.async in flight beyond the one being issued: at training batch sizes a CTA streams only a handful of
    // chunks and each costs a full L2 round trip, so the pipeline runs as deep as the stages allow
    p.lag = p.S >= 6 ? 4 : (p.S == 5 ? 3 : (p.S == 4 ? 2 : (p.S == 3 ? 1 : 0)));
    { static const int lmax = [] { const char* e = getenv("TAMGCN_W2_LAG"); return e ? atoi(e) : 4; }(); if (p.lag > lmax) p.lag = lmax; }
    p.off_hdr = (uint32_t)p.S * p.stage_bytes;
    p.off_coef = p.off_hdr + szH;
    const size_t sm = (size_t)p.off_coef + szC + 1024;
    const int gx = (g.Cout + 127) / 128, gy = (g.Cin + NT - 1) / NT;
    // TAMGCN_W2_ZMUL (experiment): more, shorter CTAs than SM slots, so that SMs return to the (higher-priority) data-gradient
    // chain more often; costs split-K atomics and prologues
    static const int zmul = [] { const char* e = getenv("TAMGCN_W2_ZMUL"); const int v = e ? atoi(e) : 1; return v < 1 ? 1 : v; }();
    long long Z = (long long)zmul * occ * wgrad_sms() / (gx * gy);
    if (Z < 1) Z = 1;
    if (Z > units) Z = units;
    if (Z > 65535) Z = 65535;
    static SmemLimit lim;
    ensure_smem(conv_wg2_kernel, lim, sm);
    dim3 grid(gx, gy, (unsigned)Z);
    conv_wg2_kernel<<<grid, W2_THREADS, sm, st>>>(p, dy, x, dW, dbias);
    count_launch();
    const int rc = check_launch("conv_wgrad(tcgen05)");
    return rc < 0 ? rc : 1;
}

}  // namespace tamgcn
