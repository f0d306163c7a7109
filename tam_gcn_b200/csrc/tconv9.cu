// tconv9.cu — (k x 1) temporal convolution, k <= 9, stride 1, for V = 25 joints on tcgen05: forward and data gradient
// of the ST-GCN temporal convolution (reference models/stgcn.py:76-82, nn.Conv2d(C, C, (9, 1), padding (4, 0))).
//
//   D[oc, (t, v)] = sum_{tap} sum_{ic} Wp[tap][oc][ic] * X[ic][(t + tap - pad, v)]
//
// The k taps read the SAME input rows shifted in time.  conv_tc2.cu flattens the taps into the GEMM K dimension and
// fetches the activation once per tap, element by element because a 25-joint row is 50 bytes; here an input tile is
// fetched ONCE and the taps are descriptor offsets:
//   B operand (activations): a stage holds 16 input channels x (16 + k - 1) time steps; every (channel, time step) row
//       is padded from 25 to 32 joints = 64 bytes, so a time step of 8 channels is one 512-byte SWIZZLE_64B atom of an
//       MN-major operand ([channel half][time step][8 rows x 64 B]).  Tap j of a 16-step output tile is the same tile
//       with the start address advanced by j atoms.  Rows are fetched as the aligned 16-byte words that hold them,
//       realigned in registers (tc_realign16), run through the lazy operand f(a P + b Q + c) and stored with four
//       16-byte shared-memory stores.  Time steps outside [0, T) are zero rows (the convolution's padding).
//   A operand (weights): MN-major SWIZZLE_128B blocks of 128 output channels x 16 input channels per tap, pre-packed
//       (tconv9_pack.cuh); the k blocks of a stage arrive with one bulk async copy.
//   D: TMEM, 128 lanes (output channels) x 512 columns (16 time steps x 32 padded joints); 2 k MMAs (N = 256, K = 16)
//       per stage.  The 7 padding columns of a time step are computed and dropped.
//   epilogue: 8 warps, thread = channel: bias / ReLU mask, BatchNorm sums of the values as stored, bf16; two time steps
//       (100 contiguous bytes per channel) are staged per warp and leave as 16-byte stores.
// Persistent warp-specialised CTA, one per SM: warps 0-7 epilogue (one accumulator: the epilogue of a tile and the MMAs
// of the next never overlap, so lane 0 of warp 0 is also the MMA issuer), warps 8-19 loaders (a loader thread
// owns the same (channel, time step) slot of every stage and requests the row of the next stage before it transforms and
// stores the current one, so a stage's global-memory latency hides behind the previous stage).
#include "tc_common.cuh"
#include "tconv9_pack.cuh"
#include <cstdio>
#include <cstdlib>

namespace tamgcn {

#define T9_V 25
#define T9_TT 16
#define T9_ROWS (T9_TT + T9_MAXK - 1)            // 24 time steps per stage
#define T9_B_BYTES (2 * T9_ROWS * 512)           // [channel half][time step][8 x 64 B]
#define T9_SMAX 4
#define T9_EPI_W 8
#define T9_LD_W 12                             // 384 loader threads: one (channel, time step) row of a stage each
#define T9_LD_W0 T9_EPI_W
#define T9_THREADS ((T9_EPI_W + T9_LD_W) * 32)    // 640: 20 warps leave 96 registers per thread
#define T9_STG_ROW 144                           // staging pitch: 2 time steps x 50 B + up to 14 B of misalignment; 144 = 36
                                                 // words keeps a warp's 16-byte row reads conflict-free
#define T9_STG_BYTES (T9_EPI_W * 32 * T9_STG_ROW)

struct T9P {
    int N, IC, OC, T, k, pad;     // T = time steps of the OUTPUT tensor
    int Tin;                      // time steps of the input tensor
    int var;                      // 0: stride 1;  1: forward, stride 2;  2: data gradient of a stride-2 convolution
    int tt;                       // output time steps per tile (16, variant 1: 8)
    int n_mt, n_kc, n_tb, n_tiles;
    int S, mode, dbg;
    long long ons;
    uint32_t a_bytes, stage_bytes, off_hdr, off_stg;
};
struct T9Epi {
    const float* bias;
    double* s1;
    double* s2;
    int stat_c0;
    const bf16* maskp;
    long long maskns;
    const float* maska;
    const float* maskc;
    int has_mask;
};
struct T9Hdr {
    uint64_t full[T9_SMAX], empty[T9_SMAX], tfull, tempty;
    uint32_t tmem_base;
    volatile uint32_t error;
};

template <typename H>
__device__ __forceinline__ bool t9_wait(H* hdr, uint64_t* bar, uint32_t parity) {
    if (hdr->error) return false;
    if (!mbar_wait(bar, parity)) { hdr->error = 1; return false; }
    return true;
}
// MN-major SWIZZLE_128B (weights): LBO = distance between 64-channel halves, SBO = 1024 (8 K rows x 128 B)
__device__ __forceinline__ uint64_t t9_desc_a(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_64B (activations): LBO = distance between time steps (512 B atoms), SBO = distance between the two
// 8-channel halves
__device__ __forceinline__ uint64_t t9_desc_b(uint32_t saddr, uint32_t lbo = 512u) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)((T9_ROWS * 512) >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}

// One row of 25 bf16 at a 2-byte aligned address: the four 16-byte granules that hold it (50 + misalignment <= 64 bytes,
// so exactly four, each containing bytes of the row), realigned to three chunks of 8 elements + the last element.
__device__ __forceinline__ void t9_load_row(const bf16* rowp, uint4 (&c)[3], uint32_t& e24) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(rowp);
    const uint32_t sft = (uint32_t)(a & 15);
    const uint4* q = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
    const uint4 w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3);
    if (sft == 0) {
        c[0] = w0; c[1] = w1; c[2] = w2;
        e24 = w3.x & 0xffffu;
    } else {
        c[0] = tc_realign16(w0, w1, sft);
        c[1] = tc_realign16(w1, w2, sft);
        c[2] = tc_realign16(w2, w3, sft);
        const uint32_t ws = sft >> 2;
        const uint32_t w = ws == 0 ? w3.x : (ws == 1 ? w3.y : (ws == 2 ? w3.z : w3.w));
        e24 = (sft & 2u) ? (w >> 16) : (w & 0xffffu);
    }
}
// the same in two halves: request the granules early, split when the values are needed
__device__ __forceinline__ uint32_t t9_row_request(const bf16* rowp, uint4 (&w)[4]) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(rowp);
    const uint4* q = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
    w[0] = __ldg(q); w[1] = __ldg(q + 1); w[2] = __ldg(q + 2); w[3] = __ldg(q + 3);
    return (uint32_t)(a & 15);
}
__device__ __forceinline__ void t9_row_split(const uint4 (&w)[4], uint32_t sft, uint4 (&c)[3], uint32_t& e24) {
    if (sft == 0) {
        c[0] = w[0]; c[1] = w[1]; c[2] = w[2];
        e24 = w[3].x & 0xffffu;
    } else {
        c[0] = tc_realign16(w[0], w[1], sft);
        c[1] = tc_realign16(w[1], w[2], sft);
        c[2] = tc_realign16(w[2], w[3], sft);
        const uint32_t ws = sft >> 2;
        const uint32_t ww = ws == 0 ? w[3].x : (ws == 1 ? w[3].y : (ws == 2 ? w[3].z : w[3].w));
        e24 = (sft & 2u) ? (ww >> 16) : (ww & 0xffffu);
    }
}
__device__ __forceinline__ float t9_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float t9_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t t9_xf2(uint32_t p, uint32_t q, const OpCoef& cf, bool has_q, bool relu) {
    float lo = fmaf(cf.a, t9_lo(p), cf.c), hi = fmaf(cf.a, t9_hi(p), cf.c);
    if (has_q) { lo = fmaf(cf.b, t9_lo(q), lo); hi = fmaf(cf.b, t9_hi(q), hi); }
    if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
    return pack_bf16(lo, hi);
}
__device__ __forceinline__ uint4 t9_xf8(const uint4& p, const uint4& q, const OpCoef& cf, bool has_q, bool relu) {
    return make_uint4(t9_xf2(p.x, q.x, cf, has_q, relu), t9_xf2(p.y, q.y, cf, has_q, relu), t9_xf2(p.z, q.z, cf, has_q, relu),
                      t9_xf2(p.w, q.w, cf, has_q, relu));
}

// 25 bf16 given as 13 words (pairs, the last word's high half unused) -> shared memory at element offset E (any parity)
__device__ __forceinline__ void t9_stage_row(unsigned char* row, int E, const uint32_t (&W)[13]) {
    if ((E & 1) == 0) {
        uint32_t* d = reinterpret_cast<uint32_t*>(row) + (E >> 1);
#pragma unroll
        for (int i = 0; i < 12; ++i) d[i] = W[i];
        reinterpret_cast<unsigned short*>(row)[E + 24] = (unsigned short)(W[12] & 0xffffu);
    } else {
        reinterpret_cast<unsigned short*>(row)[E] = (unsigned short)(W[0] & 0xffffu);
        uint32_t* d = reinterpret_cast<uint32_t*>(row) + ((E + 1) >> 1);
#pragma unroll
        for (int i = 0; i < 12; ++i) d[i] = __funnelshift_r(W[i], W[i + 1], 16);
    }
}

template <int MODE, bool HASQ>
__global__ void __launch_bounds__(T9_THREADS, 1)
tconv9_kernel(T9P p, Opnd x, const unsigned char* __restrict__ wpack, bf16* __restrict__ out, T9Epi ep) {
    extern __shared__ unsigned char t9_smem[];
    const uint32_t raw = smem_u32(t9_smem);
    const uint32_t s0 = (raw + 1023u) & ~1023u;
    unsigned char* sbase = t9_smem + (s0 - raw);
    T9Hdr* hdr = reinterpret_cast<T9Hdr*>(sbase + p.off_hdr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int V = T9_V;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.S; ++s) { mbar_init(&hdr->full[s], T9_LD_W); mbar_init(&hdr->empty[s], 1); }
        mbar_init(&hdr->tfull, 1);
        mbar_init(&hdr->tempty, T9_EPI_W);
        hdr->error = 0;
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&hdr->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = hdr->tmem_base;
    const int tile0 = blockIdx.x, tstep = gridDim.x;

    if (warp < T9_EPI_W) {
        // =============================== epilogue ===============================
        const int q = warp & 3, hh = warp >> 2;                  // TMEM lane quarter, half of the 16 time steps
        unsigned char* stg = sbase + p.off_stg + (size_t)warp * 32 * T9_STG_ROW;
        unsigned char* myrow = stg + (size_t)lane * T9_STG_ROW;
        float st1[2] = {0.f, 0.f}, st2[2] = {0.f, 0.f};
        const uint32_t idesc = umma_idesc_bf16(128, 256) | (1u << 15) | (1u << 16);     // A and B MN-major
        int it = 0, mstg = 0, mph = 0;
        long long tw_full = 0, tw_tempty = 0, t_epi = 0, t_mma = 0;
        const long long t_begin = clock64();
        for (int tile = tile0; tile < p.n_tiles; tile += tstep, ++it) {
            const int mt = tile % p.n_mt, rest = tile / p.n_mt, tb = rest % p.n_tb, n = rest / p.n_tb;
            const int t0 = tb * p.tt;
            if (warp == 0) {
                // ---- MMA issue for this tile (one thread) ----
                const long long tq0 = clock64();
                if (lane == 0 && t9_wait(hdr, &hdr->tempty, (uint32_t)((it & 1) ^ 1))) {
                    tw_tempty += clock64() - tq0;
                    const long long tm0 = clock64();
                    tc_fence_after();
                    bool ok = true;
                    for (int kc = 0; kc < p.n_kc; ++kc) {
                        const long long tf0 = clock64();
                        if (!mbar_wait_spin(&hdr->full[mstg], (uint32_t)mph)) { hdr->error = 1; ok = false; break; }
                        tw_full += clock64() - tf0;
                        tc_fence_after();
                        const uint32_t sa = s0 + (uint32_t)mstg * p.stage_bytes, sb = sa + p.a_bytes;
                        if (p.var == 0) {
                            for (int j = 0; j < p.k; ++j) {
                                const uint64_t ad = t9_desc_a(sa + (uint32_t)j * T9_BLK_BYTES);
#pragma unroll
                                for (int h = 0; h < 2; ++h)
                                    umma_bf16(tmem + (uint32_t)(h * 256), ad, t9_desc_b(sb + (uint32_t)(j + 8 * h) * 512u), idesc,
                                              (kc > 0 || j > 0) ? 1u : 0u);
                            }
                        } else if (p.var == 1) {
                            // stride 2: output step tau reads row 2 tau + j — every second atom (LBO = 2 atoms), 8 steps per tile
                            for (int j = 0; j < p.k; ++j)
                                umma_bf16(tmem, t9_desc_a(sa + (uint32_t)j * T9_BLK_BYTES), t9_desc_b(sb + (uint32_t)j * 512u, 1024u), idesc,
                                          (kc > 0 || j > 0) ? 1u : 0u);
                        } else {
                            // gradient of a stride-2 convolution: output steps 2 u + e (e = 0, 1) use the taps j = e + 2 m on the
                            // rows u + e - 2 + m of dY; phase e accumulates in columns e * 256 (8 values of u per tile)
                            for (int e = 0; e < 2; ++e)
                                for (int m = 0; e + 2 * m < p.k; ++m)
                                    umma_bf16(tmem + (uint32_t)(e * 256), t9_desc_a(sa + (uint32_t)(e + 2 * m) * T9_BLK_BYTES),
                                              t9_desc_b(sb + (uint32_t)(e + m) * 512u), idesc, (kc > 0 || m > 0) ? 1u : 0u);
                        }
                        umma_commit(&hdr->empty[mstg]);
                        if (++mstg == p.S) { mstg = 0; mph ^= 1; }
                    }
                    if (ok) umma_commit(&hdr->tfull);
                    t_mma += clock64() - tm0;
                }
                __syncwarp();
            }
            const long long tt0 = clock64();
            if (!t9_wait(hdr, &hdr->tfull, (uint32_t)(it & 1))) break;
            const long long te0 = clock64();
            if (warp == 0) t_mma += te0 - tt0;
            tc_fence_after();
            const int chw = mt * 128 + q * 32, ch = chw + lane;
            const bool chv = ch < p.OC;
            if (chw < p.OC) {
                const float bias_ = (MODE == 0 && ep.bias && chv) ? __ldg(ep.bias + ch) : 0.f;
                float ma_ = 1.f, mc_ = 0.f;
                if (ep.has_mask && chv) {
                    if (ep.maska) ma_ = __ldg(ep.maska + ch);
                    if (ep.maskc) mc_ = __ldg(ep.maskc + ch);
                }
                float s1acc = 0.f, s2acc = 0.f;
                const int half_steps = p.tt >> 1;
#pragma unroll 1
                for (int pp = 0; 2 * pp < half_steps; ++pp) {
                    const int tau0 = hh * half_steps + 2 * pp;
                    if (t0 + tau0 >= p.T) break;
                    const int nst = min(2, p.T - (t0 + tau0));
                    const long long e_base = ((long long)ch * p.T + t0 + tau0) * V;        // inside a sample
                    const int mis = (int)(reinterpret_cast<uintptr_t>(out + (long long)n * p.ons + e_base) & 15);
#pragma unroll 1
                    for (int s = 0; s < nst; ++s) {
                        float acc[32];
                        const int sg = tau0 + s;
                        const int col = (p.var == 2) ? ((sg & 1) * 256 + (sg >> 1) * 32) : sg * 32;
                        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col, acc);
                        uint32_t W[13];
                        if (MODE == 0) {
#pragma unroll
                            for (int i = 0; i < 13; ++i) {
                                const uint32_t w = pack_bf16(acc[2 * i] + bias_, (2 * i + 1 < V) ? acc[2 * i + 1] + bias_ : 0.f);
                                const float lo = t9_lo(w), hi = t9_hi(w);
                                s1acc += lo + hi;
                                s2acc = fmaf(lo, lo, fmaf(hi, hi, s2acc));
                                W[i] = w;
                            }
                        } else if (ep.has_mask) {
                            uint4 mk[3];
                            uint32_t m24 = 0u;
                            if (chv) t9_load_row(ep.maskp + (long long)n * ep.maskns + e_base + (long long)s * V, mk, m24);
                            else { mk[0] = mk[1] = mk[2] = make_uint4(0u, 0u, 0u, 0u); }
                            const uint32_t mw[13] = {mk[0].x, mk[0].y, mk[0].z, mk[0].w, mk[1].x, mk[1].y, mk[1].z, mk[1].w,
                                                     mk[2].x, mk[2].y, mk[2].z, mk[2].w, m24};
#pragma unroll
                            for (int i = 0; i < 13; ++i) {
                                const float m0 = t9_lo(mw[i]), m1 = t9_hi(mw[i]);
                                float v0 = acc[2 * i], v1 = (2 * i + 1 < V) ? acc[2 * i + 1] : 0.f;
                                if (!(fmaf(ma_, m0, mc_) > 0.f)) v0 = 0.f;
                                if (!(fmaf(ma_, m1, mc_) > 0.f)) v1 = 0.f;
                                const uint32_t w = pack_bf16(v0, v1);
                                const float lo = t9_lo(w), hi = (2 * i + 1 < V) ? t9_hi(w) : 0.f;
                                s1acc += lo + hi;
                                s2acc = fmaf(lo, m0, fmaf(hi, m1, s2acc));
                                W[i] = w;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 13; ++i) W[i] = pack_bf16(acc[2 * i], (2 * i + 1 < V) ? acc[2 * i + 1] : 0.f);
                        }
                        t9_stage_row(myrow, (mis >> 1) + s * V, W);
                    }
                    // copy-out (thread-private row): whole 16-byte chunks as vector stores, the ragged head / tail by element
                    if (chv) {
                        const int nbytes = nst * V * 2;
                        unsigned char* dst = reinterpret_cast<unsigned char*>(out + (long long)n * p.ons + e_base) - mis;
#pragma unroll
                        for (int c8 = 0; c8 < 8; ++c8) {
                            const int o = c8 * 16;
                            if (o >= mis && o + 16 <= mis + nbytes) {
                                *reinterpret_cast<uint4*>(dst + o) = *reinterpret_cast<const uint4*>(myrow + o);
                            } else if (o + 16 > mis && o < mis + nbytes) {
#pragma unroll
                                for (int b2 = 0; b2 < 16; b2 += 2)
                                    if (o + b2 >= mis && o + b2 < mis + nbytes)
                                        *reinterpret_cast<unsigned short*>(dst + o + b2) = *reinterpret_cast<const unsigned short*>(myrow + o + b2);
                            }
                        }
                    }
                }
                if (chv) { st1[mt & 1] += s1acc; st2[mt & 1] += s2acc; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr->tempty);
            t_epi += clock64() - te0;
        }
        if ((p.dbg & 8) && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 5))
            printf("tconv9 block 0 warp %d: tiles %d total %lld | wait tempty %lld  issue+wait full %lld (blocked on full %lld)  epilogue %lld\n",
                   warp, it, clock64() - t_begin, tw_tempty, t_mma, tw_full, t_epi);
        if (ep.s1 && !hdr->error) {
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int ch = m * 128 + q * 32 + lane;
                if (m < p.n_mt && ch < p.OC && ch >= ep.stat_c0 && (st1[m] != 0.f || st2[m] != 0.f)) {
                    atomicAdd(ep.s1 + (ch - ep.stat_c0), (double)st1[m]);
                    atomicAdd(ep.s2 + (ch - ep.stat_c0), (double)st2[m]);
                }
            }
        }
    } else {
        // =============================== loaders ===============================
        const int ltid = threadIdx.x - T9_LD_W0 * 32;
        constexpr bool has_q = HASQ;
        const bool relu = x.relu != 0, lazy = x.a || x.c || has_q || relu;
        // this thread's slot: row s of the stage, channel cl (channel fastest: the 8 lanes of a store phase then write the 8
        // rows of one 512-byte atom, which the swizzle spreads over all banks)
        const int s = ltid >> 4, cl = ltid & 15;
        const uint32_t r = (uint32_t)cl & 7u, sw = (r >> 1) & 3u;
        const uint32_t slot = p.a_bytes + ((uint32_t)cl >> 3) * (T9_ROWS * 512u) + (uint32_t)s * 512u + r * 64u;
        // raw granules of a row in flight
        uint4 pw[4], qw[HASQ ? 4 : 1], npw[4], nqw[HASQ ? 4 : 1];
        uint32_t sftp = 0, sftq = 0, nsftp = 0, nsftq = 0;
        bool live = false, nlive = false;
        int ci_cur = 0, ci_nxt = 0;
        OpCoef cf = {1.f, 0.f, 0.f}, ncf = {1.f, 0.f, 0.f};
        // (tile, kc) of the row to request next
        int l_tile = tile0, l_kc = 0;
        auto request = [&]() {
            nlive = false;
            if (l_tile < p.n_tiles) {
                const int rest = l_tile / p.n_mt, tb = rest % p.n_tb, n = rest / p.n_tb;
                const int t0 = tb * p.tt;
                const int nout = min(p.tt, p.T - t0);
                int nrows, t;
                if (p.var == 0) { nrows = nout + p.k - 1; t = t0 - p.pad + s; }
                else if (p.var == 1) { nrows = 2 * (nout - 1) + p.k; t = 2 * t0 - p.pad + s; }
                else { nrows = 12; t = (t0 >> 1) - 2 + s; }
                ci_nxt = l_kc * 16 + cl;
                if (s < nrows && t >= 0 && t < p.Tin) {
                    nlive = true;
                    if (lazy) ncf = opnd_coef(x, ci_nxt);
                    const long long e = ((long long)ci_nxt * p.Tin + t) * V;
                    {
                        const uintptr_t a = reinterpret_cast<uintptr_t>((const bf16*)x.p + (long long)n * x.pns + e);
                        nsftp = (uint32_t)(a & 15);
                        const uint4* g = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
                        npw[0] = __ldg(g); npw[1] = __ldg(g + 1); npw[2] = __ldg(g + 2); npw[3] = __ldg(g + 3);
                    }
                    if (has_q) {
                        const uintptr_t a = reinterpret_cast<uintptr_t>((const bf16*)x.q + (long long)n * x.qns + e);
                        nsftq = (uint32_t)(a & 15);
                        const uint4* g = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
                        nqw[0] = __ldg(g); nqw[1] = __ldg(g + 1); nqw[2] = __ldg(g + 2); nqw[3] = __ldg(g + 3);
                    }
                }
                if (++l_kc == p.n_kc) { l_kc = 0; l_tile += tstep; }
            }
        };
        auto split = [&](const uint4* w, uint32_t sft, uint4 (&c)[3], uint32_t& e24) {
            if (sft == 0) {
                c[0] = w[0]; c[1] = w[1]; c[2] = w[2];
                e24 = w[3].x & 0xffffu;
            } else {
                c[0] = tc_realign16(w[0], w[1], sft);
                c[1] = tc_realign16(w[1], w[2], sft);
                c[2] = tc_realign16(w[2], w[3], sft);
                const uint32_t ws = sft >> 2;
                const uint32_t ww = ws == 0 ? w[3].x : (ws == 1 ? w[3].y : (ws == 2 ? w[3].z : w[3].w));
                e24 = (sft & 2u) ? (ww >> 16) : (ww & 0xffffu);
            }
        };
        int stg = 0, ph = 0;
        bool ok = ltid < 16 * T9_ROWS;
        long long tw_empty = 0;
        const long long tl_begin = clock64();
        request();
        for (int tile = tile0; tile < p.n_tiles && ok; tile += tstep) {
            const int mt = tile % p.n_mt;
            for (int kc = 0; kc < p.n_kc; ++kc) {
                // the requested row becomes the current one; the row of the next stage is requested right away
                live = nlive; ci_cur = ci_nxt; sftp = nsftp; sftq = nsftq; cf = ncf;
#pragma unroll
                for (int i = 0; i < 4; ++i) { pw[i] = npw[i]; if (HASQ) qw[i] = nqw[i]; }
                request();
                const long long tw0 = clock64();
                if (!t9_wait(hdr, &hdr->empty[stg], (uint32_t)(ph ^ 1))) { ok = false; break; }
                tw_empty += clock64() - tw0;
                const uint32_t sst = s0 + (uint32_t)stg * p.stage_bytes;
                if (ltid == 0) {
                    mbar_expect_tx(&hdr->full[stg], p.a_bytes);
                    bulk_g2s(sst, wpack + ((size_t)(mt * p.n_kc + kc) * p.k) * T9_BLK_BYTES, p.a_bytes, &hdr->full[stg]);
                }
                uint4 c[4];
                c[0] = c[1] = c[2] = c[3] = make_uint4(0u, 0u, 0u, 0u);
                if (live) {
                    uint4 pr[3], qr[3];
                    uint32_t p24, q24 = 0u;
                    split(pw, sftp, pr, p24);
                    if (lazy) {
                        if (has_q) split(qw, sftq, qr, q24);
                        else { qr[0] = qr[1] = qr[2] = make_uint4(0u, 0u, 0u, 0u); }
#pragma unroll
                        for (int i = 0; i < 3; ++i) c[i] = t9_xf8(pr[i], qr[i], cf, has_q, relu);
                        c[3].x = t9_xf2(p24, q24, cf, has_q, relu) & 0xffffu;
                    } else {
                        c[0] = pr[0]; c[1] = pr[1]; c[2] = pr[2];
                        c[3].x = p24;
                    }
                }
#pragma unroll
                for (uint32_t i = 0; i < 4; ++i) st_shared_v4(sst + slot + ((i ^ sw) << 4), c[i].x, c[i].y, c[i].z, c[i].w);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&hdr->full[stg]);
                if (++stg == p.S) { stg = 0; ph ^= 1; }
            }
        }
        if ((p.dbg & 8) && blockIdx.x == 0 && (ltid == 0 || ltid == 200))
            printf("tconv9 block 0 loader %d: total %lld | blocked on empty %lld\n", ltid, clock64() - tl_begin, tw_empty);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
    if (threadIdx.x == 0 && hdr->error) printf("tamgcn: tconv9 pipeline timeout in block %d\n", blockIdx.x);
}

// =====================================================================================================================
// weight gradient:  dW[co][ci][j] += sum_{n,t,v} dY[n,co,t,v] X[n,ci,t+j-pad,v],   db[co] += sum dY
//
// One CTA = one work item (128 output channels, <= 32 input channels, one sample, a segment of time steps); the k taps
// are k accumulators of 32 TMEM columns.  Both operands are K-major SWIZZLE_128B with the (time, padded joint) positions
// as K: a 128-byte row holds two time steps of one channel, so tap j of time step t is the X tile read at K offset
// (t + j) * 32 elements — a start-address offset of 64 bytes per time step inside the swizzled rows, as the per-MMA
// K advance of every K-major kernel here.  Per stage: 8 time steps of dY (4 K blocks x 128 rows) and the 16 time steps
// of X they touch (8 K blocks x 32 rows); 2 x 8 x k MMAs (M = 128, N = 32, K = 16).  The loaders are the forward
// kernel's (aligned granules, register realignment, lazy operand), with the row of the next slot requested before the
// current one is transformed.  Drain: TMEM -> the idle stage memory as [co][ci*k + j] rows (the memory order of dW for a
// 32-channel tile) -> coalesced red.global.add.v4.f32.
// =====================================================================================================================
#define W9_TS 8                                   // time steps of dY per stage
#define W9_A_BYTES (4 * 128 * 128)                // 4 K blocks x 128 rows x 128 B
#define W9_B_BYTES (8 * 32 * 128)                 // 8 K blocks x 32 rows x 128 B
#define W9_STAGE_BYTES (W9_A_BYTES + W9_B_BYTES)  // 96 KB
#define W9_S 2
#define W9_LD_WARPS 16                           // loader warps: 8-19 and 4-7
#define W9_LT (W9_LD_WARPS * 32)
#define W9_DRAIN_PITCH (32 * T9_MAXK * 4 + 16)    // staging row of the drain: 288 floats + 16 B (bank spread)

struct W9P {
    int N, Cin, Cout, T, k, pad;                  // T = input time steps
    int To, stride, ts;                           // output time steps; dY time steps per stage (8, stride 2: 4)
    int nbmax, nslots;                            // input channels per tile (32; k = 1: 256); X time-step slots per stage
    uint32_t a_bytes;                             // dY part of a stage (the X part follows)
    int n_cot, n_cit, n_seg, seg_stages;          // item = ((n * n_seg + seg) * n_cot + cot) * n_cit + cit
    uint32_t off_hdr;
};
struct W9Hdr {
    uint64_t full[W9_S], empty[W9_S], done;
    uint32_t tmem_base;
    volatile uint32_t error;
};

__device__ __forceinline__ void red_add_v4(float* addr, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool HASQ>
__global__ void __launch_bounds__(T9_THREADS, 1)
tconv9_wgrad_kernel(W9P p, Opnd dy, Opnd x, float* __restrict__ dW, float* __restrict__ db) {
    extern __shared__ unsigned char t9_smem[];
    const uint32_t raw = smem_u32(t9_smem);
    const uint32_t s0 = (raw + 1023u) & ~1023u;
    unsigned char* sbase = t9_smem + (s0 - raw);
    W9Hdr* hdr = reinterpret_cast<W9Hdr*>(sbase + p.off_hdr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int V = T9_V;

    int item = blockIdx.x;
    const int cit = item % p.n_cit; item /= p.n_cit;
    const int cot = item % p.n_cot; item /= p.n_cot;
    const int seg = item % p.n_seg;
    const int n = item / p.n_seg;
    const int NBM = p.nbmax;
    const int NB = min(NBM, p.Cin - cit * NBM);                   // input channels of this tile
    const int stage0 = seg * p.seg_stages;
    const int n_st = min(p.seg_stages, (p.To + p.ts - 1) / p.ts - stage0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < W9_S; ++s) { mbar_init(&hdr->full[s], W9_LD_WARPS); mbar_init(&hdr->empty[s], 1); }
        mbar_init(&hdr->done, 1);
        hdr->error = 0;
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&hdr->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = hdr->tmem_base;

    if (warp < 4) {
        if (warp == 0 && lane == 0) {
            // ---- MMA issue ----
            const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)NB);
            int stg = 0, ph = 0;
            bool ok = true;
            for (int st = 0; st < n_st && ok; ++st) {
                if (!mbar_wait_spin(&hdr->full[stg], (uint32_t)ph)) { hdr->error = 1; ok = false; break; }
                tc_fence_after();
                const uint32_t sa = s0 + (uint32_t)stg * W9_STAGE_BYTES, sb = sa + p.a_bytes;
                for (int tau = 0; tau < p.ts; ++tau) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint64_t ad = umma_desc_sw128(sa + (uint32_t)(tau >> 1) * 16384u + (uint32_t)((tau & 1) * 64 + h * 32));
                        for (int j = 0; j < p.k; ++j) {
                            const int sx = p.stride * tau + j;            // X time-step slot of the stage
                            const uint64_t bd = umma_desc_sw128(sb + (uint32_t)(sx >> 1) * (uint32_t)(NBM * 128) + (uint32_t)((sx & 1) * 64 + h * 32));
                            umma_bf16(tmem + (uint32_t)(j * NBM), ad, bd, idesc, (st > 0 || tau > 0 || h > 0) ? 1u : 0u);
                        }
                    }
                }
                umma_commit(&hdr->empty[stg]);
                if (++stg == W9_S) { stg = 0; ph ^= 1; }
            }
            if (ok) umma_commit(&hdr->done);
        }
        __syncwarp();
    } else {
        // =============================== loaders ===============================
        // 512 loader threads: warps 8-19 and the second drain group (warps 4-7), which is idle until the drain
        const int ltid = warp >= T9_EPI_W ? (int)threadIdx.x - T9_EPI_W * 32 : 384 + (int)threadIdx.x - 128;
        constexpr bool has_q = HASQ;
        const bool lazy_a = dy.a || dy.c || has_q || dy.relu, relu_a = dy.relu != 0;
        const bool lazy_b = x.a || x.c || x.relu, relu_b = x.relu != 0;
        // slots of this thread in a stage: A rows (co = ltid & 127, steps (ltid >> 7) + 4 i), B rows (idx = ltid + 512 i)
        const int a_co = ltid & 127, a_s0 = ltid >> 7;
        const int cha = cot * 128 + a_co;
        const bool a_chv = cha < p.Cout;
        const OpCoef cfa = a_chv && lazy_a ? opnd_coef(dy, cha) : OpCoef{1.f, 0.f, 0.f};
        float dbsum = 0.f;
        uint4 pw[4], qw[HASQ ? 4 : 1], npw[4], nqw[HASQ ? 4 : 1];
        uint32_t sftp = 0, sftq = 0, nsftp = 0, nsftq = 0;
        int kind = 0, nkind = 0;              // 0 skip, 1 zero row, 2 A row, 3 B row
        uint32_t dsto = 0, ndsto = 0;         // byte offset of the row (chunk 0, unswizzled) inside the stage
        uint32_t rsw = 0, nrsw = 0;           // row & 7 (swizzle key)
        OpCoef cfb = {1.f, 0.f, 0.f}, ncfb = {1.f, 0.f, 0.f};
        const int nA = (128 * p.ts + W9_LT - 1) / W9_LT;                // A-row slots per thread and stage; the B-row slots follow
        const int nQ = nA + (NBM * p.nslots + W9_LT - 1) / W9_LT;
        int l_st = 0, l_q = 0;                // (stage, slot) to request next
        auto request = [&]() {
            nkind = 0;
            if (l_st >= n_st) return;
            const int t_stage = (stage0 + l_st) * p.ts;
            if (l_q < nA) {
                const int step = a_s0 + 4 * l_q;
                if (step < p.ts) {
                    const int t = t_stage + step;
                    ndsto = (uint32_t)(step >> 1) * 16384u + (uint32_t)a_co * 128u + (uint32_t)(step & 1) * 64u;
                    nrsw = (uint32_t)a_co & 7u;
                    nkind = 1;
                    if (a_chv && t < p.To) {
                        nkind = 2;
                        const long long e = ((long long)cha * p.To + t) * V;
                        {
                            const uintptr_t a = reinterpret_cast<uintptr_t>((const bf16*)dy.p + (long long)n * dy.pns + e);
                            nsftp = (uint32_t)(a & 15);
                            const uint4* g = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
                            npw[0] = __ldg(g); npw[1] = __ldg(g + 1); npw[2] = __ldg(g + 2); npw[3] = __ldg(g + 3);
                        }
                        if (has_q) {
                            const uintptr_t a = reinterpret_cast<uintptr_t>((const bf16*)dy.q + (long long)n * dy.qns + e);
                            nsftq = (uint32_t)(a & 15);
                            const uint4* g = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
                            nqw[0] = __ldg(g); nqw[1] = __ldg(g + 1); nqw[2] = __ldg(g + 2); nqw[3] = __ldg(g + 3);
                        }
                    }
                }
            } else {
                const int idx = ltid + W9_LT * (l_q - nA);
                if (idx < NBM * p.nslots) {
                    const int cl = idx % NBM, sx = idx / NBM;
                    if (cl < NB) {
                        const int t = t_stage * p.stride - p.pad + sx, ci = cit * NBM + cl;
                        ndsto = p.a_bytes + (uint32_t)(sx >> 1) * (uint32_t)(NBM * 128) + (uint32_t)cl * 128u + (uint32_t)(sx & 1) * 64u;
                        nrsw = (uint32_t)cl & 7u;
                        nkind = 1;
                        if (t >= 0 && t < p.T) {
                            nkind = 3;
                            if (lazy_b) ncfb = opnd_coef(x, ci);
                            const uintptr_t a = reinterpret_cast<uintptr_t>((const bf16*)x.p + (long long)n * x.pns + ((long long)ci * p.T + t) * V);
                            nsftp = (uint32_t)(a & 15);
                            const uint4* g = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
                            npw[0] = __ldg(g); npw[1] = __ldg(g + 1); npw[2] = __ldg(g + 2); npw[3] = __ldg(g + 3);
                        }
                    }
                }
            }
        };
        auto advance = [&]() { if (++l_q == nQ) { l_q = 0; ++l_st; } };
        auto split = [&](const uint4* w, uint32_t sft, uint4 (&c)[3], uint32_t& e24) {
            if (sft == 0) {
                c[0] = w[0]; c[1] = w[1]; c[2] = w[2];
                e24 = w[3].x & 0xffffu;
            } else {
                c[0] = tc_realign16(w[0], w[1], sft);
                c[1] = tc_realign16(w[1], w[2], sft);
                c[2] = tc_realign16(w[2], w[3], sft);
                const uint32_t ws = sft >> 2;
                const uint32_t ww = ws == 0 ? w[3].x : (ws == 1 ? w[3].y : (ws == 2 ? w[3].z : w[3].w));
                e24 = (sft & 2u) ? (ww >> 16) : (ww & 0xffffu);
            }
        };
        int stg = 0, ph = 0;
        bool ok = true;
        request();
        advance();
        for (int st = 0; st < n_st && ok; ++st) {
            for (int qslot = 0; qslot < nQ; ++qslot) {
                kind = nkind; dsto = ndsto; rsw = nrsw; sftp = nsftp; sftq = nsftq; cfb = ncfb;
#pragma unroll
                for (int i = 0; i < 4; ++i) { pw[i] = npw[i]; if (HASQ) qw[i] = nqw[i]; }
                request();
                advance();
                if (qslot == 0 && !t9_wait(hdr, &hdr->empty[stg], (uint32_t)(ph ^ 1))) { ok = false; break; }
                if (kind == 0) continue;
                uint4 c[4];
                c[0] = c[1] = c[2] = c[3] = make_uint4(0u, 0u, 0u, 0u);
                if (kind == 2) {
                    uint4 pr[3], qr[3];
                    uint32_t p24, q24 = 0u;
                    split(pw, sftp, pr, p24);
                    if (lazy_a) {
                        if (has_q) split(qw, sftq, qr, q24);
                        else { qr[0] = qr[1] = qr[2] = make_uint4(0u, 0u, 0u, 0u); }
#pragma unroll
                        for (int i = 0; i < 3; ++i) c[i] = t9_xf8(pr[i], qr[i], cfa, has_q, relu_a);
                        c[3].x = t9_xf2(p24, q24, cfa, has_q, relu_a) & 0xffffu;
                    } else {
                        c[0] = pr[0]; c[1] = pr[1]; c[2] = pr[2];
                        c[3].x = p24;
                    }
                    if (db && cit == 0) {
#pragma unroll
                        for (int i = 0; i < 3; ++i)
                            dbsum += (t9_lo(c[i].x) + t9_hi(c[i].x)) + (t9_lo(c[i].y) + t9_hi(c[i].y)) + (t9_lo(c[i].z) + t9_hi(c[i].z)) +
                                     (t9_lo(c[i].w) + t9_hi(c[i].w));
                        dbsum += t9_lo(c[3].x);
                    }
                } else if (kind == 3) {
                    uint4 pr[3];
                    uint32_t p24;
                    split(pw, sftp, pr, p24);
                    if (lazy_b) {
                        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                        for (int i = 0; i < 3; ++i) c[i] = t9_xf8(pr[i], z, cfb, false, relu_b);
                        c[3].x = t9_xf2(p24, 0u, cfb, false, relu_b) & 0xffffu;
                    } else {
                        c[0] = pr[0]; c[1] = pr[1]; c[2] = pr[2];
                        c[3].x = p24;
                    }
                }
                const uint32_t rowa = s0 + (uint32_t)stg * W9_STAGE_BYTES + (dsto & ~127u);
                const uint32_t c0 = (dsto & 64u) >> 4;                   // first 16-byte chunk of the time step inside the row
#pragma unroll
                for (uint32_t i = 0; i < 4; ++i) st_shared_v4(rowa + (((c0 + i) ^ rsw) << 4), c[i].x, c[i].y, c[i].z, c[i].w);
            }
            if (!ok) break;
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr->full[stg]);
            if (++stg == W9_S) { stg = 0; ph ^= 1; }
        }
        if (db && cit == 0 && a_chv && dbsum != 0.f) atomicAdd(db + cha, dbsum);
    }
    if (warp < T9_EPI_W) {
        // ---- drain (warps 4-7 come here after their loader duty) ----
        if (t9_wait(hdr, &hdr->done, 0u)) {
            tc_fence_after();
            const int q = warp & 3, hh = warp >> 2;                  // lane quarter; the two warps of a quarter split the taps
            const int co = q * 32 + lane;
            unsigned char* rowp = sbase + (size_t)co * W9_DRAIN_PITCH;
            const int rowlen = NB * p.k;                             // floats of a dW row segment of this tile
            // 32-column groups of the k * NBM accumulator columns: column = j * NBM + ci
            for (int gcol = hh; gcol * 32 < p.k * NBM; gcol += 2) {
                const int j = (gcol * 32) / NBM, c0 = gcol * 32 - j * NBM;
                if (c0 >= NB) continue;
                float acc[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(gcol * 32), acc);
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (c0 + c < NB) reinterpret_cast<float*>(rowp)[(c0 + c) * p.k + j] = acc[c];
            }
            // both warps of a quarter must have written their taps before rows are read back
            asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(64) : "memory");
            const int nv4 = rowlen / 4;                              // NB * k is a multiple of 4 (NB = 16 / 32)
            for (int r = hh; r < 32; r += 2) {
                const int cor = cot * 128 + q * 32 + r;
                if (cor >= p.Cout) break;
                const unsigned char* src = sbase + (size_t)(q * 32 + r) * W9_DRAIN_PITCH;
                float* dst = dW + ((long long)cor * p.Cin + cit * NBM) * p.k;
                for (int i = lane; i < nv4; i += 32) red_add_v4(dst + 4 * i, *reinterpret_cast<const float4*>(src + 16 * i));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
    if (threadIdx.x == 0 && hdr->error) printf("tamgcn: tconv9 pipeline timeout in block %d\n", blockIdx.x);
}

static bool t9_disabled() {
    static const bool off = [] { const char* e = getenv("TAMGCN_DISABLE_T9"); return e && e[0] == '1'; }();
    return off;
}

// geometry covered by the kernel (shared by the dispatchers and tamgcn_conv_needs_pack)
bool tconv9_covers(int Cin, int Cout, int k, int stride, int dil, int pad, int V, int dgrad) {
    if (t9_disabled()) return false;
    const int OC = dgrad ? Cin : Cout, IC = dgrad ? Cout : Cin;
    const bool geom = (stride == 1 && 2 * pad == k - 1) || (stride == 2 && k == 9 && pad == 4);
    return V == T9_V && dil == 1 && geom && t9_eligible(OC, IC, k) && OC <= 256 && OC >= 64 && IC >= 64;
}

// mode 0: forward (in = x, out = y), mode 1: data gradient (in = dY, out = dX).  wpack9 = the tconv9 region of the
// packed-weight buffer of that direction.  returns 1 if launched, 0 if not covered, < 0 on error
int tconv9_launch(int mode, int N, int Cin, int Cout, int T, int To, int V, int k, int stride, int dil, int pad, const Opnd& in,
                  const void* wpack9, const float* bias, void* out, long long ons, const Opnd* mask, double* s1, double* s2,
                  int stat_c0, cudaStream_t st) {
    if (!wpack9 || !tconv9_covers(Cin, Cout, k, stride, dil, pad, V, mode)) return 0;
    T9P p = {};
    // T = input length of the convolution, To = its output length; the data gradient reads To steps and writes T
    p.N = N; p.IC = mode ? Cout : Cin; p.OC = mode ? Cin : Cout; p.k = k; p.pad = pad;
    p.T = mode ? T : To; p.Tin = mode ? To : T;
    p.var = stride == 1 ? 0 : (mode ? 2 : 1);
    p.tt = p.var == 1 ? 8 : T9_TT;
    p.n_mt = (p.OC + 127) / 128; p.n_kc = p.IC / 16; p.n_tb = (p.T + p.tt - 1) / p.tt;
    const long long tiles = (long long)N * p.n_tb * p.n_mt;
    if (tiles > 0x7fffffffLL) return 0;
    p.n_tiles = (int)tiles;
    p.mode = mode; p.ons = ons;
    { const char* e = getenv("TAMGCN_T9_DBG"); p.dbg = e ? atoi(e) : 0; }
    p.a_bytes = (uint32_t)k * T9_BLK_BYTES;
    p.stage_bytes = (p.a_bytes + T9_B_BYTES + 1023u) & ~1023u;
    const uint32_t budget = 227u * 1024u - 1024u;
    const uint32_t szH = (sizeof(T9Hdr) + 15u) & ~15u;
    int S = (int)((budget - szH - T9_STG_BYTES) / p.stage_bytes);
    if (S > T9_SMAX) S = T9_SMAX;
    if (S < 2) return 0;
    p.S = S;
    p.off_stg = (uint32_t)S * p.stage_bytes;
    p.off_hdr = p.off_stg + T9_STG_BYTES;
    const size_t sm = (size_t)p.off_hdr + szH + 1024;
    T9Epi ep = {};
    ep.bias = bias; ep.s1 = s1; ep.s2 = s2; ep.stat_c0 = stat_c0;
    if (mask) { ep.has_mask = 1; ep.maskp = (const bf16*)mask->p; ep.maskns = mask->pns; ep.maska = mask->a; ep.maskc = mask->c; }
    int grid = main_sms();
    if (grid > p.n_tiles) grid = p.n_tiles;
#define T9_LAUNCH(MM, QQ)                                                                                                   \
    do {                                                                                                                    \
        static SmemLimit lim;                                                                                               \
        ensure_smem(tconv9_kernel<MM, QQ>, lim, sm);                                                                        \
        tconv9_kernel<MM, QQ><<<grid, T9_THREADS, sm, st>>>(p, in, (const unsigned char*)wpack9, (bf16*)out, ep);           \
    } while (0)
    if (mode == 0) { if (in.q) T9_LAUNCH(0, true); else T9_LAUNCH(0, false); }
    else           { if (in.q) T9_LAUNCH(1, true); else T9_LAUNCH(1, false); }
#undef T9_LAUNCH
    count_launch();
    const int rc = check_launch(mode ? "conv_dgrad(tconv9)" : "conv_fwd(tconv9)");
    return rc < 0 ? rc : 1;
}

// returns 1 if launched, 0 if not covered, < 0 on error
int tconv9_wgrad_launch(int N, int Cin, int Cout, int T, int To, int V, int k, int stride, int dil, int pad, const Opnd& dy,
                        const Opnd& x, float* dW, float* db, cudaStream_t st) {
    if (t9_disabled()) return 0;
    static const bool off = [] { const char* e = getenv("TAMGCN_DISABLE_T9W"); return e && e[0] == '1'; }();
    if (off) return 0;
    // k = 1, stride 1 only where the vector-access kernel (conv_wg2.cu) has no aligned path: planes of T*V elements that
    // are not a multiple of 4 (ST-GCN T = 150 / 75 at V = 25)
    const bool geom = (stride == 1 && 2 * pad == k - 1 && (k >= 2 || ((long long)T * V) % 4 != 0)) ||
                      (stride == 2 && ((k == 9 && pad == 4) || (k == 1 && pad == 0)));
    if (!(V == T9_V && dil == 1 && geom && k <= T9_MAXK && Cin % 16 == 0 && Cin >= 32 && Cout >= 32)) return 0;
    if ((Cin * k) % 4) return 0;                           // vector reductions of the drain
    if (x.q) return 0;                                     // the X operand is a one-tensor lazy operand
    if ((reinterpret_cast<uintptr_t>(dW) & 15) != 0) return 0;
    W9P p = {};
    p.N = N; p.Cin = Cin; p.Cout = Cout; p.T = T; p.k = k; p.pad = pad;
    p.To = To; p.stride = stride; p.ts = (stride == 1 && k > 1) ? W9_TS : 4;
    p.nbmax = (k == 1 && stride == 1) ? 256 : 32;                    // one tap leaves room for 256 accumulator columns
    p.nslots = (stride * (p.ts - 1) + k + 1) & ~1;
    p.a_bytes = (uint32_t)(p.ts / 2) * 16384u;
    if (p.a_bytes + (uint32_t)(p.nslots / 2) * (uint32_t)(p.nbmax * 128) > W9_STAGE_BYTES) return 0;
    p.n_cot = (Cout + 127) / 128; p.n_cit = (Cin + p.nbmax - 1) / p.nbmax;
    const int stages = (To + p.ts - 1) / p.ts;
    // segments per sample: enough items for ~2.5 waves of CTAs, at least 4 stages per item
    const long long base_items = (long long)N * p.n_cot * p.n_cit;
    int n_seg = (int)((5LL * wgrad_sms() / 2 + base_items - 1) / base_items);
    if (n_seg > (stages + 3) / 4) n_seg = (stages + 3) / 4;
    if (n_seg < 1) n_seg = 1;
    p.seg_stages = (stages + n_seg - 1) / n_seg;
    p.n_seg = (stages + p.seg_stages - 1) / p.seg_stages;
    const long long items = base_items * p.n_seg;
    if (items > 0x7fffffffLL) return 0;
    p.off_hdr = W9_S * W9_STAGE_BYTES;
    const size_t sm = (size_t)p.off_hdr + ((sizeof(W9Hdr) + 15) & ~15u) + 1024;
    if (dy.q) {
        static SmemLimit lim;
        ensure_smem(tconv9_wgrad_kernel<true>, lim, sm);
        tconv9_wgrad_kernel<true><<<(int)items, T9_THREADS, sm, st>>>(p, dy, x, dW, db);
    } else {
        static SmemLimit lim;
        ensure_smem(tconv9_wgrad_kernel<false>, lim, sm);
        tconv9_wgrad_kernel<false><<<(int)items, T9_THREADS, sm, st>>>(p, dy, x, dW, db);
    }
    count_launch();
    const int rc = check_launch("conv_wgrad(tconv9)");
    return rc < 0 ? rc : 1;
}

}  // namespace tamgcn
