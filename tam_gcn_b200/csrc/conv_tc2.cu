// conv_tc2.cu — (k x 1) convolution forward and data-gradient on tcgen05 (bf16 storage, fp32 accumulation).
//
//   D[oc, pos] = sum_{kf} Wp[oc, kf] * X[kf, pos]        kf = tap * IC + ic  (taps flattened into K)
//     MODE 0 (fwd)  : X[kf, pos=(to,v)] = x(n, ic, to*s + tap*d - p, v)                (zero padding)
//     MODE 1 (dgrad): X[kf, pos=(t,v)]  = dY(n, ic, (t + p - tap*d)/s, v)  when divisible and in range
//
// These GEMMs are far below the tensor-pipe ridge (K = 64..320, arithmetic intensity 30-160 FLOP/B): the kernel is
// organised around streaming the activation operand with as few instructions per byte as possible.
//   A operand = weights, pre-packed bf16 K-major SWIZZLE_128B tiles (one 1-D bulk async copy per stage)
//   B operand = activations, MN-major SWIZZLE_128B: a shared-memory row is one K index (channel/tap) and 64
//               consecutive positions — exactly the memory order of an (N,C,T,V) tensor, so the operand is moved
//               with 16-byte (or 8-byte) vector accesses and never transposed.  A plain operand goes global ->
//               shared with cp.async (no registers); a lazy operand (BatchNorm-apply / ReLU / BatchNorm-backward
//               affine, tamgcn_operand) is transformed 8 elements at a time on the way.
//   D         = TMEM, 128 channel lanes x NT positions, double buffered; one CTA covers <= 256 output channels so
//               the activation tile is read once.
//   epilogue  = 8 warps: TMEM -> registers (thread = channel: BatchNorm sums stay in registers for the whole
//               kernel) -> bias / addend / mask -> bf16 -> warp-private transpose through shared memory ->
//               full-sector coalesced stores.
// Persistent warp-specialised CTA, one per SM: warps 0-7 epilogue, warp 8 MMA issue, warps 9-16 producers.
#include "tc_common.cuh"
#include "tconv9_pack.cuh"
#include <cuda.h>
#include <cstdlib>

namespace tamgcn {

struct ConvP {
    int N, Cin, Cout, T, To, V, k, s, d, p;
};

// 16 warps, 128 registers per thread (the epilogue needs ~120; spills are poison here: with ~220 KB of the SM's
// 228 KB configured as shared memory there is next to no L1 left to catch them)
// Warp roles are a launch-time split: warps [0, n_epi) epilogue (n_epi = 8 or 16), warp n_epi MMA issue, the rest
// producers.  Kernels whose epilogue has no addend / mask / broadcast extras fit 80 registers and run 24 warps; the
// others run 16 warps at 128 registers.
#define C2_THREADS_BIG 768
#define C2_THREADS_SMALL 512
#define C2_SMAX 6
#define C2_MAXU 8
#define C2_TWAIT(acc, call) do { const long long t0_ = clock64(); call; acc += clock64() - t0_; } while (0)

struct C2Epi {
    const float* bias;
    double* s1;
    double* s2;
    int stat_c0;
    const bf16* addend;
    long long addns;
    const float* bcast;
    float bscale;
    const bf16* maskp;
    long long maskns;
    const float* maska;
    const float* maskc;
    int has_mask;
};

struct C2P {
    ConvP g;
    int IC, OC, Lin, Lout, KT, KTp, nchunk;
    int MT_total, n_oct, NT, NTp, nblk, tps, n_tiles;
    int gran, upr, fast, vec, S, lag, tmem_cols, n_epi, ra8;
    uint32_t x_bytes, q_bytes, stage_bytes, off_hdr, off_coef, off_stg;
    long long ons;
    int dbg;                 // TAMGCN_C2_DBG (profiling aid): 8 = print per-role blocked cycles of block 0
    int use_tma;             // plain 1x1 operand on 16-byte aligned planes: tiles fetched by cp.async.bulk.tensor (TMA)
};

// a CUtensorMap, passed by value as a __grid_constant__ kernel parameter
struct alignas(64) C2TensorMap { unsigned long long v[16]; };

// one 3-D box (64 positions x 64 channels x 1 sample) of the activation tensor -> shared memory (SWIZZLE_128B: the image
// of the MN-major operand tile, c2_xoff), completion on `bar`
__device__ __forceinline__ void c2_tma_load(uint32_t dst, const C2TensorMap* tm, int pos, int ch, int n, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(tm)), "r"(pos), "r"(ch), "r"(n), "r"(smem_u32(bar))
                 : "memory");
}

struct C2Hdr {
    uint64_t full[C2_SMAX], empty[C2_SMAX], tfull[2], tempty[2];
    uint64_t raw[C2_SMAX];       // TMA-landed, not yet transformed (lazy operand fetched by tensor loads)
    uint32_t tmem_base;
    volatile uint32_t error;
};

__device__ __forceinline__ bool c2_wait(C2Hdr* hdr, uint64_t* bar, uint32_t parity) {
    if (!mbar_wait(bar, parity)) { hdr->error = 1; return false; }
    return true;
}
__device__ __forceinline__ void c2_cp16(uint32_t dst, const void* src, uint32_t nbytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void c2_cp8(uint32_t dst, const void* src, uint32_t nbytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void c2_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void c2_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void c2_st8(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void c2_st2(uint32_t a, unsigned short x) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(x) : "memory");
}
// K-major SW128 descriptor (weights) is umma_desc_sw128; MN-major SW128 (activations): LBO = distance between
// 64-position column blocks, SBO = distance between groups of 8 K rows (1024 B)
__device__ __forceinline__ uint64_t c2_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
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
__device__ __forceinline__ float c2_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float c2_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// byte offset of position `pu` (multiple of the access granularity) of K row r inside an activation stage
__device__ __forceinline__ uint32_t c2_xoff(uint32_t r, uint32_t pu) {
    const uint32_t blk = pu >> 6, pp = pu & 63u;
    return blk * 8192u + r * 128u + (((pp >> 3) ^ (r & 7u)) << 4) + (pp & 7u) * 2u;
}

// input element offset (inside the sample plane of one channel) feeding output position pos through tap j; -1 = zero
template <int MODE>
__device__ __forceinline__ int c2_in_off(const ConvP& g, int pos, int j) {
    const int tq = pos / g.V, v = pos - tq * g.V;
    if (MODE == 0) {
        const int t = tq * g.s + j * g.d - g.p;
        return (t >= 0 && t < g.T) ? t * g.V + v : -1;
    } else {
        const int num = tq + g.p - j * g.d;
        if (num < 0) return -1;
        const int to = num / g.s;
        return (to * g.s == num && to < g.To) ? to * g.V + v : -1;
    }
}

template <int GR> struct C2Vec;
template <> struct C2Vec<8> { typedef uint4 T; };
template <> struct C2Vec<4> { typedef uint2 T; };

// transform GR (4 or 8) bf16 elements of the lazy operand held in w[] (and q[]) and store them
template <int GR>
__device__ __forceinline__ void c2_xform_store(uint32_t dst, const uint32_t* w, const uint32_t* q, bool has_q, float a, float b,
                                               float c, int relu) {
    uint32_t o[GR / 2];
#pragma unroll
    for (int e = 0; e < GR / 2; ++e) {
        float lo = fmaf(a, c2_lo(w[e]), c), hi = fmaf(a, c2_hi(w[e]), c);
        if (has_q) { lo = fmaf(b, c2_lo(q[e]), lo); hi = fmaf(b, c2_hi(q[e]), hi); }
        if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
        o[e] = pack_bf16(lo, hi);
    }
    if (GR == 8) st_shared_v4(dst, o[0], o[1], o[2], o[3]);
    else c2_st8(dst, o[0], o[1]);
}

// 8 consecutive bf16 of a row -> floats; vec = widest aligned access (8, 4 or 1 elements), nv = valid elements
__device__ __forceinline__ void c2_ld_row8(const bf16* __restrict__ ptr, float* f, int vec, int nv) {
    if (vec == 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(ptr);
        f[0] = c2_lo(u.x); f[1] = c2_hi(u.x); f[2] = c2_lo(u.y); f[3] = c2_hi(u.y);
        f[4] = c2_lo(u.z); f[5] = c2_hi(u.z); f[6] = c2_lo(u.w); f[7] = c2_hi(u.w);
    } else if (vec == 4) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (4 * h < nv) {
                const uint2 u = *reinterpret_cast<const uint2*>(ptr + 4 * h);
                f[4 * h] = c2_lo(u.x); f[4 * h + 1] = c2_hi(u.x); f[4 * h + 2] = c2_lo(u.y); f[4 * h + 3] = c2_hi(u.y);
            }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (e < nv) f[e] = __bfloat162float(ptr[e]);
    }
}

// in-place transform of one landed unit (GR elements at dst; second tensor at dstq)
template <int GR>
__device__ __forceinline__ void c2_xform_inplace(uint32_t dst, uint32_t dstq, bool has_q, float a, float b, float c, int relu) {
    uint32_t w[4], q[4] = {0u, 0u, 0u, 0u};
    if (GR == 8) {
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(dst));
        if (has_q) asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(dstq));
    } else {
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(dst));
        if (has_q) asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(q[0]), "=r"(q[1]) : "r"(dstq));
    }
    c2_xform_store<GR>(dst, w, q, has_q, a, b, c, relu);
}

template <int MODE, int PLAIN, int EXTRA>
__global__ void __launch_bounds__(EXTRA ? C2_THREADS_SMALL : C2_THREADS_BIG, 1)
conv_tc2_kernel(C2P p, Opnd xo, const uint8_t* __restrict__ wpack, bf16* __restrict__ out, C2Epi ep,
                const __grid_constant__ C2TensorMap tmap) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    C2Hdr* hdr = (C2Hdr*)(smem + p.off_hdr);
    float* coef = (float*)(smem + p.off_coef);           // [3][IC]
    const ConvP g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C2_THREADS = (int)blockDim.x, C2_MMA_W = p.n_epi, C2_PR_W0 = p.n_epi, C2_EPI_T = p.n_epi * 32;
    const int C2_PR_T0 = (p.n_epi + 1) * 32, C2_PR_T = C2_THREADS - C2_PR_T0;
    const int S = p.S, NT = p.NT, IC = p.IC, OC = p.OC, Lin = p.Lin, Lout = p.Lout;
    const int mtmax = min(2, p.MT_total);
    const int oct = blockIdx.x % p.n_oct, mt0 = oct * 2, mt_cnt = min(2, p.MT_total - mt0);
    const int tile0 = blockIdx.x / p.n_oct, tile_step = gridDim.x / p.n_oct;
    const uint32_t w_bytes = (uint32_t)mt_cnt * 16384u;
    const int OCpad = p.MT_total * 128;
    const int NK = p.nchunk;

    if (warp == C2_MMA_W) tmem_alloc(&hdr->tmem_base, (uint32_t)p.tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < C2_SMAX; ++i) { mbar_init(&hdr->full[i], C2_PR_T); mbar_init(&hdr->empty[i], 1); mbar_init(&hdr->raw[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&hdr->tfull[i], 1); mbar_init(&hdr->tempty[i], C2_EPI_T); }
        hdr->error = 0;
        fence_mbar_init();
    }
    for (int i = tid; i < IC; i += C2_THREADS) {
        const OpCoef cf = opnd_coef(xo, i);
        coef[i] = cf.a; coef[IC + i] = cf.b; coef[2 * IC + i] = cf.c;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = hdr->tmem_base;
    const uint32_t s0 = smem_u32(smem);
    long long tw0 = 0, tw1 = 0;
    const long long t_begin = clock64();

    if (warp < C2_PR_W0) {
        // =============================== epilogue (8 warps) ===============================
        // warp -> TMEM lane quarter q (hardware rule: warp % 4); the (channel tile, 32-column block) items of a quarter
        // are dealt round-robin to its n_epi / 4 warps.  A thread owns one channel per channel tile: BatchNorm sums in registers.
        const int q = warp & 3, grp = warp >> 2;
        float st1[2] = {0.f, 0.f}, st2[2] = {0.f, 0.f}, biasr[2] = {0.f, 0.f}, mar[2] = {1.f, 1.f}, mcr[2] = {0.f, 0.f};
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int oc_ = (mt0 + mt) * 128 + q * 32 + lane;
            if (mt < mt_cnt && oc_ < OC) {
                if (MODE == 0 && ep.bias) biasr[mt] = __ldg(ep.bias + oc_);
                if (EXTRA && ep.has_mask) {
                    if (ep.maska) mar[mt] = __ldg(ep.maska + oc_);
                    if (ep.maskc) mcr[mt] = __ldg(ep.maskc + oc_);
                }
            }
        }
        const uint32_t stg = s0 + p.off_stg + (uint32_t)warp * 2560u;   // warp-private 32 rows x 80 B
        const int nb = (NT + 31) >> 5;
        const int vec = p.vec;
        int it = 0;
        for (int tile = tile0; tile < p.n_tiles; tile += tile_step, ++it) {
            const int n = tile / p.tps, pos0 = (tile - n * p.tps) * NT;
            const int buf = it & 1;
            C2_TWAIT(tw0, c2_wait(hdr, &hdr->tfull[buf], (uint32_t)((it >> 1) & 1)));
            tc_fence_after();
            for (int item = grp; item < mt_cnt * nb; item += (C2_EPI_T >> 7)) {
                const int my_mt = (item >= nb) ? 1 : 0, pb = item - my_mt * nb;
                const int ocw = (mt0 + my_mt) * 128 + q * 32;            // first channel of this warp in this channel tile
                const int oc = ocw + lane;
                const bool ocv = oc < OC;
                const uint32_t tcol = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * mtmax + my_mt) * p.NTp + pb * 32);
                const int nvalid = min(min(32, NT - pb * 32), Lout - (pos0 + pb * 32));
                if (nvalid <= 0) continue;
                const long long boff = (long long)oc * Lout + pos0 + pb * 32;    // inside a sample
                const bf16* padd = (EXTRA && ep.addend && ocv) ? ep.addend + (long long)n * ep.addns + boff : nullptr;
                const bf16* pmask = (EXTRA && ep.has_mask && ocv) ? ep.maskp + (long long)n * ep.maskns + boff : nullptr;
                const float* pbc = (EXTRA && ep.bcast && ocv) ? ep.bcast + ((long long)n * OC + oc) * g.V : nullptr;
                int vv = (pos0 + pb * 32) % g.V;
                float s1acc = 0.f, s2acc = 0.f;
                const float bias_ = my_mt ? biasr[1] : biasr[0], ma_ = my_mt ? mar[1] : mar[0], mc_ = my_mt ? mcr[1] : mcr[0];
                const uint32_t myrow = stg + (uint32_t)lane * 80u;
                // ---- phase 1 (thread = channel), 8 columns at a time: bias / addend / bcast / mask, BatchNorm sums of the
                //      values as stored; the bf16 row goes to the warp-private staging tile
#pragma unroll 1
                for (int i0 = 0; i0 < 32; i0 += 8) {
                    if (i0 >= nvalid) break;
                    float acc[8];
                    tmem_ld8_nowait(tcol + i0, acc);
                    tmem_wait_ld();
                    float ad[8], mk[8];
                    if (EXTRA) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) { ad[e] = 0.f; mk[e] = 0.f; }
                        if (padd) c2_ld_row8(padd + i0, ad, vec, nvalid - i0);
                        if (pmask) c2_ld_row8(pmask + i0, mk, vec, nvalid - i0);
                    }
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float val = acc[e];
                        const bool in = ocv && (i0 + e < nvalid);
                        if (MODE == 0) {
                            val = rnd<bf16>(val + bias_);
                            if (in) { s1acc += val; s2acc = fmaf(val, val, s2acc); }
                        } else if (EXTRA) {
                            val += ad[e];
                            if (pbc) {
                                if (in) val = fmaf(__ldg(pbc + vv), ep.bscale, val);
                                if (++vv == g.V) vv = 0;
                            }
                            if (pmask && in) {
                                if (!(fmaf(ma_, mk[e], mc_) > 0.f)) val = 0.f;
                                val = rnd<bf16>(val);
                                s1acc += val;
                                s2acc = fmaf(val, mk[e], s2acc);
                            }
                        }
                        o[e] = val;
                    }
                    st_shared_v4(myrow + (uint32_t)i0 * 2u, pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
                }
                if (my_mt) { st1[1] += s1acc; st2[1] += s2acc; } else { st1[0] += s1acc; st2[0] += s2acc; }
                __syncwarp();
                // ---- phase 2: the warp writes whole rows (full 32-byte sectors)
                const long long wbase = (long long)n * p.ons + (long long)ocw * Lout + pos0 + pb * 32;
                if (vec == 8) {
#pragma unroll
                    for (int r8 = 0; r8 < 4; ++r8) {
                        const int row = r8 * 8 + (lane >> 2), piece = lane & 3;
                        uint4 v;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(stg + (uint32_t)row * 80u + (uint32_t)piece * 16u));
                        if (ocw + row < OC && piece * 8 < nvalid)
                            *reinterpret_cast<uint4*>(out + wbase + (long long)row * Lout + piece * 8) = v;
                    }
                } else if (vec == 4) {
#pragma unroll 1
                    for (int r4 = 0; r4 < 8; ++r4) {
                        const int row = r4 * 4 + (lane >> 3), piece = lane & 7;
                        uint2 v;
                        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(stg + (uint32_t)row * 80u + (uint32_t)piece * 8u));
                        if (ocw + row < OC && piece * 4 < nvalid)
                            *reinterpret_cast<uint2*>(out + wbase + (long long)row * Lout + piece * 4) = v;
                    }
                } else {
#pragma unroll 1
                    for (int r = 0; r < 32; ++r) {
                        unsigned short v;
                        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(stg + (uint32_t)r * 80u + (uint32_t)lane * 2u));
                        if (ocw + r < OC && lane < nvalid)
                            reinterpret_cast<unsigned short*>(out)[wbase + (long long)r * Lout + lane] = v;
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(&hdr->tempty[buf]);
        }
        if (ep.s1 && !hdr->error) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int oc_ = (mt0 + mt) * 128 + q * 32 + lane;
                if (mt < mt_cnt && oc_ < OC && oc_ >= ep.stat_c0 && (st1[mt] != 0.f || st2[mt] != 0.f)) {
                    atomicAdd(ep.s1 + (oc_ - ep.stat_c0), (double)st1[mt]);
                    atomicAdd(ep.s2 + (oc_ - ep.stat_c0), (double)st2[mt]);
                }
            }
        }
    } else if (warp == C2_MMA_W) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)NT) | (1u << 16);     // B operand MN-major
            int stg = 0, ph = 0, it = 0;
            for (int tile = tile0; tile < p.n_tiles; tile += tile_step, ++it) {
                const int buf = it & 1;
                C2_TWAIT(tw0, c2_wait(hdr, &hdr->tempty[buf], (uint32_t)(((it >> 1) & 1) ^ 1)));
                tc_fence_after();
                const uint32_t td = tmem + (uint32_t)(buf * mtmax * p.NTp);
                for (int kc = 0; kc < NK; ++kc) {
                    C2_TWAIT(tw1, c2_wait(hdr, &hdr->full[stg], (uint32_t)ph));
                    tc_fence_after();
                    const uint32_t sx = s0 + (uint32_t)stg * p.stage_bytes, sw = sx + p.x_bytes + p.q_bytes;
                    const int ksteps = min(64, p.KTp - kc * 64) >> 4;
                    for (int mt = 0; mt < mt_cnt; ++mt) {
                        for (int kk = 0; kk < ksteps; ++kk)
                            umma_bf16(td + (uint32_t)(mt * p.NTp), umma_desc_sw128(sw + mt * 16384u + kk * 32u),
                                      c2_desc_mn(sx + kk * 2048u, 8192u), idesc, (kc > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(&hdr->empty[stg]);
                    if (++stg == S) { stg = 0; ph ^= 1; }
                }
                umma_commit(&hdr->tfull[buf]);
            }
        }
    } else {
        // =============================== activation producers (7 warps) ===============================
        // gran >= 4: every unit (8 or 4 consecutive positions of one K row) is copied global -> shared with cp.async,
        // `lag` chunks are kept in flight, and a lazy operand is transformed IN PLACE by the thread that copied it once
        // its copies have landed (no registers are tied up while the loads are in flight).
        const int pt = tid - C2_PR_T0;
        const int gran = p.gran, upr = p.upr;
        const unsigned umagic = 0xFFFFFFFFu / (unsigned)upr + 1u;      // idx / upr by multiply-high (idx < 2^16)
        const bf16* xp = (const bf16*)xo.p;
        const bf16* xq = (const bf16*)xo.q;
        const bool has_q = xq != nullptr;
        const int lag = p.lag;
        int stg = 0, ph = 0, cnt = 0;
        // (K base, first position) of the chunks in flight, newest first
        int kb0 = 0, kb1 = 0, kb2 = 0, ps0 = 0, ps1 = 0, ps2 = 0;
        // source element offset of a unit (row r, tile position pu) or -1 (zero fill); ic_out = its channel
        auto unit_src = [&](int kbase, int pos0, int r, int pu, int& ic_out) -> long long {
            const int kf = kbase + r, pos = pos0 + pu;
            if (kf >= p.KT || pos >= Lout) return -1;
            int ic = kf, off = pos;
            if (!p.fast) {
                const int j = kf / IC;
                ic = kf - j * IC;
                off = c2_in_off<MODE>(g, pos, j);
                if (off < 0) return -1;
            }
            ic_out = ic;
            return (long long)ic * Lin + off;
        };
        // retire the chunk issued `age` chunks ago: transform a lazy operand in place, publish to the MMA warp
        auto retire = [&](int age, int newest, int chunk) {
            int sp = newest - age; if (sp < 0) sp += S;
            if (!PLAIN && p.use_tma == 2) c2_wait(hdr, &hdr->raw[sp], (uint32_t)((chunk / S) & 1));     // the boxes have landed
            if (!PLAIN) {
                const int kbase = age == 0 ? kb0 : (age == 1 ? kb1 : kb2), pos0 = age == 0 ? ps0 : (age == 1 ? ps1 : ps2);
                const uint32_t sx = s0 + (uint32_t)sp * p.stage_bytes;
                const unsigned tot = (unsigned)(min(64, p.KTp - kbase) * upr);
#pragma unroll 1
                for (unsigned idx = pt; idx < tot; idx += C2_PR_T) {
                    const unsigned r = __umulhi(idx, umagic), pu = (idx - r * upr) * gran;
                    int ic = 0;
                    if (unit_src(kbase, pos0, (int)r, (int)pu, ic) < 0) continue;
                    const float a = coef[ic], b = coef[IC + ic], c = coef[2 * IC + ic];
                    const uint32_t ud = sx + c2_xoff(r, pu);
                    if (gran == 8) c2_xform_inplace<8>(ud, ud + p.x_bytes, has_q, a, b, c, xo.relu);
                    else c2_xform_inplace<4>(ud, ud + p.x_bytes, has_q, a, b, c, xo.relu);
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(&hdr->full[sp]);
        };
        for (int tile = tile0; tile < p.n_tiles; tile += tile_step) {
            const int n = tile / p.tps, pos0 = (tile - n * p.tps) * NT;
            const bf16* pn = xp + (long long)n * xo.pns;
            const bf16* qn = has_q ? xq + (long long)n * xo.qns : nullptr;
            for (int kc = 0; kc < NK; ++kc, ++cnt) {
                C2_TWAIT(tw0, c2_wait(hdr, &hdr->empty[stg], (uint32_t)(ph ^ 1)));
                const uint32_t sx = s0 + (uint32_t)stg * p.stage_bytes;
                if (pt == 0) {
                    mbar_expect_tx(&hdr->full[stg], w_bytes);
                    bulk_g2s(sx + p.x_bytes + p.q_bytes, wpack + ((size_t)kc * OCpad + (size_t)mt0 * 128) * 128, w_bytes, &hdr->full[stg]);
                }
                const int kbase = kc * 64, rows = min(64, p.KTp - kbase);
                if (PLAIN && p.use_tma) {
                    // TMA: one thread asks for the nblk boxes of the stage; the barrier counts their bytes.  Rows / positions
                    // outside the tensor arrive as zeros.
                    if (pt == 0) {
                        mbar_expect_tx(&hdr->full[stg], (uint32_t)p.nblk * 8192u);
                        for (int b = 0; b < p.nblk; ++b) c2_tma_load(sx + (uint32_t)b * 8192u, &tmap, pos0 + 64 * b, kbase, n, &hdr->full[stg]);
                    }
                    mbar_arrive(&hdr->full[stg]);
                } else if (gran >= 4) {
                    const unsigned tot = (p.use_tma == 2) ? 0u : (unsigned)(rows * upr);
                    if (p.use_tma == 2 && pt == 0) {
                        // lazy one-tensor operand: the raw boxes arrive by TMA on their own barrier; the producers
                        // transform them in place once landed (retire) and only then publish the stage
                        mbar_expect_tx(&hdr->raw[stg], (uint32_t)p.nblk * 8192u);
                        for (int b = 0; b < p.nblk; ++b) c2_tma_load(sx + (uint32_t)b * 8192u, &tmap, pos0 + 64 * b, kbase, n, &hdr->raw[stg]);
                        mbar_arrive(&hdr->raw[stg]);
                    }
#pragma unroll 1
                    for (unsigned idx = pt; idx < tot; idx += C2_PR_T) {
                        const unsigned r = __umulhi(idx, umagic), pu = (idx - r * upr) * gran;
                        int ic = 0;
                        const long long e = unit_src(kbase, pos0, (int)r, (int)pu, ic);
                        const bool ok = e >= 0;
                        const uint32_t dst = sx + c2_xoff(r, pu);
                        const long long ee = ok ? e : 0;
                        if (gran == 8) {
                            c2_cp16(dst, pn + ee, ok ? 16u : 0u);
                            if (has_q) c2_cp16(dst + p.x_bytes, qn + ee, ok ? 16u : 0u);
                        } else {
                            c2_cp8(dst, pn + ee, ok ? 8u : 0u);
                            if (has_q) c2_cp8(dst + p.x_bytes, qn + ee, ok ? 8u : 0u);
                        }
                    }
                    c2_commit();
                    kb2 = kb1; kb1 = kb0; kb0 = kbase;
                    ps2 = ps1; ps1 = ps0; ps0 = pos0;
                    if (cnt >= lag) {
                        if (lag == 2) c2_wait_group<2>(); else if (lag == 1) c2_wait_group<1>(); else c2_wait_group<0>();
                        retire(lag, stg, cnt - lag);
                    }
                } else if (p.ra8) {
                    // stride-1 taps whose V-row shift is not 8-byte aligned (V = 25: 50-byte rows): 8 consecutive positions of
                    // a K row are still 8 consecutive source elements (offset pos + shift * V), only 2-byte aligned.  Fetched
                    // as aligned 16-byte words and realigned in registers, transformed, stored as one 16-byte chunk — instead
                    // of eight 2-byte loads and stores.  Units cut by the sample boundary take the element path below.
                    const int upr8 = NT >> 3;
                    const unsigned total = (unsigned)(rows * upr8);
#pragma unroll 1
                    for (unsigned idx = pt; idx < total; idx += C2_PR_T) {
                        const int r = (int)(idx / (unsigned)upr8), pu = (int)(idx - (unsigned)r * upr8) * 8;
                        const int kf = kbase + r, pos = pos0 + pu;
                        const int j = p.fast ? 0 : kf / IC, ic = kf - j * IC;
                        const int off = pos + (MODE == 0 ? j * g.d - g.p : g.p - j * g.d) * g.V;
                        const uint32_t dst = sx + c2_xoff((uint32_t)r, (uint32_t)pu);
                        if (kf >= p.KT || pos >= Lout || off + 8 <= 0 || off >= Lin) {
                            st_shared_v4(dst, 0u, 0u, 0u, 0u);
                            continue;
                        }
                        const float ca = coef[ic], cb = coef[IC + ic], cc = coef[2 * IC + ic];
                        uint32_t w[4];
                        if (off >= 0 && off + 8 <= Lin && pos + 8 <= Lout) {
                            const long long e = (long long)ic * Lin + off;
                            const uint4 x = tc_ld8_unaligned(pn + e);
                            uint4 y = make_uint4(0u, 0u, 0u, 0u);
                            if (has_q) y = tc_ld8_unaligned(qn + e);
                            const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                float lo = fmaf(ca, __uint_as_float(xw[h] << 16), cc), hi = fmaf(ca, __uint_as_float(xw[h] & 0xffff0000u), cc);
                                if (has_q) {
                                    lo = fmaf(cb, __uint_as_float(yw[h] << 16), lo);
                                    hi = fmaf(cb, __uint_as_float(yw[h] & 0xffff0000u), hi);
                                }
                                if (xo.relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
                                w[h] = pack_bf16(lo, hi);
                            }
                        } else {
                            float v8[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float val = 0.f;
                                if (off + i >= 0 && off + i < Lin && pos + i < Lout) {
                                    const long long e = (long long)ic * Lin + off + i;
                                    val = fmaf(ca, ldf<bf16>(pn + e), cc);
                                    if (has_q) val = fmaf(cb, ldf<bf16>(qn + e), val);
                                    if (xo.relu) val = fmaxf(val, 0.f);
                                }
                                v8[i] = val;
                            }
#pragma unroll
                            for (int h = 0; h < 4; ++h) w[h] = pack_bf16(v8[2 * h], v8[2 * h + 1]);
                        }
                        st_shared_v4(dst, w[0], w[1], w[2], w[3]);
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&hdr->full[stg]);
                } else {
                    // element-granular path (odd plane sizes, strided V = 25 taps): 2-byte loads and stores
                    const int total = rows * NT;
#pragma unroll 1
                    for (int idx = pt; idx < total; idx += C2_PR_T) {
                        const int r = idx / NT, pu = idx - r * NT;
                        int ic = 0;
                        const long long e = unit_src(kbase, pos0, r, pu, ic);
                        float val = 0.f;
                        if (e >= 0) {
                            val = fmaf(coef[ic], ldf<bf16>(pn + e), coef[2 * IC + ic]);
                            if (has_q) val = fmaf(coef[IC + ic], ldf<bf16>(qn + e), val);
                            if (xo.relu) val = fmaxf(val, 0.f);
                        }
                        const __nv_bfloat16 hb = __float2bfloat16_rn(val);
                        c2_st2(sx + c2_xoff((uint32_t)r, (uint32_t)pu), *reinterpret_cast<const unsigned short*>(&hb));
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&hdr->full[stg]);
                }
                if (++stg == S) { stg = 0; ph ^= 1; }
            }
        }
        if (gran >= 4 && !(PLAIN && p.use_tma)) {
            c2_wait_group<0>();
            int newest = stg - 1; if (newest < 0) newest += S;
            for (int b = min(lag, cnt); b >= 1; --b) retire(b - 1, newest, cnt - b);
        }
    }
    if ((p.dbg & 8) && blockIdx.x == 0 && (tid == 0 || tid == 256 || tid == C2_MMA_W * 32 || tid == C2_PR_T0))
        printf("c2 role tid %d: total %lld clk, blocked %lld / %lld (tiles %d, chunks %d)\n", tid, clock64() - t_begin, tw0, tw1, (p.n_tiles - tile0 + tile_step - 1) / tile_step, NK);
    tc_fence_before();
    __syncthreads();
    if (warp == C2_MMA_W) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
    if (tid == 0 && hdr->error) printf("tamgcn: conv(tcgen05) pipeline timeout in block %d\n", blockIdx.x);
}

// ------------------------------------------------------------------------------------------------
// weight pre-pack: fp32 (Cout, Cin, k) -> bf16 [chunk of 64 kf][OCp rows][64 kf], kf = tap * IC + ic, rows 128 B with
// the eight 16-byte pieces XOR-swizzled by (row & 7): the exact shared-memory image of the K-major SW128 A operand.
//   fwd  : row = output channel, IC = Cin   : Wp(oc, tap*Cin + ic)  = W[oc, ic, tap]
//   dgrad: row = input channel,  IC = Cout  : Wp(ic, tap*Cout + oc) = W[oc, ic, tap]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_w2_body(const float* __restrict__ W, int Cout, int Cin, int k, uint4* __restrict__ wf,
                                             uint4* __restrict__ wd, int OCp_f, int OCp_d, long long t0, long long tstride) {
    const int CK = Cin * k;
    for (int which = 0; which < 2; ++which) {
        uint4* dst = which ? wd : wf;
        if (!dst) continue;
        const int OC = which ? Cin : Cout, IC = which ? Cout : Cin;
        const int OCp = which ? OCp_d : OCp_f, KT = k * IC, nch = (KT + 63) / 64;
        const long long units = (long long)nch * OCp * 8;
        for (long long u = t0; u < units; u += tstride) {
            const int piece = (int)(u & 7);
            const long long rowl = u >> 3;
            const int oc = (int)(rowl % OCp), ch = (int)(rowl / OCp);
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int kf = ch * 64 + piece * 8 + e;
                float w = 0.f;
                if (oc < OC && kf < KT) {
                    const int j = kf / IC, ic = kf - j * IC;
                    w = which ? __ldg(W + (long long)ic * CK + oc * k + j) : __ldg(W + (long long)oc * CK + ic * k + j);
                }
                f[e] = w;
            }
            uint4 o;
            o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
            dst[((long long)ch * OCp + oc) * 8 + (piece ^ (oc & 7))] = o;
        }
        // the tiles of the V-padded temporal-convolution kernel (tconv9.cu) follow in the same buffer
        t9_pack_region(W, Cout, Cin, k, which, dst + units, t0, tstride);
    }
}

__global__ void __launch_bounds__(256)
pack_w2_kernel(const float* __restrict__ W, int Cout, int Cin, int k, uint4* __restrict__ wf, uint4* __restrict__ wd, int OCp_f,
               int OCp_d) {
    pack_w2_body(W, Cout, Cin, k, wf, wd, OCp_f, OCp_d, (long long)blockIdx.x * blockDim.x + threadIdx.x,
                 (long long)gridDim.x * blockDim.x);
}

// every weight matrix of a model in ONE launch: job j = blockIdx.y, table row = {W, wf, wd, Cout, Cin, k} (8 x int64)
__global__ void __launch_bounds__(256)
pack_w2_batched_kernel(const long long* __restrict__ table) {
    const long long* j = table + (long long)blockIdx.y * 8;
    const float* W = reinterpret_cast<const float*>(j[0]);
    uint4* wf = reinterpret_cast<uint4*>(j[1]);
    uint4* wd = reinterpret_cast<uint4*>(j[2]);
    const int Cout = (int)j[3], Cin = (int)j[4], k = (int)j[5];
    pack_w2_body(W, Cout, Cin, k, wf, wd, (Cout + 127) & ~127, (Cin + 127) & ~127,
                 (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

static bool c2_disabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TAMGCN_DISABLE_TC");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
static int c2_num_sms() { return main_sms(); }
static int oc_pad128(int OC) { return (OC + 127) & ~127; }

// offset of the tconv9 tiles inside a packed-weight buffer = size of the conv_tc2 tiles
size_t conv_pack_t9_offset(int Cout, int Cin, int k, int dgrad) {
    const int OC = dgrad ? Cin : Cout, IC = dgrad ? Cout : Cin;
    return (size_t)((k * IC + 63) / 64) * (size_t)oc_pad128(OC) * 128;
}
size_t conv_pack_bytes(int Cout, int Cin, int k, int dgrad) {
    const int OC = dgrad ? Cin : Cout, IC = dgrad ? Cout : Cin;
    return conv_pack_t9_offset(Cout, Cin, k, dgrad) + t9_bytes(OC, IC, k);
}

int conv_pack_weights_batched(const long long* table, int njobs, cudaStream_t st) {
    dim3 grid(16, (unsigned)njobs);
    pack_w2_batched_kernel<<<grid, 256, 0, st>>>(table);
    count_launch();
    return check_launch("conv_pack_weights_batched");
}

int conv_pack_weights(const float* W, int Cout, int Cin, int k, void* wf, void* wd, cudaStream_t st) {
    const size_t units = (conv_pack_bytes(Cout, Cin, k, 0) + conv_pack_bytes(Cout, Cin, k, 1)) / 16;
    int blocks = (int)((units + 255) / 256);
    if (blocks > 592) blocks = 592;
    if (blocks < 1) blocks = 1;
    pack_w2_kernel<<<blocks, 256, 0, st>>>(W, Cout, Cin, k, (uint4*)wf, (uint4*)wd, oc_pad128(Cout), oc_pad128(Cin));
    count_launch();
    return check_launch("conv_pack_weights");
}

static bool c2_tma_enabled() {
    static const bool on = [] { const char* e = getenv("TAMGCN_C2_TMA"); return !(e && e[0] == '0'); }();
    return on;
}
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*c2_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static c2_encode_fn c2_encoder() {
    static c2_encode_fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess)
            f = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<c2_encode_fn>(f);
    }();
    return fn;
}
static bool c2_encode_tmap(C2TensorMap* out, const void* base, unsigned long long L, unsigned long long Cn, unsigned long long N,
                           unsigned long long cstride_bytes, unsigned long long nstride_bytes) {
    static_assert(sizeof(C2TensorMap) == sizeof(CUtensorMap), "tensor map size");
    c2_encode_fn enc = c2_encoder();
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (cstride_bytes & 15) || (nstride_bytes & 15)) return false;
    if (N == 1) nstride_bytes = cstride_bytes * Cn;           // any multiple of 16 bytes: the dimension has one element
    const cuuint64_t dims[3] = {L, Cn, N};
    const cuuint64_t strides[2] = {cstride_bytes, nstride_bytes};
    const cuuint32_t box[3] = {64, 64, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int MODE>
static int launch_conv_tc2(const ConvP& g, const Opnd& xo, const void* wpack, void* out, long long ons, const C2Epi& ep,
                           cudaStream_t st) {
    C2P p = {};
    p.g = g;
    p.IC = (MODE == 0) ? g.Cin : g.Cout;
    p.OC = (MODE == 0) ? g.Cout : g.Cin;
    p.Lin = (MODE == 0) ? g.T * g.V : g.To * g.V;
    p.Lout = (MODE == 0) ? g.To * g.V : g.T * g.V;
    p.KT = g.k * p.IC;
    p.KTp = (p.KT + 15) & ~15;
    p.nchunk = (p.KT + 63) / 64;
    p.MT_total = oc_pad128(p.OC) / 128;
    p.n_oct = (p.MT_total + 1) / 2;
    const int mtmax = p.MT_total < 2 ? p.MT_total : 2;
    p.fast = (g.k == 1 && g.s == 1) ? 1 : 0;
    p.ons = ons;
    { const char* e = getenv("TAMGCN_C2_DBG"); p.dbg = e ? atoi(e) : 0; }
    // access granularity of the activation operand (elements): 8 = 16-byte, 4 = 8-byte, 1 = element
    auto aligned = [&](int gr) {
        bool ok = (p.Lin % gr == 0) && (xo.pns % gr == 0) && ((((uintptr_t)xo.p) & (uintptr_t)(2 * gr - 1)) == 0);
        if (xo.q) ok = ok && (xo.qns % gr == 0) && ((((uintptr_t)xo.q) & (uintptr_t)(2 * gr - 1)) == 0);
        if (!p.fast) ok = ok && (g.V % gr == 0);          // taps / stride shift whole V-rows
        else ok = ok && (p.Lout % gr == 0);
        return ok;
    };
    p.gran = aligned(8) ? 8 : (aligned(4) ? 4 : 1);
    // positions per tile: as few tiles per sample as the TMEM budget allows, rounded to the MMA N granularity
    const int nmax = (mtmax == 2) ? 128 : 256;
    int tps = (p.Lout + nmax - 1) / nmax;
    int NT = (((p.Lout + tps - 1) / tps) + 15) & ~15;
    if (NT < 16) NT = 16;
    {
        // Wave quantisation: the persistent CTAs of an output-channel group share N * tps tiles; with a few tiles per CTA
        // the last, partly filled round costs as much as a full one.  Among tile widths up to the TMEM limit pick the one
        // with the smallest (rounds x (tile width + a fixed per-tile cost)).
        static const int ovh = [] { const char* e = getenv("TAMGCN_C2_TILE_OVH"); return e ? atoi(e) : 64; }();
        if (ovh >= 0) {
            const long long per_oct = c2_num_sms() / ((p.MT_total + 1) / 2) > 0 ? c2_num_sms() / ((p.MT_total + 1) / 2) : 1;
            long long best = -1;
            int bestNT = NT;
            for (int cand = tps; cand <= tps + 12; ++cand) {
                int nt = (((p.Lout + cand - 1) / cand) + 15) & ~15;
                if (nt < 16) nt = 16;
                if (nt > nmax) continue;
                const long long tiles_c = (long long)g.N * ((p.Lout + nt - 1) / nt);
                const long long rounds = (tiles_c + per_oct - 1) / per_oct;
                const long long cost = rounds * (nt + ovh);
                if (best < 0 || cost < best) { best = cost; bestNT = nt; }
                if (nt == 16) break;
            }
            NT = bestNT;
        }
    }
    p.NT = NT;
    p.tps = (p.Lout + NT - 1) / NT;
    p.NTp = (NT + 31) & ~31;
    p.nblk = (NT + 63) / 64;
    p.upr = p.gran >= 4 ? NT / p.gran : NT;
    static const int ra8_env = [] { const char* e = getenv("TAMGCN_C2_RA8"); return e ? atoi(e) : 1; }();
    p.ra8 = (ra8_env && p.gran == 1 && g.s == 1 && NT % 8 == 0) ? 1 : 0;
    if (p.gran >= 4 && NT % p.gran != 0) return 0;
    const long long tiles = (long long)g.N * p.tps;
    if (tiles > 0x7fffffffLL) return 0;
    p.n_tiles = (int)tiles;
    p.tmem_cols = (int)tmem_cols_pow2((uint32_t)(2 * mtmax * p.NTp));
    if (p.tmem_cols > 512) return 0;
    p.x_bytes = (uint32_t)p.nblk * 8192u;
    p.q_bytes = (xo.q && p.gran >= 4) ? p.x_bytes : 0u;       // second tensor of a lazy operand lands next to the first
    p.stage_bytes = p.x_bytes + p.q_bytes + (uint32_t)mtmax * 16384u;
    const uint32_t szH = (sizeof(C2Hdr) + 15) & ~15u, szC = (uint32_t)((3 * p.IC * 4 + 15) & ~15), szG = 16 * 2560;
    const uint32_t budget = 227u * 1024u - 1024u;
    const uint32_t fixed = szH + szC + szG;
    if (fixed + 2 * p.stage_bytes > budget) return 0;
    p.S = (int)((budget - fixed) / p.stage_bytes);
    if (p.S > C2_SMAX) p.S = C2_SMAX;
    p.lag = p.S >= 4 ? 2 : (p.S == 3 ? 1 : 0);
    p.off_hdr = (uint32_t)p.S * p.stage_bytes;
    p.off_coef = p.off_hdr + szH;
    p.off_stg = p.off_coef + szC;
    const size_t sm = (size_t)p.off_stg + szG + 1024;
    // epilogue vector width (elements) for out / addend / mask rows
    p.vec = 1;
    auto okv = [&](int v) {
        const bool base = (p.Lout % v == 0) && (ons % v == 0) && ((((uintptr_t)out) & (uintptr_t)(2 * v - 1)) == 0);
        const bool a = !ep.addend || ((ep.addns % v == 0) && ((((uintptr_t)ep.addend) & (uintptr_t)(2 * v - 1)) == 0));
        const bool m = !ep.has_mask || ((ep.maskns % v == 0) && ((((uintptr_t)ep.maskp) & (uintptr_t)(2 * v - 1)) == 0));
        return base && a && m;
    };
    if (okv(8)) p.vec = 8; else if (okv(4)) p.vec = 4;
    const bool plain = !xo.a && !xo.b && !xo.c && !xo.q && !xo.relu;
    long long per = c2_num_sms() / p.n_oct;
    if (per < 1) per = 1;
    if (per > tiles) per = tiles;
    const int grid = (int)(per * p.n_oct);
    const bool extra = (MODE == 1) && (ep.addend || ep.has_mask || ep.bcast);
    // split of the non-MMA warps between epilogue and producers by their estimated instruction load
    {
        const int nw = (extra ? C2_THREADS_SMALL : C2_THREADS_BIG) / 32;
        const double epi_work = 6.0 * (p.OC < 256 ? p.OC : 256), prod_work = (plain ? 1.5 : (xo.q ? 14.0 : 9.0)) * p.KT;
        p.n_epi = 8;
        if (nw >= 24 && epi_work > prod_work) p.n_epi = 16;
        // K-heavy shapes (conv3 data gradient: K = 3 Cout, lazy two-tensor operand): the MMA warp waits on the producers
        // for most of a tile while the epilogue runs once per 3-16 chunks: one epilogue warp per TMEM lane quarter.
        static const double epi4 = [] { const char* e = getenv("TAMGCN_C2_EPI4"); return e ? atof(e) : 2.5; }();
        if (epi4 > 0 && prod_work > epi4 * epi_work) p.n_epi = 4;
    }
    // TMA for the plain 1x1 operand: a 3-D tensor map (positions, channels, samples) whose 64 x 64 boxes land as the
    // SWIZZLE_128B tile the MMA reads.  Needs 16-byte aligned planes (global strides are multiples of 16 bytes).
    C2TensorMap tmap = {};
    p.use_tma = 0;
    // (lazy one-tensor operands through TMA + in-place transform: parity-green, no measurable gain -> opt-in)
    static const int tma_lazy = [] { const char* e = getenv("TAMGCN_C2_TMA_LAZY"); return e ? atoi(e) : 0; }();
    if ((plain || (!xo.q && tma_lazy)) && p.fast && p.gran == 8 && c2_tma_enabled()) {
        const int IC = p.IC;
        if (c2_encode_tmap(&tmap, xo.p, (unsigned long long)p.Lin, (unsigned long long)IC, (unsigned long long)g.N,
                           (unsigned long long)p.Lin * 2ull, (unsigned long long)xo.pns * 2ull))
            p.use_tma = plain ? 1 : 2;       // 2: raw boxes, transformed in place by the producers
    }
#define C2_LAUNCH(PL, EX)                                                                                              \
    do {                                                                                                               \
        static SmemLimit lim;                                                                                          \
        ensure_smem(conv_tc2_kernel<MODE, PL, EX>, lim, sm);                                                           \
        conv_tc2_kernel<MODE, PL, EX><<<grid, EX ? C2_THREADS_SMALL : C2_THREADS_BIG, sm, st>>>(p, xo, (const uint8_t*)wpack, (bf16*)out, ep, tmap); \
    } while (0)
    if (MODE == 0 || !extra) { if (plain) C2_LAUNCH(1, 0); else C2_LAUNCH(0, 0); }
    else { if (plain) C2_LAUNCH(1, 1); else C2_LAUNCH(0, 1); }
#undef C2_LAUNCH
    count_launch();
    const int rc = check_launch(MODE == 0 ? "conv_fwd(tcgen05)" : "conv_dgrad(tcgen05)");
    return rc < 0 ? rc : 1;
}

// return 1 if handled, 0 if the caller should use the SIMT kernel, <0 on error
int conv_fwd_tc(const tamgcn_conv_geom* g, const Opnd& x, const void* wpack, const float* bias, void* y, long long yns,
                double* ssum, double* ssq, int stat_c0, cudaStream_t st) {
    if (c2_disabled() || !wpack || g->Cout > 8192) return 0;
    ConvP p = {g->N, g->Cin, g->Cout, g->T, g->To, g->V, g->k, g->stride, g->dil, g->pad};
    C2Epi ep = {};
    ep.bias = bias; ep.s1 = ssum; ep.s2 = ssq; ep.stat_c0 = stat_c0;
    return launch_conv_tc2<0>(p, x, wpack, y, yns, ep, st);
}

int conv_dgrad_tc(const tamgcn_conv_geom* g, const Opnd& dy, const void* wpack, void* dx, long long dxns,
                  const void* addend, long long addns, const float* bcast, float bscale, const Opnd* mask, double* s1,
                  double* s2, cudaStream_t st) {
    if (c2_disabled() || !wpack || g->Cin > 8192) return 0;
    if (mask && mask->q) return 0;
    ConvP p = {g->N, g->Cin, g->Cout, g->T, g->To, g->V, g->k, g->stride, g->dil, g->pad};
    C2Epi ep = {};
    ep.s1 = s1; ep.s2 = s2; ep.stat_c0 = 0;
    ep.addend = (const bf16*)addend; ep.addns = addns; ep.bcast = bcast; ep.bscale = bscale;
    if (mask) {
        ep.has_mask = 1; ep.maskp = (const bf16*)mask->p; ep.maskns = mask->pns; ep.maska = mask->a; ep.maskc = mask->c;
    }
    return launch_conv_tc2<1>(p, dy, wpack, dx, dxns, ep, st);
}

}  // namespace tamgcn

extern "C" int64_t tamgcn_conv_pack_bytes(int Cout, int Cin, int k, int dgrad) {
    return (int64_t)tamgcn::conv_pack_bytes(Cout, Cin, k, dgrad);
}

extern "C" int tamgcn_conv_pack_weights_batched(const int64_t* table, int njobs, tamgcn_stream stream) {
    TG_REQUIRE(table && njobs > 0 && njobs <= 65535, "conv_pack_weights_batched: bad arguments (njobs=%d)", njobs);
    return tamgcn::conv_pack_weights_batched(reinterpret_cast<const long long*>(table), njobs, (cudaStream_t)stream);
}

extern "C" int tamgcn_conv_pack_weights(const float* W, int Cout, int Cin, int k, void* wpack_fwd, void* wpack_dgrad,
                                        tamgcn_stream stream) {
    TG_REQUIRE(W && Cout > 0 && Cin > 0 && k > 0 && (wpack_fwd || wpack_dgrad), "conv_pack_weights: bad arguments");
    return tamgcn::conv_pack_weights(W, Cout, Cin, k, wpack_fwd, wpack_dgrad, (cudaStream_t)stream);
}
