// conv_tc.cu — (k x 1) convolutions on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// bf16 activations, fp32 parameters converted to bf16 on the way into shared memory, fp32 accumulation.
// These GEMMs have small K (Cin*k = 64..320 for CTR-GCN) and arithmetic intensity well under the B200
// ridge point, so they are bound by how fast the activation operand can be streamed, not by the tensor
// pipe.  The kernels are built around that:
//
//   forward / dgrad : one CTA = 128 positions (UMMA M = 128) x ALL output channels (UMMA N = Cout <= 256), so
//                     the activation operand is read exactly once; K is walked in 64-wide chunks, double
//                     buffered in shared memory; the CUDA cores gather + transform (lazy operand: BN-apply,
//                     ReLU, BN-backward affine, conv taps with zero padding / stride) + convert the operand
//                     into the canonical K-major SWIZZLE_128B layout while the previous chunk's MMAs run.
//   epilogue        : TMEM -> registers (tcgen05.ld 32x32b), bias / addend / mask, bf16 stores that are
//                     coalesced along positions, BatchNorm statistics via a 31-shuffle column reduction.
//   wgrad           : M = Cout tile, N = Cin tile per tap, K = positions (both operands natively K-major),
//                     split-K over CTAs, fp32 atomics from TMEM into dW.
#include "tc_common.cuh"
#include <cstdlib>

namespace tamgcn {

struct ConvP {
    int N, Cin, Cout, T, To, V, k, s, d, p;
};

#define TC_BM 128
#define TC_BK 64
#define TC_STAGES 2          // wgrad kernel
#define TC_THREADS 256       // wgrad kernel

// ------------------------------------------------------------------------------------------------
// weight pre-pack: fp32 (Cout, Cin, k) -> bf16 tiles in the exact shared-memory image the MMA wants
//   fwd  pack: [tap j][chunk of 64 ic][OCp rows oc][64 ic]   B(oc, ic) = W[oc, ic, j]      OC = Cout, IC = Cin
//   dgrad pack: [tap j][chunk of 64 ic][OCp rows oc][64 ic]  B(oc, ic) = W[ic, oc, j]      OC = Cin,  IC = Cout
// rows are 128 bytes, 16-byte chunks XOR-swizzled with (row & 7): a (tap, chunk) block of rows is copied into
// shared memory with ONE bulk async copy (cp.async.bulk, the TMA engine) per pipeline stage.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_w_kernel(const float* __restrict__ W, int Cout, int Cin, int k, uint4* __restrict__ wf, uint4* __restrict__ wd,
              int OCp_f, int OCp_d) {
    const int CK = Cin * k;
    for (int which = 0; which < 2; ++which) {
        uint4* dst = which ? wd : wf;
        if (!dst) continue;
        const int OC = which ? Cin : Cout, IC = which ? Cout : Cin;
        const int OCp = which ? OCp_d : OCp_f, nch = (IC + 63) / 64;
        const long long units = (long long)k * nch * OCp * 8;
        for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (long long)gridDim.x * blockDim.x) {
            const int chunk = (int)(u & 7);
            const long long rowl = u >> 3;
            const int oc = (int)(rowl % OCp);
            const long long blk = rowl / OCp;
            const int ch = (int)(blk % nch), j = (int)(blk / nch);
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ic = ch * 64 + chunk * 8 + e;
                float w = 0.f;
                if (oc < OC && ic < IC)
                    w = which ? __ldg(W + (long long)ic * CK + oc * k + j) : __ldg(W + (long long)oc * CK + ic * k + j);
                f[e] = w;
            }
            uint4 o;
            o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
            dst[(blk * OCp + oc) * 8 + (chunk ^ (oc & 7))] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward and data-gradient share one kernel:  D[oc, pos] = sum_{tap, ic} Wp(oc, tap, ic) * X(pos, tap, ic)
//   MODE 0 (fwd)  : ic = input channel, oc = output channel, X = x(n, ic, to*s + j*d - p, v)
//   MODE 1 (dgrad): ic = conv output channel, oc = conv input channel,
//                   X = dY(n, ic, (t + p - j*d)/s, v) when divisible and in range
// UMMA orientation: M = 128 output channels (A operand = pre-packed weights), N = 128 positions (B operand =
// activations), so a TMEM lane is a channel and its columns are consecutive positions: an epilogue thread owns
// one channel, stores 64 contiguous bytes per 32 positions and keeps that channel's BatchNorm sums in registers.
//
// Persistent, warp-specialised CTA (one per SM), 13 warps:
//   warps 0-7  : activation producers.  Thread (row = tid & 127, khalf = tid >> 7) gathers 32 channels of one
//                position (coalesced across the warp), applies the lazy-operand transform, converts to bf16 and
//                writes four 16-byte swizzled chunks of the K-major tile; thread 0 also launches the bulk async
//                copy (TMA engine) of the pre-packed weight tile of that K chunk.
//   warp 8     : lane 0 issues tcgen05.mma (128 x 128 x 16) and the commits.
//   warps 9-12 : epilogue.  TMEM (double-buffered accumulators) -> registers -> bias / addend / mask -> bf16.
// Pipelines: full/empty mbarriers over a ring of KS operand stages; tfull/tempty over the two TMEM buffers.
// ------------------------------------------------------------------------------------------------
#define KS 4
#define NPOS 128
#define NPROD 256
#define TC2_THREADS (NPROD + 32 + 128)

struct TcHdr {
    uint64_t full[KS], empty[KS], tfull[2], tempty[2];
    uint32_t tmem_base;
    volatile uint32_t error;
};

struct TcEpi {
    const float* bias;        // fwd
    double* s1;               // fwd: sum,  dgrad: sum dX
    double* s2;               // fwd: sumsq, dgrad: sum dX * maskP
    int stat_c0;
    const bf16* addend;       // dgrad
    long long addns;
    const float* bcast;
    float bscale;
    const bf16* maskp;        // dgrad mask operand P (a*P + c > 0)
    long long maskns;
    const float* maska;
    const float* maskc;
    int has_mask;
};

template <int VEC> struct BfVec;
template <> struct BfVec<8> { typedef uint4 T; };
template <> struct BfVec<4> { typedef uint2 T; };
template <> struct BfVec<1> { typedef unsigned short T; };

template <int VEC>
__device__ __forceinline__ void ld_bf16_vec(const bf16* p, float* f) {
    if (VEC == 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    } else if (VEC == 4) {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
        f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
    } else {
        f[0] = __bfloat162float(*p);
    }
}
template <int VEC>
__device__ __forceinline__ void st_bf16_vec(bf16* p, const float* f) {
    if (VEC == 8) {
        uint4 u;
        u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
        *reinterpret_cast<uint4*>(p) = u;
    } else if (VEC == 4) {
        uint2 u;
        u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
        *reinterpret_cast<uint2*>(p) = u;
    } else {
        *p = __float2bfloat16_rn(f[0]);
    }
}

// epilogue of one channel row x 32 positions.  Returns the two BatchNorm partial sums through s/q.
template <int MODE, int VEC>
__device__ __forceinline__ void epi_block(float (&acc)[32], int nvalid, bf16* __restrict__ po, float bias,
                                          const TcEpi& ep, const bf16* __restrict__ padd, const bf16* __restrict__ pmask,
                                          const float* __restrict__ pbc, int v0, int V, float ma, float mc, float& s,
                                          float& q) {
#pragma unroll
    for (int i0 = 0; i0 < 32; i0 += VEC) {
        if (i0 < nvalid) {                       // nvalid is a multiple of VEC on the vector paths
            float o[VEC], ad[VEC], mk[VEC];
            if (MODE == 1) {
                if (padd) ld_bf16_vec<VEC>(padd + i0, ad);
                if (pmask) ld_bf16_vec<VEC>(pmask + i0, mk);
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                float val = acc[i0 + e];
                if (MODE == 0) {
                    val = rnd<bf16>(val + bias);
                    s += val;
                    q = fmaf(val, val, q);
                } else {
                    if (padd) val += ad[e];
                    if (pbc) {
                        int vv = v0 + i0 + e;
                        vv -= (vv / V) * V;
                        val = fmaf(__ldg(pbc + vv), ep.bscale, val);
                    }
                    if (pmask) {
                        if (!(fmaf(ma, mk[e], mc) > 0.f)) val = 0.f;
                        val = rnd<bf16>(val);
                        s += val;
                        q = fmaf(val, mk[e], q);
                    }
                }
                o[e] = val;
            }
            st_bf16_vec<VEC>(po + i0, o);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_tc_kernel(ConvP g, Opnd xo, const uint8_t* __restrict__ wpack, bf16* __restrict__ out, long long ons, TcEpi ep,
               int IC, int OC, int MT_total, int n_oct, int Lin, int Lout, int tiles_per_sample, int n_tiles,
               int tmem_cols, int vec) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int mtmax = min(2, MT_total);
    const uint32_t x_bytes = NPOS * 128, w_bytes_max = (uint32_t)mtmax * 128 * 128, stage_bytes = x_bytes + w_bytes_max;
    TcHdr* hdr = (TcHdr*)(smem + KS * stage_bytes);
    float* coef = (float*)(hdr + 1);             // [3][IC] lazy-operand coefficients

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunk_c = (IC + TC_BK - 1) / TC_BK;
    const int NK = g.k * nchunk_c;
    // this CTA always works on the same group of <= 2 channel tiles (so BN sums can live in registers)
    const int oct = blockIdx.x % n_oct, mt0 = oct * 2, mt_cnt = min(2, MT_total - mt0);
    const int tile0 = blockIdx.x / n_oct, tile_step = gridDim.x / n_oct;
    const uint32_t w_bytes = (uint32_t)mt_cnt * 128 * 128;
    const int OCpad = MT_total * 128;

    if (warp == 8) tmem_alloc(&hdr->tmem_base, (uint32_t)tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < KS; ++i) { mbar_init(&hdr->full[i], NPROD); mbar_init(&hdr->empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&hdr->tfull[i], 1); mbar_init(&hdr->tempty[i], 128); }
        hdr->error = 0;
        fence_mbar_init();
    }
    for (int i = tid; i < IC; i += TC2_THREADS) {
        const OpCoef cf = opnd_coef(xo, i);
        coef[i] = cf.a; coef[IC + i] = cf.b; coef[2 * IC + i] = cf.c;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = hdr->tmem_base;

    if (warp < 8) {
        // =========================== activation producers (+ weight-tile bulk copies) ===========================
        const int r = tid & 127, khalf = tid >> 7;
        int stg = 0, ph = 0;
        bool ok = true;
        for (int tile = tile0; tile < n_tiles && ok; tile += tile_step) {
            const int n = tile / tiles_per_sample, pos0 = (tile - n * tiles_per_sample) * NPOS;
            const int pos = pos0 + r;
            const bool rvalid = pos < Lout;
            const int tq = pos / g.V, v = pos - tq * g.V;
            for (int kc = 0; kc < NK; ++kc) {
                const int j = kc / nchunk_c, cc = kc - j * nchunk_c, c0 = cc * TC_BK;
                if (!mbar_wait(&hdr->empty[stg], (uint32_t)(ph ^ 1))) hdr->error = 1;
                if (hdr->error) { ok = false; break; }
                const uint32_t sx = smem_u32(smem + stg * stage_bytes);
                if (tid == 0) {
                    mbar_expect_tx(&hdr->full[stg], w_bytes);
                    const uint8_t* src = wpack + ((size_t)(j * nchunk_c + cc) * OCpad + (size_t)mt0 * 128) * 128;
                    bulk_g2s(sx + x_bytes, src, w_bytes, &hdr->full[stg]);
                }
                long long aoff = -1;
                if (rvalid) {
                    if (MODE == 0) {
                        const int t = tq * g.s + j * g.d - g.p;
                        if (t >= 0 && t < g.T) aoff = (long long)t * g.V + v;
                    } else {
                        const int num = tq + g.p - j * g.d;
                        if (num >= 0) {
                            const int to = num / g.s;
                            if (to * g.s == num && to < g.To) aoff = (long long)to * g.V + v;
                        }
                    }
                }
                const bf16* pp = (const bf16*)xo.p + (long long)n * xo.pns + aoff;
                const bf16* pq = xo.q ? (const bf16*)xo.q + (long long)n * xo.qns + aoff : nullptr;
#pragma unroll
                for (int gch = 0; gch < 4; ++gch) {
                    const int chunk = khalf * 4 + gch, cb = c0 + chunk * 8;
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int ci = cb + e;
                        float val = 0.f;
                        if (aoff >= 0 && ci < IC) {
                            val = fmaf(coef[ci], ldf<bf16>(pp + (long long)ci * Lin), coef[2 * IC + ci]);
                            if (pq) val = fmaf(coef[IC + ci], ldf<bf16>(pq + (long long)ci * Lin), val);
                            if (xo.relu) val = fmaxf(val, 0.f);
                        }
                        f[e] = val;
                    }
                    st_shared_v4(sx + sw128_off(r, chunk), pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                                 pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                }
                fence_proxy_async_smem();
                mbar_arrive(&hdr->full[stg]);
                if (++stg == KS) { stg = 0; ph ^= 1; }
            }
        }
    } else if (warp == 8) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, NPOS);
            int stg = 0, ph = 0, it = 0;
            for (int tile = tile0; tile < n_tiles; tile += tile_step, ++it) {
                const int buf = it & 1;
                if (!mbar_wait(&hdr->tempty[buf], (uint32_t)(((it >> 1) & 1) ^ 1))) hdr->error = 1;
                if (hdr->error) break;
                tc_fence_after();
                const uint32_t td = tmem + (uint32_t)(buf * mtmax * NPOS);
                for (int kc = 0; kc < NK; ++kc) {
                    if (!mbar_wait(&hdr->full[stg], (uint32_t)ph)) hdr->error = 1;
                    if (hdr->error) break;
                    tc_fence_after();
                    const uint32_t sx = smem_u32(smem + stg * stage_bytes), sw = sx + x_bytes;
                    for (int mt = 0; mt < mt_cnt; ++mt) {
#pragma unroll
                        for (int kk = 0; kk < TC_BK / 16; ++kk)
                            umma_bf16(td + (uint32_t)(mt * NPOS), umma_desc_sw128(sw + mt * (128 * 128) + kk * 32),
                                      umma_desc_sw128(sx + kk * 32), idesc, (kc > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(&hdr->empty[stg]);
                    if (++stg == KS) { stg = 0; ph ^= 1; }
                }
                umma_commit(&hdr->tfull[buf]);
            }
        }
    } else {
        // =========================== epilogue: one thread = one channel ===========================
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        float st1[2] = {0.f, 0.f}, st2[2] = {0.f, 0.f}, biasr[2] = {0.f, 0.f}, mar[2] = {1.f, 1.f}, mcr[2] = {0.f, 0.f};
        int ocr[2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            ocr[mt] = (mt0 + mt) * 128 + q * 32 + lane;
            if (mt < mt_cnt && ocr[mt] < OC) {
                if (MODE == 0 && ep.bias) biasr[mt] = __ldg(ep.bias + ocr[mt]);
                if (MODE == 1 && ep.has_mask) {
                    if (ep.maska) mar[mt] = __ldg(ep.maska + ocr[mt]);
                    if (ep.maskc) mcr[mt] = __ldg(ep.maskc + ocr[mt]);
                }
            }
        }
        int it = 0;
        for (int tile = tile0; tile < n_tiles; tile += tile_step, ++it) {
            const int n = tile / tiles_per_sample, pos0 = (tile - n * tiles_per_sample) * NPOS;
            const int buf = it & 1;
            if (!mbar_wait(&hdr->tfull[buf], (uint32_t)((it >> 1) & 1))) hdr->error = 1;
            if (hdr->error) break;
            tc_fence_after();
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                if (mt < mt_cnt) {
                    const int oc = ocr[mt];
                    const bool ocv = oc < OC;
                    const long long rowoff = (long long)oc * Lout + pos0;
#pragma unroll 1
                    for (int pb = 0; pb < NPOS / 32; ++pb) {
                        float acc[32];
                        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * mtmax + mt) * NPOS + pb * 32), acc);
                        const int nvalid = min(32, Lout - (pos0 + pb * 32));
                        if (ocv && nvalid > 0) {
                            bf16* po = out + (long long)n * ons + rowoff + pb * 32;
                            const bf16* padd = (MODE == 1 && ep.addend) ? ep.addend + (long long)n * ep.addns + rowoff + pb * 32 : nullptr;
                            const bf16* pmask = (MODE == 1 && ep.has_mask) ? ep.maskp + (long long)n * ep.maskns + rowoff + pb * 32 : nullptr;
                            const float* pbc = (MODE == 1 && ep.bcast) ? ep.bcast + ((long long)n * OC + oc) * g.V : nullptr;
                            const int v0 = (pos0 + pb * 32) % g.V;
                            if (vec == 8)
                                epi_block<MODE, 8>(acc, nvalid, po, biasr[mt], ep, padd, pmask, pbc, v0, g.V, mar[mt], mcr[mt], st1[mt], st2[mt]);
                            else if (vec == 4)
                                epi_block<MODE, 4>(acc, nvalid, po, biasr[mt], ep, padd, pmask, pbc, v0, g.V, mar[mt], mcr[mt], st1[mt], st2[mt]);
                            else
                                epi_block<MODE, 1>(acc, nvalid, po, biasr[mt], ep, padd, pmask, pbc, v0, g.V, mar[mt], mcr[mt], st1[mt], st2[mt]);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&hdr->tempty[buf]);
        }
        if (ep.s1 && !hdr->error) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                if (mt < mt_cnt && ocr[mt] < OC && ocr[mt] >= ep.stat_c0) {
                    atomicAdd(ep.s1 + (ocr[mt] - ep.stat_c0), (double)st1[mt]);
                    atomicAdd(ep.s2 + (ocr[mt] - ep.stat_c0), (double)st2[mt]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

static bool tc_disabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TAMGCN_DISABLE_TC");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// output channels are padded to whole 128-row UMMA tiles in the packed weights
static int oc_pad128(int OC) { return (OC + 127) & ~127; }

size_t conv_pack_bytes(int Cout, int Cin, int k, int dgrad) {
    const int OC = dgrad ? Cin : Cout, IC = dgrad ? Cout : Cin;
    return (size_t)k * ((IC + 63) / 64) * (size_t)oc_pad128(OC) * 128;
}

int conv_pack_weights(const float* W, int Cout, int Cin, int k, void* wf, void* wd, cudaStream_t st) {
    const size_t units = (conv_pack_bytes(Cout, Cin, k, 0) + conv_pack_bytes(Cout, Cin, k, 1)) / 16;
    int blocks = (int)((units + 255) / 256);
    if (blocks > 592) blocks = 592;
    if (blocks < 1) blocks = 1;
    pack_w_kernel<<<blocks, 256, 0, st>>>(W, Cout, Cin, k, (uint4*)wf, (uint4*)wd, oc_pad128(Cout), oc_pad128(Cin));
    count_launch();
    return check_launch("conv_pack_weights");
}



template <int MODE>
static int launch_conv_tc(const ConvP& p, const Opnd& xo, const void* wpack, void* out, long long ons, const TcEpi& ep,
                          cudaStream_t st) {
    const int IC = (MODE == 0) ? p.Cin : p.Cout, OC = (MODE == 0) ? p.Cout : p.Cin;
    const int MT_total = oc_pad128(OC) / 128, n_oct = (MT_total + 1) / 2, mtmax = MT_total < 2 ? MT_total : 2;
    const int Lin = (MODE == 0) ? p.T * p.V : p.To * p.V, Lout = (MODE == 0) ? p.To * p.V : p.T * p.V;
    const int cols = (int)tmem_cols_pow2((uint32_t)(2 * mtmax * NPOS));
    const size_t sm = 1024 + (size_t)KS * (NPOS * 128 + (size_t)mtmax * 128 * 128) + sizeof(TcHdr) + sizeof(float) * 3 * IC;
    TG_REQUIRE(sm <= 227 * 1024, "conv(tcgen05): shared memory %zu too large", sm);
    static int cur = 48 * 1024;
    if ((int)sm > cur) {
        cudaFuncSetAttribute(conv_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cur = (int)sm;
    }
    // widest bf16 vector the epilogue may use for out / addend / mask rows
    int vec = 1;
    auto ok = [&](int v) {
        const bool base = (Lout % v == 0) && (ons % v == 0) && ((((uintptr_t)out) & (2 * v - 1)) == 0);
        const bool a = !ep.addend || ((ep.addns % v == 0) && ((((uintptr_t)ep.addend) & (2 * v - 1)) == 0));
        const bool m = !ep.has_mask || ((ep.maskns % v == 0) && ((((uintptr_t)ep.maskp) & (2 * v - 1)) == 0));
        return base && a && m;
    };
    if (ok(8)) vec = 8; else if (ok(4)) vec = 4;
    const int tps = cdiv(Lout, NPOS);
    const long long tiles = (long long)p.N * tps;
    long long per = num_sms() / n_oct;             // CTAs per channel-tile group
    if (per < 1) per = 1;
    if (per > tiles) per = tiles;
    const int grid = (int)(per * n_oct);
    conv_tc_kernel<MODE><<<grid, TC2_THREADS, sm, st>>>(p, xo, (const uint8_t*)wpack, (bf16*)out, ons, ep, IC, OC, MT_total,
                                                         n_oct, Lin, Lout, tps, (int)tiles, cols, vec);
    count_launch();
    return check_launch(MODE == 0 ? "conv_fwd(tcgen05)" : "conv_dgrad(tcgen05)");
}

// return 1 if handled, 0 if the caller should use the SIMT kernel, <0 on error
int conv_fwd_tc(const tamgcn_conv_geom* g, const Opnd& x, const void* wpack, const float* bias, void* y, long long yns,
                double* ssum, double* ssq, int stat_c0, cudaStream_t st) {
    if (tc_disabled() || !wpack || g->Cout < 8 || g->Cout > 4096) return 0;
    ConvP p = {g->N, g->Cin, g->Cout, g->T, g->To, g->V, g->k, g->stride, g->dil, g->pad};
    TcEpi ep = {};
    ep.bias = bias; ep.s1 = ssum; ep.s2 = ssq; ep.stat_c0 = stat_c0;
    const int rc = launch_conv_tc<0>(p, x, wpack, y, yns, ep, st);
    return rc < 0 ? rc : 1;
}

int conv_dgrad_tc(const tamgcn_conv_geom* g, const Opnd& dy, const void* wpack, void* dx, long long dxns,
                  const void* addend, long long addns, const float* bcast, float bscale, const Opnd* mask, double* s1,
                  double* s2, cudaStream_t st) {
    if (tc_disabled() || !wpack || g->Cin < 8 || g->Cin > 4096) return 0;
    if (mask && mask->q) return 0;
    ConvP p = {g->N, g->Cin, g->Cout, g->T, g->To, g->V, g->k, g->stride, g->dil, g->pad};
    TcEpi ep = {};
    ep.s1 = s1; ep.s2 = s2; ep.stat_c0 = 0;
    ep.addend = (const bf16*)addend; ep.addns = addns; ep.bcast = bcast; ep.bscale = bscale;
    if (mask) {
        ep.has_mask = 1; ep.maskp = (const bf16*)mask->p; ep.maskns = mask->pns; ep.maska = mask->a; ep.maskc = mask->c;
    }
    const int rc = launch_conv_tc<1>(p, dy, wpack, dx, dxns, ep, st);
    return rc < 0 ? rc : 1;
}

// ------------------------------------------------------------------------------------------------
// weight gradient:  dW[co, ci, j] += sum_{n,to,v} dY(n,co,to,v) * X(n,ci,to*s + j*d - p, v)
//   M = output channels (128-row tile), N = NT input channels per tap (tap j accumulates in TMEM columns
//   [j*NTp, (j+1)*NTp)), K = 64 output positions of one sample per step.  Both operands are K-major in global
//   memory already (positions are contiguous), so the producers issue plain row-segment loads.
//   Split-K: CTA z handles work units z, z+S, ... (unit = 64 positions of one sample); fp32 atomics at the end.
// ------------------------------------------------------------------------------------------------
struct TcWgHdr {
    uint64_t mma_done[TC_STAGES];
    uint32_t tmem_base;
    uint32_t error;
};

__global__ void __launch_bounds__(TC_THREADS)
conv_wgrad_tc_kernel(ConvP g, Opnd dyo, Opnd xo, float* __restrict__ dW, float* __restrict__ dbias, int NT, int NTp,
                     int nchunk, int tmem_cols) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = TC_BM * 128, b_bytes = (uint32_t)NTp * 128;
    const uint32_t stage_bytes = a_bytes + (uint32_t)g.k * b_bytes;
    TcWgHdr* hdr = (TcWgHdr*)(smem + TC_STAGES * stage_bytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int co0 = blockIdx.x * TC_BM, ci0 = blockIdx.y * NT;
    const int nco = min(TC_BM, g.Cout - co0), nci = min(NT, g.Cin - ci0);
    const int TV = g.T * g.V, Lo = g.To * g.V, CK = g.Cin * g.k;

    if (warp == 0) tmem_alloc(&hdr->tmem_base, (uint32_t)tmem_cols);
    if (tid == 32) {
        for (int i = 0; i < TC_STAGES; ++i) mbar_init(&hdr->mma_done[i], 1);
        hdr->error = 0;
        fence_mbar_init();
    }
    // rows that are never written (co >= nco, ci >= nci) must read as zero
    for (uint32_t i = tid; i < TC_STAGES * stage_bytes / 16; i += TC_THREADS)
        st_shared_v4(smem_u32(smem) + i * 16, 0u, 0u, 0u, 0u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = hdr->tmem_base;
    const uint32_t idesc = umma_idesc_bf16(TC_BM, (uint32_t)NTp);

    float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
    const int units = g.N * nchunk;
    int it = 0;
    for (int u = blockIdx.z; u < units; u += gridDim.z, ++it) {
        const int st = it % TC_STAGES;
        const int n = u / nchunk, p0 = (u - n * nchunk) * TC_BK;
        if (it >= TC_STAGES) {
            if (!mbar_wait(&hdr->mma_done[st], (uint32_t)((it / TC_STAGES - 1) & 1))) hdr->error = 1;
        }
        const uint32_t sa = smem_u32(smem + st * stage_bytes);
        // ---- A = dY rows (co) x 64 positions
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int un = tid + i * TC_THREADS, row = un >> 3, chunk = un & 7;
            if (row < nco) {
                const OpCoef cf = opnd_coef(dyo, co0 + row);
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int pos = p0 + chunk * 8 + e;
                    f[e] = (pos < Lo) ? opnd_val<bf16>(dyo, cf, n, (long long)(co0 + row) * Lo + pos) : 0.f;
                    dbacc[i] += f[e];
                }
                st_shared_v4(sa + sw128_off(row, chunk), pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                             pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            }
        }
        // ---- B_j = X rows (ci) x 64 positions, shifted by tap j
        for (int j = 0; j < g.k; ++j) {
            const uint32_t sb = sa + a_bytes + (uint32_t)j * b_bytes;
            for (int un = tid; un < nci * 8; un += TC_THREADS) {
                const int row = un >> 3, chunk = un & 7;
                const OpCoef cf = opnd_coef(xo, ci0 + row);
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int pos = p0 + chunk * 8 + e;
                    float val = 0.f;
                    if (pos < Lo) {
                        const int to = pos / g.V, v = pos - to * g.V;
                        const int t = to * g.s + j * g.d - g.p;
                        if (t >= 0 && t < g.T) val = opnd_val<bf16>(xo, cf, n, (long long)(ci0 + row) * TV + t * g.V + v);
                    }
                    f[e] = val;
                }
                st_shared_v4(sb + sw128_off(row, chunk), pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                             pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            for (int j = 0; j < g.k; ++j) {
                const uint32_t sb = sa + a_bytes + (uint32_t)j * b_bytes;
#pragma unroll
                for (int kk = 0; kk < TC_BK / 16; ++kk)
                    umma_bf16(tmem + (uint32_t)(j * NTp), umma_desc_sw128(sa + kk * 32), umma_desc_sw128(sb + kk * 32),
                              idesc, (it > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&hdr->mma_done[st]);
        }
    }
    if (it > 0) {
        const int last = it - 1;
        if (!mbar_wait(&hdr->mma_done[last % TC_STAGES], (uint32_t)((last / TC_STAGES) & 1))) hdr->error = 1;
        tc_fence_after();
        // ---- epilogue: lanes = co rows, columns = (tap, ci)
        const int q = warp & 3, hsel = warp >> 2;
        const int row = q * 32 + lane;
        const int nblk = (g.k * NTp + 31) / 32;
        for (int b = hsel; b < nblk; b += 2) {
            float acc[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 32), acc);
            if (row < nco) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int col = b * 32 + c, j = col / NTp, ci = col - j * NTp;
                    if (j < g.k && ci < nci) atomicAdd(dW + (long long)(co0 + row) * CK + (ci0 + ci) * g.k + j, acc[c]);
                }
            }
        }
    }
    if (dbias && blockIdx.y == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float s = dbacc[i];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            const int row = (tid + i * TC_THREADS) >> 3;
            if ((tid & 7) == 0 && row < nco) atomicAdd(dbias + co0 + row, s);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

int conv_wgrad_tc(const tamgcn_conv_geom* g, const Opnd& dy, const Opnd& x, float* dW, float* dbias, cudaStream_t st) {
    if (tc_disabled() || g->k > 16) return 0;
    ConvP p = {g->N, g->Cin, g->Cout, g->T, g->To, g->V, g->k, g->stride, g->dil, g->pad};
    const int Cinp = (p.Cin + 15) & ~15;
    int NT = (512 / p.k) & ~15;
    if (NT > 256) NT = 256;
    if (NT > Cinp) NT = Cinp;
    if (NT < 16) return 0;
    const int NTp = NT;
    const int cols = (int)tmem_cols_pow2((uint32_t)(p.k * NTp));
    if (cols > 512) return 0;
    const int nchunk = cdiv((long long)p.To * p.V, TC_BK);
    const long long units = (long long)p.N * nchunk;
    const int gx = cdiv(p.Cout, TC_BM), gy = cdiv(p.Cin, NT);
    const size_t sm = 1024 + TC_STAGES * ((size_t)TC_BM * 128 + (size_t)p.k * NTp * 128) + sizeof(TcWgHdr);
    if (sm > 227 * 1024) return 0;
    long long S = (148LL * 2 + gx * gy - 1) / (gx * gy);
    if (S > units) S = units;
    if (S < 1) S = 1;
    if (S > 65535) S = 65535;
    static int cur = 48 * 1024;
    if ((int)sm > cur) {
        cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cur = (int)sm;
    }
    dim3 grid(gx, gy, (unsigned)S);
    conv_wgrad_tc_kernel<<<grid, TC_THREADS, sm, st>>>(p, dy, x, dW, dbias, NT, NTp, nchunk, cols);
    count_launch();
    const int rc = check_launch("conv_wgrad(tcgen05)");
    return rc < 0 ? rc : 1;
}

}  // namespace tamgcn

extern "C" int64_t tamgcn_conv_pack_bytes(int Cout, int Cin, int k, int dgrad) {
    return (int64_t)tamgcn::conv_pack_bytes(Cout, Cin, k, dgrad);
}

extern "C" int tamgcn_conv_pack_weights(const float* W, int Cout, int Cin, int k, void* wpack_fwd, void* wpack_dgrad,
                                        tamgcn_stream stream) {
    TG_REQUIRE(W && Cout > 0 && Cin > 0 && k > 0 && (wpack_fwd || wpack_dgrad), "conv_pack_weights: bad arguments");
    return tamgcn::conv_pack_weights(W, Cout, Cin, k, wpack_fwd, wpack_dgrad, (cudaStream_t)stream);
}
