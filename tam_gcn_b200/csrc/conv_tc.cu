// conv_tc.cu — (k x 1) convolution weight gradient on tcgen05 (forward / data gradient: conv_tc2.cu).
//
//   wgrad : M = Cout tile, N = Cin tile per tap, K = positions (both operands natively K-major),
//           split-K over CTAs, fp32 atomics from TMEM into dW.
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

static bool tc_disabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TAMGCN_DISABLE_TC");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
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
    static SmemLimit lim;
    ensure_smem(conv_wgrad_tc_kernel, lim, sm);
    dim3 grid(gx, gy, (unsigned)S);
    conv_wgrad_tc_kernel<<<grid, TC_THREADS, sm, st>>>(p, dy, x, dW, dbias, NT, NTp, nchunk, cols);
    count_launch();
    const int rc = check_launch("conv_wgrad(tcgen05)");
    return rc < 0 ? rc : 1;
}

}  // namespace tamgcn
