// tc_common.cuh — tcgen05 / TMEM / mbarrier primitives (inline PTX, sm_100a) shared by the tensor-core kernels.
//
// Operand tiles live in shared memory in the canonical K-major SWIZZLE_128B layout: a tile is R rows of
// 64 bf16 (128 bytes); row r starts at byte r*128 of a 1024-byte aligned buffer and its eight 16-byte
// chunks are stored at chunk index (c ^ (r & 7)).  This is the layout TMA's 128B swizzle produces; here the
// CUDA cores produce it themselves because every activation operand is transformed on the fly
// (BatchNorm-apply / ReLU / BatchNorm-backward affine, see tamgcn_operand) between global memory and the MMA.
#pragma once
#include "common.cuh"

namespace tamgcn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared (the TMA engine, no tensor map); completion is signalled on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// L2 prefetch of a contiguous global range (16-byte aligned address, size a multiple of 16): one instruction, no
// destination, no completion to wait for.  Persistent CTAs use it a few tiles ahead of their loaders so that the loads
// proper see L2 latency instead of HBM latency.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// Bounded wait (a wedged pipeline must not hang the GPU): returns false after ~2 s.  The suspend-time hint lets the
// hardware park the warp for a while instead of re-polling at once, but the parking time is bounded by the
// implementation and failed probes are common, so the retry path is kept to three instructions (probe, two
// branches, a counter): the wall-clock watchdog is only consulted once every 2048 failed probes.
__device__ __forceinline__ bool mbar_wait_spin(uint64_t* bar, uint32_t parity);
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef TAMGCN_SPIN_WAIT
    return mbar_wait_spin(bar, parity);
#endif
    const uint32_t addr = smem_u32(bar);
    long long t0 = 0;
#pragma unroll 1
    for (uint32_t round = 0;; ++round) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .pred q;\n\t.reg .u32 n;\n\t"
            "mov.u32 n, 2048;\n\t"
            "MBW_LOOP:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "@p bra MBW_DONE;\n\t"
            "sub.u32 n, n, 1;\n\t"
            "setp.ne.u32 q, n, 0;\n\t"
            "@q bra MBW_LOOP;\n\t"
            "MBW_DONE:\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(0x989680u)
            : "memory");
        if (done) return true;
        if (round == 0) t0 = clock64();
        else if (clock64() - t0 > 4000000000LL) return false;
    }
}

// Same bounded wait without the suspend-time hint: the warp re-probes as soon as the hardware's own (short) try_wait
// window expires.  For the single latency-critical poller of a role (the MMA issuer, the one polling warp of the
// other roles): a hinted wait wakes up hundreds of cycles after the phase flips, and in a pipeline whose roles hand
// tiles to each other every ~2000 cycles that wake-up latency sits on the critical path of every hand-off.
__device__ __forceinline__ bool mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    long long t0 = 0;
#pragma unroll 1
    for (uint32_t round = 0;; ++round) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .pred q;\n\t.reg .u32 n;\n\t"
            "mov.u32 n, 4096;\n\t"
            "MBS_LOOP:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "@p bra MBS_DONE;\n\t"
            "sub.u32 n, n, 1;\n\t"
            "setp.ne.u32 q, n, 0;\n\t"
            "@q bra MBS_LOOP;\n\t"
            "MBS_DONE:\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return true;
        if (round == 0) t0 = clock64();
        else if (clock64() - t0 > 4000000000LL) return false;
    }
}

// ---- proxies / tcgen05 fences ---------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ------------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t tmem_cols_pow2(uint32_t n) {
    uint32_t c = 32;
    while (c < n) c <<= 1;
    return c;
}
// one full warp; writes the TMEM base address to *dst (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- UMMA descriptors ------------------------------------------------------------------------------
// K-major, SWIZZLE_128B: start address >> 4 | LBO (unused for swizzled K-major, canonical value 1) |
// SBO = 1024 B (eight 128-byte rows) | descriptor version 1 (sm_100) | layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D = fp32, A = B = bf16, both K-major, M x N
__host__ __device__ inline uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on `bar` when all tcgen05.mma issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---- swizzled tile addressing ------------------------------------------------------------------------
// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a K-major SW128 tile
__device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t c) { return r * 128u + ((c ^ (r & 7u)) << 4); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 16 bytes starting `sft` (even, 0..14) bytes into the 32-byte pair (lo, hi): the realignment of a 2-byte aligned run of
// 8 bf16 that was fetched as two 16-byte aligned words
__device__ __forceinline__ uint4 tc_realign16(const uint4& lo, const uint4& hi, uint32_t sft) {
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    const uint32_t ws = sft >> 2;
    uint32_t o[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) o[i] = ws == 0 ? w[i] : (ws == 1 ? w[i + 1] : (ws == 2 ? w[i + 2] : w[i + 3]));
    if (sft & 2u) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __funnelshift_r(o[i], o[i + 1], 16);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}
// 8 consecutive bf16 at p (2-byte aligned): two aligned 16-byte loads + realignment.  The second word is only touched
// when the run extends into it, so no byte outside the 16-byte granules that hold valid data is read.
__device__ __forceinline__ uint4 tc_ld8_unaligned(const bf16* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t sft = (uint32_t)(a & 15);
    const uint4* q = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
    const uint4 lo = __ldg(q);
    if (sft == 0) return lo;
    const uint4 hi = __ldg(q + 1);
    return tc_realign16(lo, hi, sft);
}

// Per-column sum of a 32 (lanes = rows) x 32 (registers = columns) tile: after the call lane L holds the sum of
// column L over the 32 lanes.  31 shuffles instead of 160.
__device__ __forceinline__ float warp_colsum32(float (&v)[32]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

}  // namespace tamgcn
