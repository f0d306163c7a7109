// common.cuh — shared device/host helpers for libtamgcn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/tamgcn.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libtamgcn targets sm_100a (B200) only"
#endif

namespace tamgcn {

typedef __nv_bfloat16 bf16;

// ---- error plumbing ---------------------------------------------------------------------------
int set_error(const char* fmt, ...);          // returns -1
void count_launch(int n = 1);
int check_launch(const char* what);           // cudaGetLastError -> 0 / -2

// ---- per-device launch state ------------------------------------------------------------------
// The opt-in dynamic shared-memory limit (cudaFuncSetAttribute) and the SM count belong to a DEVICE, while this
// library lives once per PROCESS: under nn.DataParallel (processor/io.py:85-87) one process drives several devices
// from several threads, so both are cached per device ordinal, and raising a limit is serialised by a mutex.
constexpr int TG_MAX_DEVICES = 64;
int current_device();                          // cudaGetDevice, clamped to [0, TG_MAX_DEVICES)
int num_sms();                                 // SM count of the current device (cached per device)
struct SmemLimit {                             // one per kernel instantiation (function-local static)
    std::atomic<int> cur[TG_MAX_DEVICES];
    SmemLimit() { for (int i = 0; i < TG_MAX_DEVICES; ++i) cur[i].store(48 * 1024, std::memory_order_relaxed); }
};
void raise_smem_limit(const void* kernel, SmemLimit& lim, int dev, size_t bytes);   // slow path, mutex inside
// raise the dynamic shared-memory limit of `kernel` on the current device only when a larger size than ever before
// is needed there (warm-up calls do it; replays / CUDA-graph captures then issue no attribute call)
template <typename K>
static inline void ensure_smem(K kernel, SmemLimit& lim, size_t bytes) {
    if (bytes <= 48 * 1024) return;
    const int dev = current_device();
    if ((int)bytes > lim.cur[dev].load(std::memory_order_acquire)) raise_smem_limit((const void*)kernel, lim, dev, bytes);
}

// Share (percent, 1..100) of the SMs that the persistent weight-gradient kernels of the calling thread may occupy.
// The engine launches them on a side stream next to the data-gradient chain: at 100 % they grab every SM and the
// two streams merely alternate, at ~50 % both run at once (tamgcn_set_wgrad_sm_share / tamgcn_set_main_sm_share:
// process-wide, default 100; the main share applies to the persistent convolution forward / data-gradient kernels).
int wgrad_sm_share();
int main_sm_share();
static inline int main_sms() {
    const int n = num_sms() * main_sm_share() / 100;
    return n < 1 ? 1 : n;
}
static inline int wgrad_sms() {
    const int n = num_sms() * wgrad_sm_share() / 100;
    return n < 1 ? 1 : n;
}

#define TG_REQUIRE(cond, ...)                       \
    do {                                            \
        if (!(cond)) return tamgcn::set_error(__VA_ARGS__); \
    } while (0)

// ---- element access ---------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) {
    return __bfloat162float(__ldg(p));
}
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
// value as it will be read back from storage (so BN statistics describe the stored tensor)
template <typename T> __device__ __forceinline__ float rnd(float v);
template <> __device__ __forceinline__ float rnd<float>(float v) { return v; }
template <> __device__ __forceinline__ float rnd<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// ---- lazy operand (see tamgcn_operand) --------------------------------------------------------
struct Opnd {
    const void* p;
    const void* q;
    const float* a;
    const float* b;
    const float* c;
    long long pns, qns;
    int relu;
};
inline Opnd make_opnd(const tamgcn_operand* o) {
    Opnd r;
    r.p = o->p; r.q = (o->b != nullptr) ? o->q : nullptr;
    r.a = o->a; r.b = (o->q != nullptr) ? o->b : nullptr; r.c = o->c;
    r.pns = o->p_nstride; r.qns = o->q_nstride; r.relu = o->relu;
    return r;
}
inline Opnd plain_opnd(const void* p, long long ns) {
    Opnd r; r.p = p; r.q = nullptr; r.a = r.b = r.c = nullptr; r.pns = ns; r.qns = 0; r.relu = 0;
    return r;
}
// per-channel coefficients of an operand, hoisted out of inner loops
struct OpCoef { float a, b, c; };
__device__ __forceinline__ OpCoef opnd_coef(const Opnd& o, int ch) {
    OpCoef k;
    k.a = o.a ? __ldg(o.a + ch) : 1.f;
    k.b = o.b ? __ldg(o.b + ch) : 0.f;
    k.c = o.c ? __ldg(o.c + ch) : 0.f;
    return k;
}
// off = ch*TV + t*V + v (element offset inside the sample)
template <typename T>
__device__ __forceinline__ float opnd_val(const Opnd& o, const OpCoef& k, int n, long long off) {
    float v = k.a * ldf<T>((const T*)o.p + (long long)n * o.pns + off) + k.c;
    if (o.q) v = fmaf(k.b, ldf<T>((const T*)o.q + (long long)n * o.qns + off), v);
    if (o.relu) v = fmaxf(v, 0.f);
    return v;
}

// ---- reductions -------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum of up to NV values per thread; result valid in thread 0.  scratch: NV*32 floats.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) scratch[i * 32 + w] = v[i];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float x = (lane < nw) ? scratch[i * 32 + lane] : 0.f;
            v[i] = warp_sum(x);
        }
    }
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace tamgcn
