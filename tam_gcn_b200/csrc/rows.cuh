// rows.cuh — vectorised loads/stores of one V-joint row (V = 20 or 25) into registers.
#pragma once
#include "common.cuh"

namespace tamgcn {

// ---- row loads -----------------------------------------------------------------------------------
template <typename T, int V, int VP>
__device__ __forceinline__ void load_row(const T* __restrict__ p, float (&r)[VP]);

template <> __device__ __forceinline__ void load_row<float, 20, 20>(const float* __restrict__ p, float (&r)[20]) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float4 a = __ldg(p4 + i);
        r[4 * i] = a.x; r[4 * i + 1] = a.y; r[4 * i + 2] = a.z; r[4 * i + 3] = a.w;
    }
}
template <> __device__ __forceinline__ void load_row<bf16, 20, 20>(const bf16* __restrict__ p, float (&r)[20]) {
    const uint2* p2 = reinterpret_cast<const uint2*>(p);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const uint2 a = __ldg(p2 + i);
        r[4 * i] = __uint_as_float(a.x << 16);
        r[4 * i + 1] = __uint_as_float(a.x & 0xffff0000u);
        r[4 * i + 2] = __uint_as_float(a.y << 16);
        r[4 * i + 3] = __uint_as_float(a.y & 0xffff0000u);
    }
}
template <> __device__ __forceinline__ void load_row<float, 25, 28>(const float* __restrict__ p, float (&r)[28]) {
#pragma unroll
    for (int i = 0; i < 25; ++i) r[i] = __ldg(p + i);
    r[25] = r[26] = r[27] = 0.f;
}
template <> __device__ __forceinline__ void load_row<bf16, 25, 28>(const bf16* __restrict__ p, float (&r)[28]) {
#pragma unroll
    for (int i = 0; i < 25; ++i) r[i] = __bfloat162float(__ldg(p + i));
    r[25] = r[26] = r[27] = 0.f;
}

template <typename T, int V>
__device__ __forceinline__ void store_row(T* __restrict__ p, const float* r);
template <> __device__ __forceinline__ void store_row<float, 20>(float* __restrict__ p, const float* r) {
    float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int i = 0; i < 5; ++i) p4[i] = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
}
template <> __device__ __forceinline__ void store_row<bf16, 20>(bf16* __restrict__ p, const float* r) {
    uint2* p2 = reinterpret_cast<uint2*>(p);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(r[4 * i], r[4 * i + 1]);
        __nv_bfloat162 hi = __floats2bfloat162_rn(r[4 * i + 2], r[4 * i + 3]);
        uint2 o;
        o.x = *reinterpret_cast<unsigned*>(&lo);
        o.y = *reinterpret_cast<unsigned*>(&hi);
        p2[i] = o;
    }
}
template <> __device__ __forceinline__ void store_row<float, 25>(float* __restrict__ p, const float* r) {
#pragma unroll
    for (int i = 0; i < 25; ++i) p[i] = r[i];
}
template <> __device__ __forceinline__ void store_row<bf16, 25>(bf16* __restrict__ p, const float* r) {
#pragma unroll
    for (int i = 0; i < 25; ++i) p[i] = __float2bfloat16_rn(r[i]);
}

template <int V> struct VPad { static const int VP = (V + 3) & ~3; static const int DP = V | 1; };


}  // namespace tamgcn
