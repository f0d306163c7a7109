// tconv9_pack.cuh — weight-tile layout of the V-padded temporal-convolution kernels (tconv9.cu), shared with the
// packing kernels of conv_tc2.cu (the tiles live behind the conv_tc2 tiles in the same wpack buffer).
//
// Role names: OC = channels on the MMA M dimension (forward: Cout, data gradient: Cin), IC = contraction channels.
// One block = one (128-channel M tile, 16-channel K chunk, tap): the A operand of one tcgen05.mma, MN-major (channels
// contiguous), SWIZZLE_128B: [64-channel half b][K row r (16)][128 B], 16-byte chunk index XOR (r & 7).
// Blocks are ordered [M tile][K chunk][tap]: the k blocks of a pipeline stage are contiguous (one bulk copy).
#pragma once
#include <cstddef>
#include <cstdint>

namespace tamgcn {

#define T9_BLK_BYTES 4096          // 128 channels x 16 K rows x 2 bytes
#define T9_MAXK 9

__host__ __device__ inline bool t9_eligible(int OC, int IC, int k) { return k >= 2 && k <= T9_MAXK && IC % 16 == 0 && OC >= 32; }
__host__ __device__ inline size_t t9_bytes(int OC, int IC, int k) {
    return t9_eligible(OC, IC, k) ? (size_t)((OC + 127) / 128) * (size_t)(IC / 16) * (size_t)k * T9_BLK_BYTES : 0;
}

#ifdef __CUDACC__
// which = 0: forward tiles  A[m = co][kk = ci], tap j      = W[co][ci*k + j]
// which = 1: dgrad tiles    A[m = ci][kk = co], tap j      = W[co][ci*k + (k-1-j)]      (taps flipped)
// one unit = one 16-byte chunk (8 consecutive channels of one K row)
__device__ __forceinline__ void t9_pack_region(const float* __restrict__ W, int Cout, int Cin, int k, int which,
                                               uint4* __restrict__ dst, long long t0, long long tstride) {
    const int OC = which ? Cin : Cout, IC = which ? Cout : Cin;
    if (!t9_eligible(OC, IC, k)) return;
    const int n_kc = IC / 16, CK = Cin * k;
    const long long units = (long long)((OC + 127) / 128) * n_kc * k * (T9_BLK_BYTES / 16);
    for (long long u = t0; u < units; u += tstride) {
        const int within = (int)(u & 255);                    // chunk inside the 4096-byte block
        long long blk = u >> 8;
        const int tap = (int)(blk % k); blk /= k;
        const int kc = (int)(blk % n_kc);
        const int mt = (int)(blk / n_kc);
        const int b = within >> 7, r = (within >> 3) & 15, cs = within & 7;
        const int c8 = cs ^ (r & 7);                          // logical chunk stored at swizzled slot cs
        const int m0 = mt * 128 + b * 64 + c8 * 8, kk = kc * 16 + r;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int m = m0 + e;
            float w = 0.f;
            if (m < OC) w = which ? __ldg(W + (long long)kk * CK + m * k + (k - 1 - tap)) : __ldg(W + (long long)m * CK + kk * k + tap);
            f[e] = w;
        }
        uint4 o;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]), p1 = __floats2bfloat162_rn(f[2], f[3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]), p3 = __floats2bfloat162_rn(f[6], f[7]);
        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
        dst[u] = o;
    }
}
#endif

}  // namespace tamgcn
