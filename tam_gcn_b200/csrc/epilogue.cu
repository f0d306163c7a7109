// epilogue.cu — fused BatchNorm-apply / tanh / ReLU / residual epilogues and their backward passes,
// plus the (3x1) max-pool branch of MultiScale_TemporalConv.
//
// Every kernel here is a single streaming pass over (N, C, T*V) activations: HBM-bound.  A CTA owns
// one channel (so the BatchNorm coefficients are scalars in registers and the BatchNorm-backward
// sums stay in registers until one fp64 atomic per CTA) and strides over samples.
//
// reference: models/ctrgcn.py:255-261 (unit_gcn tail), :113-119 (max-pool branch), :145-146 (cat + res),
//            :283 (TCN_GCN_unit tail); models/stgcn.py:98-99 (st_gcn tail).
#include "common.cuh"
#include <cstdlib>
#include <initializer_list>

namespace tamgcn {

// grid = (C, NG): CTA (c, j) handles samples j, j+NG, ...
static inline dim3 ew_grid(int N, int C) {
    int ng = (148 * 8 + C - 1) / C;
    if (ng > N) ng = N;
    if (ng < 1) ng = 1;
    if (ng > 65535) ng = 65535;
    return dim3(C, ng);
}

template <int NV>
__device__ __forceinline__ void flush_stats(float (&v)[NV], double* const (&dst)[NV], int c) {
    __shared__ float scratch[NV * 32];
    block_sum<NV>(v, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (dst[i]) atomicAdd(dst[i] + c, (double)v[i]);
    }
}

// out = relu( sg*y0+hg + tanh(so*z+ho) + res )
template <typename T>
__global__ void __launch_bounds__(256)
gcn_epilogue_fwd_kernel(int N, int C, int TV, const T* __restrict__ y0, const float* __restrict__ sg,
                        const float* __restrict__ hg, const T* __restrict__ z, const float* __restrict__ so,
                        const float* __restrict__ ho, int res_mode, const T* __restrict__ r, long long rns,
                        const float* __restrict__ sr, const float* __restrict__ hr, T* __restrict__ out) {
    const int c = blockIdx.x;
    const float a_g = sg[c], b_g = hg[c], a_o = so[c], b_o = ho[c];
    const float a_r = (res_mode == TAMGCN_RES_AFFINE) ? sr[c] : 1.f;
    const float b_r = (res_mode == TAMGCN_RES_AFFINE) ? hr[c] : 0.f;
    for (int n = blockIdx.y; n < N; n += gridDim.y) {
        const long long base = ((long long)n * C + c) * TV;
        const T* pr = (res_mode != TAMGCN_RES_NONE) ? r + (long long)n * rns + (long long)c * TV : nullptr;
        for (int e = threadIdx.x; e < TV; e += blockDim.x) {
            float v = fmaf(a_g, ldf<T>(y0 + base + e), b_g) + tanhf(fmaf(a_o, ldf<T>(z + base + e), b_o));
            if (pr) v += fmaf(a_r, ldf<T>(pr + e), b_r);
            stf<T>(out + base + e, fmaxf(v, 0.f));
        }
    }
}

// G = g*[out>0];  DZ = G*(1-o^2), o = tanh(so*z+ho);  s1o += sum DZ;  s2o += sum DZ*z
template <typename T>
__global__ void __launch_bounds__(256)
gcn_epilogue_bwd_kernel(int N, int C, int TV, const T* __restrict__ g, const T* __restrict__ out,
                        const T* __restrict__ z, const float* __restrict__ so, const float* __restrict__ ho,
                        T* __restrict__ G, T* __restrict__ DZ, double* s1o, double* s2o) {
    const int c = blockIdx.x;
    const float a_o = so[c], b_o = ho[c];
    float acc[2] = {0.f, 0.f};
    for (int n = blockIdx.y; n < N; n += gridDim.y) {
        const long long base = ((long long)n * C + c) * TV;
        for (int e = threadIdx.x; e < TV; e += blockDim.x) {
            const float gv = (ldf<T>(out + base + e) > 0.f) ? ldf<T>(g + base + e) : 0.f;
            const float zv = ldf<T>(z + base + e);
            const float o = tanhf(fmaf(a_o, zv, b_o));
            const float dz = rnd<T>(gv * (1.f - o * o));
            stf<T>(G + base + e, gv);
            stf<T>(DZ + base + e, dz);
            acc[0] += dz;
            acc[1] = fmaf(dz, zv, acc[1]);
        }
    }
    double* const dst[2] = {s1o, s2o};
    flush_stats<2>(acc, dst, c);
}

// DY = G - DD (in place over G);  DR = G + DD (+ extra) -> dr;  BN-backward sums for bn (y0) and down.bn (r)
template <typename T>
__global__ void __launch_bounds__(256)
gcn_mid_bwd_kernel(int N, int C, int TV, T* __restrict__ G, const T* __restrict__ DD, T* __restrict__ dr,
                   long long drns, const T* __restrict__ y0, const T* __restrict__ r, long long rns, double* s1g,
                   double* s2g, double* s1d, double* s2d, const T* __restrict__ extra, long long exns) {
    const int c = blockIdx.x;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int n = blockIdx.y; n < N; n += gridDim.y) {
        const long long base = ((long long)n * C + c) * TV;
        T* pdr = dr ? dr + (long long)n * drns + (long long)c * TV : nullptr;
        const T* pr = r ? r + (long long)n * rns + (long long)c * TV : nullptr;
        for (int e = threadIdx.x; e < TV; e += blockDim.x) {
            const float gv = ldf<T>(G + base + e), dd = ldf<T>(DD + base + e);
            const float ex = extra ? ldf<T>(extra + (long long)n * exns + (long long)c * TV + e) : 0.f;
            const float dy = rnd<T>(gv - dd), drv = rnd<T>(gv + dd + ex);
            stf<T>(G + base + e, dy);
            acc[0] += dy;
            acc[1] = fmaf(dy, ldf<T>(y0 + base + e), acc[1]);
            if (pdr) stf<T>(pdr + e, drv);
            if (pr) {
                acc[2] += drv;
                acc[3] = fmaf(drv, ldf<T>(pr + e), acc[3]);
            }
        }
    }
    double* const dst[4] = {s1g, s2g, s1d, s2d};
    flush_stats<4>(acc, dst, c);
}

// out = f( su*u+hu + res )
template <typename T>
__global__ void __launch_bounds__(256)
tcn_epilogue_fwd_kernel(int N, int C, int TV, const T* __restrict__ u, long long uns, const float* __restrict__ su,
                        const float* __restrict__ hu, int res_mode, const T* __restrict__ r, long long rns,
                        const float* __restrict__ sr, const float* __restrict__ hr, int relu, T* __restrict__ out) {
    const int c = blockIdx.x;
    const float a_u = su[c], b_u = hu[c];
    const float a_r = (res_mode == TAMGCN_RES_AFFINE) ? sr[c] : 1.f;
    const float b_r = (res_mode == TAMGCN_RES_AFFINE) ? hr[c] : 0.f;
    for (int n = blockIdx.y; n < N; n += gridDim.y) {
        const long long base = ((long long)n * C + c) * TV;
        const T* pu = u + (long long)n * uns + (long long)c * TV;
        const T* pr = (res_mode != TAMGCN_RES_NONE) ? r + (long long)n * rns + (long long)c * TV : nullptr;
        for (int e = threadIdx.x; e < TV; e += blockDim.x) {
            float v = fmaf(a_u, ldf<T>(pu + e), b_u);
            if (pr) v += fmaf(a_r, ldf<T>(pr + e), b_r);
            if (relu) v = fmaxf(v, 0.f);
            stf<T>(out + base + e, v);
        }
    }
}

// G = relu ? g*[out>0] : g;  s1 += sum G;  s2u += sum G*u;  s2r += sum G*r
template <typename T>
__global__ void __launch_bounds__(256)
tcn_epilogue_bwd_kernel(int N, int C, int TV, const T* __restrict__ g, const T* __restrict__ out, int relu,
                        const T* __restrict__ u, long long uns, const T* __restrict__ r, long long rns,
                        T* __restrict__ G, double* s1, double* s2u, double* s2r) {
    const int c = blockIdx.x;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int n = blockIdx.y; n < N; n += gridDim.y) {
        const long long base = ((long long)n * C + c) * TV;
        const T* pu = u + (long long)n * uns + (long long)c * TV;
        const T* pr = r ? r + (long long)n * rns + (long long)c * TV : nullptr;
        for (int e = threadIdx.x; e < TV; e += blockDim.x) {
            float gv = ldf<T>(g + base + e);
            if (relu && !(ldf<T>(out + base + e) > 0.f)) gv = 0.f;
            if (G) stf<T>(G + base + e, gv);
            acc[0] += gv;
            acc[1] = fmaf(gv, ldf<T>(pu + e), acc[1]);
            if (pr) acc[2] = fmaf(gv, ldf<T>(pr + e), acc[2]);
        }
    }
    double* const dst[3] = {s1, s2u, s2r};
    flush_stats<3>(acc, dst, c);
}

// ------------------------------------------------------------------------------------------------
// bf16 vector variants of the five streaming kernels above: VEC (8 or 4) elements per 16- / 8-byte access, the
// samples of a CTA walked as one flat index space (a (T*V)/VEC-vector plane is smaller than the CTA at the late
// layers).  Same arithmetic and rounding points as the element-wise kernels; tanh is the hardware approximation
// (2^-11 relative, below the bf16 rounding of the stored result).  Selected by the host when every pointer and
// stride allows the vector width.
// ------------------------------------------------------------------------------------------------
template <int VEC> struct BVec { uint32_t w[VEC / 2]; };
template <int VEC> __device__ __forceinline__ BVec<VEC> bv_ld(const bf16* p) {
    BVec<VEC> r;
    if (VEC == 8) { const uint4 u = __ldg(reinterpret_cast<const uint4*>(p)); r.w[0] = u.x; r.w[1] = u.y; r.w[VEC / 2 - 2] = u.z; r.w[VEC / 2 - 1] = u.w; }
    else { const uint2 u = __ldg(reinterpret_cast<const uint2*>(p)); r.w[0] = u.x; r.w[1] = u.y; }
    return r;
}
template <int VEC> __device__ __forceinline__ void bv_st(bf16* p, const BVec<VEC>& r) {
    if (VEC == 8) *reinterpret_cast<uint4*>(p) = make_uint4(r.w[0], r.w[1], r.w[VEC / 2 - 2], r.w[VEC / 2 - 1]);
    else *reinterpret_cast<uint2*>(p) = make_uint2(r.w[0], r.w[1]);
}
__device__ __forceinline__ float bv_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bv_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t bv_pack(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float tanh_hw(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// flat walk over the (sample, vector) pairs of CTA (c, blockIdx.y): sample n = blockIdx.y + k * gridDim.y, vector e
#define BV_WALK_BEGIN(N_, TVv_)                                                                   \
    const int nmine_ = ((N_) - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;          \
    int k_ = 0, e_ = threadIdx.x;                                                                 \
    while (e_ >= (TVv_)) { e_ -= (TVv_); ++k_; }                                                  \
    while (k_ < nmine_) {                                                                         \
        const int n = blockIdx.y + k_ * gridDim.y;                                                \
        const int e = e_;
#define BV_WALK_END(TVv_)                                                                         \
        e_ += blockDim.x;                                                                         \
        while (e_ >= (TVv_)) { e_ -= (TVv_); ++k_; }                                              \
    }

template <int VEC>
__global__ void __launch_bounds__(256)
gcn_epilogue_fwd_vec_kernel(int N, int C, int TV, const bf16* __restrict__ y0, const float* __restrict__ sg,
                            const float* __restrict__ hg, const bf16* __restrict__ z, const float* __restrict__ so,
                            const float* __restrict__ ho, int res_mode, const bf16* __restrict__ r, long long rns,
                            const float* __restrict__ sr, const float* __restrict__ hr, bf16* __restrict__ out) {
    const int c = blockIdx.x, TVv = TV / VEC;
    const float a_g = sg[c], b_g = hg[c], a_o = so[c], b_o = ho[c];
    const float a_r = (res_mode == TAMGCN_RES_AFFINE) ? sr[c] : 1.f;
    const float b_r = (res_mode == TAMGCN_RES_AFFINE) ? hr[c] : 0.f;
    BV_WALK_BEGIN(N, TVv)
        const long long off = ((long long)n * C + c) * TV + (long long)e * VEC;
        const BVec<VEC> vy = bv_ld<VEC>(y0 + off), vz = bv_ld<VEC>(z + off);
        BVec<VEC> vr, vo;
        if (res_mode != TAMGCN_RES_NONE) vr = bv_ld<VEC>(r + (long long)n * rns + (long long)c * TV + (long long)e * VEC);
#pragma unroll
        for (int j = 0; j < VEC / 2; ++j) {
            float lo = fmaf(a_g, bv_lo(vy.w[j]), b_g) + tanh_hw(fmaf(a_o, bv_lo(vz.w[j]), b_o));
            float hi = fmaf(a_g, bv_hi(vy.w[j]), b_g) + tanh_hw(fmaf(a_o, bv_hi(vz.w[j]), b_o));
            if (res_mode != TAMGCN_RES_NONE) { lo += fmaf(a_r, bv_lo(vr.w[j]), b_r); hi += fmaf(a_r, bv_hi(vr.w[j]), b_r); }
            vo.w[j] = bv_pack(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
        }
        bv_st<VEC>(out + off, vo);
    BV_WALK_END(TVv)
}

template <int VEC>
__global__ void __launch_bounds__(256)
gcn_epilogue_bwd_vec_kernel(int N, int C, int TV, const bf16* __restrict__ g, const bf16* __restrict__ out,
                            const bf16* __restrict__ z, const float* __restrict__ so, const float* __restrict__ ho,
                            bf16* __restrict__ G, bf16* __restrict__ DZ, double* s1o, double* s2o) {
    const int c = blockIdx.x, TVv = TV / VEC;
    const float a_o = so[c], b_o = ho[c];
    float acc[2] = {0.f, 0.f};
    BV_WALK_BEGIN(N, TVv)
        const long long off = ((long long)n * C + c) * TV + (long long)e * VEC;
        const BVec<VEC> vg = bv_ld<VEC>(g + off), vout = bv_ld<VEC>(out + off), vz = bv_ld<VEC>(z + off);
        BVec<VEC> oG, oD;
#pragma unroll
        for (int j = 0; j < VEC / 2; ++j) {
            const float g0 = bv_lo(vout.w[j]) > 0.f ? bv_lo(vg.w[j]) : 0.f, g1 = bv_hi(vout.w[j]) > 0.f ? bv_hi(vg.w[j]) : 0.f;
            const float z0 = bv_lo(vz.w[j]), z1 = bv_hi(vz.w[j]);
            const float o0 = tanh_hw(fmaf(a_o, z0, b_o)), o1 = tanh_hw(fmaf(a_o, z1, b_o));
            oG.w[j] = bv_pack(g0, g1);
            oD.w[j] = bv_pack(g0 * (1.f - o0 * o0), g1 * (1.f - o1 * o1));
            const float d0 = bv_lo(oD.w[j]), d1 = bv_hi(oD.w[j]);       // the rounded values, as stored
            acc[0] += d0 + d1;
            acc[1] = fmaf(d0, z0, fmaf(d1, z1, acc[1]));
        }
        bv_st<VEC>(G + off, oG);
        bv_st<VEC>(DZ + off, oD);
    BV_WALK_END(TVv)
    double* const dst[2] = {s1o, s2o};
    flush_stats<2>(acc, dst, c);
}

template <int VEC>
__global__ void __launch_bounds__(256)
gcn_mid_bwd_vec_kernel(int N, int C, int TV, bf16* __restrict__ G, const bf16* __restrict__ DD, bf16* __restrict__ dr,
                       long long drns, const bf16* __restrict__ y0, const bf16* __restrict__ r, long long rns, double* s1g,
                       double* s2g, double* s1d, double* s2d, const bf16* __restrict__ extra, long long exns) {
    const int c = blockIdx.x, TVv = TV / VEC;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    BV_WALK_BEGIN(N, TVv)
        const long long off = ((long long)n * C + c) * TV + (long long)e * VEC;
        const BVec<VEC> vg = bv_ld<VEC>(G + off), vd = bv_ld<VEC>(DD + off), vy = bv_ld<VEC>(y0 + off);
        BVec<VEC> vr, vx, oY, oR;
        if (r) vr = bv_ld<VEC>(r + (long long)n * rns + (long long)c * TV + (long long)e * VEC);
        if (extra) vx = bv_ld<VEC>(extra + (long long)n * exns + (long long)c * TV + (long long)e * VEC);
#pragma unroll
        for (int j = 0; j < VEC / 2; ++j) {
            const float g0 = bv_lo(vg.w[j]), g1 = bv_hi(vg.w[j]), d0 = bv_lo(vd.w[j]), d1 = bv_hi(vd.w[j]);
            const float x0 = extra ? bv_lo(vx.w[j]) : 0.f, x1 = extra ? bv_hi(vx.w[j]) : 0.f;
            oY.w[j] = bv_pack(g0 - d0, g1 - d1);
            oR.w[j] = bv_pack(g0 + d0 + x0, g1 + d1 + x1);
            const float y0v = bv_lo(oY.w[j]), y1v = bv_hi(oY.w[j]);
            acc[0] += y0v + y1v;
            acc[1] = fmaf(y0v, bv_lo(vy.w[j]), fmaf(y1v, bv_hi(vy.w[j]), acc[1]));
            if (r) {
                const float r0 = bv_lo(oR.w[j]), r1 = bv_hi(oR.w[j]);
                acc[2] += r0 + r1;
                acc[3] = fmaf(r0, bv_lo(vr.w[j]), fmaf(r1, bv_hi(vr.w[j]), acc[3]));
            }
        }
        bv_st<VEC>(G + off, oY);
        if (dr) bv_st<VEC>(dr + (long long)n * drns + (long long)c * TV + (long long)e * VEC, oR);
    BV_WALK_END(TVv)
    double* const dst[4] = {s1g, s2g, s1d, s2d};
    flush_stats<4>(acc, dst, c);
}

template <int VEC>
__global__ void __launch_bounds__(256)
tcn_epilogue_fwd_vec_kernel(int N, int C, int TV, const bf16* __restrict__ u, long long uns, const float* __restrict__ su,
                            const float* __restrict__ hu, int res_mode, const bf16* __restrict__ r, long long rns,
                            const float* __restrict__ sr, const float* __restrict__ hr, int relu, bf16* __restrict__ out) {
    const int c = blockIdx.x, TVv = TV / VEC;
    const float a_u = su[c], b_u = hu[c];
    const float a_r = (res_mode == TAMGCN_RES_AFFINE) ? sr[c] : 1.f;
    const float b_r = (res_mode == TAMGCN_RES_AFFINE) ? hr[c] : 0.f;
    BV_WALK_BEGIN(N, TVv)
        const long long co = (long long)c * TV + (long long)e * VEC;
        const BVec<VEC> vu = bv_ld<VEC>(u + (long long)n * uns + co);
        BVec<VEC> vr, vo;
        if (res_mode != TAMGCN_RES_NONE) vr = bv_ld<VEC>(r + (long long)n * rns + co);
#pragma unroll
        for (int j = 0; j < VEC / 2; ++j) {
            float lo = fmaf(a_u, bv_lo(vu.w[j]), b_u), hi = fmaf(a_u, bv_hi(vu.w[j]), b_u);
            if (res_mode != TAMGCN_RES_NONE) { lo += fmaf(a_r, bv_lo(vr.w[j]), b_r); hi += fmaf(a_r, bv_hi(vr.w[j]), b_r); }
            if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
            vo.w[j] = bv_pack(lo, hi);
        }
        bv_st<VEC>(out + (long long)n * C * TV + co, vo);
    BV_WALK_END(TVv)
}

template <int VEC>
__global__ void __launch_bounds__(256)
tcn_epilogue_bwd_vec_kernel(int N, int C, int TV, const bf16* __restrict__ g, const bf16* __restrict__ out, int relu,
                            const bf16* __restrict__ u, long long uns, const bf16* __restrict__ r, long long rns,
                            bf16* __restrict__ G, double* s1, double* s2u, double* s2r) {
    const int c = blockIdx.x, TVv = TV / VEC;
    float acc[3] = {0.f, 0.f, 0.f};
    BV_WALK_BEGIN(N, TVv)
        const long long co = (long long)c * TV + (long long)e * VEC;
        const long long off = (long long)n * C * TV + co;
        BVec<VEC> vg = bv_ld<VEC>(g + off), vo, vr;
        const BVec<VEC> vu = bv_ld<VEC>(u + (long long)n * uns + co);
        if (relu) vo = bv_ld<VEC>(out + off);
        if (r) vr = bv_ld<VEC>(r + (long long)n * rns + co);
#pragma unroll
        for (int j = 0; j < VEC / 2; ++j) {
            float g0 = bv_lo(vg.w[j]), g1 = bv_hi(vg.w[j]);
            if (relu) {
                if (!(bv_lo(vo.w[j]) > 0.f)) g0 = 0.f;
                if (!(bv_hi(vo.w[j]) > 0.f)) g1 = 0.f;
                vg.w[j] = bv_pack(g0, g1);
            }
            acc[0] += g0 + g1;
            acc[1] = fmaf(g0, bv_lo(vu.w[j]), fmaf(g1, bv_hi(vu.w[j]), acc[1]));
            if (r) acc[2] = fmaf(g0, bv_lo(vr.w[j]), fmaf(g1, bv_hi(vr.w[j]), acc[2]));
        }
        if (G) bv_st<VEC>(G + off, vg);
    BV_WALK_END(TVv)
    double* const dst[3] = {s1, s2u, s2r};
    flush_stats<3>(acc, dst, c);
}

// lazy operand on 4 consecutive bf16 elements (8-byte accesses): f(a*p + b*q + c)
__device__ __forceinline__ void bv_opnd4(const Opnd& o, const OpCoef& k, int n, long long off, float (&v)[4]) {
    const BVec<4> p = bv_ld<4>((const bf16*)o.p + (long long)n * o.pns + off);
    v[0] = k.a * bv_lo(p.w[0]) + k.c; v[1] = k.a * bv_hi(p.w[0]) + k.c;
    v[2] = k.a * bv_lo(p.w[1]) + k.c; v[3] = k.a * bv_hi(p.w[1]) + k.c;
    if (o.q) {
        const BVec<4> q = bv_ld<4>((const bf16*)o.q + (long long)n * o.qns + off);
        v[0] = fmaf(k.b, bv_lo(q.w[0]), v[0]); v[1] = fmaf(k.b, bv_hi(q.w[0]), v[1]);
        v[2] = fmaf(k.b, bv_lo(q.w[1]), v[2]); v[3] = fmaf(k.b, bv_hi(q.w[1]), v[3]);
    }
    if (o.relu) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    }
}

// max-pool forward, 4 joints per thread (V % 4 == 0): three 8-byte row loads per output vector
__global__ void __launch_bounds__(256)
maxpool_fwd_vec_kernel(int N, int C, int Tn, int To, int V, int s, Opnd x, bf16* __restrict__ y, long long yns, double* ssum,
                       double* ssq) {
    const int c = blockIdx.x, VQ = V / 4, TVv = To * VQ;
    const OpCoef cf = opnd_coef(x, c);
    float acc[2] = {0.f, 0.f};
    BV_WALK_BEGIN(N, TVv)
        const int to = e / VQ, vq = e - to * VQ;
        float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) {
            const int t = to * s - 1 + dt;
            if (t >= 0 && t < Tn) {
                float v[4];
                bv_opnd4(x, cf, n, ((long long)c * Tn + t) * V + 4 * vq, v);
#pragma unroll
                for (int j = 0; j < 4; ++j) best[j] = fmaxf(best[j], v[j]);
            }
        }
        BVec<4> o;
        o.w[0] = bv_pack(best[0], best[1]); o.w[1] = bv_pack(best[2], best[3]);
        bv_st<4>(y + (long long)n * yns + ((long long)c * To + to) * V + 4 * vq, o);
        const float r0 = bv_lo(o.w[0]), r1 = bv_hi(o.w[0]), r2 = bv_lo(o.w[1]), r3 = bv_hi(o.w[1]);
        acc[0] += (r0 + r1) + (r2 + r3);
        acc[1] = fmaf(r0, r0, fmaf(r1, r1, fmaf(r2, r2, fmaf(r3, r3, acc[1]))));
    BV_WALK_END(TVv)
    double* const dst[2] = {ssum, ssq};
    flush_stats<2>(acc, dst, c);
}

// max-pool backward, 4 joints per thread: the five input rows t-2..t+2 that any window containing t can touch are
// loaded once (8 bytes each), then the (first) arg-max of each window is taken per joint exactly as above
__global__ void __launch_bounds__(256)
maxpool_bwd_vec_kernel(int N, int C, int Tn, int To, int V, int s, Opnd dy, Opnd x, bf16* __restrict__ dh, long long dhns,
                       double* s1, double* s2) {
    const int c = blockIdx.x, VQ = V / 4, TVv = Tn * VQ;
    const OpCoef cf = opnd_coef(x, c), df = opnd_coef(dy, c);
    float acc[2] = {0.f, 0.f};
    BV_WALK_BEGIN(N, TVv)
        const int t = e / VQ, vq = e - t * VQ;
        const long long xoff = (long long)c * Tn * V + 4 * vq;
        float xv[5][4];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const int tt = t - 2 + r;
            if (tt >= 0 && tt < Tn) bv_opnd4(x, cf, n, xoff + (long long)tt * V, xv[r]);
            else { xv[r][0] = xv[r][1] = xv[r][2] = xv[r][3] = -INFINITY; }
        }
        const BVec<4> pw = bv_ld<4>((const bf16*)x.p + (long long)n * x.pns + xoff + (long long)t * V);
        const float pv[4] = {bv_lo(pw.w[0]), bv_hi(pw.w[0]), bv_lo(pw.w[1]), bv_hi(pw.w[1])};
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        // a window containing t starts at row t-2, t-1 or t (xv rows r0 = 0, 1, 2): compile-time register indices
#pragma unroll
        for (int r0 = 0; r0 < 3; ++r0) {
            const int tt0 = t - 2 + r0;                        // first row of the window = to * s - 1
            if ((tt0 + 1) % s != 0) continue;
            const int to = (tt0 + 1) / s;
            if (tt0 + 1 < 0 || to >= To) continue;
            float dyv[4];
            bv_opnd4(dy, df, n, ((long long)c * To + to) * V + 4 * vq, dyv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float best = -INFINITY;
                int arg = -1;
#pragma unroll
                for (int dt = 0; dt < 3; ++dt) {
                    const int tt = tt0 + dt;
                    const float val = xv[r0 + dt][j];
                    if (tt >= 0 && tt < Tn && val > best) { best = val; arg = tt; }
                }
                if (arg == t) d[j] += dyv[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (!(fmaf(cf.a, pv[j], cf.c) > 0.f)) d[j] = 0.f;
        BVec<4> o;
        o.w[0] = bv_pack(d[0], d[1]); o.w[1] = bv_pack(d[2], d[3]);
        bv_st<4>(dh + (long long)n * dhns + xoff + (long long)t * V, o);
        const float q0 = bv_lo(o.w[0]), q1 = bv_hi(o.w[0]), q2 = bv_lo(o.w[1]), q3 = bv_hi(o.w[1]);
        acc[0] += (q0 + q1) + (q2 + q3);
        acc[1] = fmaf(q0, pv[0], fmaf(q1, pv[1], fmaf(q2, pv[2], fmaf(q3, pv[3], acc[1]))));
    BV_WALK_END(TVv)
    double* const dst[2] = {s1, s2};
    flush_stats<2>(acc, dst, c);
}

// vector width usable for bf16 rows of TV elements: every pointer 2*VEC-byte aligned, every sample stride a multiple of VEC
static inline int bv_width(int TV, std::initializer_list<const void*> ptrs, std::initializer_list<long long> strides) {
    for (int v = 8; v >= 4; v >>= 1) {
        bool ok = TV % v == 0;
        for (const void* q : ptrs) ok = ok && (q == nullptr || (reinterpret_cast<uintptr_t>(q) & (uintptr_t)(2 * v - 1)) == 0);
        for (long long st : strides) ok = ok && (st % v == 0);
        if (ok) return v;
    }
    return 1;
}
static inline bool bv_disabled() {
    static const int v = [] { const char* e = getenv("TAMGCN_DISABLE_VEC_EPI"); return (e && e[0] == '1') ? 1 : 0; }();
    return v == 1;
}

// MaxPool2d((3,1), stride (s,1), padding (1,0)) over the lazily transformed input
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(int N, int C, int Tn, int To, int V, int s, Opnd x, T* __restrict__ y, long long yns, double* ssum,
                   double* ssq) {
    const int c = blockIdx.x;
    const OpCoef cf = opnd_coef(x, c);
    float acc[2] = {0.f, 0.f};
    for (int n = blockIdx.y; n < N; n += gridDim.y) {
        T* py = y + (long long)n * yns + (long long)c * To * V;
        for (int e = threadIdx.x; e < To * V; e += blockDim.x) {
            const int to = e / V, v = e - to * V;
            float best = -INFINITY;
#pragma unroll
            for (int dt = 0; dt < 3; ++dt) {
                const int t = to * s - 1 + dt;
                if (t >= 0 && t < Tn) best = fmaxf(best, opnd_val<T>(x, cf, n, ((long long)c * Tn + t) * V + v));
            }
            best = rnd<T>(best);
            stf<T>(py + e, best);
            acc[0] += best;
            acc[1] = fmaf(best, best, acc[1]);
        }
    }
    double* const dst[2] = {ssum, ssq};
    flush_stats<2>(acc, dst, c);
}

// dh[t] = [x.a*P+x.c > 0] * sum over windows `to` whose FIRST arg-max is t of dY(to)
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(int N, int C, int Tn, int To, int V, int s, Opnd dy, Opnd x, T* __restrict__ dh, long long dhns,
                   double* s1, double* s2) {
    const int c = blockIdx.x;
    const OpCoef cf = opnd_coef(x, c), df = opnd_coef(dy, c);
    float acc[2] = {0.f, 0.f};
    for (int n = blockIdx.y; n < N; n += gridDim.y) {
        T* pd = dh + (long long)n * dhns + (long long)c * Tn * V;
        for (int e = threadIdx.x; e < Tn * V; e += blockDim.x) {
            const int t = e / V, v = e - t * V;
            const long long xoff = (long long)c * Tn * V;
            const float pv = ldf<T>((const T*)x.p + (long long)n * x.pns + xoff + e);
            float d = 0.f;
            if (fmaf(cf.a, pv, cf.c) > 0.f) {
                // windows containing t: to*s-1 <= t <= to*s+1
                int to_lo = (t - 1 + s - 1) / s;  // ceil((t-1)/s), t-1 >= -1
                if (t - 1 < 0) to_lo = 0;
                const int to_hi = min(To - 1, (t + 1) / s);
                for (int to = to_lo; to <= to_hi; ++to) {
                    float best = -INFINITY;
                    int arg = -1;
#pragma unroll
                    for (int dt = 0; dt < 3; ++dt) {
                        const int tt = to * s - 1 + dt;
                        if (tt >= 0 && tt < Tn) {
                            const float val = opnd_val<T>(x, cf, n, xoff + (long long)tt * V + v);
                            if (val > best) { best = val; arg = tt; }
                        }
                    }
                    if (arg == t) d += opnd_val<T>(dy, df, n, ((long long)c * To + to) * V + v);
                }
            }
            d = rnd<T>(d);
            stf<T>(pd + e, d);
            acc[0] += d;
            acc[1] = fmaf(d, pv, acc[1]);
        }
    }
    double* const dst[2] = {s1, s2};
    flush_stats<2>(acc, dst, c);
}

}  // namespace tamgcn

using namespace tamgcn;

#define DISPATCH(dtype, who, KERNEL, grid, st, ...)                                        \
    do {                                                                                   \
        if ((dtype) == TAMGCN_F32) KERNEL<float><<<grid, 256, 0, st>>>(__VA_ARGS__);       \
        else if ((dtype) == TAMGCN_BF16) KERNEL<bf16><<<grid, 256, 0, st>>>(__VA_ARGS__);  \
        else return set_error(who ": bad dtype %d", (dtype));                              \
        count_launch();                                                                    \
        return check_launch(who);                                                          \
    } while (0)

#define TP(T_, p) ((T_*)(p))

extern "C" int tamgcn_gcn_epilogue_fwd(int dtype, int N, int C, int TV, const void* y0, const float* sg,
                                       const float* hg, const void* z, const float* so, const float* ho, int res_mode,
                                       const void* r, int64_t r_nstride, const float* sr, const float* hr, void* out,
                                       tamgcn_stream stream) {
    TG_REQUIRE(N > 0 && C > 0 && TV > 0 && y0 && sg && hg && z && so && ho && out, "gcn_epilogue_fwd: bad arguments");
    TG_REQUIRE(res_mode == TAMGCN_RES_NONE || r, "gcn_epilogue_fwd: residual tensor missing");
    TG_REQUIRE(res_mode != TAMGCN_RES_AFFINE || (sr && hr), "gcn_epilogue_fwd: residual coefficients missing");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid = ew_grid(N, C);
    if (dtype == TAMGCN_BF16 && !bv_disabled()) {
        const int vw = bv_width(TV, {y0, z, r, out}, {(long long)r_nstride});
        if (vw > 1) {
            if (vw == 8) gcn_epilogue_fwd_vec_kernel<8><<<grid, 256, 0, st>>>(N, C, TV, (const bf16*)y0, sg, hg, (const bf16*)z, so, ho, res_mode, (const bf16*)r, r_nstride, sr, hr, (bf16*)out);
            else gcn_epilogue_fwd_vec_kernel<4><<<grid, 256, 0, st>>>(N, C, TV, (const bf16*)y0, sg, hg, (const bf16*)z, so, ho, res_mode, (const bf16*)r, r_nstride, sr, hr, (bf16*)out);
            count_launch();
            return check_launch("gcn_epilogue_fwd");
        }
    }
    if (dtype == TAMGCN_F32)
        gcn_epilogue_fwd_kernel<float><<<grid, 256, 0, st>>>(N, C, TV, (const float*)y0, sg, hg, (const float*)z, so, ho,
                                                             res_mode, (const float*)r, r_nstride, sr, hr, (float*)out);
    else if (dtype == TAMGCN_BF16)
        gcn_epilogue_fwd_kernel<bf16><<<grid, 256, 0, st>>>(N, C, TV, (const bf16*)y0, sg, hg, (const bf16*)z, so, ho,
                                                            res_mode, (const bf16*)r, r_nstride, sr, hr, (bf16*)out);
    else return set_error("gcn_epilogue_fwd: bad dtype %d", dtype);
    count_launch();
    return check_launch("gcn_epilogue_fwd");
}

extern "C" int tamgcn_gcn_epilogue_bwd(int dtype, int N, int C, int TV, const void* g, const void* out, const void* z,
                                       const float* so, const float* ho, void* G, void* DZ, double* s1o, double* s2o,
                                       tamgcn_stream stream) {
    TG_REQUIRE(N > 0 && C > 0 && TV > 0 && g && out && z && so && ho && G && DZ && s1o && s2o,
               "gcn_epilogue_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid = ew_grid(N, C);
    if (dtype == TAMGCN_BF16 && !bv_disabled()) {
        const int vw = bv_width(TV, {g, out, z, G, DZ}, {});
        if (vw > 1) {
            if (vw == 8) gcn_epilogue_bwd_vec_kernel<8><<<grid, 256, 0, st>>>(N, C, TV, (const bf16*)g, (const bf16*)out, (const bf16*)z, so, ho, (bf16*)G, (bf16*)DZ, s1o, s2o);
            else gcn_epilogue_bwd_vec_kernel<4><<<grid, 256, 0, st>>>(N, C, TV, (const bf16*)g, (const bf16*)out, (const bf16*)z, so, ho, (bf16*)G, (bf16*)DZ, s1o, s2o);
            count_launch();
            return check_launch("gcn_epilogue_bwd");
        }
    }
    if (dtype == TAMGCN_F32)
        gcn_epilogue_bwd_kernel<float><<<grid, 256, 0, st>>>(N, C, TV, (const float*)g, (const float*)out,
                                                             (const float*)z, so, ho, (float*)G, (float*)DZ, s1o, s2o);
    else if (dtype == TAMGCN_BF16)
        gcn_epilogue_bwd_kernel<bf16><<<grid, 256, 0, st>>>(N, C, TV, (const bf16*)g, (const bf16*)out, (const bf16*)z,
                                                            so, ho, (bf16*)G, (bf16*)DZ, s1o, s2o);
    else return set_error("gcn_epilogue_bwd: bad dtype %d", dtype);
    count_launch();
    return check_launch("gcn_epilogue_bwd");
}

extern "C" int tamgcn_gcn_mid_bwd(int dtype, int N, int C, int TV, void* G, const void* DD, void* dr,
                                  int64_t dr_nstride, const void* y0, const void* r, int64_t r_nstride, double* s1g,
                                  double* s2g, double* s1d, double* s2d, const void* extra, int64_t extra_nstride,
                                  tamgcn_stream stream) {
    TG_REQUIRE(N > 0 && C > 0 && TV > 0 && G && DD && y0 && s1g && s2g, "gcn_mid_bwd: bad arguments");
    TG_REQUIRE(!extra || (dr && !r), "gcn_mid_bwd: `extra` is added to the dr output of the identity-residual case only");
    TG_REQUIRE(!r || (s1d && s2d), "gcn_mid_bwd: residual BN sums missing");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid = ew_grid(N, C);
    if (dtype == TAMGCN_BF16 && !bv_disabled()) {
        const int vw = bv_width(TV, {G, DD, dr, y0, r, extra}, {(long long)dr_nstride, (long long)r_nstride, (long long)extra_nstride});
        if (vw > 1) {
            if (vw == 8) gcn_mid_bwd_vec_kernel<8><<<grid, 256, 0, st>>>(N, C, TV, (bf16*)G, (const bf16*)DD, (bf16*)dr, dr_nstride, (const bf16*)y0, (const bf16*)r, r_nstride, s1g, s2g, s1d, s2d, (const bf16*)extra, extra_nstride);
            else gcn_mid_bwd_vec_kernel<4><<<grid, 256, 0, st>>>(N, C, TV, (bf16*)G, (const bf16*)DD, (bf16*)dr, dr_nstride, (const bf16*)y0, (const bf16*)r, r_nstride, s1g, s2g, s1d, s2d, (const bf16*)extra, extra_nstride);
            count_launch();
            return check_launch("gcn_mid_bwd");
        }
    }
    if (dtype == TAMGCN_F32)
        gcn_mid_bwd_kernel<float><<<grid, 256, 0, st>>>(N, C, TV, (float*)G, (const float*)DD, (float*)dr, dr_nstride,
                                                        (const float*)y0, (const float*)r, r_nstride, s1g, s2g, s1d, s2d, (const float*)extra, extra_nstride);
    else if (dtype == TAMGCN_BF16)
        gcn_mid_bwd_kernel<bf16><<<grid, 256, 0, st>>>(N, C, TV, (bf16*)G, (const bf16*)DD, (bf16*)dr, dr_nstride,
                                                       (const bf16*)y0, (const bf16*)r, r_nstride, s1g, s2g, s1d, s2d, (const bf16*)extra, extra_nstride);
    else return set_error("gcn_mid_bwd: bad dtype %d", dtype);
    count_launch();
    return check_launch("gcn_mid_bwd");
}

extern "C" int tamgcn_tcn_epilogue_fwd(int dtype, int N, int C, int TV, const void* u, int64_t u_nstride,
                                       const float* su, const float* hu, int res_mode, const void* r,
                                       int64_t r_nstride, const float* sr, const float* hr, int relu, void* out,
                                       tamgcn_stream stream) {
    TG_REQUIRE(N > 0 && C > 0 && TV > 0 && u && su && hu && out, "tcn_epilogue_fwd: bad arguments");
    TG_REQUIRE(res_mode == TAMGCN_RES_NONE || r, "tcn_epilogue_fwd: residual tensor missing");
    TG_REQUIRE(res_mode != TAMGCN_RES_AFFINE || (sr && hr), "tcn_epilogue_fwd: residual coefficients missing");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid = ew_grid(N, C);
    if (dtype == TAMGCN_BF16 && !bv_disabled()) {
        const int vw = bv_width(TV, {u, r, out}, {(long long)u_nstride, (long long)r_nstride});
        if (vw > 1) {
            if (vw == 8) tcn_epilogue_fwd_vec_kernel<8><<<grid, 256, 0, st>>>(N, C, TV, (const bf16*)u, u_nstride, su, hu, res_mode, (const bf16*)r, r_nstride, sr, hr, relu, (bf16*)out);
            else tcn_epilogue_fwd_vec_kernel<4><<<grid, 256, 0, st>>>(N, C, TV, (const bf16*)u, u_nstride, su, hu, res_mode, (const bf16*)r, r_nstride, sr, hr, relu, (bf16*)out);
            count_launch();
            return check_launch("tcn_epilogue_fwd");
        }
    }
    if (dtype == TAMGCN_F32)
        tcn_epilogue_fwd_kernel<float><<<grid, 256, 0, st>>>(N, C, TV, (const float*)u, u_nstride, su, hu, res_mode,
                                                             (const float*)r, r_nstride, sr, hr, relu, (float*)out);
    else if (dtype == TAMGCN_BF16)
        tcn_epilogue_fwd_kernel<bf16><<<grid, 256, 0, st>>>(N, C, TV, (const bf16*)u, u_nstride, su, hu, res_mode,
                                                            (const bf16*)r, r_nstride, sr, hr, relu, (bf16*)out);
    else return set_error("tcn_epilogue_fwd: bad dtype %d", dtype);
    count_launch();
    return check_launch("tcn_epilogue_fwd");
}

extern "C" int tamgcn_tcn_epilogue_bwd(int dtype, int N, int C, int TV, const void* g, const void* out, int relu,
                                       const void* u, int64_t u_nstride, const void* r, int64_t r_nstride, void* G,
                                       double* s1, double* s2u, double* s2r, tamgcn_stream stream) {
    TG_REQUIRE(N > 0 && C > 0 && TV > 0 && g && u && s1 && s2u, "tcn_epilogue_bwd: bad arguments");
    TG_REQUIRE(!relu || (out && G), "tcn_epilogue_bwd: relu needs out and G");
    TG_REQUIRE(!r || s2r, "tcn_epilogue_bwd: residual BN sum missing");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid = ew_grid(N, C);
    if (dtype == TAMGCN_BF16 && !bv_disabled()) {
        const int vw = bv_width(TV, {g, out, u, r, G}, {(long long)u_nstride, (long long)r_nstride});
        if (vw > 1) {
            if (vw == 8) tcn_epilogue_bwd_vec_kernel<8><<<grid, 256, 0, st>>>(N, C, TV, (const bf16*)g, (const bf16*)out, relu, (const bf16*)u, u_nstride, (const bf16*)r, r_nstride, (bf16*)G, s1, s2u, s2r);
            else tcn_epilogue_bwd_vec_kernel<4><<<grid, 256, 0, st>>>(N, C, TV, (const bf16*)g, (const bf16*)out, relu, (const bf16*)u, u_nstride, (const bf16*)r, r_nstride, (bf16*)G, s1, s2u, s2r);
            count_launch();
            return check_launch("tcn_epilogue_bwd");
        }
    }
    if (dtype == TAMGCN_F32)
        tcn_epilogue_bwd_kernel<float><<<grid, 256, 0, st>>>(N, C, TV, (const float*)g, (const float*)out, relu,
                                                             (const float*)u, u_nstride, (const float*)r, r_nstride,
                                                             (float*)G, s1, s2u, s2r);
    else if (dtype == TAMGCN_BF16)
        tcn_epilogue_bwd_kernel<bf16><<<grid, 256, 0, st>>>(N, C, TV, (const bf16*)g, (const bf16*)out, relu,
                                                            (const bf16*)u, u_nstride, (const bf16*)r, r_nstride,
                                                            (bf16*)G, s1, s2u, s2r);
    else return set_error("tcn_epilogue_bwd: bad dtype %d", dtype);
    count_launch();
    return check_launch("tcn_epilogue_bwd");
}

extern "C" int tamgcn_maxpool_fwd(int dtype, int N, int C, int T, int To, int V, int stride, const tamgcn_operand* x,
                                  void* y, int64_t y_nstride, double* stat_sum, double* stat_sumsq,
                                  tamgcn_stream stream) {
    TG_REQUIRE(N > 0 && C > 0 && T > 0 && V > 0 && stride >= 1 && x && x->p && y, "maxpool_fwd: bad arguments");
    TG_REQUIRE(To == (T + 2 - 3) / stride + 1, "maxpool_fwd: To=%d inconsistent with T=%d stride=%d", To, T, stride);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid = ew_grid(N, C);
    const Opnd xo = make_opnd(x);
    if (dtype == TAMGCN_BF16 && !bv_disabled() && V % 4 == 0 &&
        bv_width(4, {xo.p, xo.q, y}, {xo.pns, xo.q ? xo.qns : 0, (long long)y_nstride}) == 4) {
        maxpool_fwd_vec_kernel<<<grid, 256, 0, st>>>(N, C, T, To, V, stride, xo, (bf16*)y, y_nstride, stat_sum, stat_sumsq);
        count_launch();
        return check_launch("maxpool_fwd");
    }
    if (dtype == TAMGCN_F32)
        maxpool_fwd_kernel<float><<<grid, 256, 0, st>>>(N, C, T, To, V, stride, xo, (float*)y, y_nstride, stat_sum, stat_sumsq);
    else if (dtype == TAMGCN_BF16)
        maxpool_fwd_kernel<bf16><<<grid, 256, 0, st>>>(N, C, T, To, V, stride, xo, (bf16*)y, y_nstride, stat_sum, stat_sumsq);
    else return set_error("maxpool_fwd: bad dtype %d", dtype);
    count_launch();
    return check_launch("maxpool_fwd");
}

extern "C" int tamgcn_maxpool_bwd(int dtype, int N, int C, int T, int To, int V, int stride, const tamgcn_operand* dy,
                                  const tamgcn_operand* x, void* dh, int64_t dh_nstride, double* s1, double* s2,
                                  tamgcn_stream stream) {
    TG_REQUIRE(N > 0 && C > 0 && T > 0 && V > 0 && stride >= 1 && dy && dy->p && x && x->p && dh,
               "maxpool_bwd: bad arguments");
    TG_REQUIRE(To == (T + 2 - 3) / stride + 1, "maxpool_bwd: To=%d inconsistent with T=%d stride=%d", To, T, stride);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid = ew_grid(N, C);
    const Opnd xo = make_opnd(x), dyo = make_opnd(dy);
    if (dtype == TAMGCN_BF16 && !bv_disabled() && V % 4 == 0 &&
        bv_width(4, {xo.p, xo.q, dyo.p, dyo.q, dh}, {xo.pns, xo.q ? xo.qns : 0, dyo.pns, dyo.q ? dyo.qns : 0, (long long)dh_nstride}) == 4) {
        maxpool_bwd_vec_kernel<<<grid, 256, 0, st>>>(N, C, T, To, V, stride, dyo, xo, (bf16*)dh, dh_nstride, s1, s2);
        count_launch();
        return check_launch("maxpool_bwd");
    }
    if (dtype == TAMGCN_F32)
        maxpool_bwd_kernel<float><<<grid, 256, 0, st>>>(N, C, T, To, V, stride, dyo, xo, (float*)dh, dh_nstride, s1, s2);
    else if (dtype == TAMGCN_BF16)
        maxpool_bwd_kernel<bf16><<<grid, 256, 0, st>>>(N, C, T, To, V, stride, dyo, xo, (bf16*)dh, dh_nstride, s1, s2);
    else return set_error("maxpool_bwd: bad dtype %d", dtype);
    count_launch();
    return check_launch("maxpool_bwd");
}
