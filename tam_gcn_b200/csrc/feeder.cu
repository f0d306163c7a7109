// feeder.cu — GPU-side skeleton feeder for NW-UCLA: the arithmetic of
// feeder/feeder_nucla_gcn.py:85-130 for a whole batch in one launch, writing the (N, 3, T, V, 1) fp32 tensor that
// Model.forward consumes.  At > 10^5 samples/s the reference's per-sample numpy feeder is the bottleneck of a run.
//
// Per sample (one CTA), with the random draws supplied by the caller (so the result is reproducible and testable):
//   value -= value[0, 1, :]                                   (centre on joint 1 of the first frame,  :98-99)
//   X = value . (Ry . Rx . S)                                 (view rotation by (agx, agy) degrees, scale s, :76-84,:100)
//   X = 2 (X - min) / (max - min + 1e-6) - 1                  (per-axis min / max over all frames and joints, :102-104)
//   data[t] = X[frame_idx[t]]                                 (temporal resampling to T = 52 frames, :107-117)
//   bone:   data[:, a] -= data[:, b] for the 20 (a, b) pairs  (:119-123; every joint appears once as `a`)
//   motion: data[t] = data[t+1] - data[t], last frame 0       (:124-127)
//   out[c, t, v, 0] = data[t, v, c]                           (:129-130)
// Pass 1 reduces min / max over the sequence, pass 2 recomputes the few points each output element needs: the raw
// sequence (<= a few KB) stays in L1/L2, nothing intermediate is written.
#include "common.cuh"

namespace tamgcn {

struct FeederP {
    int B, Lmax, V, T, mode;      // mode 0 joint, 1 bone, 2 motion
};

__device__ __forceinline__ void feeder_point(const float* __restrict__ raw, int l, int v, int V, const float* c0,
                                             const float (&Mx)[9], float (&q)[3]) {
    const float* p = raw + ((long long)l * V + v) * 3;
    const float x = __ldg(p) - c0[0], y = __ldg(p + 1) - c0[1], z = __ldg(p + 2) - c0[2];
    q[0] = x * Mx[0] + y * Mx[3] + z * Mx[6];
    q[1] = x * Mx[1] + y * Mx[4] + z * Mx[7];
    q[2] = x * Mx[2] + y * Mx[5] + z * Mx[8];
}

__global__ void __launch_bounds__(256)
feeder_nucla_kernel(FeederP p, const float* __restrict__ raw, const int* __restrict__ length,
                    const long long* __restrict__ sample, const float* __restrict__ view, const int* __restrict__ frame_idx,
                    const int* __restrict__ bone_parent, float* __restrict__ out) {
    __shared__ float red[2][3][8];
    __shared__ float s_lo[3], s_sc[3], s_M[9], s_c0[3];
    const int b = blockIdx.x;
    const long long s = sample[b];
    const float* rs = raw + s * (long long)p.Lmax * p.V * 3;
    const int L = length[s];
    if (threadIdx.x == 0) {
        // M = Ry . Rx . S  (feeder/feeder_nucla_gcn.py:76-84), built in double
        const double ax = (double)view[b * 3 + 0] * 3.14159265358979323846 / 180.0;
        const double ay = (double)view[b * 3 + 1] * 3.14159265358979323846 / 180.0;
        const double sc = (double)view[b * 3 + 2];
        const double cx = cos(ax), sx = sin(ax), cy = cos(ay), sy = sin(ay);
        const double Rx[9] = {1, 0, 0, 0, cx, sx, 0, -sx, cx};
        const double Ry[9] = {cy, 0, -sy, 0, 1, 0, sy, 0, cy};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double a = 0;
                for (int k = 0; k < 3; ++k) a += Ry[i * 3 + k] * Rx[k * 3 + j];
                s_M[i * 3 + j] = (float)(a * sc);
            }
        for (int c = 0; c < 3; ++c) s_c0[c] = __ldg(rs + (0 * p.V + 1) * 3 + c);
    }
    __syncthreads();
    float Mx[9], c0[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) Mx[i] = s_M[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) c0[i] = s_c0[i];
    // pass 1: per-axis min / max of the transformed sequence
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < L * p.V; i += blockDim.x) {
        float q[3];
        feeder_point(rs, i / p.V, i % p.V, p.V, c0, Mx, q);
#pragma unroll
        for (int c = 0; c < 3; ++c) { lo[c] = fminf(lo[c], q[c]); hi[c] = fmaxf(hi[c], q[c]); }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if (lane == 0) { red[0][c][w] = lo[c]; red[1][c][w] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float a = INFINITY, z = -INFINITY;
        for (int i = 0; i < 8; ++i) { a = fminf(a, red[0][threadIdx.x][i]); z = fmaxf(z, red[1][threadIdx.x][i]); }
        s_lo[threadIdx.x] = a;
        s_sc[threadIdx.x] = 2.f / (z - a + 1e-6f);
    }
    __syncthreads();
    const float l0 = s_lo[0], l1 = s_lo[1], l2 = s_lo[2], k0 = s_sc[0], k1 = s_sc[1], k2 = s_sc[2];
    // pass 2: one thread per output (t, v), all three coordinates
    const int* fi = frame_idx + (long long)b * p.T;
    float* ob = out + (long long)b * 3 * p.T * p.V;
    for (int i = threadIdx.x; i < p.T * p.V; i += blockDim.x) {
        const int t = i / p.V, v = i - t * p.V;
        float q[3], r[3] = {0.f, 0.f, 0.f};
        feeder_point(rs, fi[t], v, p.V, c0, Mx, q);
        float d0 = (q[0] - l0) * k0 - 1.f, d1 = (q[1] - l1) * k1 - 1.f, d2 = (q[2] - l2) * k2 - 1.f;
        if (p.mode == 1) {
            const int pv = bone_parent[v];
            if (pv >= 0) {
                feeder_point(rs, fi[t], pv, p.V, c0, Mx, r);
                d0 -= (r[0] - l0) * k0 - 1.f; d1 -= (r[1] - l1) * k1 - 1.f; d2 -= (r[2] - l2) * k2 - 1.f;
            } else {
                d0 = d1 = d2 = 0.f;
            }
        } else if (p.mode == 2) {
            if (t + 1 < p.T) {
                feeder_point(rs, fi[t + 1], v, p.V, c0, Mx, r);
                d0 = ((r[0] - l0) * k0 - 1.f) - d0; d1 = ((r[1] - l1) * k1 - 1.f) - d1; d2 = ((r[2] - l2) * k2 - 1.f) - d2;
            } else {
                d0 = d1 = d2 = 0.f;
            }
        }
        ob[(0 * p.T + t) * p.V + v] = d0;
        ob[(1 * p.T + t) * p.V + v] = d1;
        ob[(2 * p.T + t) * p.V + v] = d2;
    }
}

}  // namespace tamgcn

using namespace tamgcn;

extern "C" int tamgcn_feeder_nucla(const float* raw, const int32_t* length, const int64_t* sample, const float* view,
                                   const int32_t* frame_idx, const int32_t* bone_parent, int B, int Lmax, int V, int T,
                                   int mode, float* out, tamgcn_stream stream) {
    TG_REQUIRE(B > 0 && Lmax > 0 && V > 1 && T > 0, "feeder_nucla: empty shape B=%d Lmax=%d V=%d T=%d", B, Lmax, V, T);
    TG_REQUIRE(raw && length && sample && view && frame_idx && out, "feeder_nucla: null pointer");
    TG_REQUIRE(mode >= 0 && mode <= 2, "feeder_nucla: mode %d (0 joint, 1 bone, 2 motion)", mode);
    TG_REQUIRE(mode != 1 || bone_parent, "feeder_nucla: bone mode needs the parent table");
    FeederP p = {B, Lmax, V, T, mode};
    feeder_nucla_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(p, raw, length, (const long long*)sample, view, frame_idx,
                                                            bone_parent, out);
    count_launch();
    return check_launch("feeder_nucla");
}
