// bn.cu — BatchNorm bookkeeping kernels (nn.BatchNorm2d / BatchNorm1d defaults).
//
// The heavy work of a BatchNorm never runs here: the per-channel sums are accumulated by the
// epilogue of the kernel that PRODUCES the tensor (fp64 atomics), and the normalisation is applied
// by the prologue of the kernel that CONSUMES it (tamgcn_operand).  These two tiny kernels turn the
// sums into the per-channel coefficients in between, for up to 8 BatchNorms per launch.
//
//   forward :  scale = gamma*invstd, shift = beta - mean*scale, running-stat update (momentum, unbiased var)
//   backward:  dY = A*dYhat + B*Y + C  with  A = gamma*invstd,
//              B = -gamma*invstd^2 * s2hat/count,  C = -gamma*invstd*s1/count - B*mean... (see below)
#include "common.cuh"

namespace tamgcn {

#define MAX_BN 8
struct BnPack { tamgcn_bn d[MAX_BN]; };
struct BnBwdPack { tamgcn_bn_bwd d[MAX_BN]; };

__global__ void __launch_bounds__(256)
bn_finalize_kernel(BnPack pk, double count, float momentum, float eps, int train) {
    const tamgcn_bn& d = pk.d[blockIdx.x];
    for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
        double mean, var;
        if (train) {
            mean = d.sum[c] / count;
            var = d.sumsq[c] / count - mean * mean;
            if (var < 0.0) var = 0.0;
            if (d.rmean) {
                const double unb = (count > 1.0) ? var * count / (count - 1.0) : var;
                d.rmean[c] = (float)((1.0 - (double)momentum) * (double)d.rmean[c] + (double)momentum * mean);
                d.rvar[c] = (float)((1.0 - (double)momentum) * (double)d.rvar[c] + (double)momentum * unb);
            }
        } else {
            mean = (double)d.rmean[c];
            var = (double)d.rvar[c];
        }
        const float invstd = (float)(1.0 / sqrt(var + (double)eps));
        const float gamma = d.gamma ? d.gamma[c] : 1.f;
        const float beta = d.beta ? d.beta[c] : 0.f;
        const float scale = gamma * invstd;
        d.scale[c] = scale;
        d.shift[c] = beta - (float)mean * scale;
        if (d.mean) d.mean[c] = (float)mean;
        if (d.invstd) d.invstd[c] = invstd;
    }
    if (train && d.nbt && threadIdx.x == 0) *d.nbt += 1;
}

// With xhat = (Y - mean)*invstd, s1 = sum dYhat, s2 = sum dYhat*Y:
//   sum dYhat*xhat = invstd*(s2 - mean*s1) =: sx             (= dgamma)
//   train: dY = gamma*invstd*(dYhat - s1/count - xhat*sx/count)
//             = A*dYhat + B*Y + C,  A = gamma*invstd, B = -A*invstd*sx/count, C = -A*s1/count - B*mean
//   eval : dY = A*dYhat
__global__ void __launch_bounds__(256)
bn_bwd_coef_kernel(BnBwdPack pk, double count, int train) {
    const tamgcn_bn_bwd& d = pk.d[blockIdx.x];
    for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
        const double mean = (double)d.mean[c], invstd = (double)d.invstd[c];
        const double s1 = d.s1[c], s2 = d.s2[c];
        const double sx = invstd * (s2 - mean * s1);
        const double gamma = d.gamma ? (double)d.gamma[c] : 1.0;
        const double A = gamma * invstd;
        double B = 0.0, Cc = 0.0;
        if (train) {
            B = -A * invstd * sx / count;
            Cc = -A * s1 / count - B * mean;
        }
        d.A[c] = (float)A;
        d.B[c] = (float)B;
        d.Cc[c] = (float)Cc;
        if (d.dgamma) d.dgamma[c] = (float)sx;
        if (d.dbeta) d.dbeta[c] = (float)s1;
    }
}

// coefficients of the lazily formed difference  res - y  of unit_gcn (models/ctrgcn.py:256-259):
//   nb = -sb,  c = (ha ? ha : 0) - hb
__global__ void __launch_bounds__(256)
coef_diff_kernel(int C, const float* __restrict__ sb, const float* __restrict__ ha, const float* __restrict__ hb,
                 float* __restrict__ nb, float* __restrict__ c) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C; i += gridDim.x * blockDim.x) {
        nb[i] = -sb[i];
        c[i] = (ha ? ha[i] : 0.f) - hb[i];
    }
}

}  // namespace tamgcn

using namespace tamgcn;

extern "C" int tamgcn_coef_diff(int C, const float* sb, const float* ha, const float* hb, float* nb, float* c,
                                tamgcn_stream stream) {
    TG_REQUIRE(C > 0 && sb && hb && nb && c, "coef_diff: bad arguments");
    coef_diff_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(C, sb, ha, hb, nb, c);
    count_launch();
    return check_launch("coef_diff");
}

extern "C" int tamgcn_bn_finalize(int n_bn, const tamgcn_bn* bns, double count, float momentum, float eps, int train,
                                  tamgcn_stream stream) {
    TG_REQUIRE(n_bn >= 1 && n_bn <= MAX_BN && bns, "bn_finalize: n_bn=%d out of range 1..%d", n_bn, MAX_BN);
    TG_REQUIRE(count >= 1.0, "bn_finalize: empty batch");
    BnPack pk;
    for (int i = 0; i < n_bn; ++i) {
        pk.d[i] = bns[i];
        TG_REQUIRE(bns[i].C > 0 && bns[i].scale && bns[i].shift, "bn_finalize: descriptor %d incomplete", i);
        if (train) TG_REQUIRE(bns[i].sum && bns[i].sumsq, "bn_finalize: descriptor %d needs batch sums in train mode", i);
        else TG_REQUIRE(bns[i].rmean && bns[i].rvar, "bn_finalize: descriptor %d needs running stats in eval mode", i);
    }
    bn_finalize_kernel<<<n_bn, 256, 0, (cudaStream_t)stream>>>(pk, count, momentum, eps, train);
    count_launch();
    return check_launch("bn_finalize");
}

extern "C" int tamgcn_bn_bwd_coef(int n_bn, const tamgcn_bn_bwd* bns, double count, int train, tamgcn_stream stream) {
    TG_REQUIRE(n_bn >= 1 && n_bn <= MAX_BN && bns, "bn_bwd_coef: n_bn=%d out of range 1..%d", n_bn, MAX_BN);
    BnBwdPack pk;
    for (int i = 0; i < n_bn; ++i) {
        pk.d[i] = bns[i];
        TG_REQUIRE(bns[i].C > 0 && bns[i].s1 && bns[i].s2 && bns[i].mean && bns[i].invstd && bns[i].A && bns[i].B &&
                       bns[i].Cc, "bn_bwd_coef: descriptor %d incomplete", i);
    }
    bn_bwd_coef_kernel<<<n_bn, 256, 0, (cudaStream_t)stream>>>(pk, count, train);
    count_launch();
    return check_launch("bn_bwd_coef");
}
