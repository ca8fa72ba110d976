// Parameters of the fused PatchNorm + LFQ kernels (fused_lfq.cu, dct_fold.cu).
#pragma once
#include "common.cuh"

namespace dcta {

constexpr float kSqrt2f = 1.41421356237309504880f;

struct LfqNormParams {
    const float* median;   // (C, H, W, z)
    const float* b;        // (C, H, W, z)
    int C, H, W, z;
    float eps, lo, hi;     // PatchNorm eps / clamp
    int c, d;              // LFQ codebooks x bits, c * d == z
    float scale;           // LFQ codebook_scale
};

// sign of clamp((x - m) / (b*sqrt2 + eps)): decided by the numerator whenever that is safe
// (patchnorm.py:161-163 followed by lfq.py:175)
__device__ __forceinline__ unsigned norm_sign_bit(float xv, float mv, float bv, const LfqNormParams& q) {
    const float sd = __fadd_rn(__fmul_rn(bv, kSqrt2f), q.eps);
    const float diff = __fsub_rn(xv, mv);
    if (sd > 0.0f && sd < 1e30f && fabsf(diff) > 1e-30f && q.lo < 0.0f && q.hi > 0.0f) return diff > 0.0f;
    float y = __fdiv_rn(diff, sd);
    y = y < q.lo ? q.lo : (y > q.hi ? q.hi : y);
    return y > 0.0f;
}

// *flag = nonzero iff every b[i] is finite and in [0, 1e18] (fused_lfq.cu)
int launch_b_tame(const float* b, int64_t n, int32_t* flag, cudaStream_t st);

}  // namespace dcta
