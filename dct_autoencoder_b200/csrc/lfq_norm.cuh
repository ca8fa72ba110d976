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

}  // namespace dcta
