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

// Row of the statistics tables a token reads: (channel, tile row, tile column) with torch's negative-index wrap,
// clamped into the table (patchnorm.py:157-160)
__device__ __forceinline__ int clamped_position(const int64_t* channels, const int64_t* positions, int64_t tok, int C,
                                                int H, int W) {
    int64_t c = channels[tok], h = positions[2 * tok], w = positions[2 * tok + 1];
    // torch indexing wraps negative indices; out-of-range indices raise in the reference.
    if (c < 0) c += C;
    if (h < 0) h += H;
    if (w < 0) w += W;
    c = min(max(c, (int64_t)0), (int64_t)C - 1);
    h = min(max(h, (int64_t)0), (int64_t)H - 1);
    w = min(max(w, (int64_t)0), (int64_t)W - 1);
    return (int)((c * H + h) * W + w);
}

// PatchNorm forward (frozen) / inverse of one value, in the reference's operation order with explicit roundings
template <bool kInverse>
__device__ __forceinline__ float patchnorm_value(float xv, float mv, float bv, float eps, float lo, float hi) {
    const float sd = __fadd_rn(__fmul_rn(bv, kSqrt2f), eps);
    if (kInverse) return __fadd_rn(__fmul_rn(xv, sd), mv);          // PN:177
    const float y = __fdiv_rn(__fsub_rn(xv, mv), sd);                // PN:161
    return y < lo ? lo : (y > hi ? hi : y);                          // PN:163 (NaN propagates, as torch.clamp)
}

// PatchNorm forward fused into a consumer of token rows (glue.cu: the operand split of the LFQ projection)
struct PatchNormRows {
    const int64_t* channels;    // (n_rows); nullptr = no normalisation
    const int64_t* positions;   // (n_rows, 2)
    const float* median;        // (C, H, W, z)
    const float* b;
    int C, H, W;
    float eps, lo, hi;
    const int32_t* row_src;     // nullable: row t of the operand is row row_src[t] of x (-1: a row of zeros) -- the token
                                // gather of the packed batch (dcta_pack_tiles_index) folded into this pass
};

// *flag = nonzero iff every b[i] is finite and in [0, 1e18] (fused_lfq.cu)
int launch_b_tame(const float* b, int64_t n, int32_t* flag, cudaStream_t st);

}  // namespace dcta
