// Model glue around the quantiser (reference: modeling_dct_autoencoder.py:41-64, 85-112, 129-178):
// the row-wise pieces of
//   to_patch_embedding = Linear(p*p -> F, no bias) + LayerNorm(F, eps 1e-4), + (channel, h, w) position embeddings
//   proj_out           = LayerNorm(F, eps 1e-4) + Linear(F -> p*p, no bias)
//   add_pos_embedding_decoder_, and the biases of the quantiser's project_in / project_out
// on the packed (rows, slots, F) token layout.  The contractions themselves run through the split-precision
// tcgen05 GEMM (dcta_gemm_split); the kernels here prepare its operands and finish its output, one warp per token
// row, 128-bit accesses, the row held in registers between the statistics and the write (one read, one write).
#include <cuda_fp16.h>

#include "common.cuh"
#include "lfq_norm.cuh"

namespace dcta {

constexpr int kMaxChunks = 8;      // a lane holds up to 8 float4 of its row: rows of up to 1024 floats stay in registers

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// mean and 1/sqrt(var + eps) of a row held as v[0..n4) float4 per lane (biased variance, two-pass: torch's
// layer_norm semantics, modeling_dct_autoencoder.py:62, 86)
__device__ __forceinline__ void row_stats(const float4* v, int n4, int f, float eps, float& mean, float& rstd) {
    float s = 0.f;
    for (int i = 0; i < n4; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    mean = warp_sum(s) / (float)f;
    float q = 0.f;
    for (int i = 0; i < n4; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    rstd = rsqrtf(warp_sum(q) / (float)f + eps);
}

// out[t] = LN(x[t]) * gamma + beta (+ pos_c[channel[t]] + pos_h[pos[t,0]] + pos_w[pos[t,1]]); LN skipped when gamma
// is null, positions skipped when pos_c is null; bias (F) added when given.  f % 128 == 0, f <= 1024.
template <int NCH>
__global__ void __launch_bounds__(256) ln_pos_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps,
                                                     const float* __restrict__ bias, const float* __restrict__ pos_c,
                                                     const float* __restrict__ pos_h, const float* __restrict__ pos_w,
                                                     const int64_t* __restrict__ channels,
                                                     const int64_t* __restrict__ positions, float* __restrict__ out,
                                                     int64_t n_rows, int f) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp0; t < n_rows; t += n_warps) {
        const float* src = x + t * f;
        float4 v[NCH];
#pragma unroll
        for (int i = 0; i < NCH; ++i) v[i] = ld_stream(reinterpret_cast<const float4*>(src) + i * 32 + lane);
        if (gamma != nullptr) {
            float mean, rstd;
            row_stats(v, NCH, f, eps, mean, rstd);
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const float4 g = ld4(gamma + (i * 32 + lane) * 4), b = ld4(beta + (i * 32 + lane) * 4);
                v[i].x = fmaf((v[i].x - mean) * rstd, g.x, b.x);
                v[i].y = fmaf((v[i].y - mean) * rstd, g.y, b.y);
                v[i].z = fmaf((v[i].z - mean) * rstd, g.z, b.z);
                v[i].w = fmaf((v[i].w - mean) * rstd, g.w, b.w);
            }
        }
        if (bias != nullptr) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const float4 b = ld4(bias + (i * 32 + lane) * 4);
                v[i].x += b.x; v[i].y += b.y; v[i].z += b.z; v[i].w += b.w;
            }
        }
        if (pos_c != nullptr) {
            // reference order of the additions (modeling_dct_autoencoder.py:109): ((x + h_pos) + w_pos) + c_pos
            const float* pc = pos_c + channels[t] * f;
            const float* ph = pos_h + positions[2 * t] * f;
            const float* pw = pos_w + positions[2 * t + 1] * f;
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int o = (i * 32 + lane) * 4;
                const float4 a = ld4(ph + o), b = ld4(pw + o), c = ld4(pc + o);
                v[i].x = ((v[i].x + a.x) + b.x) + c.x;
                v[i].y = ((v[i].y + a.y) + b.y) + c.y;
                v[i].z = ((v[i].z + a.z) + b.z) + c.z;
                v[i].w = ((v[i].w + a.w) + b.w) + c.w;
            }
        }
#pragma unroll
        for (int i = 0; i < NCH; ++i) st_stream(reinterpret_cast<float4*>(out + t * f) + i * 32 + lane, v[i]);
    }
}

// the same for any row length (scalar accesses, the row is re-read instead of held): small feature sizes in tests,
// and rows that are not a multiple of 128 floats (the quantiser's 208-wide projection, p*p = 196)
__global__ void __launch_bounds__(256) ln_pos_generic_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps,
                                                             const float* __restrict__ bias, const float* __restrict__ pos_c,
                                                             const float* __restrict__ pos_h, const float* __restrict__ pos_w,
                                                             const int64_t* __restrict__ channels,
                                                             const int64_t* __restrict__ positions, float* __restrict__ out,
                                                             int64_t n_rows, int f) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp0; t < n_rows; t += n_warps) {
        const float* src = x + t * f;
        float mean = 0.f, rstd = 1.f;
        if (gamma != nullptr) {
            float s = 0.f;
            for (int i = lane; i < f; i += 32) s += src[i];
            mean = warp_sum(s) / (float)f;
            float q = 0.f;
            for (int i = lane; i < f; i += 32) { const float d = src[i] - mean; q += d * d; }
            rstd = rsqrtf(warp_sum(q) / (float)f + eps);
        }
        const float* pc = pos_c ? pos_c + channels[t] * f : nullptr;
        const float* ph = pos_c ? pos_h + positions[2 * t] * f : nullptr;
        const float* pw = pos_c ? pos_w + positions[2 * t + 1] * f : nullptr;
        for (int i = lane; i < f; i += 32) {
            float v = src[i];
            if (gamma != nullptr) v = fmaf((v - mean) * rstd, gamma[i], beta[i]);
            if (bias != nullptr) v += bias[i];
            if (pc != nullptr) v = ((v + ph[i]) + pw[i]) + pc[i];
            out[t * f + i] = v;
        }
    }
}

// rows (n, d) fp32 [-> LayerNorm] -> fp16 hi/lo planes (n, ld) of y * s_row with a power-of-two scale PER ROW
// (max|y| * s_row in [2^9, 2^10]: 22 significand bits of every element that matters, whatever the row's range), and
// row_scale[t] = post / s_row for the GEMM epilogue (post = 1 / scale of the other operand).
__global__ void __launch_bounds__(256) split_rows_rowscale_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float eps,
                                                                  __half* __restrict__ hi, __half* __restrict__ lo,
                                                                  float* __restrict__ row_scale, float post, int64_t n_rows,
                                                                  int d, int64_t ld) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp0; t < n_rows; t += n_warps) {
        const float* src = x + t * d;
        float mean = 0.f, rstd = 1.f;
        if (gamma != nullptr) {
            float s = 0.f;
            for (int i = lane; i < d; i += 32) s += src[i];
            mean = warp_sum(s) / (float)d;
            float q = 0.f;
            for (int i = lane; i < d; i += 32) { const float e = src[i] - mean; q += e * e; }
            rstd = rsqrtf(warp_sum(q) / (float)d + eps);
        }
        float amax = 0.f;
        for (int i = lane; i < d; i += 32) {
            float v = src[i];
            if (gamma != nullptr) v = fmaf((v - mean) * rstd, gamma[i], beta[i]);
            amax = fmaxf(amax, fabsf(v));
        }
        amax = warp_max(amax);
        // 2^k with amax * 2^k in [2^9, 2^10); all-zero (or non-finite) rows use 1
        int e = 0;
        if (amax > 0.f && amax < 3.0e38f) { (void)frexpf(amax, &e); e = 10 - e; }
        e = max(-100, min(100, e));
        const float s_row = ldexpf(1.0f, e);
        if (lane == 0) row_scale[t] = post * ldexpf(1.0f, -e);
        __half* dh = hi + t * ld;
        __half* dl = lo + t * ld;
        for (int i = lane; i < (int)ld; i += 32) {
            float v = 0.f;
            if (i < d) {
                v = src[i];
                if (gamma != nullptr) v = fmaf((v - mean) * rstd, gamma[i], beta[i]);
                v *= s_row;
            }
            const __half h = __float2half_rn(v);
            dh[i] = h;
            if (lo != nullptr) dl[i] = __float2half_rn(v - __half2float(h));
        }
    }
}

// The same with the row held in registers: d a multiple of 4 and at most 128 * CH (CH float4 per lane), 16-byte aligned
// rows; one pass over global memory, 128-bit loads, 64-bit stores of four halves per plane.
template <int CH>
__global__ void __launch_bounds__(256) split_rows_rowscale_vec_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                      const float* __restrict__ beta, float eps,
                                                                      __half* __restrict__ hi, __half* __restrict__ lo,
                                                                      float* __restrict__ row_scale, float post, int64_t n_rows,
                                                                      int d, int64_t ld, PatchNormRows pn) {
    const int lane = threadIdx.x & 31;
    const int d4 = d >> 2, ld4 = (int)(ld >> 2);
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp0; t < n_rows; t += n_warps) {
        const int64_t t_src = pn.row_src ? (int64_t)__ldg(pn.row_src + t) : t;
        const float4* src = reinterpret_cast<const float4*>(x + (t_src < 0 ? 0 : t_src) * d);
        float4 v[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int i = lane + 32 * c;
            v[c] = (i < d4 && t_src >= 0) ? ld_stream(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (pn.channels != nullptr) {
            // PatchNorm.forward, frozen statistics (patchnorm.py:157-165), on the row in registers: the normalised
            // patches are never written (padding rows read the statistics at (0, 0, 0), like the reference)
            const int64_t row = (int64_t)clamped_position(pn.channels, pn.positions, t, pn.C, pn.H, pn.W) * d4;
            const float4* m4 = reinterpret_cast<const float4*>(pn.median) + row;
            const float4* b4 = reinterpret_cast<const float4*>(pn.b) + row;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int i = lane + 32 * c;
                if (i < d4) {
                    const float4 mv = __ldg(m4 + i), bv = __ldg(b4 + i);
                    v[c].x = patchnorm_value<false>(v[c].x, mv.x, bv.x, pn.eps, pn.lo, pn.hi);
                    v[c].y = patchnorm_value<false>(v[c].y, mv.y, bv.y, pn.eps, pn.lo, pn.hi);
                    v[c].z = patchnorm_value<false>(v[c].z, mv.z, bv.z, pn.eps, pn.lo, pn.hi);
                    v[c].w = patchnorm_value<false>(v[c].w, mv.w, bv.w, pn.eps, pn.lo, pn.hi);
                }
            }
        }
        if (gamma != nullptr) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < CH; ++c) s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
            const float mean = warp_sum(s) / (float)d;
            float q = 0.f;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (lane + 32 * c < d4) {
                    const float e0 = v[c].x - mean, e1 = v[c].y - mean, e2 = v[c].z - mean, e3 = v[c].w - mean;
                    q += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
                }
            }
            const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const int i = lane + 32 * c;
                if (i < d4) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i), b = __ldg(reinterpret_cast<const float4*>(beta) + i);
                    v[c].x = fmaf((v[c].x - mean) * rstd, g.x, b.x);
                    v[c].y = fmaf((v[c].y - mean) * rstd, g.y, b.y);
                    v[c].z = fmaf((v[c].z - mean) * rstd, g.z, b.z);
                    v[c].w = fmaf((v[c].w - mean) * rstd, g.w, b.w);
                }
            }
        }
        float amax = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[c].x), fabsf(v[c].y)), fmaxf(fabsf(v[c].z), fabsf(v[c].w))));
        amax = warp_max(amax);
        int e = 0;
        if (amax > 0.f && amax < 3.0e38f) { (void)frexpf(amax, &e); e = 10 - e; }
        e = max(-100, min(100, e));
        const float s_row = ldexpf(1.0f, e);
        if (lane == 0) row_scale[t] = post * ldexpf(1.0f, -e);
        uint2* dh = reinterpret_cast<uint2*>(hi + t * ld);
        uint2* dl = lo ? reinterpret_cast<uint2*>(lo + t * ld) : nullptr;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int i = lane + 32 * c;
            if (i >= ld4) continue;
            const float a[4] = {v[c].x * s_row, v[c].y * s_row, v[c].z * s_row, v[c].w * s_row};      // zero past d
            const __half2 h01 = __floats2half2_rn(a[0], a[1]), h23 = __floats2half2_rn(a[2], a[3]);
            dh[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
            if (dl) {
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn(a[0] - f01.x, a[1] - f01.y), l23 = __floats2half2_rn(a[2] - f23.x, a[3] - f23.y);
                dl[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
            }
        }
    }
}

}  // namespace dcta

extern "C" int dcta_ln_pos_rows(const float* x, const float* gamma, const float* beta, float eps, const float* bias,
                                const float* pos_c, const float* pos_h, const float* pos_w, const int64_t* channels,
                                const int64_t* positions, float* out, int64_t n_rows, int f, void* stream) {
    using namespace dcta;
    DCTA_REQUIRE(x && out && n_rows >= 0 && f > 0, "ln_pos_rows: bad arguments");
    DCTA_REQUIRE((gamma == nullptr) == (beta == nullptr), "ln_pos_rows: gamma and beta go together");
    DCTA_REQUIRE(pos_c == nullptr || (pos_h && pos_w && channels && positions), "ln_pos_rows: incomplete position tables");
    if (n_rows == 0) return DCTA_OK;
    const int grid = grid_for(n_rows, 8);
    cudaStream_t st = as_stream(stream);
    const bool vec = f % 128 == 0 && f / 128 <= kMaxChunks && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
#define DCTA_LN_CASE(N)                                                                                                        \
    case N: ln_pos_kernel<N><<<grid, 256, 0, st>>>(x, gamma, beta, eps, bias, pos_c, pos_h, pos_w, channels, positions, out,  \
                                                     n_rows, f); break;
    if (vec) {
        switch (f / 128) {
            DCTA_LN_CASE(1) DCTA_LN_CASE(2) DCTA_LN_CASE(3) DCTA_LN_CASE(4) DCTA_LN_CASE(5) DCTA_LN_CASE(6) DCTA_LN_CASE(7) DCTA_LN_CASE(8)
        }
    } else {
        ln_pos_generic_kernel<<<grid, 256, 0, st>>>(x, gamma, beta, eps, bias, pos_c, pos_h, pos_w, channels, positions, out, n_rows, f);
    }
#undef DCTA_LN_CASE
    return check_launch("ln_pos_rows");
}

static int launch_split_rows(const float* x, const float* gamma, const float* beta, float eps, void* hi, void* lo,
                             float* row_scale, float post, int64_t n_rows, int d, int64_t ld, const dcta::PatchNormRows& pn,
                             void* stream, const char* who) {
    using namespace dcta;
    DCTA_REQUIRE(x && hi && row_scale && n_rows >= 0 && d > 0 && ld >= d && ld % 8 == 0, "%s: bad arguments", who);
    DCTA_REQUIRE((gamma == nullptr) == (beta == nullptr), "%s: gamma and beta go together", who);
    if (n_rows == 0) return DCTA_OK;
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
                           reinterpret_cast<uintptr_t>(pn.median) | reinterpret_cast<uintptr_t>(pn.b)) & 15) == 0 &&
                         ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 7) == 0;
    const int ch = (int)((ld / 4 + 31) / 32);          // float4 per lane (the zero padding up to ld included)
    cudaStream_t st = as_stream(stream);
    const int grid = grid_for(n_rows, 8);
#define DCTA_SPLIT_CASE(N)                                                                                                  \
    case N: split_rows_rowscale_vec_kernel<N><<<grid, 256, 0, st>>>(x, gamma, beta, eps, (__half*)hi, (__half*)lo, row_scale, \
                                                                     post, n_rows, d, ld, pn); break;
    if (d % 4 == 0 && aligned && ch >= 1 && ch <= 8) {
        switch (ch) {
            DCTA_SPLIT_CASE(1) DCTA_SPLIT_CASE(2) DCTA_SPLIT_CASE(3) DCTA_SPLIT_CASE(4)
            DCTA_SPLIT_CASE(5) DCTA_SPLIT_CASE(6) DCTA_SPLIT_CASE(7) DCTA_SPLIT_CASE(8)
        }
    } else {
        if (pn.channels != nullptr) {
            set_error("%s: the fused PatchNorm needs rows of a multiple of 4 (at most 1024) 16-byte aligned floats", who);
            return DCTA_ERR_UNSUPPORTED;
        }
        split_rows_rowscale_kernel<<<grid, 256, 0, st>>>(x, gamma, beta, eps, (__half*)hi, (__half*)lo, row_scale, post, n_rows, d, ld);
    }
#undef DCTA_SPLIT_CASE
    return check_launch(who);
}

extern "C" int dcta_split_rows_rowscale(const float* x, const float* gamma, const float* beta, float eps, void* hi, void* lo,
                                        float* row_scale, float post, int64_t n_rows, int d, int64_t ld, void* stream) {
    return launch_split_rows(x, gamma, beta, eps, hi, lo, row_scale, post, n_rows, d, ld, dcta::PatchNormRows{}, stream,
                             "split_rows_rowscale");
}

extern "C" int dcta_split_rows_patchnorm(const float* x, const int32_t* row_src, const int64_t* channels,
                                         const int64_t* positions, const float* median, const float* b, int C, int H, int W,
                                         float eps, float clamp_lo, float clamp_hi, void* hi, void* lo, float* row_scale,
                                         float post, int64_t n_rows, int d, int64_t ld, void* stream) {
    using namespace dcta;
    DCTA_REQUIRE(channels && positions && median && b && C > 0 && H > 0 && W > 0, "split_rows_patchnorm: bad arguments");
    const PatchNormRows pn{channels, positions, median, b, C, H, W, eps, clamp_lo, clamp_hi, row_src};
    return launch_split_rows(x, nullptr, nullptr, 0.0f, hi, lo, row_scale, post, n_rows, d, ld, pn, stream, "split_rows_patchnorm");
}
