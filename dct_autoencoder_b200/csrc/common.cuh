// Shared helpers for libdcta.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dcta.h"

namespace dcta {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const char* fmt, ...);

// dcta_profile_begin / dcta_profile_end (api.cu): while a profile is open, every check_launch() records a CUDA event
// on the stream of the launch it checks, so that a caller gets device times per launch group without a profiler.
extern bool g_profile_on;
extern cudaStream_t g_last_stream;       // stream of the most recent as_stream() (one thread drives the library)
void profile_mark(const char* what);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return DCTA_ERR_LAUNCH;
    }
    if (g_profile_on) profile_mark(what);
    return DCTA_OK;
}

#define DCTA_REQUIRE(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            dcta::set_error(__VA_ARGS__);    \
            return DCTA_ERR_INVALID_ARG;     \
        }                                    \
    } while (0)

inline cudaStream_t as_stream(void* s) {
    g_last_stream = reinterpret_cast<cudaStream_t>(s);
    return g_last_stream;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid for a grid-stride loop: enough CTAs for `work` items at `per_cta`, capped at a multiple
// of the SM count so the tail wave is full.
inline int grid_for(int64_t work, int per_cta, int ctas_per_sm = 8) {
    int64_t g = ceil_div(work, per_cta);
    int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

struct Mat3 {
    float m[9];
};

// streaming 128-bit accesses (read-once / write-once data must not pollute L1)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// uint8 pixel -> float in [0, 1], bit-identical to torch's `u8_tensor / 255` (an IEEE fp32 division; what the
// reference's callers do with torchvision.io.read_image, decode_gif.py:22, testpipe.py:17): q = u * fl(1/255) is off
// by one ulp for 126 of the 256 values; one FMA residual + one FMA correction rounds all 256 correctly.
__device__ __forceinline__ float u8_to_unit(uint32_t u) {
    const float uf = (float)u, r = 1.0f / 255.0f;
    const float q = uf * r;
    return fmaf(fmaf(-q, 255.0f, uf), r, q);
}
// float -> uint8 as torchvision.utils.save_image stores it (what the reference's callers do with a reconstruction,
// testpipe.py:74-75): clamp(0, 1) * 255 + 0.5, clamp to [0, 255], truncate.  NaN -> 0.
__device__ __forceinline__ uint32_t unit_to_u8(float v) {
    const float c = fminf(fmaxf(v, 0.0f), 1.0f);
    return (uint32_t)__float2int_rz(__fadd_rn(__fmul_rn(c, 255.0f), 0.5f));
}
// four consecutive pixels of one channel plane, in units of 4 pixels (float4 for fp32 planes, one 32-bit word for uint8)
__device__ __forceinline__ float4 ld_px4(const float* plane, int64_t i4) {
    return ld_stream(reinterpret_cast<const float4*>(plane) + i4);
}
__device__ __forceinline__ float4 ld_px4(const uint8_t* plane, int64_t i4) {
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(plane) + i4);
    return make_float4(u8_to_unit(v & 255u), u8_to_unit((v >> 8) & 255u), u8_to_unit((v >> 16) & 255u), u8_to_unit(v >> 24));
}
__device__ __forceinline__ void st_px4(float* plane, int64_t i4, const float4& v) {
    st_stream(reinterpret_cast<float4*>(plane) + i4, v);
}
__device__ __forceinline__ void st_px4(uint8_t* plane, int64_t i4, const float4& v) {
    reinterpret_cast<uint32_t*>(plane)[i4] = unit_to_u8(v.x) | (unit_to_u8(v.y) << 8) | (unit_to_u8(v.z) << 16) | (unit_to_u8(v.w) << 24);
}

// sign(v) * |v|^g as exp2(g * log2|v|) on the special-function unit (MUFU.LG2 / MUFU.EX2):
// |log2 error| <= 2^-21.4 on [0.5, 2] and 2 ulp elsewhere, exp2 2 ulp, i.e. <= ~4e-7 relative for the
// exponents used here (0.43 and 1/0.43) at a tenth of the instructions of powf(), which keeps the colour
// kernels HBM-bound.  |v| = 0 -> log2 = -inf -> exp2 = 0.  (reference: util.py:76-78, util.py:93-95)
__device__ __forceinline__ float signed_pow(float v, float g) {
    float l, a;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fabsf(v)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(g * l));
    return v < 0.0f ? -a : a;
}

// The same with the exp2 evaluated by a degree-6 polynomial on the FMA pipe (truncation 2e-9, fp32 Horner rounding
// <= 1e-7, pseudo-random) and the product g * log2|v| carried with its FMA residual.  MUFU.EX2's 2^-22 relative error
// is a smooth function of its argument: over a natural image it is spatially correlated, and the whole-image DCT sums
// correlated input errors coherently into the low-frequency coefficients (measured: 4.4e-7 * max|Y| on the
// reference's images/ at 256^2 with MUFU.EX2, against 4e-7 stated).  Used by the FORWARD colour transform only; the
// decode side ends at the pixels, where nothing accumulates.
// CLAMP = false: for |g| < 1 and finite v the exponent g * log2|v| lies in [-126 g, 128 g] (subnormals flush to -inf and
// are caught by the last line), so the range clamp is dead code; round(g*l) comes out of one FMA against the magic
// constant and the residual r = g*l - n out of a second one (a single rounding of the exact difference: the same value
// as the hi/lo sequence below yields); the sign is copied with one logic instruction.  14 instructions instead of 21.
template <bool CLAMP = true>
__device__ __forceinline__ float signed_pow_fwd(float v, float g) {
    float l;
    const float av = fabsf(v);
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(av));
    const float magic = 12582912.0f;                  // 1.5 * 2^23: the sum holds round(t) in its low mantissa bits
    float nm, r;
    if (CLAMP) {
        float t_hi = g * l;
        const float t_lo = fmaf(g, l, -t_hi);
        t_hi = fminf(fmaxf(t_hi, -126.0f), 126.0f);
        nm = t_hi + magic;
        r = (t_hi - (nm - magic)) + t_lo;             // [-0.5, 0.5]
    } else {
        nm = fmaf(g, l, magic);
        r = fmaf(g, l, -(nm - magic));
    }
    float p = 1.5370705e-4f;
    p = fmaf(p, r, 1.3399848e-3f);
    p = fmaf(p, r, 9.6183736e-3f);
    p = fmaf(p, r, 5.5503290e-2f);
    p = fmaf(p, r, 2.4022648e-1f);
    p = fmaf(p, r, 6.9314718e-1f);
    p = fmaf(p, r, 1.0f);
    const float a = __int_as_float(__float_as_int(p) + (__float_as_int(nm) << 23));
    if (CLAMP) return l < -125.0f ? 0.0f : (v < 0.0f ? -a : a);   // |v| = 0 (or subnormal): log2 = -inf -> 0, as pow does
    const float z = l < -125.0f ? 0.0f : a;
    return __uint_as_float(__float_as_uint(z) | (__float_as_uint(v) & 0x80000000u));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace dcta
