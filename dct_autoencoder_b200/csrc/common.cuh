// Shared helpers for libdcta.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dcta.h"

namespace dcta {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return DCTA_ERR_LAUNCH;
    }
    return DCTA_OK;
}

#define DCTA_REQUIRE(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            dcta::set_error(__VA_ARGS__);    \
            return DCTA_ERR_INVALID_ARG;     \
        }                                    \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

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
