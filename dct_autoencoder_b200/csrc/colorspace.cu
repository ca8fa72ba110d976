// RGB <-> IPT colour space (reference: util.py:46-47, 56-97).  Streaming, HBM-bound:
// 12 B read + 12 B written per pixel, 128-bit accesses on each of the three channel planes.
#include "common.cuh"

namespace dcta {

void set_error(const char* fmt, ...);


template <bool kToIpt>
__device__ __forceinline__ void convert_px(float r, float g, float b, const Mat3& A, const Mat3& B,
                                           float& o0, float& o1, float& o2) {
    // out = B . signed_pow(A . in)
    float l = fmaf(A.m[2], b, fmaf(A.m[1], g, A.m[0] * r));
    float m = fmaf(A.m[5], b, fmaf(A.m[4], g, A.m[3] * r));
    float s = fmaf(A.m[8], b, fmaf(A.m[7], g, A.m[6] * r));
    const float gamma = kToIpt ? 0.43f : (float)(1.0 / 0.43);
    if (kToIpt) {      // forward: accurate exp2 (errors here are summed coherently by the whole-image DCT)
        l = signed_pow_fwd(l, gamma);
        m = signed_pow_fwd(m, gamma);
        s = signed_pow_fwd(s, gamma);
    } else {
        l = signed_pow(l, gamma);
        m = signed_pow(m, gamma);
        s = signed_pow(s, gamma);
    }
    o0 = fmaf(B.m[2], s, fmaf(B.m[1], m, B.m[0] * l));
    o1 = fmaf(B.m[5], s, fmaf(B.m[4], m, B.m[3] * l));
    o2 = fmaf(B.m[8], s, fmaf(B.m[7], m, B.m[6] * l));
}

template <bool kToIpt>
__global__ void __launch_bounds__(256) colorspace_vec4(const float* __restrict__ in,
                                                       float* __restrict__ out, int64_t n_img,
                                                       int64_t plane4, Mat3 A, Mat3 B) {
    // one thread = 4 consecutive pixels of one image (three 128-bit loads, three 128-bit stores)
    const int64_t total = n_img * plane4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t img = i / plane4, q = i - img * plane4;
        const float4* src = reinterpret_cast<const float4*>(in) + img * 3 * plane4 + q;
        float4* dst = reinterpret_cast<float4*>(out) + img * 3 * plane4 + q;
        float4 c0 = ld_stream(src), c1 = ld_stream(src + plane4), c2 = ld_stream(src + 2 * plane4);
        float4 o0, o1, o2;
        convert_px<kToIpt>(c0.x, c1.x, c2.x, A, B, o0.x, o1.x, o2.x);
        convert_px<kToIpt>(c0.y, c1.y, c2.y, A, B, o0.y, o1.y, o2.y);
        convert_px<kToIpt>(c0.z, c1.z, c2.z, A, B, o0.z, o1.z, o2.z);
        convert_px<kToIpt>(c0.w, c1.w, c2.w, A, B, o0.w, o1.w, o2.w);
        st_stream(dst, o0);
        st_stream(dst + plane4, o1);
        st_stream(dst + 2 * plane4, o2);
    }
}

template <bool kToIpt>
__global__ void __launch_bounds__(256) colorspace_scalar(const float* __restrict__ in,
                                                         float* __restrict__ out, int64_t n_img,
                                                         int64_t plane, Mat3 A, Mat3 B) {
    const int64_t total = n_img * plane;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t img = i / plane, q = i - img * plane;
        const float* src = in + img * 3 * plane + q;
        float* dst = out + img * 3 * plane + q;
        float o0, o1, o2;
        convert_px<kToIpt>(src[0], src[plane], src[2 * plane], A, B, o0, o1, o2);
        dst[0] = o0;
        dst[plane] = o1;
        dst[2 * plane] = o2;
    }
}

template <bool kToIpt>
int launch_colorspace(const float* in, float* out, int64_t n_img, int64_t plane, const float* a,
                      const float* b, void* stream) {
    DCTA_REQUIRE(in && out && a && b, "colorspace: null pointer");
    DCTA_REQUIRE(n_img >= 0 && plane >= 0, "colorspace: negative size");
    if (n_img == 0 || plane == 0) return DCTA_OK;
    Mat3 A, B;
    for (int i = 0; i < 9; ++i) {
        A.m[i] = a[i];
        B.m[i] = b[i];
    }
    const bool vec = (plane % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec) {
        int grid = grid_for(n_img * (plane / 4), 256);
        colorspace_vec4<kToIpt><<<grid, 256, 0, as_stream(stream)>>>(in, out, n_img, plane / 4, A, B);
    } else {
        int grid = grid_for(n_img * plane, 256);
        colorspace_scalar<kToIpt><<<grid, 256, 0, as_stream(stream)>>>(in, out, n_img, plane, A, B);
    }
    return check_launch("colorspace");
}

// 8-bit pixels <-> unit-range floats (the conversions the reference's callers do around the path; common.cuh)
__global__ void __launch_bounds__(256) u8_to_unit_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int64_t n) {
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        st_stream(reinterpret_cast<float4*>(out) + i, ld_px4(in, i));
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[n4 * 4 + threadIdx.x] = u8_to_unit(in[n4 * 4 + threadIdx.x]);
}
__global__ void __launch_bounds__(256) unit_to_u8_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, int64_t n) {
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        st_px4(out, i, ld_stream(reinterpret_cast<const float4*>(in) + i));
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[n4 * 4 + threadIdx.x] = (uint8_t)unit_to_u8(in[n4 * 4 + threadIdx.x]);
}

}  // namespace dcta

extern "C" int dcta_u8_to_unit_f32(const uint8_t* in, float* out, int64_t n, void* stream) {
    DCTA_REQUIRE(in && out && n >= 0, "u8_to_unit_f32: bad arguments");
    DCTA_REQUIRE((reinterpret_cast<uintptr_t>(in) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "u8_to_unit_f32: needs a 4-byte aligned input and a 16-byte aligned output");
    if (n == 0) return DCTA_OK;
    dcta::u8_to_unit_kernel<<<dcta::grid_for(n / 4 + 1, 256), 256, 0, dcta::as_stream(stream)>>>(in, out, n);
    return dcta::check_launch("u8_to_unit_f32");
}

extern "C" int dcta_unit_f32_to_u8(const float* in, uint8_t* out, int64_t n, void* stream) {
    DCTA_REQUIRE(in && out && n >= 0, "unit_f32_to_u8: bad arguments");
    DCTA_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
                 "unit_f32_to_u8: needs a 16-byte aligned input and a 4-byte aligned output");
    if (n == 0) return DCTA_OK;
    dcta::unit_to_u8_kernel<<<dcta::grid_for(n / 4 + 1, 256), 256, 0, dcta::as_stream(stream)>>>(in, out, n);
    return dcta::check_launch("unit_f32_to_u8");
}

extern "C" int dcta_rgb_to_ipt(const float* rgb, float* ipt, int64_t n_img, int64_t plane,
                               const float* m_rgb2lms_host, const float* m_ipt_host, void* stream) {
    return dcta::launch_colorspace<true>(rgb, ipt, n_img, plane, m_rgb2lms_host, m_ipt_host, stream);
}

extern "C" int dcta_ipt_to_rgb(const float* ipt, float* rgb, int64_t n_img, int64_t plane,
                               const float* m_ipt_inv_host, const float* m_lms2rgb_host,
                               void* stream) {
    return dcta::launch_colorspace<false>(ipt, rgb, n_img, plane, m_ipt_inv_host, m_lms2rgb_host,
                                          stream);
}
