// Antialiased bilinear resize, the `crop` step of the reference's data loader (dataset.py:59-73:
// torchvision.transforms.Resize(min(h, w), antialias=True) on a float tensor = F.interpolate(mode="bilinear",
// antialias=True, align_corners=False)): a separable triangle filter whose support grows with the down-scaling factor,
// weights normalised per output sample.  One thread per output pixel (all channels), weights formed on the fly.
#include "common.cuh"

namespace dcta {

struct AaAxis {
    float scale, support, invscale;
    int in_size;
};

__device__ __forceinline__ void aa_window(const AaAxis& ax, int i, int& xmin, int& xsize, float& center) {
    center = ax.scale * ((float)i + 0.5f);
    xmin = max((int)(center - ax.support + 0.5f), 0);
    xsize = min((int)(center + ax.support + 0.5f), ax.in_size) - xmin;
}
__device__ __forceinline__ float aa_weight(const AaAxis& ax, int j, int xmin, float center) {
    const float x = ((float)(j + xmin) - center + 0.5f) * ax.invscale;
    const float a = fabsf(x);
    return a < 1.0f ? 1.0f - a : 0.0f;
}

__device__ __forceinline__ float px_in(const float* p, int64_t i) { return __ldg(p + i); }
__device__ __forceinline__ float px_in(const uint8_t* p, int64_t i) { return u8_to_unit(__ldg(p + i)); }

template <typename TIn>
__global__ void __launch_bounds__(256) resize_aa_kernel(const TIn* __restrict__ in, float* __restrict__ out, int64_t n_planes,
                                                        int ih, int iw, int oh, int ow, AaAxis ay, AaAxis ax) {
    const int64_t total = (int64_t)oh * ow;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int oy = (int)(i / ow), oxp = (int)(i - (int64_t)oy * ow);
        int xmin, xsize, ymin, ysize;
        float cx, cy;
        aa_window(ax, oxp, xmin, xsize, cx);
        aa_window(ay, oy, ymin, ysize, cy);
        float wx_total = 0.f, wy_total = 0.f;
        for (int j = 0; j < xsize; ++j) wx_total += aa_weight(ax, j, xmin, cx);
        for (int j = 0; j < ysize; ++j) wy_total += aa_weight(ay, j, ymin, cy);
        for (int64_t pl = 0; pl < n_planes; ++pl) {
            const TIn* src = in + pl * (int64_t)ih * iw;
            float acc = 0.f;
            for (int jy = 0; jy < ysize; ++jy) {
                const TIn* row = src + (int64_t)(ymin + jy) * iw + xmin;
                float r = 0.f;
                for (int jx = 0; jx < xsize; ++jx) r += (aa_weight(ax, jx, xmin, cx) / wx_total) * px_in(row, jx);
                acc += (aa_weight(ay, jy, ymin, cy) / wy_total) * r;
            }
            out[pl * total + i] = acc;
        }
    }
}

static AaAxis make_axis(int in_size, int out_size) {
    AaAxis a;
    a.scale = (float)in_size / (float)out_size;
    a.support = a.scale >= 1.0f ? a.scale : 1.0f;          // interp_size / 2 = 1 for the triangle filter
    a.invscale = a.scale >= 1.0f ? 1.0f / a.scale : 1.0f;
    a.in_size = in_size;
    return a;
}

}  // namespace dcta

using namespace dcta;

extern "C" int dcta_resize_bilinear_aa(const float* in, float* out, int64_t n_planes, int ih, int iw, int oh, int ow,
                                       void* stream) {
    DCTA_REQUIRE(in && out && ih > 0 && iw > 0 && oh > 0 && ow > 0 && n_planes >= 0, "resize_bilinear_aa: bad args");
    if (n_planes == 0) return DCTA_OK;
    resize_aa_kernel<float><<<grid_for((int64_t)oh * ow, 256), 256, 0, as_stream(stream)>>>(in, out, n_planes, ih, iw, oh, ow,
                                                                                           make_axis(ih, oh), make_axis(iw, ow));
    return check_launch("resize_bilinear_aa");
}

extern "C" int dcta_resize_bilinear_aa_u8(const uint8_t* in, float* out, int64_t n_planes, int ih, int iw, int oh, int ow,
                                          void* stream) {
    DCTA_REQUIRE(in && out && ih > 0 && iw > 0 && oh > 0 && ow > 0 && n_planes >= 0, "resize_bilinear_aa_u8: bad args");
    if (n_planes == 0) return DCTA_OK;
    resize_aa_kernel<uint8_t><<<grid_for((int64_t)oh * ow, 256), 256, 0, as_stream(stream)>>>(in, out, n_planes, ih, iw, oh, ow,
                                                                                             make_axis(ih, oh), make_axis(iw, ow));
    return check_launch("resize_bilinear_aa_u8");
}
