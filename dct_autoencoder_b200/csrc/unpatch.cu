// Un-patchify: packed token rows -> zero-filled coefficient planes
// (reference: feature_extraction_dct_autoencoder.py:607-656 revert_patching).
// Gather formulation: a tiny (image, channel, tile) -> token map is scattered first, then every
// plane element is written exactly once, coalesced (no memset + scatter, no write races).
#include "common.cuh"

namespace dcta {

// slot_map[img, c, h, w] = max flat token index at that position == the LAST token in sequence
// order, which is what the reference's sequential "image[channel,h,w] = token" loop keeps (FE:639-643).
__global__ void slot_map_kernel(const int64_t* __restrict__ channels,
                                const int64_t* __restrict__ positions,
                                const int64_t* __restrict__ image_ids,
                                const uint8_t* __restrict__ pad,
                                const int32_t* __restrict__ row_img_base, int n_rows, int s,
                                int64_t n_img, int C, int th, int tw, int32_t* __restrict__ slot_map) {
    const int64_t total = (int64_t)n_rows * s;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        if (pad && pad[t]) continue;
        const int row = (int)(t / s);
        const int64_t img = row_img_base[row] + (image_ids ? image_ids[t] : 0);
        const int64_t c = channels[t], h = positions[2 * t], w = positions[2 * t + 1];
        if (img < 0 || img >= n_img || c < 0 || c >= C || h < 0 || h >= th || w < 0 || w >= tw) continue;
        atomicMax(slot_map + ((img * C + c) * th + h) * tw + w, (int32_t)t);
    }
}

// one thread per 4 consecutive plane columns (or 1 in the scalar variant)
template <int kVec>
__global__ void __launch_bounds__(256) unpatchify_kernel(const float* __restrict__ patches,
                                                         const int32_t* __restrict__ slot_map,
                                                         const int32_t* __restrict__ img_sel,
                                                         int64_t n_sel, int C, int th, int tw, int p,
                                                         int rows, int cols,
                                                         float* __restrict__ planes) {
    const int z = p * p;
    const int cols_v = cols / kVec;
    const int64_t total = n_sel * C * rows * cols_v;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int xv = (int)(i % cols_v);
        int64_t r = i / cols_v;
        const int y = (int)(r % rows);
        r /= rows;
        const int c = (int)(r % C);
        const int64_t sel = r / C;
        const int64_t img = img_sel ? img_sel[sel] : sel;
        const int ty = y / p, py = y - ty * p;
        float v[kVec];
#pragma unroll
        for (int j = 0; j < kVec; ++j) {
            const int x = xv * kVec + j;
            const int tx = x / p, px = x - tx * p;
            float val = 0.0f;
            if (ty < th && tx < tw) {
                const int32_t slot = __ldg(slot_map + ((img * C + c) * th + ty) * tw + tx);
                if (slot >= 0) val = __ldg(patches + (int64_t)slot * z + py * p + px);
            }
            v[j] = val;
        }
        float* dst = planes + ((sel * C + c) * rows + y) * (int64_t)cols + (int64_t)xv * kVec;
        if (kVec == 4) st_stream(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
        else dst[0] = v[0];
    }
}

// planes -> token grid (FE:374-380 rearrange + the max_patch clip of FE:393-394); used when the
// transform is supplied by the caller (the reference lets _transform_image_in be replaced).
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ planes,
                                                       float* __restrict__ tiles, int64_t n_img,
                                                       int C, int rows, int cols, int th, int tw,
                                                       int p) {
    const int z = p * p;
    const int64_t total = n_img * th * tw * C * z;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i % z);
        int64_t r = i / z;
        const int c = (int)(r % C);
        r /= C;
        const int x = (int)(r % tw);
        r /= tw;
        const int y = (int)(r % th);
        const int64_t img = r / th;
        const int py = e / p, px = e - py * p;
        tiles[i] = __ldg(planes + ((img * C + c) * rows + (y * p + py)) * (int64_t)cols + x * p + px);
    }
}

}  // namespace dcta

using namespace dcta;

extern "C" int dcta_patchify(const float* planes, float* tiles, int64_t n_img, int channels_n,
                             int rows, int cols, int th, int tw, int p, void* stream) {
    DCTA_REQUIRE(planes && tiles, "patchify: null pointer");
    DCTA_REQUIRE(n_img >= 0 && channels_n > 0 && p > 0 && th > 0 && tw > 0 && th * p <= rows && tw * p <= cols,
                 "patchify: bad sizes");
    if (n_img == 0) return DCTA_OK;
    const int64_t total = n_img * th * tw * channels_n * p * p;
    patchify_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(planes, tiles, n_img, channels_n, rows, cols, th, tw, p);
    return check_launch("patchify");
}

extern "C" int dcta_build_slot_map(const int64_t* channels, const int64_t* positions,
                                   const int64_t* image_ids, const uint8_t* key_pad_mask,
                                   const int32_t* row_img_base, int n_rows, int s, int64_t n_img,
                                   int channels_n, int th, int tw, int32_t* slot_map, void* stream) {
    DCTA_REQUIRE(channels && positions && row_img_base && slot_map, "build_slot_map: null pointer");
    DCTA_REQUIRE(n_rows >= 0 && s > 0 && n_img >= 0 && channels_n > 0 && th > 0 && tw > 0,
                 "build_slot_map: bad sizes");
    DCTA_REQUIRE((int64_t)n_rows * s < (1ll << 31), "build_slot_map: more than 2^31 token slots");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(slot_map, 0xff, sizeof(int32_t) * n_img * channels_n * th * tw, st);
    if (n_rows == 0) return DCTA_OK;
    slot_map_kernel<<<grid_for((int64_t)n_rows * s, 256), 256, 0, st>>>(
        channels, positions, image_ids, key_pad_mask, row_img_base, n_rows, s, n_img, channels_n, th, tw, slot_map);
    return check_launch("build_slot_map");
}

extern "C" int dcta_unpatchify(const float* patches, const int32_t* slot_map, const int32_t* img_sel,
                               int64_t n_sel, int channels_n, int th, int tw, int p, int rows,
                               int cols, float* planes, void* stream) {
    DCTA_REQUIRE(patches && slot_map && planes, "unpatchify: null pointer");
    DCTA_REQUIRE(n_sel >= 0 && channels_n > 0 && th > 0 && tw > 0 && p > 0 && rows > 0 && cols > 0,
                 "unpatchify: bad sizes");
    if (n_sel == 0) return DCTA_OK;
    const bool vec = (cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(planes) & 15) == 0);
    const int64_t total = n_sel * channels_n * rows * (vec ? cols / 4 : cols);
    if (vec)
        unpatchify_kernel<4><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(patches, slot_map, img_sel, n_sel, channels_n, th, tw, p, rows, cols, planes);
    else
        unpatchify_kernel<1><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(patches, slot_map, img_sel, n_sel, channels_n, th, tw, p, rows, cols, planes);
    return check_launch("unpatchify");
}
