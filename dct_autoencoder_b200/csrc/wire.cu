// Compact wire format of quantised tokens.
//
// The reference serialises an encoded image as a list of per-token dicts {c, h, w, data:[codes]}
// built with one .item() per field (dct_patches.py:54-87) for the autoregressive dataset builder
// (prepare_autoregressive_dataset.py:51-66).  Here every token becomes one fixed-size record
//
//     bytes 0..1   little-endian u16:  channel << 12 | h << 6 | w        (channel < 16, h, w < 64)
//     bytes 2..    the token's c code words, d bits each, most significant bit first, concatenated;
//                  the last byte is zero-padded in its low bits           (rec = 2 + ceil(c*d/8) bytes)
//
// written by one kernel for the whole batch (padding slots included, so the output is a dense
// (rows, s, rec) byte tensor); the number of tokens of every image of every row is counted in the
// same launch so that the host can cut the rows into per-image byte strings without looking at the
// bookkeeping tensors.  HBM-bound: reads 8c + 24 (+9) bytes per token, writes rec.
#include "common.cuh"

namespace dcta {

constexpr int kWireTokens = 128;   // tokens per CTA (one thread each)

__device__ __forceinline__ void coop_copy16(void* dst, const void* src, int bytes, int tid, int nthreads) {
    // both 16-byte aligned, bytes a multiple of 8
    const int n16 = bytes >> 4;
    for (int i = tid; i < n16; i += nthreads) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    if ((bytes & 15) && tid == 0)
        reinterpret_cast<uint2*>(dst)[n16 * 2] = reinterpret_cast<const uint2*>(src)[n16 * 2];
}

__global__ void __launch_bounds__(kWireTokens)
wire_pack_kernel(const int64_t* __restrict__ codes, const int64_t* __restrict__ positions,
                 const int64_t* __restrict__ channels, const int64_t* __restrict__ image_ids,
                 const uint8_t* __restrict__ key_pad_mask, int64_t n_tok, int s, int c, int d, int rec,
                 uint8_t* __restrict__ out, int32_t* __restrict__ counts) {
    extern __shared__ __align__(16) uint8_t smem[];
    int64_t* s_codes = reinterpret_cast<int64_t*>(smem);
    uint8_t* s_rec = smem + (size_t)kWireTokens * c * 8;
    const int tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * kWireTokens;
    const int n_here = (int)min((int64_t)kWireTokens, n_tok - t0);
    coop_copy16(s_codes, codes + t0 * c, n_here * c * 8, tid, kWireTokens);
    __syncthreads();
    if (tid < n_here) {
        const int64_t t = t0 + tid;
        const longlong2 hw = reinterpret_cast<const longlong2*>(positions)[t];
        const unsigned head = ((unsigned)channels[t] & 15u) << 12 | ((unsigned)hw.x & 63u) << 6 | ((unsigned)hw.y & 63u);
        uint8_t* r = s_rec + tid * rec;
        r[0] = (uint8_t)head;
        r[1] = (uint8_t)(head >> 8);
        const unsigned long long mask = (1ull << d) - 1ull;
        unsigned long long acc = 0;     // pending bits, right-aligned
        int have = 0, o = 2;
        for (int i = 0; i < c; ++i) {
            unsigned long long v = (unsigned long long)s_codes[tid * c + i] & mask;
            // d <= 32 and have < 8, so acc never overflows 64 bits
            acc = (acc << d) | v;
            have += d;
            while (have >= 8) {
                have -= 8;
                r[o++] = (uint8_t)(acc >> have);
            }
            acc &= (1ull << have) - 1ull;
        }
        if (have) r[o] = (uint8_t)(acc << (8 - have));
        if (counts != nullptr && !key_pad_mask[t]) atomicAdd(&counts[(t / s) * s + image_ids[t]], 1);
    }
    __syncthreads();
    // kWireTokens * rec is a multiple of 4, so every CTA's slice of `out` starts 4-byte aligned
    uint8_t* dst = out + t0 * rec;
    const int bytes = n_here * rec;
    const int n4 = bytes >> 2;
    for (int i = tid; i < n4; i += kWireTokens) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(s_rec)[i];
    for (int i = (n4 << 2) + tid; i < bytes; i += kWireTokens) dst[i] = s_rec[i];
}

__global__ void __launch_bounds__(kWireTokens)
wire_unpack_kernel(const uint8_t* __restrict__ in, int64_t n_tok, int c, int d, int rec,
                   int64_t* __restrict__ codes, int64_t* __restrict__ positions, int64_t* __restrict__ channels) {
    extern __shared__ __align__(16) uint8_t smem[];
    int64_t* s_codes = reinterpret_cast<int64_t*>(smem);
    uint8_t* s_rec = smem + (size_t)kWireTokens * c * 8;
    const int tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * kWireTokens;
    const int n_here = (int)min((int64_t)kWireTokens, n_tok - t0);
    const uint8_t* src = in + t0 * rec;
    const int bytes = n_here * rec;
    const int n4 = bytes >> 2;
    for (int i = tid; i < n4; i += kWireTokens) reinterpret_cast<uint32_t*>(s_rec)[i] = reinterpret_cast<const uint32_t*>(src)[i];
    for (int i = (n4 << 2) + tid; i < bytes; i += kWireTokens) s_rec[i] = src[i];
    __syncthreads();
    if (tid < n_here) {
        const int64_t t = t0 + tid;
        const uint8_t* r = s_rec + tid * rec;
        const unsigned head = (unsigned)r[0] | (unsigned)r[1] << 8;
        channels[t] = head >> 12;
        reinterpret_cast<longlong2*>(positions)[t] = make_longlong2((head >> 6) & 63u, head & 63u);
        unsigned long long acc = 0;
        int have = 0, o = 2;
        for (int i = 0; i < c; ++i) {
            while (have < d) {
                acc = (acc << 8) | r[o++];
                have += 8;
            }
            have -= d;
            s_codes[tid * c + i] = (int64_t)(acc >> have);
            acc &= (1ull << have) - 1ull;
        }
    }
    __syncthreads();
    coop_copy16(codes + t0 * c, s_codes, n_here * c * 8, tid, kWireTokens);
}

static int wire_check(int64_t n_tok, int c, int d, int rec, size_t* smem) {
    DCTA_REQUIRE(n_tok >= 0 && c >= 1 && d >= 1 && d <= 32, "wire: need n_tok >= 0, c >= 1, 1 <= d <= 32");
    DCTA_REQUIRE(rec == 2 + (c * d + 7) / 8, "wire: rec must be 2 + ceil(c*d/8) = %d, got %d", 2 + (c * d + 7) / 8, rec);
    *smem = (size_t)kWireTokens * c * 8 + (((size_t)kWireTokens * rec + 15) & ~(size_t)15);
    DCTA_REQUIRE(*smem <= 48 * 1024, "wire: c = %d code words of %d bits do not fit the staging buffer", c, d);
    return DCTA_OK;
}

}  // namespace dcta

using namespace dcta;

extern "C" int dcta_wire_pack(const int64_t* codes, const int64_t* positions, const int64_t* channels,
                              const int64_t* image_ids, const uint8_t* key_pad_mask, int64_t n_rows, int s,
                              int c, int d, int rec, uint8_t* out, int32_t* counts, void* stream) {
    const int64_t n_tok = n_rows * s;
    size_t smem;
    if (int rc = wire_check(n_tok, c, d, rec, &smem)) return rc;
    DCTA_REQUIRE(codes && positions && channels && out, "wire_pack: null pointer");
    DCTA_REQUIRE(counts == nullptr || (image_ids && key_pad_mask), "wire_pack: counts need image_ids and key_pad_mask");
    if (n_tok == 0) return DCTA_OK;
    if (counts) cudaMemsetAsync(counts, 0, sizeof(int32_t) * n_tok, as_stream(stream));
    wire_pack_kernel<<<(unsigned)ceil_div(n_tok, kWireTokens), kWireTokens, smem, as_stream(stream)>>>(
        codes, positions, channels, image_ids, key_pad_mask, n_tok, s, c, d, rec, out, counts);
    return check_launch("wire_pack_kernel");
}

extern "C" int dcta_wire_unpack(const uint8_t* in, int64_t n_tok, int c, int d, int rec, int64_t* codes,
                                int64_t* positions, int64_t* channels, void* stream) {
    size_t smem;
    if (int rc = wire_check(n_tok, c, d, rec, &smem)) return rc;
    DCTA_REQUIRE(in && codes && positions && channels, "wire_unpack: null pointer");
    DCTA_REQUIRE((reinterpret_cast<uintptr_t>(in) & 3) == 0, "wire_unpack: input must be 4-byte aligned");
    if (n_tok == 0) return DCTA_OK;
    wire_unpack_kernel<<<(unsigned)ceil_div(n_tok, kWireTokens), kWireTokens, smem, as_stream(stream)>>>(
        in, n_tok, c, d, rec, codes, positions, channels);
    return check_launch("wire_unpack_kernel");
}
