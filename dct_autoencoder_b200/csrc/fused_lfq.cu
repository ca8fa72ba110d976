// Fused encode-to-codes and decode-from-codes kernels for the PatchNorm + LFQ bottleneck
// (reference: feature_extraction_dct_autoencoder.py:437-452, 516-605; patchnorm.py:157-177;
//  lfq.py:105-134, 168-187; feature_extraction_dct_autoencoder.py:635-653).
//
// The staged API materialises gathered patches, normalised patches, quantised patches and
// de-normalised patches (4 x 2.4 MB per image, each written once and read once).  When the
// quantiser is a projection-free LFQ and PatchNorm is frozen, the same arithmetic can run in
// registers between two tensors that must exist anyway:
//   encode: token grid (GEMM output) --[gather by rank, normalise, sign, bit-pack]--> codes
//   decode: codes --[slot map, unpack bit, *std + median, fp16 split]--> coefficient planes (GEMM input)
// Every fp32 operation is the one the staged kernels execute (explicit _rn intrinsics, same
// order), so codes and planes are bit-identical to the staged path.
#include <cuda_fp16.h>

#include "common.cuh"
#include "lfq_norm.cuh"

namespace dcta {

// one warp per output slot (row, s)
__global__ void __launch_bounds__(256) pack_codes_kernel(
    const float* __restrict__ tiles, const int32_t* __restrict__ order, const dcta_segment* __restrict__ segs,
    const int32_t* __restrict__ row_seg_start, int n_rows, int s, int tw, int channels, int n_tok_img,
    LfqNormParams q, int64_t* __restrict__ codes, int64_t* __restrict__ positions, int64_t* __restrict__ channels_out,
    int64_t* __restrict__ image_ids, uint8_t* __restrict__ key_pad_mask) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = (int64_t)n_rows * s;
    const int z = q.z;
    for (int64_t slot = warp0; slot < total; slot += n_warps) {
        const int row = (int)(slot / s);
        const int off = (int)(slot - (int64_t)row * s);
        int lo = row_seg_start[row], hi = row_seg_start[row + 1];
        int seg = -1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const int so = segs[mid].offset;
            if (off < so) hi = mid;
            else if (off >= so + segs[mid].k) lo = mid + 1;
            else { seg = mid; break; }
        }
        const float* src = nullptr;
        int ph = 0, pw = 0, pc = 0, image_id = 0;
        if (seg >= 0) {
            const dcta_segment sg = segs[seg];
            const int tok = order[sg.img * n_tok_img + (off - sg.offset)];
            src = tiles + (sg.img * n_tok_img + tok) * z;
            pc = tok % channels;
            const int tile = tok / channels;
            ph = tile / tw;
            pw = tile - ph * tw;
            image_id = sg.image_id;
        }
        // padding slots are quantised like the reference does: zeros normalised with the statistics
        // at (0, 0, 0)  (patchnorm.py:157-161 runs on padding rows too)
        const int64_t pid = ((int64_t)pc * q.H + ph) * q.W + pw;
        const float* ms = q.median + pid * z;
        const float* bs = q.b + pid * z;
        // sign bits of up to 256 elements: lane k keeps word k (bit (e & 31) of word e >> 5)
        unsigned my_word = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int e = lane + 32 * k;
            bool bit = false;
            if (k * 32 < z && e < z) {
                const float xv = src ? __ldg(src + e) : 0.0f;
                const float sd = __fadd_rn(__fmul_rn(__ldg(bs + e), kSqrt2f), q.eps);
                const float diff = __fsub_rn(xv, __ldg(ms + e));
                // bit = clamp(diff / sd) > 0 (patchnorm.py:161-163, lfq.py:175).  For a finite positive
                // sd and a numerator far from the denormal range the quotient has the numerator's sign;
                // only otherwise is the IEEE division (and the clamp) evaluated.
                if (sd > 0.0f && sd < 1e30f && fabsf(diff) > 1e-30f && q.lo < 0.0f && q.hi > 0.0f) {
                    bit = diff > 0.0f;
                } else {
                    float y = __fdiv_rn(diff, sd);
                    y = y < q.lo ? q.lo : (y > q.hi ? q.hi : y);
                    bit = y > 0.0f;
                }
            }
            const unsigned w = __ballot_sync(0xffffffffu, bit);
            if (lane == k) my_word = w;
        }
        // lfq.py:187: code[cb] = sum_i bit(cb*d + i) << (d-1-i); lane cb assembles codebook cb
        for (int cb0 = 0; cb0 < q.c; cb0 += 32) {
            const int cb = cb0 + lane;
            unsigned long long code = 0;
            for (int i = 0; i < q.d; ++i) {
                const int e = min(cb * q.d + i, z - 1);
                const unsigned w = __shfl_sync(0xffffffffu, my_word, e >> 5);
                code = (code << 1) | ((w >> (e & 31)) & 1u);
            }
            if (cb < q.c) codes[slot * q.c + cb] = (int64_t)code;
        }
        if (lane == 0) {
            positions[slot * 2] = ph;
            positions[slot * 2 + 1] = pw;
            channels_out[slot] = pc;
            if (image_ids) image_ids[slot] = image_id;
            if (key_pad_mask) key_pad_mask[slot] = seg < 0;
        }
    }
}

// Vectorised variant (z % 4 == 0, z <= 256, d <= 32).  A warp takes 32 consecutive slots: the per-slot
// metadata chain (segment search -> order -> token address; three dependent global loads) runs ONCE,
// lane-parallel, and the bookkeeping outputs are written coalesced; then the 32 tokens are streamed two at
// a time with 128-bit loads (one quad of 4 elements per lane and pass), sign nibbles OR-reduced inside
// groups of 8 lanes into element-ordered 32-bit words, and the codes cut out of the bit string with a
// funnel shift + bit reversal.
struct TokenBits { unsigned w[2]; };

// tame: every PatchNorm b is finite and in [0, 1e18] (checked on the device by b_tame_kernel), i.e. the divisor
// b*sqrt2 + eps is a positive normal number: the sign of the clamped quotient is then the sign of x - median
// unless that difference is tiny (< 1e-20: the quotient could underflow) -- only those quads read b and divide.
__global__ void __launch_bounds__(256) b_tame_kernel(const float* __restrict__ b, int64_t n, int32_t* __restrict__ flag) {
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(b + i);
        bad |= !(v >= 0.0f && v <= 1e18f);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicExch(flag, 0);
}

__device__ __forceinline__ void load_token(const float4* src, const float4* ms, int lane, int z4,
                                           float4 (&xv)[2], float4 (&mv)[2]) {
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int qd = lane + 32 * pass;
        if (qd < z4) {
            xv[pass] = src ? __ldg(src + qd) : make_float4(0.f, 0.f, 0.f, 0.f);
            mv[pass] = __ldg(ms + qd);
        }
    }
}

__device__ __forceinline__ TokenBits token_bits(const float4 (&xv)[2], const float4 (&mv)[2], const float4* bs,
                                                bool tame, int lane, int z4, const LfqNormParams& q) {
    TokenBits r;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        unsigned nib = 0;
        if (lane + 32 * pass < z4) {
            bool fast = false;
            if (tame) {
                const float d0 = xv[pass].x - mv[pass].x, d1 = xv[pass].y - mv[pass].y;
                const float d2 = xv[pass].z - mv[pass].z, d3 = xv[pass].w - mv[pass].w;
                fast = fminf(fminf(fabsf(d0), fabsf(d1)), fminf(fabsf(d2), fabsf(d3))) > 1e-20f && q.eps > 1e-12f &&
                       q.lo < 0.0f && q.hi > 0.0f;
                if (fast) nib = (d0 > 0.0f) | ((d1 > 0.0f) << 1) | ((d2 > 0.0f) << 2) | ((d3 > 0.0f) << 3);
            }
            if (!fast) {
                const float4 bv = __ldg(bs + lane + 32 * pass);
                nib = norm_sign_bit(xv[pass].x, mv[pass].x, bv.x, q) | (norm_sign_bit(xv[pass].y, mv[pass].y, bv.y, q) << 1) |
                      (norm_sign_bit(xv[pass].z, mv[pass].z, bv.z, q) << 2) | (norm_sign_bit(xv[pass].w, mv[pass].w, bv.w, q) << 3);
            }
        }
        unsigned v = nib << (4 * (lane & 7));
        v |= __shfl_xor_sync(0xffffffffu, v, 1);
        v |= __shfl_xor_sync(0xffffffffu, v, 2);
        v |= __shfl_xor_sync(0xffffffffu, v, 4);
        r.w[pass] = v;          // lanes 8j .. 8j+7 hold elements 128*pass + 32*j .. +31 (bit = element & 31)
    }
    return r;
}

// lfq.py:187: code[cb] = bits [cb*d, cb*d + d) of the element-ordered string, first element = MSB
__device__ __forceinline__ void write_codes(const TokenBits& tb, int lane, const LfqNormParams& q, unsigned dmask,
                                            int64_t* __restrict__ dst) {
    for (int cb0 = 0; cb0 < q.c; cb0 += 32) {
        const int cb = min(cb0 + lane, q.c - 1);
        const int e0 = cb * q.d;
        const int k0 = e0 >> 5, k1 = min(k0 + 1, 7), sh = e0 & 31;
        const unsigned a0 = __shfl_sync(0xffffffffu, tb.w[0], 8 * (k0 & 3));
        const unsigned a1 = __shfl_sync(0xffffffffu, tb.w[1], 8 * (k0 & 3));
        const unsigned b0 = __shfl_sync(0xffffffffu, tb.w[0], 8 * (k1 & 3));
        const unsigned b1 = __shfl_sync(0xffffffffu, tb.w[1], 8 * (k1 & 3));
        const unsigned wlo = (k0 >> 2) ? a1 : a0;
        const unsigned whi = (k1 >> 2) ? b1 : b0;
        const unsigned bits = __funnelshift_r(wlo, whi, sh) & dmask;      // LSB = first element
        const unsigned code = __brev(bits) >> (32 - q.d);
        if (cb0 + lane < q.c) dst[cb] = (int64_t)code;
    }
}

__global__ void __launch_bounds__(256, 4) pack_codes_vec_kernel(
    const float* __restrict__ tiles, const int32_t* __restrict__ order, const dcta_segment* __restrict__ segs,
    const int32_t* __restrict__ row_seg_start, int n_rows, int s, int tw, int channels, int n_tok_img,
    float inv_channels, float inv_tw, LfqNormParams q, const int32_t* __restrict__ tame_flag,
    int64_t* __restrict__ codes, int64_t* __restrict__ positions,
    int64_t* __restrict__ channels_out, int64_t* __restrict__ image_ids, uint8_t* __restrict__ key_pad_mask) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = (int64_t)n_rows * s;
    const int z4 = q.z >> 2;
    const unsigned dmask = q.d == 32 ? 0xffffffffu : ((1u << q.d) - 1u);
    const bool tame = tame_flag != nullptr && __ldg(tame_flag) != 0;
    for (int64_t slot0 = warp0 * 32; slot0 < total; slot0 += n_warps * 32) {
        // ---- metadata of slot0 + lane
        const int64_t slot = slot0 + lane;
        int64_t src_off = -1;          // element offset of the token in `tiles`, -1 = padding slot
        int pid = 0;                   // PatchNorm position (pc * H + ph) * W + pw
        if (slot < total) {
            const int row = (int)(slot / s);
            const int off = (int)(slot - (int64_t)row * s);
            int lo = row_seg_start[row], hi = row_seg_start[row + 1];
            int seg = -1;
            if (hi - lo == 1) {                       // the common case: one image per row
                seg = (off >= segs[lo].offset && off < segs[lo].offset + segs[lo].k) ? lo : -1;
            } else {
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const int so = segs[mid].offset;
                    if (off < so) hi = mid;
                    else if (off >= so + segs[mid].k) lo = mid + 1;
                    else { seg = mid; break; }
                }
            }
            int ph = 0, pw = 0, pc = 0, image_id = 0;
            if (seg >= 0) {
                const dcta_segment sg = segs[seg];
                const int tok = order[sg.img * n_tok_img + (off - sg.offset)];
                src_off = (sg.img * n_tok_img + tok) * (int64_t)q.z;
                const int tile = (int)(((float)tok + 0.5f) * inv_channels);   // exact for tok < 2^22
                pc = tok - tile * channels;
                ph = (int)(((float)tile + 0.5f) * inv_tw);
                pw = tile - ph * tw;
                image_id = sg.image_id;
            }
            // padding slots are quantised like the reference does: zeros normalised with the statistics at (0,0,0)
            pid = (pc * q.H + ph) * q.W + pw;
            reinterpret_cast<longlong2*>(positions)[slot] = make_longlong2(ph, pw);
            channels_out[slot] = pc;
            if (image_ids) image_ids[slot] = image_id;
            if (key_pad_mask) key_pad_mask[slot] = seg < 0;
        }
        // ---- the tokens, two in flight
        const int n_here = (int)min((int64_t)32, total - slot0);
        for (int t = 0; t < n_here; t += 2) {
            float4 xv[2][2], mv[2][2];
            const float4* bsp[2];
            const bool second = t + 1 < n_here;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int64_t so = __shfl_sync(0xffffffffu, src_off, t + u);
                const int pd = __shfl_sync(0xffffffffu, pid, t + u);
                bsp[u] = reinterpret_cast<const float4*>(q.b + (int64_t)pd * q.z);
                if (u == 0 || second)
                    load_token(so >= 0 ? reinterpret_cast<const float4*>(tiles + so) : nullptr,
                               reinterpret_cast<const float4*>(q.median + (int64_t)pd * q.z), lane, z4, xv[u], mv[u]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 0 || second) {
                    const TokenBits tb = token_bits(xv[u], mv[u], bsp[u], tame, lane, z4, q);
                    write_codes(tb, lane, q, dmask, codes + (slot0 + t + u) * q.c);
                }
            }
        }
    }
}

int launch_b_tame(const float* b, int64_t n, int32_t* flag, cudaStream_t st) {
    cudaMemsetAsync(flag, 1, sizeof(int32_t), st);
    b_tame_kernel<<<grid_for(n, 1024, 1), 256, 0, st>>>(b, n, flag);
    return check_launch("b_tame");
}

// Code words of a padding slot: zeros normalised with the statistics at (0, 0, 0) like the reference does
// (patchnorm.py:157-161 runs on padding rows too).  One warp; pad_codes (c) int64.
__global__ void pad_codes_kernel(LfqNormParams q, int64_t* __restrict__ pad_codes) {
    for (int cb = threadIdx.x; cb < q.c; cb += blockDim.x) {
        unsigned long long code = 0;
        for (int i = 0; i < q.d; ++i) {
            const int e = cb * q.d + i;
            code = (code << 1) | norm_sign_bit(0.0f, q.median[e], q.b[e], q);
        }
        pad_codes[cb] = (int64_t)code;
    }
}

// Gather of the code words computed by the forward DCT epilogue (dct_fold.cu): code_grid (n_img, th, tw, C, c)
// int32 in token-grid order -> codes (rows, s, c) int64 in packed, sorted order, + the bookkeeping outputs.
// A warp takes 32 consecutive slots: the metadata chain runs once, lane-parallel; the 32 * c code words are
// then copied with the lanes on consecutive OUTPUT words (fully coalesced 256-byte stores, independent loads).
__global__ void __launch_bounds__(256) pack_codes_grid_kernel(
    const int32_t* __restrict__ code_grid, const int32_t* __restrict__ order, const dcta_segment* __restrict__ segs,
    const int32_t* __restrict__ row_seg_start, int n_rows, int s, int tw, int channels, int n_tok_img,
    float inv_channels, float inv_tw, int c, float inv_c, const int64_t* __restrict__ pad_codes,
    int64_t* __restrict__ codes, int64_t* __restrict__ positions, int64_t* __restrict__ channels_out,
    int64_t* __restrict__ image_ids, uint8_t* __restrict__ key_pad_mask) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = (int64_t)n_rows * s;
    for (int64_t slot0 = warp0 * 32; slot0 < total; slot0 += n_warps * 32) {
        const int64_t slot = slot0 + lane;
        int64_t src_off = -1;                  // first code word of the token in code_grid, -1 = padding slot
        if (slot < total) {
            const int row = (int)(slot / s);
            const int off = (int)(slot - (int64_t)row * s);
            int lo = row_seg_start[row], hi = row_seg_start[row + 1];
            int seg = -1;
            if (hi - lo == 1) {
                seg = (off >= segs[lo].offset && off < segs[lo].offset + segs[lo].k) ? lo : -1;
            } else {
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const int so = segs[mid].offset;
                    if (off < so) hi = mid;
                    else if (off >= so + segs[mid].k) lo = mid + 1;
                    else { seg = mid; break; }
                }
            }
            int ph = 0, pw = 0, pc = 0, image_id = 0;
            if (seg >= 0) {
                const dcta_segment sg = segs[seg];
                const int tok = order[sg.img * n_tok_img + (off - sg.offset)];
                const int tile = (int)(((float)tok + 0.5f) * inv_channels);   // exact for tok < 2^22
                pc = tok - tile * channels;
                ph = (int)(((float)tile + 0.5f) * inv_tw);
                pw = tile - ph * tw;
                image_id = sg.image_id;
                src_off = (sg.img * n_tok_img + tok) * (int64_t)c;
            }
            reinterpret_cast<longlong2*>(positions)[slot] = make_longlong2(ph, pw);
            channels_out[slot] = pc;
            if (image_ids) image_ids[slot] = image_id;
            if (key_pad_mask) key_pad_mask[slot] = seg < 0;
        }
        const int n_words = (int)min((int64_t)32, total - slot0) * c;
        int64_t* dst = codes + slot0 * c;
#pragma unroll 4
        for (int it = 0; it < c; ++it) {
            const int idx = it * 32 + lane;
            const int sl = min((int)(((float)idx + 0.5f) * inv_c), 31), wd = idx - sl * c;
            const int64_t so = __shfl_sync(0xffffffffu, src_off, sl);
            if (idx < n_words) dst[idx] = so >= 0 ? (int64_t)__ldg(code_grid + so + wd) : __ldg(pad_codes + wd);
        }
    }
}

__device__ __forceinline__ void split16f(float v, float scale, __half& h, __half& l) {
    const float s = v * scale;
    h = __float2half_rn(s);
    l = __float2half_rn(s - __half2float(h));
}

// One CTA per tile-row of one plane (p consecutive plane rows): the (image, channel, tile-row)
// decomposition is block-uniform, each thread owns 4 consecutive columns and walks the p rows, so the
// slot-map entry is loaded once per tile and the statistics pointers advance by p per row.  Planes are
// the inverse GEMM's fp16 hi/lo operand.
__global__ void __launch_bounds__(128) decode_codes_split_kernel(
    const int64_t* __restrict__ codes, const int32_t* __restrict__ slot_map, const int32_t* __restrict__ img_sel,
    int th, int tw, int p, int rows, int cols, int ld, LfqNormParams q, __half* __restrict__ hi,
    __half* __restrict__ lo, float* __restrict__ dc, float dc_factor, float scale) {
    const int C = q.C;
    const int tile_rows = rows / p;
    const unsigned id = blockIdx.x;
    const int ty = (int)(id % (unsigned)tile_rows);
    const unsigned t = id / (unsigned)tile_rows;
    const int c = (int)(t % (unsigned)C);
    const int sel = (int)(t / (unsigned)C);
    const int64_t img = img_sel ? img_sel[sel] : sel;
    const bool row_in = ty < th;
    const int32_t* smap = slot_map + ((img * C + c) * th + (row_in ? ty : 0)) * tw;
    const float* med_row = q.median + (((int64_t)c * q.H + ty) * q.W) * q.z;   // + tx * z + py * p + px
    const float* b_row = q.b + (((int64_t)c * q.H + ty) * q.W) * q.z;
    const bool row_codebook = (q.d == p);       // one codebook per patch row (14 x 14 bits at patch 14)
    const int64_t plane_row0 = ((int64_t)(sel * C + c) * rows + (int64_t)ty * p) * ld;
    for (int xv = threadIdx.x; xv < ld / 4; xv += blockDim.x) {
        const int x0 = xv * 4;
        const int tx0 = x0 / p, px0 = x0 - tx0 * p;
        // the 4 columns touch at most two tiles
        const int tx1 = tx0 + 1;
        const int n_first = min(4, p - px0);                       // columns that belong to tile tx0
        const int32_t slot0 = (row_in && tx0 < tw) ? __ldg(smap + tx0) : -1;
        const int32_t slot1 = (row_in && n_first < 4 && tx1 < tw) ? __ldg(smap + tx1) : -1;
        for (int py = 0; py < p; ++py) {
            __half oh[4], ol[4];
            long long code0 = 0, code1 = 0;
            if (row_codebook) {
                if (slot0 >= 0) code0 = __ldg(codes + (int64_t)slot0 * q.c + py);
                if (slot1 >= 0) code1 = __ldg(codes + (int64_t)slot1 * q.c + py);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool first = j < n_first;
                const int32_t slot = first ? slot0 : slot1;
                const int tx = first ? tx0 : tx1;
                const int px = first ? px0 + j : j - n_first;
                float val = 0.0f;
                if (slot >= 0 && x0 + j < cols) {
                    const int e = py * p + px;
                    long long code;
                    int bi;
                    if (row_codebook) { code = first ? code0 : code1; bi = px; }
                    else { const int cb = e / q.d; bi = e - cb * q.d; code = __ldg(codes + (int64_t)slot * q.c + cb); }
                    const float qv = ((code >> (q.d - 1 - bi)) & 1) ? q.scale : -q.scale;     // lfq.py:118-120
                    const int pe = tx * q.z + e;
                    const float sd = __fadd_rn(__fmul_rn(__ldg(b_row + pe), kSqrt2f), q.eps);
                    val = __fadd_rn(__fmul_rn(qv, sd), __ldg(med_row + pe));                    // patchnorm.py:177
                }
                if (ty == 0 && py == 0 && x0 + j == 0) {
                    dc[sel * C + c] = val * dc_factor;
                    val = 0.0f;
                }
                split16f(val, scale, oh[j], ol[j]);
            }
            const int64_t o = (plane_row0 + (int64_t)py * ld) / 4 + xv;
            reinterpret_cast<uint2*>(hi)[o] = *reinterpret_cast<const uint2*>(oh);
            reinterpret_cast<uint2*>(lo)[o] = *reinterpret_cast<const uint2*>(ol);
        }
    }
}

}  // namespace dcta

using namespace dcta;

static int check_params(const char* who, const float* median, const float* b, int C, int H, int W, int z, int c, int d) {
    DCTA_REQUIRE(median && b, "%s: null statistics", who);
    DCTA_REQUIRE(C > 0 && H > 0 && W > 0 && z > 0 && z <= 256, "%s: z=%d outside 1..256", who, z);
    DCTA_REQUIRE(c > 0 && d > 0 && d <= 62 && c * d == z, "%s: needs a projection-free LFQ (c*d == z)", who);
    return DCTA_OK;
}

extern "C" int dcta_pack_codes_lfq(const float* tiles, const int32_t* order, const dcta_segment* segs,
                                   const int32_t* row_seg_start, int n_rows, int s, int th, int tw, int channels,
                                   int z, const float* median, const float* b, int H, int W, float eps, float lo,
                                   float hi, int c, int d, float scale, int32_t* tame_scratch, int64_t* codes,
                                   int64_t* positions, int64_t* channels_out, int64_t* image_ids,
                                   uint8_t* key_pad_mask, void* stream) {
    DCTA_REQUIRE(tiles && order && segs && row_seg_start && codes && positions && channels_out, "pack_codes_lfq: null pointer");
    DCTA_REQUIRE(n_rows >= 0 && s > 0 && th > 0 && tw > 0 && channels > 0 && th <= H && tw <= W,
                 "pack_codes_lfq: bad sizes (token grid must fit the PatchNorm tables)");
    int rc = check_params("pack_codes_lfq", median, b, channels, H, W, z, c, d);
    if (rc) return rc;
    if (n_rows == 0) return DCTA_OK;
    LfqNormParams q{median, b, channels, H, W, z, eps, lo, hi, c, d, scale};
    const bool aligned = ((reinterpret_cast<uintptr_t>(tiles) | reinterpret_cast<uintptr_t>(median) |
                           reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(positions)) & 15) == 0;
    if (z % 4 == 0 && d <= 32 && aligned && th * tw * channels < (1 << 22)) {
        if (tame_scratch)       // is every b a tame divisor?  (2.4 MB table, L2 resident: a few microseconds)
            launch_b_tame(b, (int64_t)channels * H * W * z, tame_scratch, as_stream(stream));
        pack_codes_vec_kernel<<<grid_for((int64_t)n_rows * s, 256), 256, 0, as_stream(stream)>>>(
            tiles, order, segs, row_seg_start, n_rows, s, tw, channels, th * tw * channels, 1.0f / (float)channels,
            1.0f / (float)tw, q, tame_scratch, codes, positions, channels_out, image_ids, key_pad_mask);
        return check_launch("pack_codes_lfq");
    }
    pack_codes_kernel<<<grid_for((int64_t)n_rows * s, 8), 256, 0, as_stream(stream)>>>(
        tiles, order, segs, row_seg_start, n_rows, s, tw, channels, th * tw * channels, q, codes, positions,
        channels_out, image_ids, key_pad_mask);
    return check_launch("pack_codes_lfq");
}

extern "C" int dcta_decode_codes_split(const int64_t* codes, const int32_t* slot_map, const int32_t* img_sel,
                                       int64_t n_img, int channels_n, int th, int tw, int p, int rows, int cols,
                                       int64_t ld, int out_h, int out_w, const float* median, const float* b, int H,
                                       int W, float eps, int c, int d, float scale, void* y_hi, void* y_lo,
                                       float* dc, void* stream) {
    DCTA_REQUIRE(codes && slot_map && y_hi && y_lo && dc, "decode_codes_split: null pointer");
    DCTA_REQUIRE(ld % 8 == 0 && ld >= cols && rows > 0 && cols > 0 && p > 0 && out_h > 0 && out_w > 0 && th <= H && tw <= W,
                 "decode_codes_split: bad sizes");
    int rc = check_params("decode_codes_split", median, b, channels_n, H, W, p * p, c, d);
    if (rc) return rc;
    if (n_img == 0) return DCTA_OK;
    LfqNormParams q{median, b, channels_n, H, W, p * p, eps, 0.f, 0.f, c, d, scale};
    DCTA_REQUIRE(rows % p == 0, "decode_codes_split: plane rows must be a multiple of the patch size");
    const int64_t n_rows_total = n_img * channels_n * (rows / p);
    DCTA_REQUIRE(n_rows_total < (1ll << 31) && ld < (1ll << 30), "decode_codes_split: too many plane rows for one launch");
    decode_codes_split_kernel<<<(unsigned)n_rows_total, 128, 0, as_stream(stream)>>>(
        codes, slot_map, img_sel, th, tw, p, rows, cols, (int)ld, q, (__half*)y_hi, (__half*)y_lo, dc,
        1.0f / sqrtf((float)out_h * (float)out_w), 16.0f);
    return check_launch("decode_codes_split");
}

extern "C" int dcta_pack_codes_grid(const int32_t* code_grid, const int32_t* order, const dcta_segment* segs,
                                    const int32_t* row_seg_start, int n_rows, int s, int th, int tw, int channels,
                                    const float* median, const float* b, int H, int W, float eps, float lo, float hi,
                                    int c, int d, int64_t* pad_scratch, int64_t* codes, int64_t* positions,
                                    int64_t* channels_out, int64_t* image_ids, uint8_t* key_pad_mask, void* stream) {
    DCTA_REQUIRE(code_grid && order && segs && row_seg_start && pad_scratch && codes && positions && channels_out,
                 "pack_codes_grid: null pointer");
    DCTA_REQUIRE(n_rows >= 0 && s > 0 && th > 0 && tw > 0 && channels > 0 && th * tw * channels < (1 << 22),
                 "pack_codes_grid: bad sizes");
    int rc = check_params("pack_codes_grid", median, b, channels, H, W, c * d, c, d);
    if (rc) return rc;
    if (n_rows == 0) return DCTA_OK;
    LfqNormParams q{median, b, channels, H, W, c * d, eps, lo, hi, c, d, 1.0f};
    pad_codes_kernel<<<1, 32, 0, as_stream(stream)>>>(q, pad_scratch);
    pack_codes_grid_kernel<<<grid_for((int64_t)n_rows * s, 256), 256, 0, as_stream(stream)>>>(
        code_grid, order, segs, row_seg_start, n_rows, s, tw, channels, th * tw * channels, 1.0f / (float)channels,
        1.0f / (float)tw, c, 1.0f / (float)c, pad_scratch, codes, positions, channels_out, image_ids, key_pad_mask);
    return check_launch("pack_codes_grid");
}
