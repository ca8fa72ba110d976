// Token importance scores, per-image descending sort, and gather + pad-and-stack packing
// (reference: feature_extraction_dct_autoencoder.py:403-452 and 455-605, util.py:149-164).
// All three are HBM-bound streaming kernels: one warp per token, 128-bit accesses.
#include "common.cuh"

namespace dcta {

struct Importances {
    float v[8];
};

// ------------------------------------------------------------------------------ scores
// FE:409-416.  fp32 arithmetic in the reference's order with NO fma contraction, so that the
// scores (and therefore the selection order) are bit-identical for identical coefficients:
//   mags = amax|tile| * w ;  dist = float(-(th+tw)) / imp[c] ;  score = mags + dist
template <int kVec>
__global__ void __launch_bounds__(256) tile_scores_kernel(const float* __restrict__ tiles,
                                                          float* __restrict__ scores,
                                                          int64_t n_tokens_total, int th, int tw,
                                                          int channels, int z, float mag_weight,
                                                          Importances imp) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int per_img = th * tw * channels;
    for (int64_t tok = warp0; tok < n_tokens_total; tok += n_warps) {
        const float* src = tiles + tok * z;
        float m = 0.0f;
        if (kVec == 4) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            for (int i = lane; i < z / 4; i += 32) {
                float4 v = ld_stream(s4 + i);
                m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            }
        } else {
            for (int i = lane; i < z; i += 32) m = fmaxf(m, fabsf(__ldg(src + i)));
        }
        m = warp_max(m);
        if (lane == 0) {
            const int t = (int)(tok % per_img);
            const int c = t % channels;
            const int tile = t / channels;
            const int h = tile / tw, w = tile - h * tw;
            const float mags = __fmul_rn(m, mag_weight);
            const float dist = __fdiv_rn((float)(-(h + w)), imp.v[c]);
            scores[tok] = __fadd_rn(mags, dist);
        }
    }
}

// ------------------------------------------------------------------------------ sort
// Per-image bitonic sort in shared memory on 64-bit keys  (~orderable(score)) << 32 | index :
// ascending key order == descending score, ties by ascending flat index.  NaN scores sort first
// (torch.sort treats NaN as the largest value).
__device__ __forceinline__ uint32_t orderable(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// kFromMax: `scores` holds amax|tile| per token and the score of FE:409-416 is formed here (same
// arithmetic as tile_scores_kernel); scores_out [nullable] receives it.
struct ScoreParams {
    int tw, channels;
    float mag_weight;
    Importances imp;
    float* scores_out;
};

template <bool kFromMax>
__global__ void __launch_bounds__(1024) sort_tokens_kernel(const float* __restrict__ scores,
                                                           int32_t* __restrict__ order, int n_tok,
                                                           int n_pad, ScoreParams sp) {
    extern __shared__ unsigned long long keys[];
    const int64_t img = blockIdx.x;
    const float* s = scores + img * n_tok;
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        unsigned long long k = ~0ull;
        if (i < n_tok) {
            float sc = s[i];
            if (kFromMax) {
                const int c = i % sp.channels;
                const int tile = i / sp.channels;
                const int h = tile / sp.tw, w = tile - h * sp.tw;
                const float mags = __fmul_rn(sc, sp.mag_weight);
                const float dist = __fdiv_rn((float)(-(h + w)), sp.imp.v[c]);
                sc = __fadd_rn(mags, dist);
                if (sp.scores_out) sp.scores_out[img * n_tok + i] = sc;
            }
            k = ((unsigned long long)(~orderable(sc)) << 32) | (unsigned)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (n_pad >> 1); t += blockDim.x) {
                const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
                const int hi = lo | stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a > b) == up) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n_tok; i += blockDim.x)
        order[img * n_tok + i] = (int32_t)(keys[i] & 0xffffffffu);
}

// Faster variant for 128 <= n_pad <= 4096: four consecutive keys per thread.  Compare-exchange distances 1 and 2
// stay in registers, distances 4..64 are partner lanes of the same warp (shuffles, no barrier), only distances
// >= 128 go through (double-buffered) shared memory: 15 block barriers for 4096 keys instead of 78.
__device__ __forceinline__ void cmp_swap(unsigned long long& a, unsigned long long& b, bool up) {
    if ((a > b) == up) { const unsigned long long t = a; a = b; b = t; }
}

template <bool kFromMax>
__global__ void __launch_bounds__(1024) sort_tokens_fast_kernel(const float* __restrict__ scores,
                                                                int32_t* __restrict__ order, int n_tok, int n_pad,
                                                                ScoreParams sp) {
    extern __shared__ unsigned long long keys[];          // [2][n_pad]
    const int64_t img = blockIdx.x;
    const float* s = scores + img * n_tok;
    const int t = threadIdx.x, lane = t & 31;
    unsigned long long k[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = 4 * t + e;
        k[e] = ~0ull;
        if (i < n_tok) {
            float sc = s[i];
            if (kFromMax) {
                const int c = i % sp.channels;
                const int tile = i / sp.channels;
                const int h = tile / sp.tw, w = tile - h * sp.tw;
                const float mags = __fmul_rn(sc, sp.mag_weight);
                const float dist = __fdiv_rn((float)(-(h + w)), sp.imp.v[c]);
                sc = __fadd_rn(mags, dist);
                if (sp.scores_out) sp.scores_out[img * n_tok + i] = sc;
            }
            k[e] = ((unsigned long long)(~orderable(sc)) << 32) | (unsigned)i;
        }
    }
    int buf = 0;
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 128) {
                unsigned long long* kb = keys + buf * n_pad;
                buf ^= 1;
                reinterpret_cast<ulonglong2*>(kb)[2 * t] = make_ulonglong2(k[0], k[1]);
                reinterpret_cast<ulonglong2*>(kb)[2 * t + 1] = make_ulonglong2(k[2], k[3]);
                __syncthreads();
                const int pt = t ^ (stride >> 2);                     // the partner thread holds the 4 partner keys
                const ulonglong2 o01 = reinterpret_cast<const ulonglong2*>(kb)[2 * pt];
                const ulonglong2 o23 = reinterpret_cast<const ulonglong2*>(kb)[2 * pt + 1];
                const unsigned long long o[4] = {o01.x, o01.y, o23.x, o23.y};
                const bool take_min = (((4 * t) & stride) == 0) == (((4 * t) & size) == 0);
#pragma unroll
                for (int e = 0; e < 4; ++e) k[e] = take_min ? (k[e] < o[e] ? k[e] : o[e]) : (k[e] > o[e] ? k[e] : o[e]);
            } else if (stride >= 4) {
                const int d = stride >> 2;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = 4 * t + e;
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, k[e], d);
                    const bool up = (i & size) == 0, low = (lane & d) == 0;
                    const bool take_min = (low == up);
                    k[e] = take_min ? (k[e] < o ? k[e] : o) : (k[e] > o ? k[e] : o);
                }
            } else {
                const bool up = ((4 * t) & size) == 0;       // size >= 4: the same direction for the 4 keys of a thread
                if (stride == 2) {
                    cmp_swap(k[0], k[2], up);
                    cmp_swap(k[1], k[3], up);
                } else if (size == 2) {                       // directions alternate between the two pairs
                    cmp_swap(k[0], k[1], true);
                    cmp_swap(k[2], k[3], false);
                } else {
                    cmp_swap(k[0], k[1], up);
                    cmp_swap(k[2], k[3], up);
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = 4 * t + e;
        if (i < n_tok) order[img * n_tok + i] = (int32_t)(k[e] & 0xffffffffu);
    }
}

// Block radix sort for 1024 < n_tok <= 4096 (the 3072 tokens of a 512^2 image): the order is decided by the 32-bit
// key ~orderable(score) alone, ties by ascending index, i.e. a STABLE sort of the keys in index order -- which is
// what least-significant-digit radix sort delivers.  Four passes of 8 bits; 16, 24 or 32 warps hold 128 elements
// each (768 threads for 3072 tokens: no padding; warp w owns elements 128w .. 128w+127, slot e of lane l is element
// 128w + 32e + l, so rank order inside a warp is index order).  Per pass: every warp ranks its elements per digit with ballots (no atomics), an exclusive
// scan over (digit, warp) turns the counts into offsets, elements move to their place in shared memory.
// ~1/4 of the instructions of the 4096-slot bitonic network on 64-bit keys.
constexpr int kRadixThreads = 1024, kRadixSlots = 4;

template <bool kFromMax>      // blockDim.x = 512, 768 or 1024
__global__ void __launch_bounds__(kRadixThreads, 2) sort_tokens_radix_kernel(const float* __restrict__ scores,
                                                                             int32_t* __restrict__ order, int n_tok,
                                                                             ScoreParams sp) {
    extern __shared__ __align__(16) uint32_t radix_smem[];
    uint32_t* keys0 = radix_smem;                                     // [4096]
    uint32_t* keys1 = keys0 + 4096;
    uint16_t* idx0 = reinterpret_cast<uint16_t*>(keys1 + 4096);       // [4096]
    uint16_t* idx1 = idx0 + 4096;
    uint32_t* hist = reinterpret_cast<uint32_t*>(idx1 + 4096);        // [n_warps][256 digits]
    uint32_t* warp_tot = hist + 32 * 256;                             // [32]
    const int n_thr = blockDim.x, n_warps = n_thr >> 5;               // n_warps is a multiple of 8
    const int64_t img = blockIdx.x;
    const float* s = scores + img * n_tok;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint32_t key[kRadixSlots];
    uint32_t idx[kRadixSlots];
#pragma unroll
    for (int e = 0; e < kRadixSlots; ++e) {
        const int i = warp * 128 + e * 32 + lane;
        key[e] = 0xffffffffu;                                         // padding sorts last (stable: after every token)
        idx[e] = (uint32_t)i;
        if (i < n_tok) {
            float sc = s[i];
            if (kFromMax) {
                const int c = i % sp.channels;
                const int tile = i / sp.channels;
                const int h = tile / sp.tw, w = tile - h * sp.tw;
                const float mags = __fmul_rn(sc, sp.mag_weight);
                const float dist = __fdiv_rn((float)(-(h + w)), sp.imp.v[c]);
                sc = __fadd_rn(mags, dist);
                if (sp.scores_out) sp.scores_out[img * n_tok + i] = sc;
            }
            key[e] = ~orderable(sc);
        }
    }
    uint32_t* kin = keys0;
    uint32_t* kout = keys1;
    uint16_t* iin = idx0;
    uint16_t* iout = idx1;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        for (int i = t; i < n_warps * 256; i += n_thr) hist[i] = 0u;
        __syncthreads();
        // rank of every element among the elements of its warp with the same digit, in index order
        uint32_t rank[kRadixSlots], dig[kRadixSlots];
        uint32_t* my_hist = hist + warp * 256;
#pragma unroll
        for (int e = 0; e < kRadixSlots; ++e) {
            dig[e] = (key[e] >> shift) & 0xffu;
            // lanes holding the same digit: eight ballots (match.any costs one round per distinct value, ~30 here)
            unsigned peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const unsigned bal = __ballot_sync(0xffffffffu, (dig[e] >> b) & 1u);
                peers &= ((dig[e] >> b) & 1u) ? bal : ~bal;
            }
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (lane == leader) {
                base = my_hist[dig[e]];
                my_hist[dig[e]] = base + __popc(peers);
            }
            base = __shfl_sync(0xffffffffu, base, leader);
            rank[e] = base + __popc(peers & lt_mask);
            __syncwarp();
        }
        __syncthreads();
        // exclusive scan of the counts in (digit, warp) order: thread t owns the 8 consecutive entries 8t .. 8t+7 of
        // that order (256 * n_warps = 8 * n_thr of them), which share their digit because 8 divides n_warps
        {
            const int d = (8 * t) / n_warps, w0 = 8 * t - d * n_warps;
            uint32_t sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += hist[(w0 + j) * 256 + d];
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                uint32_t v = lane < n_warps ? warp_tot[lane] : 0u, inc2 = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, inc2, o);
                    if (lane >= o) inc2 += u;
                }
                warp_tot[lane] = inc2 - v;                            // exclusive
            }
            __syncthreads();
            uint32_t run = warp_tot[warp] + incl - sum;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t cnt = hist[(w0 + j) * 256 + d];
                hist[(w0 + j) * 256 + d] = run;
                run += cnt;
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < kRadixSlots; ++e) {
            const uint32_t pos = my_hist[dig[e]] + rank[e];
            kout[pos] = key[e];
            iout[pos] = (uint16_t)idx[e];
        }
        __syncthreads();
        if (pass < 3) {
#pragma unroll
            for (int e = 0; e < kRadixSlots; ++e) {
                const int i = warp * 128 + e * 32 + lane;
                key[e] = kout[i];
                idx[e] = iout[i];
            }
        }
        uint32_t* tk = kin; kin = kout; kout = tk;
        uint16_t* ti = iin; iin = iout; iout = ti;
    }
    for (int i = t; i < n_tok; i += n_thr) order[img * n_tok + i] = (int32_t)iin[i];
}

// ------------------------------------------------------------------------------ pack
// One warp per output slot (row, s).  kTiles: source is the token grid + sort order;
// otherwise per-image token lists through pointer tables.
template <bool kTiles, int kVec>
__global__ void __launch_bounds__(256) pack_kernel(
    const float* __restrict__ tiles, const int32_t* __restrict__ order,
    const float* const* __restrict__ src_patches, const int64_t* const* __restrict__ src_positions,
    const int64_t* const* __restrict__ src_channels, const dcta_segment* __restrict__ segs,
    const int32_t* __restrict__ row_seg_start, int n_rows, int s, int tw, int channels,
    int n_tok_img, int z, float* __restrict__ patches, int64_t* __restrict__ positions,
    int64_t* __restrict__ channels_out, int64_t* __restrict__ image_ids,
    uint8_t* __restrict__ key_pad_mask) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = (int64_t)n_rows * s;
    for (int64_t slot = warp0; slot < total; slot += n_warps) {
        const int row = (int)(slot / s);
        const int off = (int)(slot - (int64_t)row * s);
        // locate the segment of this row that covers `off` (rows hold few segments)
        int lo = row_seg_start[row], hi = row_seg_start[row + 1];
        int seg = -1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const int so = segs[mid].offset;
            if (off < so) hi = mid;
            else if (off >= so + segs[mid].k) lo = mid + 1;
            else { seg = mid; break; }
        }
        float* dst = patches + slot * z;
        if (seg < 0) {  // padding slot (UT:155-157 zeros; FE:574-576 mask = True)
            if (kVec == 4) {
                float4* d4 = reinterpret_cast<float4*>(dst);
                for (int i = lane; i < z / 4; i += 32) st_stream(d4 + i, make_float4(0.f, 0.f, 0.f, 0.f));
            } else {
                for (int i = lane; i < z; i += 32) dst[i] = 0.0f;
            }
            if (lane == 0) {
                positions[slot * 2] = 0;
                positions[slot * 2 + 1] = 0;
                channels_out[slot] = 0;
                if (image_ids) image_ids[slot] = 0;
                if (key_pad_mask) key_pad_mask[slot] = 1;
            }
            continue;
        }
        const dcta_segment sg = segs[seg];
        const int j = off - sg.offset;
        const float* src;
        int64_t ph, pw, pc;
        if (kTiles) {
            const int tok = order[sg.img * n_tok_img + j];
            src = tiles + (sg.img * n_tok_img + tok) * z;
            pc = tok % channels;
            const int tile = tok / channels;
            ph = tile / tw;
            pw = tile - (tile / tw) * tw;
        } else {
            src = src_patches[sg.img] + (int64_t)j * z;
            ph = pw = pc = 0;
            if (lane == 0) {
                ph = src_positions[sg.img][2 * j];
                pw = src_positions[sg.img][2 * j + 1];
                pc = src_channels[sg.img][j];
            }
        }
        if (kVec == 4) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4* d4 = reinterpret_cast<float4*>(dst);
            for (int i = lane; i < z / 4; i += 32) st_stream(d4 + i, ld_stream(s4 + i));
        } else {
            for (int i = lane; i < z; i += 32) dst[i] = __ldg(src + i);
        }
        if (lane == 0) {
            positions[slot * 2] = ph;
            positions[slot * 2 + 1] = pw;
            channels_out[slot] = pc;
            if (image_ids) image_ids[slot] = sg.image_id;
            if (key_pad_mask) key_pad_mask[slot] = 0;
        }
    }
}

// 128-bit variant (z % 4 == 0, z <= 1024, aligned bases): a warp takes 32 consecutive output slots.  The
// metadata chain (segment search -> sort order -> source address) runs once, lane-parallel; the 32 * z/4
// output float4s, which are contiguous, are then copied with the lanes on consecutive OUTPUT words: every
// store instruction writes 512 contiguous bytes, every load reads at most two tokens' contiguous bytes, and
// kUnroll independent loads are in flight per lane.
template <bool kTiles>
__global__ void __launch_bounds__(256) pack_rows_kernel(
    const float* __restrict__ tiles, const int32_t* __restrict__ order,
    const float* const* __restrict__ src_patches, const int64_t* const* __restrict__ src_positions,
    const int64_t* const* __restrict__ src_channels, const dcta_segment* __restrict__ segs,
    const int32_t* __restrict__ row_seg_start, int n_rows, int s, int tw, int channels,
    int n_tok_img, int z, float* __restrict__ patches, int64_t* __restrict__ positions,
    int64_t* __restrict__ channels_out, int64_t* __restrict__ image_ids,
    uint8_t* __restrict__ key_pad_mask, int32_t* __restrict__ src_index = nullptr) {
    // src_index (kTiles, nullable): the token-grid row each slot reads (-1: padding); patches may then be NULL --
    // the consumer gathers the rows itself (dcta_split_rows_patchnorm) and the packed patches are never written
    constexpr int kUnroll = 7;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = (int64_t)n_rows * s;
    const int z4 = z >> 2;
    const float inv_z4 = 1.0f / (float)z4;
    for (int64_t slot0 = warp0 * 32; slot0 < total; slot0 += n_warps * 32) {
        const int64_t slot = slot0 + lane;
        const float4* src = nullptr;             // nullptr = padding slot (UT:155-157 zeros; FE:574-576 mask = True)
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
            int64_t ph = 0, pw = 0, pc = 0, image_id = 0;
            int32_t src_row = -1;
            if (seg >= 0) {
                const dcta_segment sg = segs[seg];
                const int j = off - sg.offset;
                image_id = sg.image_id;
                if (kTiles) {
                    const int tok = order[sg.img * n_tok_img + j];
                    src_row = (int32_t)(sg.img * n_tok_img + tok);
                    src = reinterpret_cast<const float4*>(tiles + (sg.img * n_tok_img + tok) * z);
                    const int tile = tok / channels;
                    pc = tok - tile * channels;
                    ph = tile / tw;
                    pw = tile - (int)ph * tw;
                } else {
                    src = reinterpret_cast<const float4*>(src_patches[sg.img] + (int64_t)j * z);
                    const longlong2 hw = *reinterpret_cast<const longlong2*>(src_positions[sg.img] + 2 * j);
                    ph = hw.x;
                    pw = hw.y;
                    pc = src_channels[sg.img][j];
                }
            }
            reinterpret_cast<longlong2*>(positions)[slot] = make_longlong2(ph, pw);
            channels_out[slot] = pc;
            if (image_ids) image_ids[slot] = image_id;
            if (key_pad_mask) key_pad_mask[slot] = seg < 0;
            if (src_index) src_index[slot] = src_row;
        }
        if (patches == nullptr) continue;
        const int n4 = (int)min((int64_t)32, total - slot0) * z4;
        float4* dst = reinterpret_cast<float4*>(patches + slot0 * z);
        const unsigned long long src_bits = (unsigned long long)src;
        for (int it = 0; it < z4; it += kUnroll) {
            float4 v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int idx = (it + u) * 32 + lane;
                const int sl = min((int)(((float)idx + 0.5f) * inv_z4), 31), wd = idx - sl * z4;   // exact: idx < 2^13
                const float4* sp = (const float4*)__shfl_sync(0xffffffffu, src_bits, sl);
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (it + u < z4 && idx < n4 && sp != nullptr) v[u] = ld_stream(sp + wd);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int idx = (it + u) * 32 + lane;
                if (it + u < z4 && idx < n4) st_stream(dst + idx, v[u]);
            }
        }
    }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace dcta

using namespace dcta;

extern "C" int dcta_tile_scores(const float* tiles, float* scores, int64_t n_img, int th, int tw,
                                int channels, int z, float mag_weight,
                                const float* channel_importances_host, void* stream) {
    DCTA_REQUIRE(tiles && scores && channel_importances_host, "tile_scores: null pointer");
    DCTA_REQUIRE(channels >= 1 && channels <= 8, "tile_scores: 1..8 channels supported, got %d", channels);
    DCTA_REQUIRE(th > 0 && tw > 0 && z > 0 && n_img >= 0, "tile_scores: bad sizes");
    if (n_img == 0) return DCTA_OK;
    Importances imp;
    for (int i = 0; i < 8; ++i) imp.v[i] = i < channels ? channel_importances_host[i] : 1.0f;
    const int64_t total = n_img * th * tw * channels;
    const int grid = grid_for(total, 8);
    if (z % 4 == 0 && al16(tiles))
        tile_scores_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(tiles, scores, total, th, tw, channels, z, mag_weight, imp);
    else
        tile_scores_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(tiles, scores, total, th, tw, channels, z, mag_weight, imp);
    return check_launch("tile_scores");
}

static int launch_sort(const float* scores, int32_t* order, int64_t n_img, int n_tok, bool from_max, const ScoreParams& sp,
                       void* stream) {
    DCTA_REQUIRE(scores && order, "sort_tokens: null pointer");
    DCTA_REQUIRE(n_tok > 0 && n_tok <= 16384, "sort_tokens: n_tok=%d outside 1..16384", n_tok);
    if (n_img == 0) return DCTA_OK;
    int n_pad = 2;
    while (n_pad < n_tok) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * sizeof(unsigned long long);
    if (smem > 48 * 1024) {  // per-device attribute: set on every call that needs it (cheap)
        cudaFuncSetAttribute(sort_tokens_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
        cudaFuncSetAttribute(sort_tokens_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
    }
    if (n_tok > 1024 && n_tok <= 4096) {
        const size_t smem_radix = 2 * 4096 * 4 + 2 * 4096 * 2 + 32 * 256 * 4 + 32 * 4;
        const int threads = (int)ceil_div(n_tok, 1024) * 256;          // 128 elements per warp, a multiple of 8 warps
        if (from_max) {
            cudaFuncSetAttribute(sort_tokens_radix_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_radix);
            sort_tokens_radix_kernel<true><<<(unsigned)n_img, threads, smem_radix, as_stream(stream)>>>(scores, order, n_tok, sp);
        } else {
            cudaFuncSetAttribute(sort_tokens_radix_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_radix);
            sort_tokens_radix_kernel<false><<<(unsigned)n_img, threads, smem_radix, as_stream(stream)>>>(scores, order, n_tok, sp);
        }
        return check_launch("sort_tokens");
    }
    if (n_pad >= 128 && n_pad <= 4096) {
        const size_t smem_fast = 2 * smem;
        if (smem_fast > 48 * 1024) {
            cudaFuncSetAttribute(sort_tokens_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * 8);
            cudaFuncSetAttribute(sort_tokens_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * 8);
        }
        if (from_max)
            sort_tokens_fast_kernel<true><<<(unsigned)n_img, n_pad / 4, smem_fast, as_stream(stream)>>>(scores, order, n_tok, n_pad, sp);
        else
            sort_tokens_fast_kernel<false><<<(unsigned)n_img, n_pad / 4, smem_fast, as_stream(stream)>>>(scores, order, n_tok, n_pad, sp);
        return check_launch("sort_tokens");
    }
    int threads = n_pad / 2;
    if (threads > 1024) threads = 1024;
    if (threads < 32) threads = 32;
    if (from_max)
        sort_tokens_kernel<true><<<(unsigned)n_img, threads, smem, as_stream(stream)>>>(scores, order, n_tok, n_pad, sp);
    else
        sort_tokens_kernel<false><<<(unsigned)n_img, threads, smem, as_stream(stream)>>>(scores, order, n_tok, n_pad, sp);
    return check_launch("sort_tokens");
}

extern "C" int dcta_sort_tokens(const float* scores, int32_t* order, int64_t n_img, int n_tok,
                                void* stream) {
    ScoreParams sp{};
    return launch_sort(scores, order, n_img, n_tok, false, sp, stream);
}

extern "C" int dcta_sort_tokens_maxabs(const float* maxabs, float* scores, int32_t* order, int64_t n_img, int th, int tw,
                                       int channels, float mag_weight, const float* channel_importances_host,
                                       void* stream) {
    DCTA_REQUIRE(channel_importances_host, "sort_tokens_maxabs: null pointer");
    DCTA_REQUIRE(channels >= 1 && channels <= 8 && th > 0 && tw > 0, "sort_tokens_maxabs: bad sizes");
    ScoreParams sp{};
    sp.tw = tw; sp.channels = channels; sp.mag_weight = mag_weight; sp.scores_out = scores;
    for (int i = 0; i < 8; ++i) sp.imp.v[i] = i < channels ? channel_importances_host[i] : 1.0f;
    return launch_sort(maxabs, order, n_img, th * tw * channels, true, sp, stream);
}

extern "C" int dcta_pack_tiles(const float* tiles, const int32_t* order, const dcta_segment* segs,
                               const int32_t* row_seg_start, int n_rows, int s, int th, int tw,
                               int channels, int z, float* patches, int64_t* positions,
                               int64_t* channels_out, int64_t* image_ids, uint8_t* key_pad_mask,
                               void* stream) {
    DCTA_REQUIRE(tiles && order && segs && row_seg_start && patches && positions && channels_out,
                 "pack_tiles: null pointer");
    DCTA_REQUIRE(n_rows >= 0 && s > 0 && th > 0 && tw > 0 && channels > 0 && z > 0, "pack_tiles: bad sizes");
    if (n_rows == 0) return DCTA_OK;
    const int grid = grid_for((int64_t)n_rows * s, 8);
    const int n_tok_img = th * tw * channels;
    if (z % 4 == 0 && z <= 1024 && al16(tiles) && al16(patches))
        pack_rows_kernel<true><<<grid_for((int64_t)n_rows * s, 256), 256, 0, as_stream(stream)>>>(tiles, order, nullptr, nullptr, nullptr, segs, row_seg_start, n_rows, s, tw, channels, n_tok_img, z, patches, positions, channels_out, image_ids, key_pad_mask);
    else if (z % 4 == 0 && al16(tiles) && al16(patches))
        pack_kernel<true, 4><<<grid, 256, 0, as_stream(stream)>>>(tiles, order, nullptr, nullptr, nullptr, segs, row_seg_start, n_rows, s, tw, channels, n_tok_img, z, patches, positions, channels_out, image_ids, key_pad_mask);
    else
        pack_kernel<true, 1><<<grid, 256, 0, as_stream(stream)>>>(tiles, order, nullptr, nullptr, nullptr, segs, row_seg_start, n_rows, s, tw, channels, n_tok_img, z, patches, positions, channels_out, image_ids, key_pad_mask);
    return check_launch("pack_tiles");
}

extern "C" int dcta_pack_tiles_index(const int32_t* order, const dcta_segment* segs, const int32_t* row_seg_start, int n_rows,
                                     int s, int th, int tw, int channels, int64_t n_img, int64_t* positions,
                                     int64_t* channels_out, int64_t* image_ids, uint8_t* key_pad_mask, int32_t* src_index,
                                     void* stream) {
    DCTA_REQUIRE(order && segs && row_seg_start && positions && channels_out && src_index, "pack_tiles_index: null pointer");
    DCTA_REQUIRE(n_rows >= 0 && s > 0 && th > 0 && tw > 0 && channels > 0, "pack_tiles_index: bad sizes");
    DCTA_REQUIRE(n_img * th * tw * channels < (1ll << 31), "pack_tiles_index: more than 2^31 tokens in the grid");
    if (n_rows == 0) return DCTA_OK;
    const int n_tok_img = th * tw * channels;
    pack_rows_kernel<true><<<grid_for((int64_t)n_rows * s, 256), 256, 0, as_stream(stream)>>>(
        nullptr, order, nullptr, nullptr, nullptr, segs, row_seg_start, n_rows, s, tw, channels, n_tok_img, 4, nullptr, positions,
        channels_out, image_ids, key_pad_mask, src_index);
    return check_launch("pack_tiles_index");
}

extern "C" int dcta_pack_lists(const float* const* src_patches, const int64_t* const* src_positions,
                               const int64_t* const* src_channels, const dcta_segment* segs,
                               const int32_t* row_seg_start, int n_rows, int s, int z,
                               float* patches, int64_t* positions, int64_t* channels_out,
                               int64_t* image_ids, uint8_t* key_pad_mask, void* stream) {
    DCTA_REQUIRE(src_patches && src_positions && src_channels && segs && row_seg_start && patches &&
                     positions && channels_out, "pack_lists: null pointer");
    DCTA_REQUIRE(n_rows >= 0 && s > 0 && z > 0, "pack_lists: bad sizes");
    if (n_rows == 0) return DCTA_OK;
    const int grid = grid_for((int64_t)n_rows * s, 8);
    // per-image sources are torch allocations (>= 16-byte aligned at offset 0) but may be views:
    // the 128-bit path needs z*4 % 16 == 0 AND every base aligned; the host shim guarantees the
    // latter by passing contiguous tensors, else it asks for the scalar path via z's alignment.
    if (z % 4 == 0 && z <= 1024 && al16(patches))
        pack_rows_kernel<false><<<grid_for((int64_t)n_rows * s, 256), 256, 0, as_stream(stream)>>>(nullptr, nullptr, src_patches, src_positions, src_channels, segs, row_seg_start, n_rows, s, 0, 0, 0, z, patches, positions, channels_out, image_ids, key_pad_mask);
    else if (z % 4 == 0 && al16(patches))
        pack_kernel<false, 4><<<grid, 256, 0, as_stream(stream)>>>(nullptr, nullptr, src_patches, src_positions, src_channels, segs, row_seg_start, n_rows, s, 0, 0, 0, z, patches, positions, channels_out, image_ids, key_pad_mask);
    else
        pack_kernel<false, 1><<<grid, 256, 0, as_stream(stream)>>>(nullptr, nullptr, src_patches, src_positions, src_channels, segs, row_seg_start, n_rows, s, 0, 0, 0, z, patches, positions, channels_out, image_ids, key_pad_mask);
    return check_launch("pack_lists");
}
