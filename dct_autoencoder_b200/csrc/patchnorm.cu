// PatchNorm: per-(channel, tile-row, tile-col, coefficient) median / mean-abs-deviation
// normalisation, its inverse, and statistic fitting (reference: patchnorm.py:9-177).
//
// normalise / denormalise are HBM-bound streams (x read once, out written once; the two stat
// tables are 2 x C*H*W*z*4 B = 4.8 MB at patch 14 / 32x32x3 and stay L2-resident).
// All fp32 arithmetic uses explicit round-to-nearest intrinsics in the reference's operation
// order (no fma contraction) so results are bit-identical to the reference's CPU fp32 path.
#include "common.cuh"
#include "lfq_norm.cuh"

namespace dcta {

// ------------------------------------------------------------------------------ apply
// (clamped_position and patchnorm_value live in lfq_norm.cuh: the fused consumers use the same arithmetic)

// kVec == 4 (z % 4 == 0, z <= 256): a warp takes 32 consecutive tokens, the position lookups (three dependent
// 8-byte loads per token) run once, lane-parallel, and the tokens' words are then streamed with 128-bit
// accesses, 9 independent loads in flight per lane at 32 resident warps per SM.  kVec == 1: one warp per token, scalar.
template <bool kInverse, int kVec>
__global__ void __launch_bounds__(256, 4) patchnorm_apply_kernel(
    const float* __restrict__ x, const int64_t* __restrict__ channels,
    const int64_t* __restrict__ positions, const float* __restrict__ median,
    const float* __restrict__ b, float* __restrict__ out, int64_t n_tok, int z, int C, int H, int W,
    float eps, float lo, float hi) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    if (kVec == 4 && z <= 256) {
        // the 32 * z/4 float4s of the warp's tokens are contiguous in x and out: lanes run over consecutive
        // words (all 32 lanes busy whatever z is), the statistics row of a word's token comes from a shuffle
        constexpr int kUnroll = 3;
        const int z4 = z >> 2;
        const float inv_z4 = 1.0f / (float)z4;
        const float4* m4 = reinterpret_cast<const float4*>(median);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        for (int64_t tok0 = warp0 * 32; tok0 < n_tok; tok0 += n_warps * 32) {
            const int row_mine = (tok0 + lane < n_tok ? clamped_position(channels, positions, tok0 + lane, C, H, W) : 0) * z4;
            const int n4 = (int)min((int64_t)32, n_tok - tok0) * z4;
            const float4* xs = reinterpret_cast<const float4*>(x + tok0 * z);
            float4* os = reinterpret_cast<float4*>(out + tok0 * z);
            for (int it = 0; it < z4; it += kUnroll) {
                float4 xv[kUnroll], mv[kUnroll], bv[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int idx = (it + u) * 32 + lane;
                    const int sl = min((int)(((float)idx + 0.5f) * inv_z4), 31), wd = idx - sl * z4;   // exact: idx < 2^11
                    const int row = __shfl_sync(0xffffffffu, row_mine, sl);
                    if (it + u < z4 && idx < n4) {
                        xv[u] = ld_stream(xs + idx);
                        mv[u] = __ldg(m4 + row + wd);
                        bv[u] = __ldg(b4 + row + wd);
                    }
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int idx = (it + u) * 32 + lane;
                    if (it + u < z4 && idx < n4) {
                        float4 o;
                        o.x = patchnorm_value<kInverse>(xv[u].x, mv[u].x, bv[u].x, eps, lo, hi);
                        o.y = patchnorm_value<kInverse>(xv[u].y, mv[u].y, bv[u].y, eps, lo, hi);
                        o.z = patchnorm_value<kInverse>(xv[u].z, mv[u].z, bv[u].z, eps, lo, hi);
                        o.w = patchnorm_value<kInverse>(xv[u].w, mv[u].w, bv[u].w, eps, lo, hi);
                        st_stream(os + idx, o);
                    }
                }
            }
        }
        return;
    }
    for (int64_t tok = warp0; tok < n_tok; tok += n_warps) {
        const int64_t pid = clamped_position(channels, positions, tok, C, H, W);
        const float* xs = x + tok * z;
        const float* ms = median + pid * z;
        const float* bs = b + pid * z;
        float* os = out + tok * z;
        if (kVec == 4) {
            for (int i = lane; i < z / 4; i += 32) {
                const float4 xv = ld_stream(reinterpret_cast<const float4*>(xs) + i);
                const float4 mv = __ldg(reinterpret_cast<const float4*>(ms) + i);
                const float4 bv = __ldg(reinterpret_cast<const float4*>(bs) + i);
                float4 o;
                o.x = patchnorm_value<kInverse>(xv.x, mv.x, bv.x, eps, lo, hi);
                o.y = patchnorm_value<kInverse>(xv.y, mv.y, bv.y, eps, lo, hi);
                o.z = patchnorm_value<kInverse>(xv.z, mv.z, bv.z, eps, lo, hi);
                o.w = patchnorm_value<kInverse>(xv.w, mv.w, bv.w, eps, lo, hi);
                st_stream(reinterpret_cast<float4*>(os) + i, o);
            }
        } else {
            for (int i = lane; i < z; i += 32)
                os[i] = patchnorm_value<kInverse>(__ldg(xs + i), __ldg(ms + i), __ldg(bs + i), eps, lo, hi);
        }
    }
}

// ------------------------------------------------------------------------------ token lists
__device__ __forceinline__ int position_id(const int64_t* channels, const int64_t* positions,
                                           int64_t tok, int C, int H, int W) {
    const int64_t c = channels[tok], h = positions[2 * tok], w = positions[2 * tok + 1];
    if (c < 0 || c >= C || h < 0 || h >= H || w < 0 || w >= W) return -1;
    return (int)((c * H + h) * W + w);
}

__global__ void count_positions_kernel(const int64_t* __restrict__ channels,
                                       const int64_t* __restrict__ positions,
                                       const uint8_t* __restrict__ pad, int64_t n_tok, int C, int H,
                                       int W, int32_t* __restrict__ counts) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_tok;
         t += (int64_t)gridDim.x * blockDim.x) {
        if (pad && pad[t]) continue;
        const int pid = position_id(channels, positions, t, C, H, W);
        if (pid >= 0) atomicAdd(counts + pid, 1);
    }
}

// single-CTA exclusive scan of counts -> offsets (n_pos + 1) and cursor (copy of offsets)
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int32_t* __restrict__ counts,
                                                           int n_pos, int32_t* __restrict__ offsets,
                                                           int32_t* __restrict__ cursor) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n_pos; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < n_pos ? counts[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int t = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += u;
            }
            warp_tot[lane] = t;  // inclusive totals of warps
        }
        __syncthreads();
        const int before = carry + (wid ? warp_tot[wid - 1] : 0) + inc - v;
        if (i < n_pos) {
            offsets[i] = before;
            cursor[i] = before;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n_pos] = carry;
}

__global__ void fill_lists_kernel(const int64_t* __restrict__ channels,
                                  const int64_t* __restrict__ positions,
                                  const uint8_t* __restrict__ pad, int64_t n_tok, int C, int H, int W,
                                  int32_t* __restrict__ cursor, int32_t* __restrict__ unsorted) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_tok;
         t += (int64_t)gridDim.x * blockDim.x) {
        if (pad && pad[t]) continue;
        const int pid = position_id(channels, positions, t, C, H, W);
        if (pid >= 0) unsorted[atomicAdd(cursor + pid, 1)] = (int32_t)t;
    }
}

// rank sort of each position's list (ascending token index): makes every later reduction run in
// the reference's token order, hence deterministic and bit-reproducible.
__global__ void __launch_bounds__(128) sort_lists_kernel(const int32_t* __restrict__ offsets,
                                                         const int32_t* __restrict__ unsorted,
                                                         int32_t* __restrict__ sorted) {
    const int pid = blockIdx.x;
    const int beg = offsets[pid], n = offsets[pid + 1] - beg;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int32_t v = unsorted[beg + i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (__ldg(unsorted + beg + j) < v);
        sorted[beg + rank] = v;
    }
}

// ------------------------------------------------------------------------------ batch median
__device__ __forceinline__ uint32_t orderable_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Batch median per (position, coefficient): the element of rank (n-1)/2 of the position's list (torch.median returns
// the LOWER middle for even n, PN:129), found by an MSB-first radix select on order-preserving keys -- bit sliced.
// One CTA of 8 warps per position, one thread per coefficient of a 32-coefficient chunk.  A thread loads its 32 keys of a
// token block into registers and transposes the 32 x 32 bit matrix in place (5 butterfly stages), so that word b holds
// bit b of all 32 tokens; the select is then one AND + POPC per block and bit on a candidate mask instead of one
// compare per token and bit.
//   n <= 64   (a fit batch of up to 64 images: a position occurs at most once per image): every warp owns whole chunks,
//             two token blocks in registers, no barrier: ~45 warp instructions per coefficient;
//   n <= 512  the 8 warps share a chunk, two blocks each; the per-bit counts meet in a double-buffered table (one
//             barrier per bit);
//   longer    counted token by token from global memory, 8 token groups x 32 lanes.
// The first version (one compare per token and bit, tokens re-read from L2 once per bit) took 1.04 ms of the 1.54 ms
// fit step at 64 images; a shared-memory staged version of it 0.76 ms (issue-bound: ~560 warp instructions per
// coefficient).
constexpr int kMedianWarps = 8;

__device__ __forceinline__ void transpose_bits32(uint32_t (&a)[32]) {
    uint32_t m = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            const uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
            a[k] ^= (t << j);
            a[k + j] ^= t;
        }
    }
}

// bit planes of the tokens [t0, t0 + 32) of the list (tokens at or beyond n read as key 0; the candidate mask hides them)
__device__ __forceinline__ void load_bit_planes(uint32_t (&p)[32], const float* __restrict__ x,
                                                const int32_t* __restrict__ list, int t0, int n, int z, int zi,
                                                bool live) {
    if (t0 >= n) {
#pragma unroll
        for (int t = 0; t < 32; ++t) p[t] = 0u;
        return;
    }
    // the block's 32 token indices: one coalesced load, then broadcast by shuffle (no dependent load per token)
    const int lane = threadIdx.x & 31;
    const int my_tok = (t0 + lane < n) ? __ldg(list + t0 + lane) : 0;
#pragma unroll
    for (int t = 0; t < 32; ++t) {
        const int tok = __shfl_sync(0xffffffffu, my_tok, t);
        p[t] = (t0 + t < n && live) ? orderable_key(__ldg(x + (int64_t)tok * z + zi)) : 0u;
    }
    transpose_bits32(p);
}

__device__ __forceinline__ uint32_t first_bits(int count) {  // mask of min(max(count, 0), 32) low bits
    return count >= 32 ? 0xFFFFFFFFu : (count <= 0 ? 0u : ((1u << count) - 1u));
}

__global__ void __launch_bounds__(32 * kMedianWarps, 3)
batch_median_kernel(const float* __restrict__ x, const int32_t* __restrict__ offsets,
                    const int32_t* __restrict__ list, int n_pos, int z, float* __restrict__ packed) {
    __shared__ int partial[2][kMedianWarps][32];
    const int pid = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunks = (z + 31) >> 5;
    const int beg = offsets[pid], n = offsets[pid + 1] - beg;
    if (threadIdx.x == 0) packed[pid] = (float)n;  // batch_n (PN:112-119)
    float* out = packed + n_pos + (int64_t)pid * z;
    list += beg;
    if (n == 0) {
        for (int zi = threadIdx.x; zi < z; zi += blockDim.x) out[zi] = 0.0f;
        return;
    }
    const float fn = (float)n;  // results are batch_median * batch_n, ready to be summed over ranks
    if (n <= 64) {
        for (int chunk = warp; chunk < chunks; chunk += kMedianWarps) {
            const int zi = chunk * 32 + lane;
            const bool live = zi < z;
            uint32_t p0[32], p1[32];
            load_bit_planes(p0, x, list, 0, n, z, zi, live);
            load_bit_planes(p1, x, list, 32, n, z, zi, live);
            uint32_t cand0 = first_bits(n), cand1 = first_bits(n - 32), prefix = 0;
            int r = (n - 1) >> 1;
#pragma unroll
            for (int bit = 31; bit >= 0; --bit) {
                const uint32_t z0 = cand0 & ~p0[bit], z1 = cand1 & ~p1[bit];
                const int cnt0 = __popc(z0) + __popc(z1);
                const bool up = r >= cnt0;
                r -= up ? cnt0 : 0;
                cand0 = up ? (cand0 & p0[bit]) : z0;
                cand1 = up ? (cand1 & p1[bit]) : z1;
                prefix |= up ? (1u << bit) : 0u;
            }
            if (live) out[zi] = __fmul_rn(from_orderable(prefix), fn);
        }
    } else if (n <= 64 * kMedianWarps) {
        for (int chunk = 0; chunk < chunks; ++chunk) {
            const int zi = chunk * 32 + lane;
            const bool live = zi < z;
            const int t0 = 32 * warp, t1 = 32 * (kMedianWarps + warp);
            uint32_t p0[32], p1[32];
            load_bit_planes(p0, x, list, t0, n, z, zi, live);
            load_bit_planes(p1, x, list, t1, n, z, zi, live);
            uint32_t cand0 = first_bits(n - t0), cand1 = first_bits(n - t1), prefix = 0;
            int r = (n - 1) >> 1;
#pragma unroll
            for (int bit = 31; bit >= 0; --bit) {
                const uint32_t z0 = cand0 & ~p0[bit], z1 = cand1 & ~p1[bit];
                int (*tab)[32] = partial[bit & 1];
                tab[warp][lane] = __popc(z0) + __popc(z1);
                __syncthreads();
                int cnt0 = 0;
#pragma unroll
                for (int g = 0; g < kMedianWarps; ++g) cnt0 += tab[g][lane];
                const bool up = r >= cnt0;
                r -= up ? cnt0 : 0;
                cand0 = up ? (cand0 & p0[bit]) : z0;
                cand1 = up ? (cand1 & p1[bit]) : z1;
                prefix |= up ? (1u << bit) : 0u;
            }
            if (warp == 0 && live) out[zi] = __fmul_rn(from_orderable(prefix), fn);
        }
    } else {
        for (int chunk = 0; chunk < chunks; ++chunk) {
            const int zi = chunk * 32 + lane;
            const bool live = zi < z;
            uint32_t prefix = 0, mask = 0;
            int r = (n - 1) >> 1;
            for (int bit = 31; bit >= 0; --bit) {
                const uint32_t bm = 1u << bit;
                int cnt0 = 0;
                if (live)
                    for (int t = warp; t < n; t += kMedianWarps) {
                        const uint32_t k = orderable_key(__ldg(x + (int64_t)__ldg(list + t) * z + zi));
                        cnt0 += ((k & mask) == prefix) && !(k & bm);
                    }
                int (*tab)[32] = partial[bit & 1];
                tab[warp][lane] = cnt0;
                __syncthreads();
                cnt0 = 0;
#pragma unroll
                for (int g = 0; g < kMedianWarps; ++g) cnt0 += tab[g][lane];
                if (r >= cnt0) {
                    r -= cnt0;
                    prefix |= bm;
                }
                mask |= bm;
            }
            if (warp == 0 && live) out[zi] = __fmul_rn(from_orderable(prefix), fn);
        }
    }
}

// median <- (median*n + sum_r batch_median_r*batch_n_r) / clamp(n + batch_n, 1)        PN:135-138
__global__ void update_median_kernel(float* __restrict__ median, const float* __restrict__ n,
                                     const float* __restrict__ packed, int n_pos, int z) {
    const int64_t total = (int64_t)n_pos * z;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int pid = (int)(i / z);
        const float nn = n[pid], bn = packed[pid];
        const float num = __fadd_rn(__fmul_rn(median[i], nn), packed[n_pos + i]);
        median[i] = __fdiv_rn(num, fmaxf(__fadd_rn(nn, bn), 1.0f));
    }
}

// abs_dev[pid, zi] = sum_t |x[t, zi] - median[pid, zi]| in ascending token order       PN:140-143
__global__ void __launch_bounds__(256) abs_dev_kernel(const float* __restrict__ x,
                                                      const int32_t* __restrict__ offsets,
                                                      const int32_t* __restrict__ list,
                                                      const float* __restrict__ median, int z,
                                                      float* __restrict__ abs_dev) {
    const int pid = blockIdx.x;
    const int beg = offsets[pid], n = offsets[pid + 1] - beg;
    for (int zi = threadIdx.x; zi < z; zi += blockDim.x) {
        const float m = median[(int64_t)pid * z + zi];
        float acc = 0.0f;
        for (int t = 0; t < n; ++t)
            acc = __fadd_rn(acc, fabsf(__fsub_rn(__ldg(x + (int64_t)__ldg(list + beg + t) * z + zi), m)));
        abs_dev[(int64_t)pid * z + zi] = acc;
    }
}

// b <- (b*n + (abs_dev / clamp(batch_n,1)) * batch_n) / clamp(n + batch_n, 1)           PN:144-148
__global__ void update_b_kernel(float* __restrict__ b, const float* __restrict__ n,
                                const float* __restrict__ packed, const float* __restrict__ abs_dev,
                                int n_pos, int z) {
    const int64_t total = (int64_t)n_pos * z;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int pid = (int)(i / z);
        const float nn = n[pid], bn = packed[pid];
        const float batch_b = __fdiv_rn(abs_dev[i], fmaxf(bn, 1.0f));
        const float num = __fadd_rn(__fmul_rn(b[i], nn), __fmul_rn(batch_b, bn));
        b[i] = __fdiv_rn(num, fmaxf(__fadd_rn(nn, bn), 1.0f));
    }
}

__global__ void update_n_kernel(float* __restrict__ n, const float* __restrict__ packed, int n_pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pos) n[i] = __fadd_rn(n[i], packed[i]);  // PN:150
}

template <int kVec>
__global__ void __launch_bounds__(256) zero_padding_kernel(const float* __restrict__ x,
                                                           const uint8_t* __restrict__ pad,
                                                           float* __restrict__ out, int64_t n_tok,
                                                           int z) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t tok = warp0; tok < n_tok; tok += n_warps) {
        const bool is_pad = pad[tok] != 0;
        if (kVec == 4) {
            const float4* s = reinterpret_cast<const float4*>(x + tok * z);
            float4* d = reinterpret_cast<float4*>(out + tok * z);
            for (int i = lane; i < z / 4; i += 32)
                st_stream(d + i, is_pad ? make_float4(0.f, 0.f, 0.f, 0.f) : ld_stream(s + i));
        } else {
            for (int i = lane; i < z; i += 32) out[tok * z + i] = is_pad ? 0.0f : x[tok * z + i];
        }
    }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace dcta

using namespace dcta;

extern "C" int dcta_patchnorm_apply(const float* x, const int64_t* channels,
                                    const int64_t* positions, const float* median, const float* b,
                                    float* out, int64_t n_tok, int z, int C, int H, int W, float eps,
                                    float lo, float hi, int inverse, void* stream) {
    DCTA_REQUIRE(x && channels && positions && median && b && out, "patchnorm_apply: null pointer");
    DCTA_REQUIRE(n_tok >= 0 && z > 0 && C > 0 && H > 0 && W > 0, "patchnorm_apply: bad sizes");
    if (n_tok == 0) return DCTA_OK;
    const bool vec = (z % 4 == 0) && al16(x) && al16(out) && al16(median) && al16(b);
    const int grid = (vec && z <= 256) ? grid_for(n_tok, 256) : grid_for(n_tok, 8);
    cudaStream_t st = as_stream(stream);
#define LAUNCH(INV, VEC) patchnorm_apply_kernel<INV, VEC><<<grid, 256, 0, st>>>(x, channels, positions, median, b, out, n_tok, z, C, H, W, eps, lo, hi)
    if (inverse) { if (vec) LAUNCH(true, 4); else LAUNCH(true, 1); }
    else         { if (vec) LAUNCH(false, 4); else LAUNCH(false, 1); }
#undef LAUNCH
    return check_launch("patchnorm_apply");
}

extern "C" int dcta_patchnorm_build_lists(const int64_t* channels, const int64_t* positions,
                                          const uint8_t* key_pad_mask, int64_t n_tok, int C, int H,
                                          int W, int32_t* counts, int32_t* offsets, int32_t* cursor,
                                          int32_t* list, void* stream) {
    DCTA_REQUIRE(channels && positions && counts && offsets && cursor && list,
                 "patchnorm_build_lists: null pointer");
    DCTA_REQUIRE(n_tok >= 0 && n_tok < (1ll << 31) && C > 0 && H > 0 && W > 0,
                 "patchnorm_build_lists: bad sizes");
    const int n_pos = C * H * W;
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(counts, 0, sizeof(int32_t) * n_pos, st);
    const int grid = grid_for(n_tok > 0 ? n_tok : 1, 256);
    // `list` holds 2*n_tok entries: [sorted | unsorted scratch]
    int32_t* unsorted = list + n_tok;
    if (n_tok > 0) count_positions_kernel<<<grid, 256, 0, st>>>(channels, positions, key_pad_mask, n_tok, C, H, W, counts);
    scan_counts_kernel<<<1, 1024, 0, st>>>(counts, n_pos, offsets, cursor);
    if (n_tok > 0) {
        fill_lists_kernel<<<grid, 256, 0, st>>>(channels, positions, key_pad_mask, n_tok, C, H, W, cursor, unsorted);
        sort_lists_kernel<<<n_pos, 128, 0, st>>>(offsets, unsorted, list);
    }
    return check_launch("patchnorm_build_lists");
}

extern "C" int dcta_patchnorm_batch_median(const float* x, const int32_t* offsets,
                                           const int32_t* list, int n_pos, int z, float* packed,
                                           void* stream) {
    DCTA_REQUIRE(x && offsets && list && packed, "patchnorm_batch_median: null pointer");
    DCTA_REQUIRE(n_pos > 0 && z > 0, "patchnorm_batch_median: bad sizes");
    batch_median_kernel<<<n_pos, 32 * kMedianWarps, 0, as_stream(stream)>>>(x, offsets, list, n_pos, z, packed);
    return check_launch("patchnorm_batch_median");
}

extern "C" int dcta_patchnorm_update_median(float* median, const float* n, const float* packed,
                                            int n_pos, int z, void* stream) {
    DCTA_REQUIRE(median && n && packed && n_pos > 0 && z > 0, "patchnorm_update_median: bad args");
    update_median_kernel<<<grid_for((int64_t)n_pos * z, 256), 256, 0, as_stream(stream)>>>(median, n, packed, n_pos, z);
    return check_launch("patchnorm_update_median");
}

extern "C" int dcta_patchnorm_abs_dev(const float* x, const int32_t* offsets, const int32_t* list,
                                      const float* median, int n_pos, int z, float* abs_dev,
                                      void* stream) {
    DCTA_REQUIRE(x && offsets && list && median && abs_dev && n_pos > 0 && z > 0,
                 "patchnorm_abs_dev: bad args");
    int threads = ((z + 31) / 32) * 32;
    if (threads > 256) threads = 256;
    abs_dev_kernel<<<n_pos, threads, 0, as_stream(stream)>>>(x, offsets, list, median, z, abs_dev);
    return check_launch("patchnorm_abs_dev");
}

extern "C" int dcta_patchnorm_update_b(float* b, float* n, const float* packed, const float* abs_dev,
                                       int n_pos, int z, void* stream) {
    DCTA_REQUIRE(b && n && packed && abs_dev && n_pos > 0 && z > 0, "patchnorm_update_b: bad args");
    cudaStream_t st = as_stream(stream);
    update_b_kernel<<<grid_for((int64_t)n_pos * z, 256), 256, 0, st>>>(b, n, packed, abs_dev, n_pos, z);
    update_n_kernel<<<(n_pos + 255) / 256, 256, 0, st>>>(n, packed, n_pos);
    return check_launch("patchnorm_update_b");
}

extern "C" int dcta_zero_padding(const float* x, const uint8_t* key_pad_mask, float* out,
                                 int64_t n_tok, int z, void* stream) {
    DCTA_REQUIRE(x && key_pad_mask && out && n_tok >= 0 && z > 0, "zero_padding: bad args");
    if (n_tok == 0) return DCTA_OK;
    const int grid = grid_for(n_tok, 8);
    if (z % 4 == 0 && al16(x) && al16(out))
        zero_padding_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(x, key_pad_mask, out, n_tok, z);
    else
        zero_padding_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(x, key_pad_mask, out, n_tok, z);
    return check_launch("zero_padding");
}
