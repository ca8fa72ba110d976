// Lookup-free (sign) quantiser, its bit-packed codes, and the training-time terms
// (reference: lfq.py:35-227; util.py:341-410).  quantize / indices_to_codes are HBM-bound
// streams: x read once (128-bit), q written once, 8 B of code per (token, codebook).
#include "common.cuh"

namespace dcta {

constexpr int kLfqTok = 32;  // tokens per CTA tile in lfq_quantize

// ------------------------------------------------------------------------------ quantize
template <int kVec>
__global__ void __launch_bounds__(256) lfq_quantize_kernel(const float* __restrict__ x,
                                                           float* __restrict__ q,
                                                           int64_t* __restrict__ indices,
                                                           int64_t n_tok, int c, int d, float scale) {
    extern __shared__ uint8_t sgn[];  // kLfqTok * c * d sign bits, one per byte
    const int cd = c * d;
    const int64_t n_tiles = (n_tok + kLfqTok - 1) / kLfqTok;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t t0 = tile * kLfqTok;
        const int nt = (int)min((int64_t)kLfqTok, n_tok - t0);
        const int ne = nt * cd;
        const float* xs = x + t0 * cd;
        float* qs = q ? q + t0 * cd : nullptr;
        if (kVec == 4) {
            // four 128-bit loads in flight per thread before the first store (the streaming accessors are volatile asm:
            // one load + its stores per iteration left a single load in flight)
            const int n4 = ne / 4;
            for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * blockDim.x;
                    if (i < n4) v[u] = ld_stream(reinterpret_cast<const float4*>(xs) + i);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * blockDim.x;
                    if (i < n4) {
                        // lfq.py:175  where(x > 0, +s, -s): zero and NaN go to -s / bit 0
                        const uchar4 b = make_uchar4(v[u].x > 0.f, v[u].y > 0.f, v[u].z > 0.f, v[u].w > 0.f);
                        reinterpret_cast<uchar4*>(sgn)[i] = b;
                        if (qs)
                            st_stream(reinterpret_cast<float4*>(qs) + i,
                                      make_float4(b.x ? scale : -scale, b.y ? scale : -scale,
                                                  b.z ? scale : -scale, b.w ? scale : -scale));
                    }
                }
            }
        } else {
            for (int i = threadIdx.x; i < ne; i += blockDim.x) {
                const bool b = xs[i] > 0.f;
                sgn[i] = b;
                if (qs) qs[i] = b ? scale : -scale;
            }
        }
        __syncthreads();
        // lfq.py:187  indices = sum_i bit_i * 2^(d-1-i)   (MSB first, lfq.py:87)
        for (int o = threadIdx.x; o < nt * c; o += blockDim.x) {
            const uint8_t* sb = sgn + (size_t)o * d;  // (tok, codebook) pairs are contiguous runs of d
            unsigned long long code = 0;
            for (int i = 0; i < d; ++i) code = (code << 1) | sb[i];
            indices[t0 * c + o] = (int64_t)code;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------ indices -> codes
template <int kVec>
__global__ void __launch_bounds__(256) lfq_codes_kernel(const int64_t* __restrict__ indices,
                                                        float* __restrict__ codes, int64_t n_elem,
                                                        int c, int d, float scale) {
    const int cd = c * d;
    const int64_t n_items = n_elem / kVec;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items;
         it += (int64_t)gridDim.x * blockDim.x) {
        float v[kVec];
#pragma unroll
        for (int j = 0; j < kVec; ++j) {
            const int64_t e = it * kVec + j;
            const int64_t tok = e / cd;
            const int k = (int)(e - tok * cd);
            const int cb = k / d, i = k - cb * d;
            const long long idx = __ldg(indices + tok * c + cb);
            v[j] = ((idx >> (d - 1 - i)) & 1) ? scale : -scale;  // lfq.py:118-120
        }
        if (kVec == 4) st_stream(reinterpret_cast<float4*>(codes) + it, make_float4(v[0], v[1], v[2], v[3]));
        else codes[it] = v[0];
    }
}

// ------------------------------------------------------------------------------ block reduce
__device__ __forceinline__ float block_sum(float v, float* red /* >= 32 floats */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = 0.f;
    if (wid == 0) {
        t = lane < (int)((blockDim.x + 31) >> 5) ? red[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float t = lane < (int)((blockDim.x + 31) >> 5) ? red[lane] : -INFINITY;
        t = warp_max(t);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

// ------------------------------------------------------------------------------ commit loss
// deterministic two-stage reduction: per-CTA partials in fixed slots, then one CTA sums them.
// One flat grid-stride pass over the (n_tok, cd) array in chunks of kVec elements (a chunk never straddles two tokens:
// cd % kVec == 0), loaded unconditionally so that four 128-bit loads are in flight per thread; masked tokens contribute
// through a 0 / 1 factor.  (One warp per token with scalar loads ran at 0.77 TB/s: 800 us for the 617 MB of config 2.)
template <int kVec>
__global__ void __launch_bounds__(256) lfq_commit_partial_kernel(const float* __restrict__ x,
                                                                 const uint8_t* __restrict__ mask,
                                                                 float* __restrict__ partial,
                                                                 int64_t n_tok, int cd, float scale) {
    __shared__ float red[32];
    const int64_t n_chunks = n_tok * cd / kVec;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float se = 0.f, cnt = 0.f;
    for (int64_t it0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it0 < n_chunks; it0 += 4 * stride) {
        float v[4][kVec];
        float m[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t it = it0 + u * stride;
            m[u] = 0.f;
#pragma unroll
            for (int j = 0; j < kVec; ++j) v[u][j] = 0.f;
            if (it < n_chunks) {
                if (kVec == 4) {
                    const float4 f = ld_stream(reinterpret_cast<const float4*>(x) + it);
                    v[u][0] = f.x; v[u][1 % kVec] = f.y; v[u][2 % kVec] = f.z; v[u][3 % kVec] = f.w;
                } else {
                    v[u][0] = x[it];
                }
                const int64_t e = it * kVec, tok = e / cd;
                m[u] = mask[tok] ? 1.f : 0.f;
                if (e - tok * cd == 0) cnt += m[u];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float s4 = 0.f;
#pragma unroll
            for (int j = 0; j < kVec; ++j) {
                const float dq = v[u][j] - (v[u][j] > 0.f ? scale : -scale);
                s4 = fmaf(dq, dq, s4);
            }
            se = fmaf(m[u], s4, se);
        }
    }
    const float s = block_sum(se, red);
    const float n = block_sum(cnt, red);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = s;
        partial[2 * blockIdx.x + 1] = n;
    }
}
__global__ void __launch_bounds__(256) lfq_commit_final_kernel(const float* __restrict__ partial,
                                                               int n_partial, int cd,
                                                               float* __restrict__ result) {
    __shared__ float red[32];
    float s = 0.f, n = 0.f;
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) {
        s += partial[2 * i];
        n += partial[2 * i + 1];
    }
    s = block_sum(s, red);
    n = block_sum(n, red);
    if (threadIdx.x == 0) result[0] = s / (n * (float)cd);  // lfq.py:199 masked_mean(dim=0).sum(0).mean()
}

// ------------------------------------------------------------------------------ distance
// lfq.py:191  distance[t, cb, j] = -2 * sum_i x[t, cb, i] * codebook[j, i],  codebook[j,i] = +-scale
__global__ void __launch_bounds__(256) lfq_distance_kernel(const float* __restrict__ x,
                                                           float* __restrict__ distance,
                                                           int64_t n_rows, int d, float scale) {
    const int n_codes = 1 << d;
    for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
        float xv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) xv[i] = i < d ? __ldg(x + row * d + i) * scale : 0.f;
        float* out = distance + row * n_codes;
        for (int j = threadIdx.x; j < n_codes; j += blockDim.x) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < d) s += ((j >> (d - 1 - i)) & 1) ? xv[i] : -xv[i];
            out[j] = -2.f * s;
        }
    }
}

// ------------------------------------------------------------------------------ entropy loss
// UT:355-387.  scratch = [acc (n_codes) | sum_rows p.logp | n_valid]
__global__ void __launch_bounds__(256) entropy_rows_kernel(const float* __restrict__ affinity,
                                                           const uint8_t* __restrict__ mask,
                                                           float* __restrict__ scratch,
                                                           int64_t n_tok, int c, int n_codes,
                                                           float temperature, float eps) {
    extern __shared__ float acc[];  // n_codes per-CTA accumulators of masked probabilities
    __shared__ float red[32];
    for (int j = threadIdx.x; j < n_codes; j += blockDim.x) acc[j] = 0.f;
    float plogp_total = 0.f, valid_total = 0.f;
    for (int64_t tok = blockIdx.x; tok < n_tok; tok += gridDim.x) {
        if (!mask[tok]) continue;  // masked_mean multiplies padding rows by 0 (UT:346-353)
        if (threadIdx.x == 0) valid_total += 1.f;
        for (int cb = 0; cb < c; ++cb) {
            const float* row = affinity + (tok * c + cb) * (int64_t)n_codes;
            float mx = -INFINITY;
            for (int j = threadIdx.x; j < n_codes; j += blockDim.x)
                mx = fmaxf(mx, __fdiv_rn(__ldg(row + j), temperature) + eps);
            mx = block_max(mx, red);
            float se = 0.f;
            for (int j = threadIdx.x; j < n_codes; j += blockDim.x)
                se += expf(__fdiv_rn(__ldg(row + j), temperature) + eps - mx);
            se = block_sum(se, red);
            const float lse = logf(se), inv = 1.f / se;
            float pl = 0.f;
            for (int j = threadIdx.x; j < n_codes; j += blockDim.x) {
                const float l = __fdiv_rn(__ldg(row + j), temperature) + eps - mx;
                const float p = expf(l) * inv;
                acc[j] += p;                   // thread-private slots: j == tid (mod blockDim)
                pl = fmaf(p, l - lse, pl);     // p * log_softmax
            }
            plogp_total += pl;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n_codes; j += blockDim.x)
        if (acc[j] != 0.f) atomicAdd(scratch + j, acc[j]);
    const float s = block_sum(plogp_total, red);
    if (threadIdx.x == 0) {
        atomicAdd(scratch + n_codes, s);
        atomicAdd(scratch + n_codes + 1, valid_total);
    }
}
__global__ void __launch_bounds__(256) entropy_final_kernel(const float* __restrict__ scratch, int c,
                                                            int n_codes, float eps,
                                                            float* __restrict__ result) {
    __shared__ float red[32];
    const float n_valid = scratch[n_codes + 1];
    float ae = 0.f;
    for (int j = threadIdx.x; j < n_codes; j += blockDim.x) {
        const float avg = scratch[j] / n_valid / (float)c;  // masked_mean(dim=0).mean(dim=0)
        ae += avg * logf(avg + eps);
    }
    ae = block_sum(ae, red);
    if (threadIdx.x == 0) {
        const float avg_entropy = -ae;
        const float sample_entropy = -(scratch[n_codes] / n_valid);
        result[0] = sample_entropy - avg_entropy;
    }
}

// ------------------------------------------------------------------------------ perplexity
__global__ void histogram_kernel(const int64_t* __restrict__ codes, int64_t n, int codebook_size,
                                 int64_t null_index, unsigned long long* __restrict__ counts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = codes[i];
        if (v == null_index || v < 0 || v >= codebook_size) continue;
        atomicAdd(counts + v, 1ull);
    }
}
__global__ void __launch_bounds__(256) perplexity_final_kernel(const unsigned long long* __restrict__ counts,
                                                               int codebook_size,
                                                               float* __restrict__ result) {
    __shared__ float red[32];
    float tot = 0.f;
    for (int j = threadIdx.x; j < codebook_size; j += blockDim.x) tot += (float)counts[j];
    tot = block_sum(tot, red);
    float e = 0.f;
    for (int j = threadIdx.x; j < codebook_size; j += blockDim.x) {
        const float p = (float)counts[j] / tot;
        if (p != 0.f) e += p * log2f(p);
    }
    e = block_sum(e, red);
    if (threadIdx.x == 0) result[0] = exp2f(-e);
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace dcta

using namespace dcta;

extern "C" int dcta_lfq_quantize(const float* x, float* q, int64_t* indices, int64_t n_tok, int c,
                                 int d, float scale, void* stream) {
    DCTA_REQUIRE(x && indices, "lfq_quantize: null pointer");
    DCTA_REQUIRE(n_tok >= 0 && c > 0 && d > 0 && d <= 62, "lfq_quantize: bad sizes c=%d d=%d", c, d);
    if (n_tok == 0) return DCTA_OK;
    const int cd = c * d;
    const size_t smem = (size_t)kLfqTok * cd;
    DCTA_REQUIRE(smem <= 200 * 1024, "lfq_quantize: c*d=%d too large", cd);
    const int grid = grid_for((n_tok + kLfqTok - 1) / kLfqTok, 1, 6);
    const bool vec = (cd % 4 == 0) && al16(x) && (!q || al16(q));
    if (vec) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(lfq_quantize_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lfq_quantize_kernel<4><<<grid, 256, smem, as_stream(stream)>>>(x, q, indices, n_tok, c, d, scale);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(lfq_quantize_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lfq_quantize_kernel<1><<<grid, 256, smem, as_stream(stream)>>>(x, q, indices, n_tok, c, d, scale);
    }
    return check_launch("lfq_quantize");
}

extern "C" int dcta_lfq_indices_to_codes(const int64_t* indices, float* codes, int64_t n_tok, int c,
                                         int d, float scale, void* stream) {
    DCTA_REQUIRE(indices && codes, "lfq_indices_to_codes: null pointer");
    DCTA_REQUIRE(n_tok >= 0 && c > 0 && d > 0 && d <= 62, "lfq_indices_to_codes: bad sizes");
    if (n_tok == 0) return DCTA_OK;
    const int64_t n_elem = n_tok * c * d;
    if ((c * d) % 4 == 0 && al16(codes))
        lfq_codes_kernel<4><<<grid_for(n_elem / 4, 256), 256, 0, as_stream(stream)>>>(indices, codes, n_elem, c, d, scale);
    else
        lfq_codes_kernel<1><<<grid_for(n_elem, 256), 256, 0, as_stream(stream)>>>(indices, codes, n_elem, c, d, scale);
    return check_launch("lfq_indices_to_codes");
}

extern "C" int dcta_lfq_commit_loss(const float* x, const uint8_t* mask, float* result,
                                    float* scratch2, int64_t n_tok, int cd, float scale,
                                    void* stream) {
    DCTA_REQUIRE(x && mask && result && scratch2 && n_tok >= 0 && cd > 0, "lfq_commit_loss: bad args");
    const bool vec = cd % 4 == 0 && al16(x);
    const int grid = grid_for(n_tok > 0 ? n_tok * cd / (vec ? 4 : 1) : 1, 1024, 4);  // <= 592 CTAs -> 1184 scratch floats
    if (vec) lfq_commit_partial_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(x, mask, scratch2, n_tok, cd, scale);
    else lfq_commit_partial_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(x, mask, scratch2, n_tok, cd, scale);
    lfq_commit_final_kernel<<<1, 256, 0, as_stream(stream)>>>(scratch2, grid, cd, result);
    return check_launch("lfq_commit_loss");
}

extern "C" int dcta_lfq_distance(const float* x, float* distance, int64_t n_tok, int c, int d,
                                 float scale, void* stream) {
    DCTA_REQUIRE(x && distance && n_tok >= 0 && c > 0, "lfq_distance: bad args");
    DCTA_REQUIRE(d > 0 && d <= 16, "lfq_distance: codebook_dim %d > 16 would need 2^d floats per row", d);
    if (n_tok == 0) return DCTA_OK;
    const int64_t rows = n_tok * c;
    lfq_distance_kernel<<<grid_for(rows, 1), 256, 0, as_stream(stream)>>>(x, distance, rows, d, scale);
    return check_launch("lfq_distance");
}

extern "C" int dcta_entropy_loss(const float* affinity, const uint8_t* mask, float* scratch,
                                 float* result, int64_t n_tok, int c, int n_codes,
                                 float temperature, float eps, void* stream) {
    DCTA_REQUIRE(affinity && mask && scratch && result, "entropy_loss: null pointer");
    DCTA_REQUIRE(n_tok >= 0 && c > 0 && n_codes > 0 && n_codes <= 49152, "entropy_loss: bad sizes");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(scratch, 0, sizeof(float) * (n_codes + 2), st);
    const size_t smem = sizeof(float) * n_codes;
    if (smem > 48 * 1024) cudaFuncSetAttribute(entropy_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (n_tok > 0)
        entropy_rows_kernel<<<grid_for(n_tok, 1, 2), 256, smem, st>>>(affinity, mask, scratch, n_tok, c, n_codes, temperature, eps);
    entropy_final_kernel<<<1, 256, 0, st>>>(scratch, c, n_codes, eps, result);
    return check_launch("entropy_loss");
}

extern "C" int dcta_perplexity(const int64_t* codes, int64_t n, int codebook_size,
                               int64_t null_index, int64_t* counts, float* result, void* stream) {
    DCTA_REQUIRE(codes && counts && result && n >= 0 && codebook_size > 0, "perplexity: bad args");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(counts, 0, sizeof(int64_t) * codebook_size, st);
    if (n > 0)
        histogram_kernel<<<grid_for(n, 256), 256, 0, st>>>(codes, n, codebook_size, null_index, reinterpret_cast<unsigned long long*>(counts));
    perplexity_final_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(counts), codebook_size, result);
    return check_launch("perplexity");
}
