// VectorQuantize codebook learning: cluster statistics of a token batch, the EMA codebook update and the
// k-means mean update.  Reference: vector_quantize.py:180-220 (kmeans), :479-500 (EMA update in EuclideanCodebook.forward),
// :38-44 (ema_inplace = lerp_), :100-102 (laplace_smoothing).
#include "common.cuh"

namespace dcta {

// counts[c] += 1, sums[c, :] += x[t, :] for every token t with mask[t] != 0 (mask nullable) and code idx[t] = c.
// One warp per token, the lanes across the row; 128-bit vector reductions when the rows allow it.
__global__ void __launch_bounds__(256) vq_cluster_stats_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx,
                                                               const uint8_t* __restrict__ mask, int64_t n_tok, int d, int n_codes,
                                                               float* __restrict__ counts, float* __restrict__ sums, int vec) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp0; t < n_tok; t += n_warps) {
        if (mask != nullptr && mask[t] == 0) continue;
        const int64_t c = idx[t];
        if (c < 0 || c >= n_codes) continue;
        if (lane == 0) atomicAdd(counts + c, 1.0f);
        if (vec) {
            const float4* src = reinterpret_cast<const float4*>(x + t * d);
            float4* dst = reinterpret_cast<float4*>(sums + c * d);
            for (int i = lane; i < (d >> 2); i += 32) atomicAdd(dst + i, __ldg(src + i));
        } else {
            for (int i = lane; i < d; i += 32) atomicAdd(sums + c * d + i, __ldg(x + t * d + i));
        }
    }
}

// torch.lerp for fp32 (ATen lerp: the branch keeps the result monotonic in the weight)
__device__ __forceinline__ float torch_lerp(float start, float end, float w) {
    const float diff = end - start;
    return w < 0.5f ? start + w * diff : end - diff * (1.0f - w);
}

// cluster_size <- lerp(cluster_size, batch counts, 1 - decay); total += sum of the new cluster sizes
__global__ void __launch_bounds__(256) vq_ema_counts_kernel(float* __restrict__ cluster_size, const float* __restrict__ counts,
                                                            int n_codes, float w, float* __restrict__ total) {
    __shared__ float red[8];
    float s = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_codes; i += gridDim.x * blockDim.x) {
        const float v = torch_lerp(cluster_size[i], counts[i], w);
        cluster_size[i] = v;
        s += v;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += red[k];
        atomicAdd(total, t);
    }
}

// embed_avg <- lerp(embed_avg, batch sums, 1 - decay); embed <- embed_avg / (laplace_smoothing(cluster_size) * total)
__global__ void __launch_bounds__(256) vq_ema_embed_kernel(float* __restrict__ embed, float* __restrict__ embed_avg,
                                                           const float* __restrict__ sums, const float* __restrict__ cluster_size,
                                                           const float* __restrict__ total, int n_codes, int d, float w, float eps) {
    const float tot = *total;
    const int64_t n = (int64_t)n_codes * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i / d);
        const float avg = torch_lerp(embed_avg[i], sums[i], w);
        embed_avg[i] = avg;
        const float smoothed = ((cluster_size[c] + eps) / (tot + (float)n_codes * eps)) * tot;
        embed[i] = avg / smoothed;
    }
}

// k-means: means[c] <- sums[c] / counts[c] where the cluster is not empty (vector_quantize.py:203-220)
__global__ void __launch_bounds__(256) vq_kmeans_means_kernel(float* __restrict__ means, const float* __restrict__ counts,
                                                              const float* __restrict__ sums, int n_codes, int d) {
    const int64_t n = (int64_t)n_codes * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float cnt = counts[i / d];
        if (cnt > 0.f) means[i] = sums[i] / cnt;
    }
}

}  // namespace dcta

using namespace dcta;

extern "C" int dcta_vq_cluster_stats(const float* x, const int64_t* indices, const uint8_t* mask, int64_t n_tok, int d,
                                     int n_codes, float* counts, float* sums, void* stream) {
    DCTA_REQUIRE(x && indices && counts && sums && d > 0 && n_codes > 0 && n_tok >= 0, "vq_cluster_stats: bad args");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(counts, 0, sizeof(float) * n_codes, st);
    cudaMemsetAsync(sums, 0, sizeof(float) * (size_t)n_codes * d, st);
    if (n_tok == 0) return DCTA_OK;
    const int vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(sums)) & 15) == 0;
    vq_cluster_stats_kernel<<<grid_for(n_tok, 8), 256, 0, st>>>(x, indices, mask, n_tok, d, n_codes, counts, sums, vec);
    return check_launch("vq_cluster_stats");
}

extern "C" int dcta_vq_ema_update(float* embed, float* cluster_size, float* embed_avg, const float* counts, const float* sums,
                                  int n_codes, int d, float decay, float eps, float* total_scratch, void* stream) {
    DCTA_REQUIRE(embed && cluster_size && embed_avg && counts && sums && total_scratch && n_codes > 0 && d > 0,
                 "vq_ema_update: bad args");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(total_scratch, 0, sizeof(float), st);
    const float w = 1.0f - decay;
    vq_ema_counts_kernel<<<grid_for(n_codes, 256, 1), 256, 0, st>>>(cluster_size, counts, n_codes, w, total_scratch);
    vq_ema_embed_kernel<<<grid_for((int64_t)n_codes * d, 256), 256, 0, st>>>(embed, embed_avg, sums, cluster_size, total_scratch,
                                                                           n_codes, d, w, eps);
    return check_launch("vq_ema_update");
}

// ------------------------------------------------------------------------------ commitment loss
// VQ:976-1003: loss = mean over the valid tokens and their dim elements of (q - x)^2, q detached; its gradient
// together with the straight-through estimator's (VQ:944-952: the output q passes its gradient to x unchanged):
//   grad_x = grad_q + grad_loss * 2 (x - q) * mask / (n_valid * dim).
// One flat 128-bit pass each (the eager expressions cost six passes and a host synchronisation for se[mask]);
// deterministic: per-CTA partial sums in fixed slots, summed by one CTA in a fixed order.
namespace dcta {
constexpr int kMseCtas = 4 * kNumSMs;

template <int kVec>
__global__ void __launch_bounds__(256) masked_mse_partial_kernel(const float* __restrict__ x, const float* __restrict__ q,
                                                                 const uint8_t* __restrict__ mask, float* __restrict__ partial,
                                                                 int64_t n_tok, int dim) {
    __shared__ float red[8];
    const int64_t n_chunks = n_tok * dim / kVec, stride = (int64_t)gridDim.x * blockDim.x;
    float se = 0.f, cnt = 0.f;
    for (int64_t it0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it0 < n_chunks; it0 += 2 * stride) {
        float a[2][kVec], b[2][kVec], m[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t it = it0 + u * stride;
            m[u] = 0.f;
#pragma unroll
            for (int j = 0; j < kVec; ++j) { a[u][j] = 0.f; b[u][j] = 0.f; }
            if (it < n_chunks) {
                if (kVec == 4) {
                    const float4 fa = ld_stream(reinterpret_cast<const float4*>(x) + it);
                    const float4 fb = ld_stream(reinterpret_cast<const float4*>(q) + it);
                    a[u][0] = fa.x; a[u][1 % kVec] = fa.y; a[u][2 % kVec] = fa.z; a[u][3 % kVec] = fa.w;
                    b[u][0] = fb.x; b[u][1 % kVec] = fb.y; b[u][2 % kVec] = fb.z; b[u][3 % kVec] = fb.w;
                } else {
                    a[u][0] = x[it];
                    b[u][0] = q[it];
                }
                const int64_t e = it * kVec, tok = e / dim;
                m[u] = (mask == nullptr || mask[tok]) ? 1.f : 0.f;
                if (e - tok * dim == 0) cnt += m[u];
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float s4 = 0.f;
#pragma unroll
            for (int j = 0; j < kVec; ++j) {
                const float dq = b[u][j] - a[u][j];
                s4 = fmaf(dq, dq, s4);
            }
            se = fmaf(m[u], s4, se);
        }
    }
    se = warp_sum(se);
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = se; }
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x == 0) for (int w = 0; w < 8; ++w) s += red[w];
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float n = 0.f;
        for (int w = 0; w < 8; ++w) n += red[w];
        partial[2 * blockIdx.x] = s;
        partial[2 * blockIdx.x + 1] = n;
    }
}

// result[0] = sum / (n_valid * dim) (0 / 0 = NaN like the mean of an empty selection), result[1] = n_valid
__global__ void masked_mse_final_kernel(const float* __restrict__ partial, int n_partial, int dim, float* __restrict__ result) {
    if (threadIdx.x != 0) return;
    float s = 0.f, n = 0.f;
    for (int i = 0; i < n_partial; ++i) { s += partial[2 * i]; n += partial[2 * i + 1]; }
    result[0] = s / (n * (float)dim);
    result[1] = n;
}

template <int kVec>
__global__ void __launch_bounds__(256) masked_mse_bwd_kernel(const float* __restrict__ x, const float* __restrict__ q,
                                                             const uint8_t* __restrict__ mask, const float* __restrict__ result,
                                                             const float* __restrict__ grad_loss, const float* __restrict__ grad_q,
                                                             float* __restrict__ grad_x, int64_t n_tok, int dim) {
    const float nv = result[1];
    const float k = 2.f * grad_loss[0] / (nv * (float)dim);
    const int64_t total = n_tok * dim / kVec;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const bool on = mask == nullptr || mask[i * kVec / dim] != 0;
        if (kVec == 4) {
            const float4 a = ld_stream(reinterpret_cast<const float4*>(x) + i);
            const float4 b = ld_stream(reinterpret_cast<const float4*>(q) + i);
            float4 g = grad_q ? ld_stream(reinterpret_cast<const float4*>(grad_q) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (on) {
                g.x = fmaf(k, a.x - b.x, g.x); g.y = fmaf(k, a.y - b.y, g.y);
                g.z = fmaf(k, a.z - b.z, g.z); g.w = fmaf(k, a.w - b.w, g.w);
            }
            st_stream(reinterpret_cast<float4*>(grad_x) + i, g);
        } else {
            float g = grad_q ? grad_q[i] : 0.f;
            if (on) g = fmaf(k, x[i] - q[i], g);
            grad_x[i] = g;
        }
    }
}
}  // namespace dcta

extern "C" int dcta_masked_mse(const float* x, const float* q, const uint8_t* mask, int64_t n_tok, int dim, float* scratch,
                               float* result, void* stream) {
    using namespace dcta;
    DCTA_REQUIRE(x && q && scratch && result && n_tok >= 0 && dim > 0, "masked_mse: bad args");
    const bool vec = dim % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(q)) & 15) == 0;
    const int64_t chunks = n_tok * dim / (vec ? 4 : 1);
    int grid = (int)ceil_div(chunks > 0 ? chunks : 1, 512);
    if (grid > kMseCtas) grid = kMseCtas;
    cudaStream_t st = as_stream(stream);
    if (vec) masked_mse_partial_kernel<4><<<grid, 256, 0, st>>>(x, q, mask, scratch, n_tok, dim);
    else masked_mse_partial_kernel<1><<<grid, 256, 0, st>>>(x, q, mask, scratch, n_tok, dim);
    masked_mse_final_kernel<<<1, 32, 0, st>>>(scratch, grid, dim, result);
    return check_launch("masked_mse");
}

extern "C" int dcta_masked_mse_scratch_floats(void) { return 2 * dcta::kMseCtas; }

extern "C" int dcta_masked_mse_backward(const float* x, const float* q, const uint8_t* mask, const float* result,
                                        const float* grad_loss, const float* grad_q, float* grad_x, int64_t n_tok, int dim,
                                        void* stream) {
    using namespace dcta;
    DCTA_REQUIRE(x && q && result && grad_loss && grad_x && n_tok >= 0 && dim > 0, "masked_mse_backward: bad args");
    if (n_tok == 0) return DCTA_OK;
    const bool vec = dim % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(q) |
                                      reinterpret_cast<uintptr_t>(grad_q) | reinterpret_cast<uintptr_t>(grad_x)) & 15) == 0;
    if (vec) masked_mse_bwd_kernel<4><<<grid_for(n_tok * dim / 4, 256), 256, 0, as_stream(stream)>>>(x, q, mask, result, grad_loss, grad_q, grad_x, n_tok, dim);
    else masked_mse_bwd_kernel<1><<<grid_for(n_tok * dim, 256), 256, 0, as_stream(stream)>>>(x, q, mask, result, grad_loss, grad_q, grad_x, n_tok, dim);
    return check_launch("masked_mse_backward");
}

extern "C" int dcta_vq_kmeans_means(float* means, const float* counts, const float* sums, int n_codes, int d, void* stream) {
    DCTA_REQUIRE(means && counts && sums && n_codes > 0 && d > 0, "vq_kmeans_means: bad args");
    vq_kmeans_means_kernel<<<grid_for((int64_t)n_codes * d, 256), 256, 0, as_stream(stream)>>>(means, counts, sums, n_codes, d);
    return check_launch("vq_kmeans_means");
}
