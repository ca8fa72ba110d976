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

extern "C" int dcta_vq_kmeans_means(float* means, const float* counts, const float* sums, int n_codes, int d, void* stream) {
    DCTA_REQUIRE(means && counts && sums && n_codes > 0 && d > 0, "vq_kmeans_means: bad args");
    vq_kmeans_means_kernel<<<grid_for((int64_t)n_codes * d, 256), 256, 0, as_stream(stream)>>>(means, counts, sums, n_codes, d);
    return check_launch("vq_kmeans_means");
}
