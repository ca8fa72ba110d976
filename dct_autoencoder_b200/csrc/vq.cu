// VectorQuantize eval path: nearest codebook entry under the reference's Euclidean distance
// (reference: vector_quantize.py:29-33 cdist, :467-469 argmax(-dist), :222-226/:477 gather).
// The (n_tok, n_codes) distance matrix is never materialised: each CTA owns 128 tokens, walks the
// codebook in 128-code tiles with a register-tiled fp32 FFMA GEMM and keeps a running arg-min.
// This is the exact-fp32 path (FMA-pipe bound); see DESIGN.md for the tensor-core path.
#include "common.cuh"

namespace dcta {

constexpr int VM = 128, VN = 128, VK = 8, VPAD = 4;

__global__ void __launch_bounds__(256) row_sumsq_kernel(const float* __restrict__ e,
                                                        float* __restrict__ e2, int n, int d) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    float s = 0.f;
    for (int i = lane; i < d; i += 32) {
        const float v = e[(int64_t)row * d + i];
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) e2[row] = s;
}

// torch.argmax(-sqrt(d2)): first minimum of sqrt(d2); NaN (d2 < 0) counts as the maximum of -dist.
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
    if (v != v) return (bv == bv) || i < bi;
    if (bv != bv) return false;
    return v < bv || (v == bv && i < bi);
}

__global__ void __launch_bounds__(256, 2) vq_nearest_kernel(const float* __restrict__ x,
                                                            const float* __restrict__ embed,
                                                            const float* __restrict__ e2,
                                                            int64_t* __restrict__ indices,
                                                            float* __restrict__ quantized,
                                                            int64_t n_tok, int n_codes, int d) {
    __shared__ __align__(16) float Xs[2][VK][VM + VPAD];
    __shared__ __align__(16) float Es[2][VK][VN + VPAD];
    __shared__ float x2s[VM];
    __shared__ int best_idx[VM];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t t0 = (int64_t)blockIdx.x * VM;

    // |x|^2 of the CTA's tokens (VQ:30)
    for (int r = wid; r < VM; r += 8) {
        float s = 0.f;
        if (t0 + r < n_tok)
            for (int i = lane; i < d; i += 32) {
                const float v = x[(t0 + r) * d + i];
                s = fmaf(v, v, s);
            }
        s = warp_sum(s);
        if (lane == 0) x2s[r] = s;
    }
    __syncthreads();

    float bv[8];
    int bi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { bv[i] = INFINITY; bi[i] = 0x7fffffff; }

    float rx[4], re[4];
    const int lr = tid >> 1, lk = (tid & 1) * 4;  // both operands are K-contiguous
    const int nk = (d + VK - 1) / VK;
    for (int n0 = 0; n0 < n_codes; n0 += VN) {
        auto load_tiles = [&](int k0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + lk + j;
                rx[j] = (t0 + lr < n_tok && k < d) ? __ldg(x + (t0 + lr) * d + k) : 0.f;
                re[j] = (n0 + lr < n_codes && k < d) ? __ldg(embed + (int64_t)(n0 + lr) * d + k) : 0.f;
            }
        };
        auto store_tiles = [&](int buf) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                Xs[buf][lk + j][lr] = rx[j];
                Es[buf][lk + j][lr] = re[j];
            }
        };
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        __syncthreads();  // previous code tile's smem reads are done
        load_tiles(0);
        store_tiles(0);
        __syncthreads();
        for (int kb = 0; kb < nk; ++kb) {
            const int cur = kb & 1;
            if (kb + 1 < nk) load_tiles((kb + 1) * VK);
#pragma unroll
            for (int kk = 0; kk < VK; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(&Xs[cur][kk][ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&Xs[cur][kk][64 + ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Es[cur][kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Es[cur][kk][64 + tx * 4]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            if (kb + 1 < nk) {
                store_tiles(cur ^ 1);
                __syncthreads();
            }
        }
        // running arg-min in the sqrt domain, reference operation order (VQ:29-33):
        //   dist = sqrt((x2 + y2) + (-2 * xy))
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (col >= n_codes) continue;
            const float y2 = __ldg(e2 + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4);
                const float d2 = __fadd_rn(__fadd_rn(x2s[r], y2), __fmul_rn(acc[i][j], -2.f));
                const float v = __fsqrt_rn(d2);
                if (better(v, col, bv[i], bi[i])) { bv[i] = v; bi[i] = col; }
            }
        }
    }
    // merge the 16 threads (tx) that share each token row
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv[i], o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi[i], o);
            if (better(ov, oi, bv[i], bi[i])) { bv[i] = ov; bi[i] = oi; }
        }
        if (tx == 0) {
            const int r = i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4);
            best_idx[r] = bi[i];
            if (t0 + r < n_tok) indices[t0 + r] = bi[i];
        }
    }
    if (!quantized) return;
    __syncthreads();
    // VQ:477 quantize = embed[ind]
    for (int r = wid; r < VM; r += 8) {
        if (t0 + r >= n_tok) break;
        const float* src = embed + (int64_t)best_idx[r] * d;
        float* dst = quantized + (t0 + r) * d;
        for (int i = lane; i < d; i += 32) dst[i] = __ldg(src + i);
    }
}

}  // namespace dcta

using namespace dcta;

extern "C" int dcta_vq_nearest(const float* x, const float* embed, float* e2, int64_t* indices,
                               float* quantized, int64_t n_tok, int n_codes, int d, void* stream) {
    DCTA_REQUIRE(x && embed && e2 && indices, "vq_nearest: null pointer");
    DCTA_REQUIRE(n_tok >= 0 && n_codes > 0 && d > 0, "vq_nearest: bad sizes");
    if (n_tok == 0) return DCTA_OK;
    cudaStream_t st = as_stream(stream);
    row_sumsq_kernel<<<(n_codes * 32 + 255) / 256, 256, 0, st>>>(embed, e2, n_codes, d);
    const int64_t grid = ceil_div(n_tok, VM);
    DCTA_REQUIRE(grid < (1ll << 31), "vq_nearest: too many tokens");
    vq_nearest_kernel<<<(unsigned)grid, 256, 0, st>>>(x, embed, e2, indices, quantized, n_tok, n_codes, d);
    return check_launch("vq_nearest");
}

extern "C" int dcta_row_sumsq(const float* x, float* out, int64_t n, int d, void* stream) {
    DCTA_REQUIRE(x && out && n >= 0 && n < (1ll << 26) && d > 0, "row_sumsq: bad args");
    if (n == 0) return DCTA_OK;
    row_sumsq_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(x, out, (int)n, d);
    return check_launch("row_sumsq");
}
