// Factorised entropy loss of the lookup-free quantiser.
//
// Reference: util.py:355-387 compute_entropy_loss applied to LFQ's `distance` (lfq.py:191):
//   affinity[t, c, j] = -2 * sum_i x[t, c, i] * codebook[j, i],  codebook[j, i] = +-s (bit d-1-i of j set: +s)
//   probs = softmax(affinity / T + eps);   loss = E_valid[sum_c H(probs)] - H(mean_{valid, c} probs)
// The reference materialises affinity as (b, n, c, 2^d): 721 GB at the benchmark shape (786 k tokens x 14 x 2^14).
// With a +-s codebook the softmax over the 2^d sign patterns FACTORISES over the d dimensions:
//   probs[j] = prod_i p_i(bit_i(j)),   p_i(1) = sigmoid(u_i),  u_i = -4 s x_i / T
// so  (a) the per-sample entropy is a sum of d binary entropies, and
//     (b) the average distribution is a mean of rank-one tensors: with j = (j1, j2) split into the D1 high and D2 low
//         bits, avg[j1, j2] = (1/N) sum_n U_n[j1] * V_n[j2], U_n / V_n the 2^D1 / 2^D2 products of each half -- an
//         A^T B contraction over the (token, codebook) pairs n, accumulated here in registers (one 8 x 8 block of the
//         128 x 128 table per thread), never touching a (T, c, 2^d) tensor.
// The backward pass uses the same factors: d avg_entropy / d u_{n,i} = (1/N) (E_n[G b_i] - p_i E_n[G]) with
// G = d avg_entropy / d avg and E_n[f] = sum_j probs_n[j] f[j], i.e. two matrix-vector products G V_n and G^T U_n per pair.
// Everything is deterministic: per-CTA partial tables in fixed slots, summed in a fixed order.
#include "common.cuh"

namespace dcta {

constexpr int kPairs = 4;          // (token, codebook) pairs per block iteration
constexpr int kMaxD = 14;          // 2^7 x 2^7 table: 64 accumulators per thread

struct EntArgs {
    const float* x;            // (n_tok, c * d)
    const uint8_t* mask;       // (n_tok) 1 = valid
    int64_t n_pairs;           // n_tok * c
    int c, d, D1, D2;
    float u_scale;             // -4 s / T
};

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
    __syncthreads();
    return red[0];
}

// per-dimension probabilities of kPairs pairs: sp[pb][i] = (p(bit=1), p(bit=0)); returns the pair's entropy
// contribution (this thread's share) -- called by threads 0 .. kPairs*d-1
__device__ __forceinline__ float pair_probs(const EntArgs& a, int64_t n0, int tid, float2 (*sp)[16], float* sval, float* su) {
    float h = 0.f;
    if (tid < kPairs * a.d) {
        const int pb = tid / a.d, i = tid - pb * a.d;
        const int64_t n = n0 + pb;
        float2 p = make_float2(0.f, 0.f);
        float u = 0.f;
        bool valid = false;
        if (n < a.n_pairs) {
            const int64_t t = n / a.c;
            valid = a.mask[t] != 0;
            if (valid) {
                u = a.u_scale * a.x[n * a.d + i];
                const float e = __expf(-fabsf(u));               // in (0, 1]
                const float big = 1.f / (1.f + e), small = e * big;   // sigmoid(|u|), sigmoid(-|u|)
                p = u >= 0.f ? make_float2(big, small) : make_float2(small, big);
                h = log1pf(e) + fabsf(u) * small;                 // binary entropy of sigmoid(u), nats
            }
        }
        sp[pb][i] = p;
        if (su) su[pb * 16 + i] = u;
        if (i == 0) sval[pb] = valid ? 1.f : 0.f;
    }
    return h;
}

// U[pb][j1], V[pb][j2]: products of the per-dimension probabilities of each half (zero for masked pairs)
__device__ __forceinline__ void pair_factors(const EntArgs& a, int tid, const float2 (*sp)[16], const float* sval,
                                             float (*sU)[128], float (*sV)[128]) {
    const int NU = 1 << a.D1, NV = 1 << a.D2;
    for (int idx = tid; idx < kPairs * (NU + NV); idx += 256) {
        const int pb = idx / (NU + NV), r = idx - pb * (NU + NV);
        float v = sval[pb];
        if (r < NU) {
            for (int i = 0; i < a.D1; ++i) v *= ((r >> (a.D1 - 1 - i)) & 1) ? sp[pb][i].x : sp[pb][i].y;
            sU[pb][r] = v;
        } else {
            const int q = r - NU;
            v = v != 0.f ? 1.f : 0.f;
            for (int i = 0; i < a.D2; ++i) v *= ((q >> (a.D2 - 1 - i)) & 1) ? sp[pb][a.D1 + i].x : sp[pb][a.D1 + i].y;
            sV[pb][q] = v;
        }
    }
}

// forward: partial[cta][2^d] = sum over this CTA's pairs of U (x) V;  stats[cta] = (entropy sum, valid pairs)
__global__ void __launch_bounds__(256) lfq_entropy_fwd_kernel(EntArgs a, float* __restrict__ partial, float* __restrict__ stats) {
    __shared__ float2 sp[kPairs][16];
    __shared__ float sval[kPairs];
    __shared__ __align__(16) float sU[kPairs][128];
    __shared__ __align__(16) float sV[kPairs][128];
    __shared__ float red[8];
    const int tid = threadIdx.x;
    const int NU = 1 << a.D1, NV = 1 << a.D2;
    const int RU = max(1, NU >> 4), RV = max(1, NV >> 4);         // tile of a thread: RU x RV, threads as 16 x 16
    const int tu = tid >> 4, tv = tid & 15;
    const bool active = tu * RU < NU && tv * RV < NV;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float h_sum = 0.f, n_valid = 0.f;
    const int64_t n_blocks = (a.n_pairs + kPairs - 1) / kPairs;
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        __syncthreads();                                  // the previous iteration's readers are done
        h_sum += pair_probs(a, blk * kPairs, tid, sp, sval, nullptr);
        __syncthreads();
        if (tid < kPairs) n_valid += sval[tid];
        pair_factors(a, tid, sp, sval, sU, sV);
        __syncthreads();
        if (active) {
#pragma unroll
            for (int pb = 0; pb < kPairs; ++pb) {
                float uu[8], vv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) uu[i] = i < RU ? sU[pb][tu * RU + i] : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) vv[j] = j < RV ? sV[pb][tv * RV + j] : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(uu[i], vv[j], acc[i][j]);
            }
        }
    }
    float* out = partial + (int64_t)blockIdx.x * ((int64_t)1 << a.d);
    if (active) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (i < RU && j < RV) out[((tu * RU + i) << a.D2) | (tv * RV + j)] = acc[i][j];
    }
    const float hs = block_sum_256(h_sum, red);
    const float nv = block_sum_256(n_valid, red);
    if (tid == 0) { stats[2 * blockIdx.x] = hs; stats[2 * blockIdx.x + 1] = nv; }
}

// avg[j] = sum_cta partial / N, G[j] = d avg_entropy / d avg[j], loss = sample_entropy - avg_entropy.
// out: tables[0 .. 2^d) = avg, tables[2^d .. 2*2^d) = G;  result[0] = loss, [1] = n_valid tokens, [2] = sample entropy,
// [3] = avg entropy
__global__ void __launch_bounds__(1024) lfq_entropy_final_kernel(const float* __restrict__ partial, const float* __restrict__ stats,
                                                                 int n_cta, int c, int d, float eps, float* __restrict__ tables,
                                                                 float* __restrict__ result) {
    __shared__ float red[32];
    __shared__ float s_h, s_n;
    const int n_codes = 1 << d;
    if (threadIdx.x == 0) {
        float h = 0.f, n = 0.f;
        for (int i = 0; i < n_cta; ++i) { h += stats[2 * i]; n += stats[2 * i + 1]; }
        s_h = h; s_n = n;
    }
    __syncthreads();
    const float n_pairs = s_n;                        // valid (token, codebook) pairs
    const float inv = n_pairs > 0.f ? 1.f / n_pairs : 0.f;
    float ent = 0.f;
    for (int j = threadIdx.x; j < n_codes; j += blockDim.x) {
        float s = 0.f;
        for (int i = 0; i < n_cta; ++i) s += partial[(int64_t)i * n_codes + j];
        const float avg = s * inv;
        tables[j] = avg;
        const float lg = logf(avg + eps);
        tables[n_codes + j] = -(lg + avg / (avg + eps));
        ent -= avg * lg;
    }
    ent = warp_sum(ent);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ent;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = red[threadIdx.x];
        t = warp_sum(t);
        if (threadIdx.x == 0) {
            const float n_tok = n_pairs / (float)c;
            const float sample = n_tok > 0.f ? s_h / n_tok : 0.f;     // masked_mean over tokens, SUM over codebooks (UT:382)
            result[0] = sample - t;
            result[1] = n_tok;
            result[2] = sample;
            result[3] = t;
        }
    }
}

// backward: grad_x[n, i] = g * du/dx * ( dS/du - dA/du ),  S = sample entropy, A = avg entropy
//   dS/du_i = -(u_i p_i (1 - p_i)) / n_tok_valid;   dA/du_i = (E_n[G b_i] - p_i E_n[G]) / N_pairs_valid
__global__ void __launch_bounds__(256) lfq_entropy_bwd_kernel(EntArgs a, const float* __restrict__ tables,
                                                              const float* __restrict__ result, const float* __restrict__ grad_out,
                                                              float* __restrict__ grad_x) {
    extern __shared__ __align__(16) float smem_g[];            // G as NU rows of (NV + 1) floats
    __shared__ float2 sp[kPairs][16];
    __shared__ float su[kPairs * 16];
    __shared__ float sval[kPairs];
    __shared__ __align__(16) float sU[kPairs][128];
    __shared__ __align__(16) float sV[kPairs][128];
    __shared__ __align__(16) float sT1[128][kPairs];           // U * (G V) per row j1
    __shared__ __align__(16) float sT2[128][kPairs];           // V * (G^T U) per column j2
    const int tid = threadIdx.x;
    const int NU = 1 << a.D1, NV = 1 << a.D2, ldg = NV + 1;
    const int n_codes = 1 << a.d;
    for (int j = tid; j < n_codes; j += 256) smem_g[(j >> a.D2) * ldg + (j & (NV - 1))] = tables[n_codes + j];
    const float n_tok = result[1];
    const float g = grad_out[0];
    const float inv_tok = n_tok > 0.f ? 1.f / n_tok : 0.f;
    const float inv_pairs = n_tok > 0.f ? 1.f / (n_tok * (float)a.c) : 0.f;
    const int64_t n_blocks = (a.n_pairs + kPairs - 1) / kPairs;
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        __syncthreads();
        (void)pair_probs(a, blk * kPairs, tid, sp, sval, su);
        __syncthreads();
        pair_factors(a, tid, sp, sval, sU, sV);
        __syncthreads();
        if (tid < 128) {
            // row r of G against V of the four pairs
            const int r = tid;
            if (r < NU) {
                float w[kPairs] = {0.f, 0.f, 0.f, 0.f};
                const float* grow = smem_g + r * ldg;
                for (int q = 0; q < NV; ++q) {
                    const float gv = grow[q];
#pragma unroll
                    for (int pb = 0; pb < kPairs; ++pb) w[pb] = fmaf(gv, sV[pb][q], w[pb]);
                }
#pragma unroll
                for (int pb = 0; pb < kPairs; ++pb) sT1[r][pb] = w[pb] * sU[pb][r];
            }
        } else {
            const int q = tid - 128;
            if (q < NV) {
                float w[kPairs] = {0.f, 0.f, 0.f, 0.f};
                for (int r = 0; r < NU; ++r) {
                    const float gv = smem_g[r * ldg + q];
#pragma unroll
                    for (int pb = 0; pb < kPairs; ++pb) w[pb] = fmaf(gv, sU[pb][r], w[pb]);
                }
#pragma unroll
                for (int pb = 0; pb < kPairs; ++pb) sT2[q][pb] = w[pb] * sV[pb][q];
            }
        }
        __syncthreads();
        if (tid < kPairs * a.d) {
            const int pb = tid / a.d, i = tid - pb * a.d;
            const int64_t n = blk * kPairs + pb;
            if (n < a.n_pairs) {
                float out = 0.f;
                if (sval[pb] != 0.f) {
                    float eg = 0.f, egb = 0.f;                   // E[G], E[G b_i]
                    if (i < a.D1) {
                        const int sh = a.D1 - 1 - i;
                        for (int r = 0; r < NU; ++r) { const float t = sT1[r][pb]; eg += t; if ((r >> sh) & 1) egb += t; }
                    } else {
                        const int sh = a.D2 - 1 - (i - a.D1);
                        for (int q = 0; q < NV; ++q) { const float t = sT2[q][pb]; eg += t; if ((q >> sh) & 1) egb += t; }
                    }
                    const float p1 = sp[pb][i].x, p0 = sp[pb][i].y, u = su[pb * 16 + i];
                    const float ds = -(u * p1 * p0) * inv_tok;
                    const float da = (egb - p1 * eg) * inv_pairs;
                    out = g * a.u_scale * (ds - da);
                }
                grad_x[n * a.d + i] = out;
            }
        }
    }
}

// lfq.py:195-200 backward of the masked commitment loss: grad_x = g * 2 (x - q) * mask / (n_valid * c * d)
__global__ void __launch_bounds__(256) lfq_commit_bwd_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask,
                                                             const float* __restrict__ grad_out, const float* __restrict__ n_valid_dev,
                                                             float* __restrict__ grad_x, int64_t n_tok, int cd, float scale) {
    const float nv = n_valid_dev[0];
    const float k = nv > 0.f ? 2.f * grad_out[0] / (nv * (float)cd) : 0.f;
    const int64_t total = n_tok * cd;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        grad_x[i] = mask[i / cd] ? k * (v - (v > 0.f ? scale : -scale)) : 0.f;
    }
}

__global__ void count_valid_kernel(const uint8_t* __restrict__ mask, int64_t n, float* __restrict__ out) {
    __shared__ float red[8];
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += mask[i] ? 1.f : 0.f;
    s = block_sum_256(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

static bool ent_args(EntArgs& a, const float* x, const uint8_t* mask, int64_t n_tok, int c, int d, float scale, float temperature) {
    a.x = x; a.mask = mask; a.n_pairs = n_tok * c; a.c = c; a.d = d;
    a.D1 = (d + 1) / 2; a.D2 = d - a.D1;
    a.u_scale = -4.0f * scale / temperature;
    return d >= 1 && d <= kMaxD;
}

}  // namespace dcta

extern "C" int dcta_lfq_entropy_ctas(void) { return 2 * dcta::kNumSMs; }

extern "C" int dcta_lfq_entropy_factorized(const float* x, const uint8_t* mask, int64_t n_tok, int c, int d, float codebook_scale,
                                           float temperature, float eps, float* partial_scratch, float* tables, float* result,
                                           void* stream) {
    using namespace dcta;
    DCTA_REQUIRE(x && mask && partial_scratch && tables && result && n_tok >= 0 && c > 0, "lfq_entropy_factorized: bad arguments");
    EntArgs a;
    DCTA_REQUIRE(ent_args(a, x, mask, n_tok, c, d, codebook_scale, temperature),
                 "lfq_entropy_factorized: codebook_dim %d outside 1..%d", d, kMaxD);
    const int64_t n_blocks = ceil_div(a.n_pairs, kPairs);
    const int grid = (int)(n_blocks < 1 ? 1 : (n_blocks < 2 * kNumSMs ? n_blocks : 2 * kNumSMs));
    float* stats = partial_scratch + (int64_t)dcta_lfq_entropy_ctas() * ((int64_t)1 << d);
    cudaStream_t st = as_stream(stream);
    lfq_entropy_fwd_kernel<<<grid, 256, 0, st>>>(a, partial_scratch, stats);
    lfq_entropy_final_kernel<<<1, 1024, 0, st>>>(partial_scratch, stats, grid, c, d, eps, tables, result);
    return check_launch("lfq_entropy_factorized");
}

extern "C" int dcta_lfq_entropy_factorized_backward(const float* x, const uint8_t* mask, int64_t n_tok, int c, int d,
                                                    float codebook_scale, float temperature, const float* tables,
                                                    const float* result, const float* grad_out, float* grad_x, void* stream) {
    using namespace dcta;
    DCTA_REQUIRE(x && mask && tables && result && grad_out && grad_x && n_tok >= 0 && c > 0,
                 "lfq_entropy_factorized_backward: bad arguments");
    EntArgs a;
    DCTA_REQUIRE(ent_args(a, x, mask, n_tok, c, d, codebook_scale, temperature),
                 "lfq_entropy_factorized_backward: codebook_dim %d outside 1..%d", d, kMaxD);
    if (n_tok == 0) return DCTA_OK;
    const int64_t n_blocks = ceil_div(a.n_pairs, kPairs);
    const int grid = (int)(n_blocks < 2 * kNumSMs ? n_blocks : 2 * kNumSMs);
    const size_t smem = sizeof(float) * ((size_t)1 << a.D1) * (((size_t)1 << a.D2) + 1);
    if (smem > 40 * 1024)
        cudaFuncSetAttribute(lfq_entropy_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lfq_entropy_bwd_kernel<<<grid, 256, smem, as_stream(stream)>>>(a, tables, result, grad_out, grad_x);
    return check_launch("lfq_entropy_factorized_backward");
}

extern "C" int dcta_lfq_commit_backward(const float* x, const uint8_t* mask, const float* grad_out, float* n_valid_scratch,
                                        float* grad_x, int64_t n_tok, int cd, float scale, void* stream) {
    using namespace dcta;
    DCTA_REQUIRE(x && mask && grad_out && n_valid_scratch && grad_x && n_tok >= 0 && cd > 0, "lfq_commit_backward: bad arguments");
    if (n_tok == 0) return DCTA_OK;
    cudaStream_t st = as_stream(stream);
    count_valid_kernel<<<1, 256, 0, st>>>(mask, n_tok, n_valid_scratch);
    lfq_commit_bwd_kernel<<<grid_for(n_tok * cd, 256), 256, 0, st>>>(x, mask, grad_out, n_valid_scratch, grad_x, n_tok, cd, scale);
    return check_launch("lfq_commit_backward");
}
