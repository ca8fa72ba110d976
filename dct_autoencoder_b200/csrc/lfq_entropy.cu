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

constexpr int kPairs = 16;         // (token, codebook) pairs per block iteration (one barrier round per 16 rank-one updates)
constexpr int kSpPitch = 17;        // float2 per pair in the probability table: an odd pitch, so that threads on consecutive
                                   // pairs (the backward pass' factor generation) do not all hit the same banks
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

// per-dimension probabilities of kP pairs: sp[pb][i] = (p(bit=1), p(bit=0)); returns the pairs' entropy
// contribution (this thread's share)
template <int kP>
__device__ __forceinline__ float pair_probs(const EntArgs& a, int64_t n0, int tid, float2 (*sp)[kSpPitch], float* sval, float* su) {
    float h = 0.f;
    for (int idx = tid; idx < kP * a.d; idx += 256) {
        const int pb = idx / a.d, i = idx - pb * a.d;
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
                h += log1pf(e) + fabsf(u) * small;                // binary entropy of sigmoid(u), nats
            }
        }
        sp[pb][i] = p;
        if (su) su[pb * 16 + i] = u;
        if (i == 0) sval[pb] = valid ? 1.f : 0.f;
    }
    return h;
}

// [j][pair] tables of the backward pass (32 pairs = 8 chunks of 4 per row): the chunk index is XORed with j & 7, so that
// the 128-bit accesses of 8 lanes on 8 consecutive rows (same pairs) fall on 8 different bank groups
__device__ __forceinline__ int pair_slot(int j, int pb) { return j * 32 + (pb ^ ((j & 7) << 2)); }

// U[pair][j1], V[pair][j2]: products of the per-dimension probabilities of each half (zero for masked pairs).
// One work item = the 2^HB values j = (m << LBITS) | g that share their LOW bits g (HB = min(D, 4) high bits m
// enumerated by doubling: 30 multiplications for 16 values instead of 16 x 7), so that consecutive threads store to
// consecutive addresses.  kTransposed: the tables are written as [j][pair] (the backward pass reads 8 pairs of one j);
// kP: pairs per round.
template <bool kTransposed, int kP>
__device__ __forceinline__ void pair_factors(const EntArgs& a, int tid, const float2 (*sp)[kSpPitch], const float* sval,
                                             float* sU, float* sV) {
    const int HB1 = min(a.D1, 4), HB2 = min(a.D2, 4);
    const int items_u = 1 << (a.D1 - HB1), items_v = 1 << (a.D2 - HB2);
    const int per_pair = items_u + items_v;
    for (int item = tid; item < kP * per_pair; item += 256) {
        // transposed tables: consecutive threads take consecutive pairs (the fastest index of the table)
        const int pb = kTransposed ? item % kP : item / per_pair;
        const int rem = kTransposed ? item / kP : item - pb * per_pair;
        const bool second = rem >= items_u;
        const int g = second ? rem - items_u : rem;
        const int D = second ? a.D2 : a.D1, HB = second ? HB2 : HB1, LBITS = D - HB;
        const float2* p = &sp[pb][second ? a.D1 : 0];       // p[i]: dimension i of this half = bit D-1-i of j
        float v = sval[pb];                                   // 1 for a valid pair, 0 for a masked one
        for (int i = 0; i < LBITS; ++i) v *= ((g >> (LBITS - 1 - i)) & 1) ? p[HB + i].x : p[HB + i].y;
        float t[16];
        t[0] = v;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < HB) {
                const float2 pk = p[k];
#pragma unroll
                for (int m = (1 << k) - 1; m >= 0; --m) {
                    const float tm = t[m];
                    t[2 * m + 1] = tm * pk.x;
                    t[2 * m] = tm * pk.y;
                }
            }
        }
        float* out = second ? sV : sU;
#pragma unroll
        for (int m = 0; m < 16; ++m)
            if (m < (1 << HB)) {
                const int j = (m << LBITS) | g;
                out[kTransposed ? pair_slot(j, pb) : pb * 128 + j] = t[m];
            }
    }
}

// forward: partial[cta][2^d] = sum over this CTA's pairs of U (x) V;  stats[cta] = (entropy sum, valid pairs).
// kRU x kRV: the register tile of a thread (threads as 16 x 16) when the halves have 128 / 64 entries (d = 13, 14: the
// operands then come as 128-bit shared-memory loads); 0: any d, sizes at run time.
template <int kRU, int kRV>
__global__ void __launch_bounds__(256, 2) lfq_entropy_fwd_kernel(EntArgs a, float* __restrict__ partial, float* __restrict__ stats) {
    __shared__ float2 sp[kPairs][kSpPitch];
    __shared__ float sval[kPairs];
    __shared__ __align__(16) float sU[kPairs][128];
    __shared__ __align__(16) float sV[kPairs][128];
    __shared__ float red[8];
    const int tid = threadIdx.x;
    const int NU = 1 << a.D1, NV = 1 << a.D2;
    const int RU = kRU ? kRU : max(1, NU >> 4), RV = kRV ? kRV : max(1, NV >> 4);   // tile of a thread: RU x RV
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
        h_sum += pair_probs<kPairs>(a, blk * kPairs, tid, sp, sval, nullptr);
        __syncthreads();
        if (tid < kPairs) n_valid += sval[tid];
        pair_factors<false, kPairs>(a, tid, sp, sval, &sU[0][0], &sV[0][0]);
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int pb = 0; pb < kPairs; ++pb) {
                // an 8-wide tile is two runs of 4, 64 apart: the 16 threads of a row then read 256 contiguous bytes per
                // 128-bit load (a contiguous run of 8 per thread costs every load a two-way bank conflict)
                float uu[8], vv[8];
                if (kRU == 8) {
                    *reinterpret_cast<float4*>(uu) = *reinterpret_cast<const float4*>(&sU[pb][tu * 4]);
                    *reinterpret_cast<float4*>(uu + 4) = *reinterpret_cast<const float4*>(&sU[pb][64 + tu * 4]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) uu[i] = i < RU ? sU[pb][tu * RU + i] : 0.f;
                }
                if (kRV == 8) {
                    *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(&sV[pb][tv * 4]);
                    *reinterpret_cast<float4*>(vv + 4) = *reinterpret_cast<const float4*>(&sV[pb][64 + tv * 4]);
                } else if (kRV == 4) {
                    *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(&sV[pb][tv * 4]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) vv[j] = j < RV ? sV[pb][tv * RV + j] : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if ((kRU == 0 || i < kRU) && (kRV == 0 || j < kRV)) acc[i][j] = fmaf(uu[i], vv[j], acc[i][j]);
            }
        }
    }
    float* out = partial + (int64_t)blockIdx.x * ((int64_t)1 << a.d);
    if (active) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (i < RU && j < RV) {
                    const int r = kRU == 8 ? (i >> 2) * 64 + tu * 4 + (i & 3) : tu * RU + i;
                    const int q = kRV == 8 ? (j >> 2) * 64 + tv * 4 + (j & 3) : tv * RV + j;
                    out[(r << a.D2) | q] = acc[i][j];
                }
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
// Per round of kPairsB = 32 pairs two small GEMMs against the resident G (shared memory, rows padded to NV + 1):
//   threads 0..127:   W1[r][pair] = sum_q G[r][q] V[q][pair]
//   threads 128..255: W2[q][pair] = sum_r G[r][q] U[r][pair]
// A thread owns 4 rows (columns) j = lane + 32 k and 8 pairs; the 32 lanes of a warp share the pairs, so the operand
// comes as two broadcast 128-bit loads and G as four conflict-free scalar loads per 32 FMAs.  (A first version with
// 2 x 8 tiles over 16 pairs spent its time in shared-memory wavefronts: ncu 78 % of the LSU peak, 43 ms.)  Then
// T1 = U * W1 and T2 = V * W2 overwrite U and V, and one thread per (pair, dimension) sums them with and without its bit.
constexpr int kPairsB = 32;

__global__ void __launch_bounds__(256) lfq_entropy_bwd_kernel(EntArgs a, const float* __restrict__ tables,
                                                              const float* __restrict__ result, const float* __restrict__ grad_out,
                                                              float* __restrict__ grad_x) {
    extern __shared__ __align__(16) float smem_g[];            // G as NU rows of (NV + 1) floats
    __shared__ float2 sp[kPairsB][kSpPitch];
    __shared__ float su[kPairsB * 16];
    __shared__ float sval[kPairsB];
    static_assert(kPairsB == 32, "pair_slot assumes rows of 32 pairs");
    __shared__ __align__(16) float sUt[128 * kPairsB];         // U[r][pair] at pair_slot(r, pair), then T1 = U * (G V)
    __shared__ __align__(16) float sVt[128 * kPairsB];         // V[q][pair], then T2 = V * (G^T U)
    __shared__ float sPart[2][4][6][kPairsB];                  // per (half, quarter of the table): total + 5 low-bit sums
    const int tid = threadIdx.x;
    const int NU = 1 << a.D1, NV = 1 << a.D2, ldg = NV + 1;
    const int n_codes = 1 << a.d;
    for (int j = tid; j < n_codes; j += 256) smem_g[(j >> a.D2) * ldg + (j & (NV - 1))] = tables[n_codes + j];
    const float n_tok = result[1];
    const float g = grad_out[0];
    const float inv_tok = n_tok > 0.f ? 1.f / n_tok : 0.f;
    const float inv_pairs = n_tok > 0.f ? 1.f / (n_tok * (float)a.c) : 0.f;
    const int64_t n_blocks = (a.n_pairs + kPairsB - 1) / kPairsB;
    const bool first = tid < 128;
    const bool fast_sums = a.D1 >= 6 && a.D2 >= 6;             // quarters of 16 or 32 entries: d = 12, 13, 14
    const int lane = tid & 31, ph = ((tid >> 5) & 3) * 8;       // rows (columns) lane + 32 k; pairs ph .. ph + 7
    const int n_own = first ? NU : NV, n_red = first ? NV : NU;
    const float* opnd = first ? sVt : sUt;                      // contracted operand
    float* mine = first ? sUt : sVt;                            // multiplies the result, then holds it
    // G[j][x] (own rows: stride ldg between j, 1 between x) or G[x][j] (own columns)
    const int jstep = first ? ldg : 1, xstep = first ? 1 : ldg;
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        __syncthreads();
        (void)pair_probs<kPairsB>(a, blk * kPairsB, tid, sp, sval, su);
        __syncthreads();
        pair_factors<true, kPairsB>(a, tid, sp, sval, sUt, sVt);
        __syncthreads();
        float w[4][8];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int e = 0; e < 8; ++e) w[k][e] = 0.f;
        {
            // rows beyond the table (small d) read row `lane` again; their results are dropped
            const float* g0 = smem_g + (lane < n_own ? lane : 0) * jstep;
            const int o1 = (lane + 32 < n_own ? 32 : 0) * jstep, o2 = (lane + 64 < n_own ? 64 : 0) * jstep,
                      o3 = (lane + 96 < n_own ? 96 : 0) * jstep;
            if ((n_red & 7) == 0) {
                // eight steps at a time: (x & 7) is then a compile-time constant and the swizzled operand addresses are
                // one XOR away from a per-block base (ncu: the per-step shifts / multiplies of pair_slot were a fifth of the
                // loop's instructions)
                for (int xb = 0; xb < n_red; xb += 8) {
                    const float* gx0 = g0 + xb * xstep;
                    const float* ob = opnd + xb * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float* gx = gx0 + j * xstep;
                        const float gv[4] = {gx[0], gx[o1], gx[o2], gx[o3]};
                        const int c0 = ph ^ (j << 2);
                        const float4 v0 = *reinterpret_cast<const float4*>(ob + j * 32 + c0);
                        const float4 v1 = *reinterpret_cast<const float4*>(ob + j * 32 + (c0 ^ 4));
                        const float o[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k)
#pragma unroll
                            for (int e = 0; e < 8; ++e) w[k][e] = fmaf(gv[k], o[e], w[k][e]);
                    }
                }
            } else {
                for (int x = 0; x < n_red; ++x) {
                    const float* gx = g0 + x * xstep;
                    const float gv[4] = {gx[0], gx[o1], gx[o2], gx[o3]};
                    const float4 v0 = *reinterpret_cast<const float4*>(opnd + pair_slot(x, ph));
                    const float4 v1 = *reinterpret_cast<const float4*>(opnd + pair_slot(x, ph + 4));
                    const float o[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int e = 0; e < 8; ++e) w[k][e] = fmaf(gv[k], o[e], w[k][e]);
                }
            }
        }
        __syncthreads();                                         // every operand has been read: U, V may be overwritten
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = lane + 32 * k;
            if (j < n_own) {
                float4* q0 = reinterpret_cast<float4*>(mine + pair_slot(j, ph));
                float4* q1 = reinterpret_cast<float4*>(mine + pair_slot(j, ph + 4));
                float4 m0 = *q0, m1 = *q1;
                m0.x *= w[k][0]; m0.y *= w[k][1]; m0.z *= w[k][2]; m0.w *= w[k][3];
                m1.x *= w[k][4]; m1.y *= w[k][5]; m1.z *= w[k][6]; m1.w *= w[k][7];
                *q0 = m0;
                *q1 = m1;
            }
        }
        __syncthreads();
        if (fast_sums) {
            // warp = (half, quarter of the table), lane = pair: the quarter's total and its sums over the entries with low
            // bit b set, b < 5, with the bit pattern of the unrolled loop known at compile time (no divergence)
            const int wq = tid >> 5, half = wq >> 2, seg = wq & 3;
            const int L = (half ? NV : NU) >> 2;
            const float* tab = half ? sVt : sUt;
            float tot = 0.f, lb[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                if (e < L) {
                    const float t = tab[pair_slot(seg * L + e, lane)];
                    tot += t;
#pragma unroll
                    for (int b = 0; b < 5; ++b) if (e & (1 << b)) lb[b] += t;
                }
            }
            sPart[half][seg][0][lane] = tot;
#pragma unroll
            for (int b = 0; b < 5; ++b) sPart[half][seg][1 + b][lane] = lb[b];
            __syncthreads();
        }
        for (int idx = tid; idx < kPairsB * a.d; idx += 256) {
            const int pb = idx / a.d, i = idx - pb * a.d;
            const int64_t n = blk * kPairsB + pb;
            if (n < a.n_pairs) {
                float out = 0.f;
                if (sval[pb] != 0.f) {
                    float eg = 0.f, egb = 0.f;                   // E[G], E[G b_i]
                    if (fast_sums) {
                        const int half = i >= a.D1, Dh = half ? a.D2 : a.D1;
                        const int sh = Dh - 1 - (half ? i - a.D1 : i), LB = Dh - 2;
#pragma unroll
                        for (int seg = 0; seg < 4; ++seg) {
                            const float tot = sPart[half][seg][0][pb];
                            eg += tot;
                            if (sh < LB) egb += sPart[half][seg][1 + sh][pb];
                            else if ((seg >> (sh - LB)) & 1) egb += tot;
                        }
                    } else if (i < a.D1) {
                        const int sh = a.D1 - 1 - i;
                        for (int r = 0; r < NU; ++r) { const float t = sUt[pair_slot(r, pb)]; eg += t; if ((r >> sh) & 1) egb += t; }
                    } else {
                        const int sh = a.D2 - 1 - (i - a.D1);
                        for (int q = 0; q < NV; ++q) { const float t = sVt[pair_slot(q, pb)]; eg += t; if ((q >> sh) & 1) egb += t; }
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
template <int kVec>
__global__ void __launch_bounds__(256) lfq_commit_bwd_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask,
                                                             const float* __restrict__ grad_out, const float* __restrict__ n_valid_dev,
                                                             float* __restrict__ grad_x, int64_t n_tok, int cd, float scale) {
    const float nv = n_valid_dev[0];
    const float k = nv > 0.f ? 2.f * grad_out[0] / (nv * (float)cd) : 0.f;
    const int64_t total = n_tok * cd / kVec;                      // chunks of kVec elements inside one token (cd % kVec == 0)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const bool on = mask[i * kVec / cd] != 0;
        if (kVec == 4) {
            const float4 v = ld_stream(reinterpret_cast<const float4*>(x) + i);
            float4 o;
            o.x = on ? k * (v.x - (v.x > 0.f ? scale : -scale)) : 0.f;
            o.y = on ? k * (v.y - (v.y > 0.f ? scale : -scale)) : 0.f;
            o.z = on ? k * (v.z - (v.z > 0.f ? scale : -scale)) : 0.f;
            o.w = on ? k * (v.w - (v.w > 0.f ? scale : -scale)) : 0.f;
            st_stream(reinterpret_cast<float4*>(grad_x) + i, o);
        } else {
            const float v = x[i];
            grad_x[i] = on ? k * (v - (v > 0.f ? scale : -scale)) : 0.f;
        }
    }
}

// number of valid tokens: one CTA (deterministic), 16 mask bytes per load, four loads in flight per thread
__global__ void __launch_bounds__(256) count_valid_kernel(const uint8_t* __restrict__ mask, int64_t n, float* __restrict__ out) {
    __shared__ float red[8];
    int cnt = 0;
    const bool vec = (reinterpret_cast<uintptr_t>(mask) & 15) == 0;
    const int64_t n16 = vec ? n / 16 : 0;
    const uint4* m16 = reinterpret_cast<const uint4*>(mask);
    for (int64_t i0 = threadIdx.x; i0 < n16; i0 += 4 * blockDim.x) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * blockDim.x;
            v[u] = i < n16 ? __ldg(m16 + i) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // bytes are 0 / non-zero: fold every byte to its low bit, then count
                uint32_t t = w[k];
                t |= t >> 4; t |= t >> 2; t |= t >> 1;
                cnt += __popc(t & 0x01010101u);
            }
        }
    }
    for (int64_t i = n16 * 16 + threadIdx.x; i < n; i += blockDim.x) cnt += mask[i] ? 1 : 0;
    const float s = block_sum_256((float)cnt, red);
    if (threadIdx.x == 0) out[0] = s;
}

static bool ent_args(EntArgs& a, const float* x, const uint8_t* mask, int64_t n_tok, int c, int d, float scale, float temperature) {
    a.x = x; a.mask = mask; a.n_pairs = n_tok * c; a.c = c; a.d = d;
    a.D1 = (d + 1) / 2; a.D2 = d - a.D1;
    a.u_scale = -4.0f * scale / temperature;
    return d >= 1 && d <= kMaxD;
}

int launch_lfq_entropy_final(const float* partial, const float* stats, int n_cta, int c, int d, float eps, float* tables,
                             float* result, cudaStream_t st) {
    lfq_entropy_final_kernel<<<1, 1024, 0, st>>>(partial, stats, n_cta, c, d, eps, tables, result);
    return check_launch("lfq_entropy_factorized");
}
// lfq_entropy_tc.cu: the forward contraction on tcgen05 (d = 13, 14)
int launch_lfq_entropy_fwd_tc(const float* x, const uint8_t* mask, int64_t n_tok, int c, int d, float u_scale, float eps,
                              float* partial_scratch, float* stats, float* tables, float* result, cudaStream_t st);

}  // namespace dcta

static bool g_entropy_tc = true;
// 1 (default): d = 14 runs the forward contraction on the tensor cores; 0: the fp32 FMA kernel (the checker of the tests)
extern "C" int dcta_lfq_entropy_use_tensor_cores(int on) {
    g_entropy_tc = on != 0;
    return DCTA_OK;
}

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
    if ((d == 14 || d == 13) && n_tok > 0 && g_entropy_tc)
        return launch_lfq_entropy_fwd_tc(x, mask, n_tok, c, d, a.u_scale, eps, partial_scratch, stats, tables, result, st);
    if (a.D1 == 7 && a.D2 == 7) lfq_entropy_fwd_kernel<8, 8><<<grid, 256, 0, st>>>(a, partial_scratch, stats);
    else if (a.D1 == 7 && a.D2 == 6) lfq_entropy_fwd_kernel<8, 4><<<grid, 256, 0, st>>>(a, partial_scratch, stats);
    else lfq_entropy_fwd_kernel<0, 0><<<grid, 256, 0, st>>>(a, partial_scratch, stats);
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
    const int64_t n_blocks = ceil_div(a.n_pairs, kPairsB);
    const int grid = (int)(n_blocks < 2 * kNumSMs ? n_blocks : 2 * kNumSMs);
    const size_t smem = sizeof(float) * ((size_t)1 << a.D1) * (((size_t)1 << a.D2) + 1);
    // static (45 KB) + dynamic shared memory exceed the 48 KB default for every d >= 9: always opt in
    if (cudaFuncSetAttribute(lfq_entropy_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return check_launch("lfq_entropy_factorized_backward (shared memory opt-in)");
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
    if (cd % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(grad_x)) & 15) == 0)
        lfq_commit_bwd_kernel<4><<<grid_for(n_tok * cd / 4, 256), 256, 0, st>>>(x, mask, grad_out, n_valid_scratch, grad_x, n_tok, cd, scale);
    else
        lfq_commit_bwd_kernel<1><<<grid_for(n_tok * cd, 256), 256, 0, st>>>(x, mask, grad_out, n_valid_scratch, grad_x, n_tok, cd, scale);
    return check_launch("lfq_commit_backward");
}
