// Factorised LFQ entropy loss, forward, on the tensor cores (d = 14: two halves of 7 bits; d = 13: 7 + 6, N = 64).
//
// The average distribution of util.py:355-387 over LFQ's factorised softmax (see lfq_entropy.cu) is
//     avg[j1, j2] = (1/N) sum_n U_n[j1] * V_n[j2],
// an A . B^T product with M = N = 128 and the (token, codebook) pair index n as the contraction: 0.36 TFLOP at
// config 2's 11 M pairs, which the fp32 FMA pipe cannot do in less than ~6.6 ms.  Here:
//   * twelve PRODUCER warps (two per ring stage: the U and the V tile), a lane = one pair: the lane forms its 7 sigmoids, enumerates its 128 half-products by
//     doubling (30 multiplications per 16 values) and writes them, scaled by 2^14 and split into fp16 hi / lo, into K-major
//     SWIZZLE_64B operand tiles [128 rows j][32 pairs] -- the 32 lanes of a warp store the 32 K-elements of one row:
//     64 contiguous bytes, one shared-memory wavefront per store;
//   * ONE thread issues tcgen05.mma (M = 128, N = 128, K = 16): lo.hi + hi.lo + hi.hi per K slice, fp32 accumulation in TMEM;
//   * the tensor core's fp32 accumulation truncates, so an accumulator is FLUSHED every 32 stages (1024 pairs: a drift of at
//     most ~10^-5 relative): four epilogue warps read it out of TMEM, un-scale by 2^-28 and add it in fp32 (round to nearest)
//     to this CTA's partial table in global memory while the MMAs continue on the second accumulator.
// The per-CTA partial tables and (entropy sum, valid pairs) go to the same slots as lfq_entropy_fwd_kernel's, so
// lfq_entropy_final_kernel and the backward pass are shared.  Deterministic: fixed stage -> CTA mapping, fixed order.
#include "tc_ptx.cuh"

namespace dcta {

namespace etc {
constexpr int kD = 14, kH = 7, kRows = 128;          // bits, bits per half, entries per half
constexpr int kStagePairs = 32;                      // K of one stage
constexpr int kTile = kRows * kStagePairs * 2;       // one fp16 operand tile: 8 KB
constexpr int kStage = 4 * kTile;                    // U hi, U lo, V hi, V lo
constexpr int kStages = 6;
// one producer warp per ring slot: a warp then waits for consecutive phases of ITS slot's barrier (with more producers than
// slots a warp skips a phase, and a parity wait that is two phases ahead passes immediately)
// Two warps per slot: one forms the U tile (dimensions 0..6), the other the V tile (dimensions 7..13) of the same 32 pairs --
// the producers are latency-bound (a lane's chain of loads, 7 sigmoids, 128 products, 256 stores), so more warps on
// shorter chains is what shortens a stage.
constexpr int kProducers = 2 * kStages, kThreads = 32 * (kProducers + 1 + 4);   // + MMA warp + 4 epilogue warps
constexpr int kFlushStages = 32;                     // stages accumulated in TMEM before a flush
constexpr float kScale = 16384.f, kUnscale = 1.f / (16384.f * 16384.f);
constexpr int kSmem = kStages * kStage + 1024;
}  // namespace etc

struct EntTcArgs {
    const float* x;            // (n_tok, c * 14)
    const uint8_t* mask;       // (n_tok)
    int64_t n_pairs;           // n_tok * c
    int c;
    float u_scale;             // -4 s / T
};

// the 128 products of 7 per-dimension probabilities (dimension 0 = the highest bit), scaled, as fp16 hi / lo into column
// `lane` of a [128][32] SWIZZLE_64B tile; base[x] = tile + ((chunk ^ x) << 4) + (lane & 7) * 2
template <int kDims>     // 7 or 6 dimensions in this half: 2^kDims rows
__device__ __forceinline__ void emit_half(const float (&p1)[7], const float (&p0)[7], float v0, const uint32_t (&base_hi)[4]) {
    constexpr int LB = kDims - 4;                    // low bits of the row index: dimensions 4 .. kDims - 1
#pragma unroll
    for (int g = 0; g < (1 << LB); ++g) {
        float v = v0;
#pragma unroll
        for (int i = 0; i < LB; ++i) v *= ((g >> (LB - 1 - i)) & 1) ? p1[4 + i] : p0[4 + i];
        float t[16];
        t[0] = v;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                // the four HIGH bits by doubling: dimensions 0..3
#pragma unroll
            for (int m = (1 << k) - 1; m >= 0; --m) {
                const float tm = t[m];
                t[2 * m + 1] = tm * p1[k];
                t[2 * m] = tm * p0[k];
            }
        }
        // two values per packed conversion (cvt.rn.f16x2.f32: the scalar f32 -> f16 conversions run at a quarter of the
        // rate and were the producers' bottleneck: 768 per lane and stage)
#pragma unroll
        for (int m = 0; m < 16; m += 2) {
            const __half2 h2 = __floats2half2_rn(t[m], t[m + 1]);
            const float2 hf = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn(t[m] - hf.x, t[m + 1] - hf.y);
            const uint32_t hw = *reinterpret_cast<const uint32_t*>(&h2), lw = *reinterpret_cast<const uint32_t*>(&l2);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = ((m + e) << LB) | g;   // row of the tile
                const uint32_t addr = base_hi[(j >> 1) & 3] + (uint32_t)j * 64u;
                const uint16_t hv = (uint16_t)(e ? hw >> 16 : hw), lv = (uint16_t)(e ? lw >> 16 : lw);
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(hv) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr + (uint32_t)etc::kTile), "h"(lv) : "memory");
            }
        }
    }
}

template <int kH2>       // bits of the second half: d = 7 + kH2 = 14 or 13
__global__ void __launch_bounds__(etc::kThreads, 1)
lfq_entropy_fwd_tc_kernel(EntTcArgs a, float* __restrict__ partial, float* __restrict__ stats) {
    using namespace etc;
    constexpr int kDd = kH + kH2, kNV = 1 << kH2;
    constexpr uint32_t kIdescN = (1u << 4) | ((uint32_t)(kNV >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float red_h[kProducers], red_n[kProducers];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_blocks = (a.n_pairs + kStagePairs - 1) / kStagePairs;
    // stages of this CTA: blocks blockIdx.x, blockIdx.x + gridDim.x, ...
    const int n_st = (int)((n_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 2);                 // the U warp and the V warp of the slot
            mbar_init(&empty_bar[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kProducers) tmem_alloc(&tmem_base_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp < kProducers) {
        // ---------------- producers: warps 2s (U tile) and 2s + 1 (V tile) take this CTA's stages s, s + kStages, ... in slot s
        float h_sum = 0.f, n_valid = 0.f;
        const int s = warp >> 1, half = warp & 1;
        const uint32_t chunk = (uint32_t)lane >> 3, in_chunk = ((uint32_t)lane & 7u) * 2u;
        uint32_t k = 0;                                                      // use count of the slot
        for (int it = s; it < n_st; it += kStages, ++k) {
            const int64_t n = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kStagePairs + lane;
            float p1[kH], p0[kH];
            float v0 = 0.f;
            bool valid = false;
            if (n < a.n_pairs) valid = a.mask[n / a.c] != 0;
            if (valid) {
                v0 = kScale;
                const float* xs = a.x + n * kDd + half * kH;
#pragma unroll
                for (int i = 0; i < kH; ++i) {
                    if (half && i >= kH2) { p1[i] = 0.f; p0[i] = 0.f; continue; }
                    const float u = a.u_scale * __ldg(xs + i);
                    const float e = __expf(-fabsf(u));                      // in (0, 1]
                    const float big = 1.f / (1.f + e), small = e * big;     // sigmoid(|u|), sigmoid(-|u|)
                    p1[i] = u >= 0.f ? big : small;
                    p0[i] = u >= 0.f ? small : big;
                    h_sum += log1pf(e) + fabsf(u) * small;                   // binary entropy of sigmoid(u), nats
                }
                if (half == 0) n_valid += 1.f;
            } else {
#pragma unroll
                for (int i = 0; i < kH; ++i) { p1[i] = 0.f; p0[i] = 0.f; }
            }
            mbar_wait(&empty_bar[s], (k & 1u) ^ 1u);
            const uint32_t tile = smem_u32(smem + s * kStage) + (uint32_t)half * 2u * (uint32_t)kTile;
            uint32_t base[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) base[q] = tile + ((chunk ^ (uint32_t)q) << 4) + in_chunk;
            if (half) emit_half<kH2>(p1, p0, v0, base);
            else emit_half<kH>(p1, p0, v0, base);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> the MMA's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive_local(&full_bar[s]);
        }
        h_sum = warp_sum(h_sum);
        n_valid = warp_sum(n_valid);
        if (lane == 0) { red_h[warp] = h_sum; red_n[warp] = n_valid; }
    } else if (warp == kProducers) {
        // ---------------- MMA issuer
        if (lane == 0) {
            for (int it = 0; it < n_st; ++it) {
                const int rr = it / kFlushStages, acc = rr & 1;
                if (it % kFlushStages == 0) {
                    mbar_wait(&tmem_empty[acc], (((uint32_t)rr >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                }
                const int s = it % kStages;
                mbar_wait(&full_bar[s], (uint32_t)(it / kStages) & 1u);
                tc_fence_after();
                const uint32_t tile = smem_u32(smem + s * kStage);
                const uint32_t tmem_acc = tmem_base + acc * 128;
#pragma unroll
                for (int k = 0; k < kStagePairs / 16; ++k) {
                    const uint32_t ko = k * 32;
                    const uint64_t u_hi = smem_desc_sw64(tile + ko), u_lo = smem_desc_sw64(tile + kTile + ko);
                    const uint64_t v_hi = smem_desc_sw64(tile + 2 * kTile + ko), v_lo = smem_desc_sw64(tile + 3 * kTile + ko);
                    umma_f16(tmem_acc, u_lo, v_hi, kIdescN, (it % kFlushStages || k) ? 1u : 0u);
                    umma_f16(tmem_acc, u_hi, v_lo, kIdescN, 1u);
                    umma_f16(tmem_acc, u_hi, v_hi, kIdescN, 1u);
                }
                umma_commit(&empty_bar[s]);
                if (it % kFlushStages == kFlushStages - 1 || it == n_st - 1) umma_commit(&tmem_full[acc]);
            }
        }
    } else {
        // ---------------- epilogue: TMEM lane quarter = warp & 3; row j1 = quarter * 32 + lane, all 2^kH2 columns j2
        const int quarter = warp & 3;
        const int n_rounds = (n_st + kFlushStages - 1) / kFlushStages;
        float* out = partial + (int64_t)blockIdx.x * (kRows * kNV) + (int64_t)(quarter * 32 + lane) * kNV;
        for (int rr = 0; rr < n_rounds; ++rr) {
            const int acc = rr & 1;
            mbar_wait(&tmem_full[acc], ((uint32_t)rr >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * 128 + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
            for (int c4 = 0; c4 < kNV / 32; ++c4) {
                uint32_t r[32];
                tmem_ld32(taddr + c4 * 32, r);
                if (c4 == kNV / 32 - 1) {                        // accumulator read out: hand it back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_local(&tmem_empty[acc]);
                }
                float4* o4 = reinterpret_cast<float4*>(out + c4 * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 v = rr ? o4[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                    v.x += __uint_as_float(r[4 * q]) * kUnscale;
                    v.y += __uint_as_float(r[4 * q + 1]) * kUnscale;
                    v.z += __uint_as_float(r[4 * q + 2]) * kUnscale;
                    v.w += __uint_as_float(r[4 * q + 3]) * kUnscale;
                    o4[q] = v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kProducers) tmem_dealloc(tmem_base, 256);
    if (threadIdx.x == 0) {
        float h = 0.f, n = 0.f;
        for (int w = 0; w < kProducers; ++w) { h += red_h[w]; n += red_n[w]; }
        stats[2 * blockIdx.x] = h;
        stats[2 * blockIdx.x + 1] = n;
    }
}

// defined in lfq_entropy.cu
int launch_lfq_entropy_final(const float* partial, const float* stats, int n_cta, int c, int d, float eps, float* tables,
                             float* result, cudaStream_t st);

int launch_lfq_entropy_fwd_tc(const float* x, const uint8_t* mask, int64_t n_tok, int c, int d, float u_scale, float eps,
                              float* partial_scratch, float* stats, float* tables, float* result, cudaStream_t st) {
    using namespace etc;
    EntTcArgs a{x, mask, n_tok * c, c, u_scale};
    const int64_t n_blocks = ceil_div(a.n_pairs, kStagePairs);
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)(n_blocks < sms ? n_blocks : sms);
    auto kernel = d == 14 ? lfq_entropy_fwd_tc_kernel<7> : lfq_entropy_fwd_tc_kernel<6>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess)
        return check_launch("lfq_entropy_factorized (tensor-core forward, shared memory opt-in)");
    kernel<<<grid, kThreads, kSmem, st>>>(a, partial_scratch, stats);
    return launch_lfq_entropy_final(partial_scratch, stats, grid, c, d, eps, tables, result, st);
}

}  // namespace dcta
