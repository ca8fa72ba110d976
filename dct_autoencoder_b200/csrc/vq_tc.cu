// Nearest-code search on persistent CTA pairs (sm_100a: TMA + tcgen05 cta_group::2 + TMEM), single pass.
// Reference: vector_quantize.py:29-33 (cdist), :467-469 (argmax of -dist), :222-226 (gather).
//
// Pass 1 (vq_pair_kernel): APPROXIMATE distances  |e_n|^2 - 2 x_t . e_n  from ONE fp16 tcgen05.mma per product (each
//   operand rounded to 11 significant bits, rows scaled by powers of two; absolute error ~1e-5 of |x||e|), reduced in
//   the epilogue to the two smallest per token and half of the code slices: four candidates per token.
// Pass 2 (vq_rerank_kernel): the candidates re-scored EXACTLY in fp32 with the reference's formula and its
//   first-index rule.  The true nearest code is lost only if three codes of one half lie within the approximation
//   error of each other, in which case the returned code's distance is within ~1e-4 relative of the minimum.
// One third of the tensor work and half of the L2 -> SM operand traffic of the split-precision (fp16 x 3) version it
// replaces.
//
// A pair of CTAs owns a tile of 256 tokens: their operand (128 tokens x K per CTA) is loaded ONCE and stays in shared
// memory while the whole codebook streams through as the B operand (256 codes per MMA, 128 per CTA and ring stage).
// Operand tiles are 128 rows x 64 k (128-byte rows, SWIZZLE_128B).
//   warp 0: TMA producer, warp 1 of the leader: tcgen05.mma issuer, warps 2..9: epilogue (lane quarter = token
//   rows, the two warps of a quarter take one 128-code slice each); two TMEM accumulators of 256 columns.
#include "tc_ptx.cuh"

namespace dcta {

constexpr int VK = 64;                       // k block: 128-byte rows (SWIZZLE_128B) -- half the TMA row requests of 64-byte rows,
                                             // which bounded the kernel at ~2.7 cycles per row per SM
constexpr int V_TILE = 128 * VK * 2;         // 128-row fp16 operand tile of one k block: 16 KB
constexpr int V_STAGE = V_TILE;              // ring stage: this CTA's 128 codes
constexpr int V_THREADS = 320;
constexpr int V_MAX_STAGES = 8;

struct VqArgs {
    int64_t n_tok;
    int n_codes, num_kb, stages, n_ctiles;   // n_ctiles: 256-code tiles
    int64_t n_ttiles;                        // 256-token tiles
    uint32_t a_bytes;                        // one CTA's resident token operand: num_kb * 8 KB
    uint32_t e2_bytes;                       // |e_n|^2 of the whole codebook, staged in shared memory (padded to 256 codes)
    const float* col_bias;                   // |e_n|^2
    const float* row_alpha;                  // (n_tok) -2 / (scale of the token's row * scale of the codebook)
    int32_t* cand;                           // (n_tok, 4): the two best codes of each half of the code slices
    float* cand_val;                         // [nullable] (n_tok, 4): their approximate |e|^2 - 2 x.e
};

// running two smallest (value, index) pairs; strict '<' keeps the earlier index among equal values
__device__ __forceinline__ void top2_insert(float v, int n, float& b1, int& i1, float& b2, int& i2) {
    const bool lt1 = v < b1, lt2 = v < b2;
    b2 = lt1 ? b1 : (lt2 ? v : b2);
    i2 = lt1 ? i1 : (lt2 ? n : i2);
    b1 = lt1 ? v : b1;
    i1 = lt1 ? n : i1;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(V_THREADS, 1)
vq_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_e, VqArgs g) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[V_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[V_MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full[2];
    __shared__ __align__(8) uint64_t tmem_empty[2];
    __shared__ __align__(8) uint64_t a_full, a_empty;
    __shared__ uint32_t tmem_base_slot;

    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_op = smem;
    uint8_t* ring = smem + g.a_bytes;
    float* s_e2 = reinterpret_cast<float*>(ring + g.stages * V_STAGE);
    // |e_n|^2 of every code: read by every epilogue thread for every token tile.  With 200+ KB of shared memory carved
    // out, L1 holds almost nothing and each of those reads was an L2 round trip on the epilogue's critical path
    // (ncu: 60 % of the stall samples on the FFMA consuming it).  +inf past the last code: those columns never win.
    for (int i = threadIdx.x; i < (int)(g.e2_bytes >> 2); i += V_THREADS) s_e2[i] = i < g.n_codes ? __ldg(g.col_bias + i) : INFINITY;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_e);
        for (int s = 0; s < V_MAX_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 16);
        }
        mbar_init(&a_full, 1);
        mbar_init(&a_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc_2sm(&tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();                                  // (also publishes s_e2)
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer
        const uint32_t a_full_leader = mapa_u32(smem_u32(&a_full), 0);
        uint32_t it = 0, tt_count = 0;
        for (int64_t tt = pair; tt < g.n_ttiles; tt += n_pairs, ++tt_count) {
            // the token operand of this tile: wait until the MMAs of the previous tile have read the old one
            mbar_wait(&a_empty, (tt_count & 1) ^ 1);
            if (rank == 0) mbar_expect_tx(&a_full, 2 * g.a_bytes);
            const int row0 = (int)(tt * 256 + rank * 128);
            for (int kb = 0; kb < g.num_kb; ++kb) tma_load_3d_2sm(&map_x, a_full_leader, a_op + kb * V_TILE, kb * VK, row0, 0);
            for (int ct = 0; ct < g.n_ctiles; ++ct) {
                const int code0 = ct * 256 + (int)rank * 128;
                for (int kb = 0; kb < g.num_kb; ++kb, ++it) {
                    const int s = it % g.stages;
                    mbar_wait(&empty_bar[s], ((it / g.stages) & 1) ^ 1);
                    const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[s]), 0);
                    if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * V_STAGE);
                    tma_load_3d_2sm(&map_e, full_leader, ring + s * V_STAGE, kb * VK, code0, 0);
                }
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ---------------- MMA issuer (leader): M = 256 tokens, N = 256 codes, one fp16 MMA per k16 step
        const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint32_t a_lo32 = smem_desc_lo(smem_u32(a_op)), ring_lo32 = smem_desc_lo(smem_u32(ring));
        uint32_t it = 0, tcount = 0, tt_count = 0;
        for (int64_t tt = pair; tt < g.n_ttiles; tt += n_pairs, ++tt_count) {
            mbar_wait_cluster(&a_full, tt_count & 1);
            tc_fence_after();
            for (int ct = 0; ct < g.n_ctiles; ++ct, ++tcount) {
                const int acc = tcount & 1;
                mbar_wait_cluster(&tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + acc * 256;
                for (int kb = 0; kb < g.num_kb; ++kb, ++it) {
                    const int s = it % g.stages;
                    mbar_wait_cluster(&full_bar[s], (it / g.stages) & 1);
                    tc_fence_after();
                    const uint32_t a16 = a_lo32 + (uint32_t)(kb * (V_TILE >> 4)), b16 = ring_lo32 + (uint32_t)(s * (V_STAGE >> 4));
#pragma unroll
                    for (int k = 0; k < VK / 16; ++k)
                        umma_f16_2sm(tmem_acc, smem_desc_sw128_from_lo(a16 + 2 * k), smem_desc_sw128_from_lo(b16 + 2 * k), idesc,
                                     (kb | k) ? 1u : 0u);
                    umma_commit_2sm(&empty_bar[s], 3);
                }
                umma_commit_2sm(&tmem_full[acc], 3);
            }
            umma_commit_2sm(&a_empty, 3);            // every MMA of this token tile has read the operand
        }
    } else if (warp >= 2) {
        // ---------------- epilogue: per token row the two smallest  |e_n|^2 + alpha_t * acc  of this warp's half of
        // the 128-code slices, carried over the whole codebook
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const uint32_t tmem_empty_leader0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
        const uint32_t tmem_empty_leader1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
        uint32_t tcount = 0;
        for (int64_t tt = pair; tt < g.n_ttiles; tt += n_pairs) {
            const int64_t tok = tt * 256 + rank * 128 + quarter * 32 + lane;
            const float alpha = tok < g.n_tok ? __ldg(g.row_alpha + tok) : 0.f;
            // Running two smallest values of the row.  After the first few hundred codes an element rarely beats the
            // current second best (probability 2 / codes seen), so a quad of columns is first reduced to its minimum
            // (3 FMNMX) and the 4 ordered inserts run only when some lane of the warp needs them (warp-uniform branch):
            // ~4 instead of ~12 instructions per element (ncu: the epilogue, not the MMA, bounded the kernel).
            float b1 = INFINITY, b2 = INFINITY;
            int i1 = 0x7fffffff, i2 = 0x7fffffff;
            auto process = [&](const uint32_t (&rr)[32], int nbase) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int nn = nbase + j;
                    const float4 cb = *reinterpret_cast<const float4*>(s_e2 + nn);
                    const float v0 = fmaf(alpha, __uint_as_float(rr[j]), cb.x), v1 = fmaf(alpha, __uint_as_float(rr[j + 1]), cb.y);
                    const float v2 = fmaf(alpha, __uint_as_float(rr[j + 2]), cb.z), v3 = fmaf(alpha, __uint_as_float(rr[j + 3]), cb.w);
                    const float m = fminf(fminf(v0, v1), fminf(v2, v3));
                    if (__any_sync(0xffffffffu, m < b2)) {
                        top2_insert(v0, nn, b1, i1, b2, i2);
                        top2_insert(v1, nn + 1, b1, i1, b2, i2);
                        top2_insert(v2, nn + 2, b1, i1, b2, i2);
                        top2_insert(v3, nn + 3, b1, i1, b2, i2);
                    }
                }
            };
            for (int ct = 0; ct < g.n_ctiles; ++ct, ++tcount) {
                const int acc = tcount & 1;
                mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + acc * 256 + half * 128 + ((uint32_t)(quarter * 32) << 16);
                const int n0 = ct * 256 + half * 128;
                uint32_t ra[32], rb[32];
                tmem_ld32_nowait(tmem_acc, ra);
#pragma unroll 1
                for (int c = 0; c < 4; c += 2) {
                    tmem_ld_wait();
                    tmem_ld32_nowait(tmem_acc + (c + 1) * 32, rb);
                    process(ra, n0 + c * 32);
                    tmem_ld_wait();
                    if (c + 2 < 4) tmem_ld32_nowait(tmem_acc + (c + 2) * 32, ra);
                    else {                             // this warp's slice is read: hand the accumulator back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster_relaxed(acc ? tmem_empty_leader1 : tmem_empty_leader0);
                    }
                    process(rb, n0 + (c + 1) * 32);
                }
            }
            if (tok < g.n_tok) {
                *reinterpret_cast<int2*>(g.cand + tok * 4 + half * 2) = make_int2(i1, i2);
                if (g.cand_val) *reinterpret_cast<float2*>(g.cand_val + tok * 4 + half * 2) = make_float2(b1, b2);
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

// torch.argmax(-dist) semantics: the first minimum wins, and a NaN (sqrt of a rounding-negative argument; the reference
// does not clamp, vector_quantize.py:29-33) beats every number
__device__ __forceinline__ bool vq_better(float v, int i, float bv, int bi) {
    if (v != v) return (bv == bv) || i < bi;
    if (bv != bv) return false;
    return v < bv || (v == bv && i < bi);
}

// exact fp32 re-rank of the (up to) four candidates of every token: dist = sqrt(|x|^2 + (|e|^2 - 2 x.e)) as in
// vector_quantize.py:29-33, first minimum wins (torch.argmax(-dist), :467-469); optional gather (:222-226).
// One warp per token.
// keep (nullable, n_tok bytes): where 0 the token keeps its own row in `quantized` (vector_quantize.py:1043-1048,
// torch.where(mask, quantize, orig_input), for the projection-free layer) -- the index is computed either way.
// VEC: d % 4 == 0 and 16-byte aligned rows: 128-bit loads, the token and its four candidates in flight together.
// cand_val / e2_max [nullable together]: the approximate values of the four candidates and max_n |e_n|^2.  Every
// approximate value is within E = 2^-8 |x| max|e| of the exact |e|^2 - 2 x.e (operands rounded to fp16: 2 * 2^-11 relative
// per product, i.e. 2^-9 |x||e| on the value by Cauchy-Schwarz; the bound used is twice that and also covers the fp32
// accumulation), and every code that is NOT a candidate has an approximate value >= the runner-up's.  So when the
// runner-up exceeds the best by more than 2E the best candidate is the exact argmin (no tie is possible) and the four
// exact distances need not be formed: only the token's own row is read (for |x|) and the winner's row gathered.
// G lanes per token (G = 32: one warp per token, scalar loads; G = 8: four tokens per warp in flight, 128-bit loads --
// the kernel is bound by the latency of its dependent loads (candidates -> rows), not by bytes).
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool VEC>
__global__ void __launch_bounds__(256) vq_rerank_kernel(const float* __restrict__ x, const float* __restrict__ embed,
                                                        const float* __restrict__ e2, const int32_t* __restrict__ cand,
                                                        const float* __restrict__ cand_val, const float* __restrict__ e2_max,
                                                        const uint8_t* __restrict__ keep, int64_t n_tok, int n_codes, int d,
                                                        int64_t* __restrict__ indices, float* __restrict__ quantized) {
    constexpr int G = VEC ? 8 : 32;                   // lanes per token
    constexpr int TPW = 32 / G;                       // tokens per warp
    const int lane = threadIdx.x & 31, gl = lane % G, sub = lane / G;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int d4 = d >> 2;
    const bool shortcut = VEC && cand_val != nullptr && e2_max != nullptr;
    const float e_max = shortcut ? sqrtf(__ldg(e2_max)) : 0.f;
    for (int64_t t0 = warp0 * TPW; t0 < n_tok; t0 += n_warps * TPW) {        // warp-uniform trip count
        const int64_t t = t0 + sub;
        const bool live = t < n_tok;
        const int64_t tt = live ? t : n_tok - 1;                              // idle groups shadow the last token
        const int4 c4 = *reinterpret_cast<const int4*>(cand + tt * 4);
        const int cs[4] = {c4.x, c4.y, c4.z, c4.w};
        float x2 = 0.f, dot[4] = {0.f, 0.f, 0.f, 0.f};
        int bi = 0x7fffffff;
        bool decided = false;
        if (shortcut) {
            const float4 a4 = *reinterpret_cast<const float4*>(cand_val + tt * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            int kb = 0;
#pragma unroll
            for (int k = 1; k < 4; ++k) if (av[k] < av[kb]) kb = k;
            float second = INFINITY;
            int cbest = cs[0];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k != kb) second = fminf(second, av[k]);
                else cbest = cs[k];
            }
            const float4* xr = reinterpret_cast<const float4*>(x + tt * d);
            float xs = 0.f;
            for (int i = gl; i < d4; i += G) {
                const float4 v = __ldg(xr + i);
                xs = fmaf(v.w, v.w, fmaf(v.z, v.z, fmaf(v.y, v.y, fmaf(v.x, v.x, xs))));
            }
            xs = group_sum<G>(xs);
            const float bound = 0.0078125f * sqrtf(xs) * e_max;              // 2E = 2 * 2^-8 |x| max|e|
            if (cbest < n_codes && second - av[kb] > bound && xs < 3.0e38f) {   // (false for NaN / inf anywhere)
                bi = cbest;
                decided = true;
            }
        }
        if (!decided) {
            if (VEC) {
                const float4* xr = reinterpret_cast<const float4*>(x + tt * d);
                for (int i = gl; i < d4; i += G) {
                    const float4 v = __ldg(xr + i);
                    float4 e[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        e[k] = cs[k] < n_codes ? __ldg(reinterpret_cast<const float4*>(embed + (int64_t)cs[k] * d) + i)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                    x2 = fmaf(v.w, v.w, fmaf(v.z, v.z, fmaf(v.y, v.y, fmaf(v.x, v.x, x2))));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        dot[k] = fmaf(v.w, e[k].w, fmaf(v.z, e[k].z, fmaf(v.y, e[k].y, fmaf(v.x, e[k].x, dot[k]))));
                }
            } else {
                for (int i = gl; i < d; i += G) {
                    const float v = __ldg(x + tt * d + i);
                    x2 = fmaf(v, v, x2);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (cs[k] < n_codes) dot[k] = fmaf(v, __ldg(embed + (int64_t)cs[k] * d + i), dot[k]);
                }
            }
        }
        // (the reductions are executed by every lane: groups that took the shortcut contribute zeros and ignore the result)
        x2 = group_sum<G>(x2);
        float bv = INFINITY;
        int bj = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float dk = group_sum<G>(dot[k]);
            if (cs[k] < n_codes) {
                const float v = __fsqrt_rn(__fadd_rn(x2, fmaf(-2.0f, dk, __ldg(e2 + cs[k]))));
                if (bj == 0x7fffffff || vq_better(v, cs[k], bv, bj)) { bv = v; bj = cs[k]; }
            }
        }
        if (!decided) bi = bj;
        if (live && gl == 0) indices[t] = bi;
        if (live && quantized) {
            const float* src = (keep != nullptr && keep[t] == 0) ? x + t * d : embed + (int64_t)bi * d;
            if (VEC) {
                for (int i = gl; i < d4; i += G)
                    reinterpret_cast<float4*>(quantized + t * d)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
            } else {
                for (int i = gl; i < d; i += G) quantized[t * d + i] = __ldg(src + i);
            }
        }
    }
}

static int make_map_rows(CUtensorMap* map, const void* ptr, int64_t k, int64_t rows, int64_t ld) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DCTA_ERR_UNSUPPORTED; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8)) {
        set_error("vq_nearest_tc: operand planes need a 16-byte aligned base and pitch");
        return DCTA_ERR_INVALID_ARG;
    }
    cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, 1};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * ld * 2};
    cuuint32_t box[3] = {VK, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", (int)r); return DCTA_ERR_LAUNCH; }
    return DCTA_OK;
}

// DCTA_ERR_UNSUPPORTED (no error message) when the token operand does not fit in shared memory
int launch_vq_pair(const void* x_hi, const void* e_hi, const float* e2, const float* row_alpha, int32_t* cand, float* cand_val,
                   int64_t n_tok, int n_codes, int d, int64_t ld, cudaStream_t st) {
    VqArgs g{};
    g.cand_val = cand_val;
    g.n_tok = n_tok;
    g.n_codes = n_codes;
    g.num_kb = (int)ceil_div(d, VK);
    g.a_bytes = (uint32_t)g.num_kb * V_TILE;
    g.e2_bytes = (uint32_t)(ceil_div(n_codes, 256) * 256 * 4);
    const int64_t stages = (227 * 1024 - 5120 - 1024 - (int64_t)g.a_bytes - (int64_t)g.e2_bytes) / V_STAGE;
    if (stages < 3 || 2ll * g.a_bytes >= (1 << 20) || n_tok >= (1ll << 31) - 256 || n_codes < 1) return DCTA_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(e2) & 15) != 0 || (reinterpret_cast<uintptr_t>(cand) & 15) != 0) return DCTA_ERR_UNSUPPORTED;
    g.stages = (int)(stages < V_MAX_STAGES ? stages : V_MAX_STAGES);
    g.n_ctiles = (int)ceil_div(n_codes, 256);
    g.n_ttiles = ceil_div(n_tok, 256);
    g.col_bias = e2;
    g.row_alpha = row_alpha;
    g.cand = cand;
    CUtensorMap mx, me;
    int rc;
    if ((rc = make_map_rows(&mx, x_hi, d, n_tok, ld))) return rc;
    if ((rc = make_map_rows(&me, e_hi, d, n_codes, ld))) return rc;
    const int smem_bytes = 1024 + (int)g.a_bytes + g.stages * V_STAGE + (int)g.e2_bytes;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t pairs = sms / 2;
    if (pairs > g.n_ttiles) pairs = g.n_ttiles;
    if (pairs < 1) pairs = 1;
    cudaError_t e = cudaFuncSetAttribute(vq_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { set_error("vq_pair: %s", cudaGetErrorString(e)); return DCTA_ERR_LAUNCH; }
    vq_pair_kernel<<<(unsigned)(2 * pairs), V_THREADS, smem_bytes, st>>>(mx, me, g);
    return check_launch("vq_pair");
}

int launch_vq_rerank(const float* x, const float* embed, const float* e2, const int32_t* cand, const float* cand_val,
                     const float* e2_max, const uint8_t* keep, int64_t n_tok, int n_codes, int d, int64_t* indices,
                     float* quantized, cudaStream_t st) {
    const bool vec = d % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(embed) |
                                     reinterpret_cast<uintptr_t>(quantized) | reinterpret_cast<uintptr_t>(cand_val)) & 15) == 0;
    if (vec) vq_rerank_kernel<true><<<grid_for(n_tok, 32), 256, 0, st>>>(x, embed, e2, cand, cand_val, e2_max, keep, n_tok, n_codes, d, indices, quantized);
    else vq_rerank_kernel<false><<<grid_for(n_tok, 8), 256, 0, st>>>(x, embed, e2, cand, cand_val, e2_max, keep, n_tok, n_codes, d, indices, quantized);
    return check_launch("vq_rerank");
}

}  // namespace dcta
