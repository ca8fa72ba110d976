// Nearest-code search distances on persistent CTA pairs (sm_100a: TMA + tcgen05 cta_group::2 + TMEM).
// Reference: vector_quantize.py:29-33 (cdist), :467-469 (argmax of -dist).
//
//   part[t, n/128] = first minimum over the 128 codes of a slice of   |e_n|^2 + alpha * x_t . e_n
//
// (vq_merge_kernel in gemm_tc.cu adds |x_t|^2, compares the slices in the sqrt domain like the reference and
// gathers).  The x . e products run through the split-precision fp16x3 scheme of the DCT kernels.
//
// A pair of CTAs owns a tile of 256 tokens: their operand (128 tokens x K per CTA, hi + lo) is loaded ONCE and
// stays in shared memory while the whole codebook streams through as the B operand (256 codes per MMA, 128 per
// CTA and ring stage).  L2 -> SM traffic per tile is the codebook once (8 MB for 8192 x 256) instead of the token
// tile once per 128 codes: the 128 x 128-tile kernel it replaces was bound by that traffic.
//   warp 0: TMA producer, warp 1 of the leader: tcgen05.mma issuer, warps 2..9: epilogue (lane quarter = token
//   rows, the two warps of a quarter take one 128-code slice each); two TMEM accumulators of 256 columns.
#include "tc_ptx.cuh"

namespace dcta {

constexpr int VK = 32;                       // k block
constexpr int V_TILE = 128 * VK * 2;         // 128-row operand tile (hi or lo) of one k block: 8 KB
constexpr int V_STAGE = 2 * V_TILE;          // ring stage: this CTA's 128 codes, hi + lo
constexpr int V_THREADS = 320;
constexpr int V_MAX_STAGES = 8;

struct VqArgs {
    int64_t n_tok;
    int n_codes, num_kb, stages, n_ctiles;   // n_ctiles: 256-code tiles
    int n_part;                              // 128-code slices per token = ceil(n_codes / 128)
    int64_t n_ttiles;                        // 256-token tiles
    uint32_t a_bytes;                        // one CTA's resident token operand plane (hi or lo): num_kb * 8 KB
    const float* col_bias;                   // |e_n|^2
    const float* alpha_dev;                  // device scalar: -2 / (scale_x * scale_e)
    float* part_val;                         // (n_tok, n_part)
    int32_t* part_idx;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(V_THREADS, 1)
vq_pair_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
               const __grid_constant__ CUtensorMap map_e_hi, const __grid_constant__ CUtensorMap map_e_lo, VqArgs g) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[V_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[V_MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full[2];
    __shared__ __align__(8) uint64_t tmem_empty[2];
    __shared__ __align__(8) uint64_t a_full, a_empty;
    __shared__ uint32_t tmem_base_slot;

    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_hi = smem;
    uint8_t* a_lo = smem + g.a_bytes;
    uint8_t* ring = smem + 2 * g.a_bytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_x_hi);
        tma_prefetch_desc(&map_x_lo);
        tma_prefetch_desc(&map_e_hi);
        tma_prefetch_desc(&map_e_lo);
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
    __syncthreads();
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
            if (rank == 0) mbar_expect_tx(&a_full, 4 * g.a_bytes);
            const int row0 = (int)(tt * 256 + rank * 128);
            for (int kb = 0; kb < g.num_kb; ++kb) {
                tma_load_3d_2sm(&map_x_hi, a_full_leader, a_hi + kb * V_TILE, kb * VK, row0, 0);
                tma_load_3d_2sm(&map_x_lo, a_full_leader, a_lo + kb * V_TILE, kb * VK, row0, 0);
            }
            for (int ct = 0; ct < g.n_ctiles; ++ct) {
                const int code0 = ct * 256 + (int)rank * 128;
                for (int kb = 0; kb < g.num_kb; ++kb, ++it) {
                    const int s = it % g.stages;
                    mbar_wait(&empty_bar[s], ((it / g.stages) & 1) ^ 1);
                    uint8_t* st = ring + s * V_STAGE;
                    const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[s]), 0);
                    if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * V_STAGE);
                    tma_load_3d_2sm(&map_e_hi, full_leader, st, kb * VK, code0, 0);
                    tma_load_3d_2sm(&map_e_lo, full_leader, st + V_TILE, kb * VK, code0, 0);
                }
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ---------------- MMA issuer (leader): M = 256 tokens, N = 256 codes
        const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo);
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
                    const uint32_t base = smem_u32(ring + s * V_STAGE);
#pragma unroll
                    for (int k = 0; k < VK / 16; ++k) {
                        const uint32_t ko = k * 32;
                        const uint64_t d_a_hi = smem_desc_sw64(ah + kb * V_TILE + ko);
                        const uint64_t d_a_lo = smem_desc_sw64(al + kb * V_TILE + ko);
                        const uint64_t d_b_hi = smem_desc_sw64(base + ko);
                        const uint64_t d_b_lo = smem_desc_sw64(base + V_TILE + ko);
                        umma_f16_2sm(tmem_acc, d_a_lo, d_b_hi, idesc, (kb | k) ? 1u : 0u);   // small terms first
                        umma_f16_2sm(tmem_acc, d_a_hi, d_b_lo, idesc, 1u);
                        umma_f16_2sm(tmem_acc, d_a_hi, d_b_hi, idesc, 1u);
                    }
                    umma_commit_2sm(&empty_bar[s], 3);
                }
                umma_commit_2sm(&tmem_full[acc], 3);
            }
            umma_commit_2sm(&a_empty, 3);            // every MMA of this token tile has read the operand
        }
    } else if (warp >= 2) {
        // ---------------- epilogue: per token row the first minimum of each 128-code slice
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const uint32_t tmem_empty_leader0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
        const uint32_t tmem_empty_leader1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
        const float alpha = __ldg(g.alpha_dev);
        uint32_t tcount = 0;
        for (int64_t tt = pair; tt < g.n_ttiles; tt += n_pairs) {
            const int64_t tok = tt * 256 + rank * 128 + quarter * 32 + lane;
            for (int ct = 0; ct < g.n_ctiles; ++ct, ++tcount) {
                const int acc = tcount & 1;
                mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + acc * 256 + half * 128 + ((uint32_t)(quarter * 32) << 16);
                const int n0 = ct * 256 + half * 128;
                float best = INFINITY;
                int bi = 0x7fffffff;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t rr[32];
                    tmem_ld32_nowait(tmem_acc + c * 32, rr);
                    tmem_ld_wait();
                    if (c == 3) {                      // this warp's slice is read: hand the accumulator back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster_relaxed(acc ? tmem_empty_leader1 : tmem_empty_leader0);
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const int nn = n0 + c * 32 + j;
                        if (nn + 3 < g.n_codes) {
                            const float4 cb = __ldg(reinterpret_cast<const float4*>(g.col_bias + nn));
                            const float v0 = fmaf(alpha, __uint_as_float(rr[j]), cb.x);
                            const float v1 = fmaf(alpha, __uint_as_float(rr[j + 1]), cb.y);
                            const float v2 = fmaf(alpha, __uint_as_float(rr[j + 2]), cb.z);
                            const float v3 = fmaf(alpha, __uint_as_float(rr[j + 3]), cb.w);
                            if (v0 < best) { best = v0; bi = nn; }
                            if (v1 < best) { best = v1; bi = nn + 1; }
                            if (v2 < best) { best = v2; bi = nn + 2; }
                            if (v3 < best) { best = v3; bi = nn + 3; }
                        } else {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (nn + u < g.n_codes) {
                                    const float v = fmaf(alpha, __uint_as_float(rr[j + u]), __ldg(g.col_bias + nn + u));
                                    if (v < best) { best = v; bi = nn + u; }
                                }
                            }
                        }
                    }
                }
                const int slice = ct * 2 + half;
                if (tok < g.n_tok && slice < g.n_part) {
                    g.part_val[tok * g.n_part + slice] = best;
                    g.part_idx[tok * g.n_part + slice] = bi;
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
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
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", (int)r); return DCTA_ERR_LAUNCH; }
    return DCTA_OK;
}

// DCTA_ERR_UNSUPPORTED (no error message) when the token operand does not fit in shared memory
int launch_vq_pair(const void* x_hi, const void* x_lo, const void* e_hi, const void* e_lo, const float* e2,
                   const float* alpha_dev, float* part_val, int32_t* part_idx, int64_t n_tok, int n_codes, int d,
                   int64_t ld, cudaStream_t st) {
    VqArgs g{};
    g.n_tok = n_tok;
    g.n_codes = n_codes;
    g.num_kb = (int)ceil_div(d, VK);
    g.a_bytes = (uint32_t)g.num_kb * V_TILE;
    const int64_t stages = (227 * 1024 - 5120 - 1024 - 2 * (int64_t)g.a_bytes) / V_STAGE;
    if (stages < 3 || 4ll * g.a_bytes >= (1 << 20) || n_tok >= (1ll << 31) - 256 || n_codes < 1) return DCTA_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(e2) & 15) != 0) return DCTA_ERR_UNSUPPORTED;
    g.stages = (int)(stages < V_MAX_STAGES ? stages : V_MAX_STAGES);
    g.n_ctiles = (int)ceil_div(n_codes, 256);
    g.n_part = (int)ceil_div(n_codes, 128);
    g.n_ttiles = ceil_div(n_tok, 256);
    g.col_bias = e2;
    g.alpha_dev = alpha_dev;
    g.part_val = part_val;
    g.part_idx = part_idx;
    CUtensorMap mx_hi, mx_lo, me_hi, me_lo;
    int rc;
    if ((rc = make_map_rows(&mx_hi, x_hi, d, n_tok, ld))) return rc;
    if ((rc = make_map_rows(&mx_lo, x_lo, d, n_tok, ld))) return rc;
    if ((rc = make_map_rows(&me_hi, e_hi, d, n_codes, ld))) return rc;
    if ((rc = make_map_rows(&me_lo, e_lo, d, n_codes, ld))) return rc;
    const int smem_bytes = 1024 + 2 * (int)g.a_bytes + g.stages * V_STAGE;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t pairs = sms / 2;
    if (pairs > g.n_ttiles) pairs = g.n_ttiles;
    if (pairs < 1) pairs = 1;
    cudaError_t e = cudaFuncSetAttribute(vq_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { set_error("vq_pair: %s", cudaGetErrorString(e)); return DCTA_ERR_LAUNCH; }
    vq_pair_kernel<<<(unsigned)(2 * pairs), V_THREADS, smem_bytes, st>>>(mx_hi, mx_lo, me_hi, me_lo, g);
    return check_launch("vq_pair");
}

}  // namespace dcta
