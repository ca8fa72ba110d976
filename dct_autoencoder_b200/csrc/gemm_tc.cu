// Split-precision tensor-core GEMM for the DCT / IDCT basis contractions (sm_100a: TMA + tcgen05 + TMEM).
//
//   D[b] (M x N, fp32) = sum_k A[b][m,k] * B[b][n,k]          (both operands K-major)
//
// Every fp32 operand v is carried as two fp16 planes  v*2^s = hi + lo  (hi = rn16(v*2^s),
// lo = rn16(v*2^s - hi): 22 significand bits, fp16 subnormals keep the absolute error at 2^-25).
// The product is accumulated in fp32 in TMEM from three tcgen05.mma per k-step:
//   hi*hi + hi*lo + lo*hi        (the dropped lo*lo term is 2^-22 relative)
// at the fp16 tensor rate, i.e. fp32-class accuracy at 1/3 of the dense fp16 peak.
//
// Kernel shape (one 128x128 output tile per CTA, 2 CTAs per SM so one CTA's epilogue overlaps the
// other's main loop):
//   warp 0 / one lane : TMA producer  -- 4 tiles (A_hi, A_lo, B_hi, B_lo) of 128 x 32 fp16 per stage,
//                       SWIZZLE_64B, 3-stage mbarrier ring (96 KB)
//   warp 1 / one lane : tcgen05.mma issuer, M=128 N=128 K=16, accumulator = 128 TMEM columns
//   all 4 warps       : epilogue: tcgen05.ld 32x32b -> registers -> (scale) -> shared-memory
//                       transpose (reusing the pipeline buffers) -> coalesced global stores in one
//                       of three layouts (fp32 planes, fp16 hi/lo planes for the next GEMM, or the
//                       token grid of feature_extraction_dct_autoencoder.py:374-380).
#include "tc_ptx.cuh"

namespace dcta {

constexpr int TM = 128, TN = 128, TK = 32, STAGES = 3;
constexpr int TILE_BYTES = TM * TK * 2;          // 8 KB: one 128 x 32 fp16 operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;      // A_hi, A_lo, B_hi, B_lo
constexpr int EPI_PITCH = TN + 4;                // fp32 staging pitch (conflict-free 128-bit accesses)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack
static_assert(TM * EPI_PITCH * 4 <= STAGES * STAGE_BYTES, "epilogue staging must fit in the ring");

struct EpiArgs {
    int mode;                 // 0: fp32 planes, 1: fp16 hi/lo planes, 2: fp32 token grid, 3: arg-min partials,
                              // 4: fp16 hi/lo planes written TRANSPOSED, out[b][n*ld + m] (persistent kernel)
                              // 5: LFQ sign (lfq.py:175-187): out_hi = +-lfq_scale as fp16 (the operand of project_out),
                              //    sign_bits[m][n tile][j] = ballot of (column n0 + 4*lane + j > 0)
    const float* col_bias;    // optional per-column (n) bias added after the scaling (modes 0, 1, 5)
    uint32_t* sign_bits;      // mode 5
    float lfq_scale;          // mode 5
    int a_lo_zero;            // the A operand has no lo plane (exact fp16 values): two MMAs per k step, no A_lo loads
    float* out_f32;
    __half* out_hi;
    __half* out_lo;
    int64_t ld, batch_stride; // element pitch / per-batch stride of the plane outputs
    const float* row_scale;   // optional per-row (m) factor
    const float* col_scale;   // optional per-column (n) factor (persistent kernel)
    float alpha;              // global factor
    const float* alpha_dev;   // optional device scalar multiplied into alpha (scale chosen on the device)
    const float* dc;          // optional per-batch constant handled outside the GEMM (see dc_mode)
    int dc_mode;              // 0: none, 1: add dc[b] to element (0,0), 2: add dc[b] to every element
    int M, N;
    int tile_p, channels, tiles_h, tiles_w;
};

// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=TN
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

// ------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(128, 2)
gemm_split_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                  int a_batched, int b_batched, int num_k_blocks, EpiArgs ep) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t acc_bar;
    __shared__ uint32_t tmem_base_slot;

    // 1024-byte alignment computed in the shared window so that the pointer stays a shared pointer
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * TN, m0 = blockIdx.y * TM, batch = blockIdx.z;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, TN);   // 128 fp32 accumulator columns
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer
        const int ab = a_batched ? batch : 0, bb = b_batched ? batch : 0;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* st = smem + s * STAGE_BYTES;
            mbar_expect_tx(&full_bar[s], ep.a_lo_zero ? STAGE_BYTES - TILE_BYTES : STAGE_BYTES);
            tma_load_3d(&map_a_hi, &full_bar[s], st, kb * TK, m0, ab);
            if (!ep.a_lo_zero) tma_load_3d(&map_a_lo, &full_bar[s], st + TILE_BYTES, kb * TK, m0, ab);
            tma_load_3d(&map_b_hi, &full_bar[s], st + 2 * TILE_BYTES, kb * TK, n0, bb);
            tma_load_3d(&map_b_lo, &full_bar[s], st + 3 * TILE_BYTES, kb * TK, n0, bb);
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer
        for (int kb = 0; kb < num_k_blocks; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < TK / 16; ++k) {
                const uint32_t ko = k * 32;  // 16 fp16 = 32 bytes along K inside the swizzle atom
                const uint64_t a_hi = smem_desc_sw64(base + ko);
                const uint64_t a_lo = smem_desc_sw64(base + TILE_BYTES + ko);
                const uint64_t b_hi = smem_desc_sw64(base + 2 * TILE_BYTES + ko);
                const uint64_t b_lo = smem_desc_sw64(base + 3 * TILE_BYTES + ko);
                if (ep.a_lo_zero) {
                    umma_f16(tmem_acc, a_hi, b_lo, kIdesc, (kb | k) ? 1u : 0u);
                } else {
                    umma_f16(tmem_acc, a_lo, b_hi, kIdesc, (kb | k) ? 1u : 0u);   // small terms first
                    umma_f16(tmem_acc, a_hi, b_lo, kIdesc, 1u);
                }
                umma_f16(tmem_acc, a_hi, b_hi, kIdesc, 1u);
            }
            umma_commit(&empty_bar[s]);          // frees the stage when these MMAs have read it
        }
        umma_commit(&acc_bar);                   // accumulator complete
    }
    __syncwarp();

    // ---------------- epilogue (all 4 warps; warp w owns TMEM lanes / tile rows 32w .. 32w+31)
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    // The ring is idle now (all MMAs have completed): reuse it as a 128 x 132 fp32 staging tile.
    // Pitch 132 keeps 128-bit accesses conflict-free in both directions: the TMEM side writes one
    // row per lane (8 lanes x 4 words x stride 4 banks = 32 banks), the store side reads one row per
    // warp with 4 consecutive columns per lane.
    const uint32_t stage = smem_u32(smem);
    const int row = warp * 32 + lane;
    const int gm = m0 + row;
    float rs = ep.alpha;
    if (ep.alpha_dev != nullptr) rs *= __ldg(ep.alpha_dev);
    if (ep.row_scale != nullptr && gm < ep.M) rs *= __ldg(ep.row_scale + gm);
#pragma unroll
    for (int c = 0; c < TN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_acc + ((uint32_t)(warp * 32) << 16) + c * 32, r);
        const uint32_t dst = stage + (uint32_t)(row * EPI_PITCH + c * 32) * 4;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + j * 4),
                         "f"(__uint_as_float(r[j]) * rs), "f"(__uint_as_float(r[j + 1]) * rs),
                         "f"(__uint_as_float(r[j + 2]) * rs), "f"(__uint_as_float(r[j + 3]) * rs) : "memory");
    }
    __syncwarp();
    const float dcv = ep.dc_mode ? __ldg(ep.dc + batch) : 0.0f;
    const int n = n0 + lane * 4;                 // this lane's 4 consecutive output columns
    const int n_valid = ep.N - n;                // >= 4: all four in range
    // per-lane column offsets of the token-grid layout (constant over the rows)
    int64_t col_off[4] = {0, 0, 0, 0};
    int64_t plane_off = 0;
    const int p = ep.tile_p, zz = p * p;
    if (ep.mode == 2) {
        const int64_t img = batch / ep.channels;
        const int ch = batch - (int)img * ep.channels;
        plane_off = (img * ep.tiles_h * ep.tiles_w * ep.channels + ch) * (int64_t)zz;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int tw = (n + j) / p;
            col_off[j] = (int64_t)tw * ep.channels * zz + ((n + j) - tw * p);
        }
    } else {
        plane_off = (int64_t)batch * ep.batch_stride + n;
    }
    const bool pair_ok = (p % 2 == 0);           // columns (n, n+1) and (n+2, n+3) never straddle a tile
    float bias4[4] = {0.f, 0.f, 0.f, 0.f};
    if (ep.col_bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < n_valid) bias4[j] = __ldg(ep.col_bias + n + j);
    }
    for (int rr = 0; rr < 32; ++rr) {
        const int m = m0 + warp * 32 + rr;
        if (m >= ep.M) break;
        if (n_valid <= 0 && ep.mode != 5) continue;
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(stage + (uint32_t)((warp * 32 + rr) * EPI_PITCH + lane * 4) * 4));
        if (ep.dc_mode == 2) { v.x += dcv; v.y += dcv; v.z += dcv; v.w += dcv; }
        else if (ep.dc_mode == 1 && m == 0 && n == 0) v.x += dcv;
        v.x += bias4[0]; v.y += bias4[1]; v.z += bias4[2]; v.w += bias4[3];
        if (ep.mode == 5) {
            // LFQ: x > 0 -> +scale and bit 1, anything else (0, NaN, negative) -> -scale and bit 0 (lfq.py:175, 187)
            const float a[4] = {v.x, v.y, v.z, v.w};
            __half h[4];
            uint32_t bal[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool pos = j < n_valid && a[j] > 0.0f;
                h[j] = __float2half_rn(pos ? ep.lfq_scale : -ep.lfq_scale);
                bal[j] = __ballot_sync(0xffffffffu, pos);
            }
            if (lane < 4) ep.sign_bits[((int64_t)m * gridDim.x + blockIdx.x) * 4 + lane] = bal[lane];
            if (n_valid > 0) {
                const int64_t o = plane_off + (int64_t)m * ep.ld;
                if (n_valid >= 4) {
                    *reinterpret_cast<uint2*>(ep.out_hi + o) = *reinterpret_cast<const uint2*>(h);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j < n_valid) ep.out_hi[o + j] = h[j];
                }
            }
            continue;
        }
        if (ep.mode == 0) {
            float* dst = ep.out_f32 + plane_off + (int64_t)m * ep.ld;
            if (n_valid >= 4 && (ep.ld & 3) == 0) {
                *reinterpret_cast<float4*>(dst) = v;
            } else {
                const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) if (j < n_valid) dst[j] = a[j];
            }
        } else if (ep.mode == 1) {
            const float a[4] = {v.x, v.y, v.z, v.w};
            __half h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                h[j] = __float2half_rn(a[j]);
                l[j] = __float2half_rn(a[j] - __half2float(h[j]));
            }
            const int64_t o = plane_off + (int64_t)m * ep.ld;   // ld % 8 == 0, n % 4 == 0: 8-byte aligned
            if (n_valid >= 4) {
                *reinterpret_cast<uint2*>(ep.out_hi + o) = *reinterpret_cast<const uint2*>(h);
                *reinterpret_cast<uint2*>(ep.out_lo + o) = *reinterpret_cast<const uint2*>(l);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (j < n_valid) { ep.out_hi[o + j] = h[j]; ep.out_lo[o + j] = l[j]; }
            }
        } else {
            const int th = m / p, pi = m - th * p;
            float* dst = ep.out_f32 + plane_off + (int64_t)th * ep.tiles_w * ep.channels * zz + pi * p;
            if (pair_ok && n_valid >= 4) {
                *reinterpret_cast<float2*>(dst + col_off[0]) = make_float2(v.x, v.y);
                *reinterpret_cast<float2*>(dst + col_off[2]) = make_float2(v.z, v.w);
            } else {
                const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) if (j < n_valid) dst[col_off[j]] = a[j];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_acc, TN);
}

// ------------------------------------------------------------------------------ host: tensor maps
// rows x k fp16 matrix with pitch `ld` elements, `batch` of them `batch_stride` elements apart
static int make_map(CUtensorMap* map, const void* ptr, int rows, int k, int64_t ld, int64_t batch, int64_t batch_stride,
                    int box_rows = TM) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DCTA_ERR_UNSUPPORTED; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8) || (batch > 1 && batch_stride % 8)) {
        set_error("gemm_split: operand planes need 16-byte aligned base, pitch and batch stride");
        return DCTA_ERR_INVALID_ARG;
    }
    cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)(batch > 0 ? batch : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? batch_stride : (int64_t)rows * ld) * 2};
    cuuint32_t box[3] = {TK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", (int)r); return DCTA_ERR_LAUNCH; }
    return DCTA_OK;
}

struct Operand {
    const __half* hi;
    const __half* lo;
    int rows;
    int64_t ld, batch_stride;   // batch_stride == 0: shared by every batch item
};

static int launch_gemm_split(const Operand& A, const Operand& B, int K, int64_t batch, const EpiArgs& ep, void* stream) {
    if (batch == 0 || ep.M == 0 || ep.N == 0) return DCTA_OK;
    if (batch > 65535) { set_error("gemm_split: more than 65535 batch items per call"); return DCTA_ERR_UNSUPPORTED; }
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    const int64_t ab = A.batch_stride ? batch : 1, bb = B.batch_stride ? batch : 1;
    if ((rc = make_map(&ma_hi, A.hi, A.rows, K, A.ld, ab, A.batch_stride))) return rc;
    if ((rc = make_map(&ma_lo, A.lo, A.rows, K, A.ld, ab, A.batch_stride))) return rc;
    if ((rc = make_map(&mb_hi, B.hi, B.rows, K, B.ld, bb, B.batch_stride))) return rc;
    if ((rc = make_map(&mb_lo, B.lo, B.rows, K, B.ld, bb, B.batch_stride))) return rc;
    cudaFuncSetAttribute(gemm_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    dim3 grid((unsigned)ceil_div(ep.N, TN), (unsigned)ceil_div(ep.M, TM), (unsigned)batch);
    gemm_split_kernel<<<grid, 128, SMEM_BYTES, as_stream(stream)>>>(ma_hi, ma_lo, mb_hi, mb_lo, A.batch_stride != 0,
                                                                      B.batch_stride != 0, (int)ceil_div(K, TK), ep);
    return check_launch("gemm_split");
}

// ------------------------------------------------------------------------------ persistent kernel
// Same arithmetic, fewer L2 bytes per flop: 128 x TNP output tiles (TNP = 256, or 224 so that N = 448
// tiles exactly), one persistent CTA per SM walking a static tile schedule, and the 512 TMEM columns
// split into two accumulators so that the epilogue of tile i overlaps the main loop of tile i+1.
//   warp 0 / one lane : TMA producer (ring of PSTAGES stages: A 128x32 hi/lo + B TNPx32 hi/lo)
//   warp 1 / one lane : tcgen05.mma issuer (M=128, N=TNP), commits free ring stages and publish accumulators
//   warps 2..9        : epilogue, two warps per TMEM lane quarter (quarter = warp & 3): they read alternate 32-column
//                       chunks out of TMEM into the quarter's staging slice, meet at a named barrier, and each stores
//                       16 of the 32 rows, four rows in flight at a time (one warp per scheduler with one row in
//                       flight left the store-out exposed: 20 k cycles per tile against 4.7 k of MMAs at K = 196);
//                       the accumulator is handed back as soon as it has been read out
constexpr int PSTAGES = 3;
constexpr int P_EPI_BYTES = TM * EPI_PITCH * 4;          // 128 x 132 fp32 staging (one 128-column half)

template <int TNP>
struct PCfg {
    static constexpr int kBTile = TNP * TK * 2;                     // one B operand tile (hi or lo)
    static constexpr int kStage = 2 * TILE_BYTES + 2 * kBTile;      // A_hi, A_lo, B_hi, B_lo
    static constexpr int kSmem = PSTAGES * kStage + P_EPI_BYTES + 1024;
    static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(TNP >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
};

constexpr int P_THREADS = 320;

template <int TNP>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_split_persistent_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                             const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                             int a_batched, int b_batched, int num_k_blocks, int tiles_m, int tiles_n, int n_batch,
                             EpiArgs ep) {
    using Cfg = PCfg<TNP>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[PSTAGES];
    __shared__ __align__(8) uint64_t empty_bar[PSTAGES];
    __shared__ __align__(8) uint64_t tmem_full[2];
    __shared__ __align__(8) uint64_t tmem_empty[2];
    __shared__ uint32_t tmem_base_slot;

    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (int64_t)tiles_m * tiles_n * n_batch;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
        for (int s = 0; s < PSTAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 8);     // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(&tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int tn = (int)(tile % tiles_n);
            const int64_t r = tile / tiles_n;
            const int tm = (int)(r % tiles_m);
            const int batch = (int)(r / tiles_m);
            const int ab = a_batched ? batch : 0, bb = b_batched ? batch : 0;
            for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                const int s = it % PSTAGES;
                mbar_wait(&empty_bar[s], ((it / PSTAGES) & 1) ^ 1);
                uint8_t* st = smem + s * Cfg::kStage;
                mbar_expect_tx(&full_bar[s], ep.a_lo_zero ? Cfg::kStage - TILE_BYTES : Cfg::kStage);
                tma_load_3d(&map_a_hi, &full_bar[s], st, kb * TK, tm * TM, ab);
                if (!ep.a_lo_zero) tma_load_3d(&map_a_lo, &full_bar[s], st + TILE_BYTES, kb * TK, tm * TM, ab);
                tma_load_3d(&map_b_hi, &full_bar[s], st + 2 * TILE_BYTES, kb * TK, tn * TNP, bb);
                tma_load_3d(&map_b_lo, &full_bar[s], st + 2 * TILE_BYTES + Cfg::kBTile, kb * TK, tn * TNP, bb);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer
        uint32_t it = 0, tcount = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const int acc = tcount & 1;
            mbar_wait(&tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + acc * 256;
            for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                const int s = it % PSTAGES;
                mbar_wait(&full_bar[s], (it / PSTAGES) & 1);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + s * Cfg::kStage);
#pragma unroll
                for (int k = 0; k < TK / 16; ++k) {
                    const uint32_t ko = k * 32;
                    const uint64_t a_hi = smem_desc_sw64(base + ko);
                    const uint64_t a_lo = smem_desc_sw64(base + TILE_BYTES + ko);
                    const uint64_t b_hi = smem_desc_sw64(base + 2 * TILE_BYTES + ko);
                    const uint64_t b_lo = smem_desc_sw64(base + 2 * TILE_BYTES + Cfg::kBTile + ko);
                    if (ep.a_lo_zero) {
                        umma_f16(tmem_acc, a_hi, b_lo, Cfg::kIdesc, (kb | k) ? 1u : 0u);
                    } else {
                        umma_f16(tmem_acc, a_lo, b_hi, Cfg::kIdesc, (kb | k) ? 1u : 0u);
                        umma_f16(tmem_acc, a_hi, b_lo, Cfg::kIdesc, 1u);
                    }
                    umma_f16(tmem_acc, a_hi, b_hi, Cfg::kIdesc, 1u);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&tmem_full[acc]);
        }
    } else if (warp >= 2) {
        // ---------------- epilogue warps
        const int quarter = warp & 3;                     // TMEM lanes 32*quarter .. +31
        const int chalf = (warp - 2) >> 2;                // which of the quarter's two warps
        const uint32_t stage = smem_u32(smem + PSTAGES * Cfg::kStage) + (uint32_t)(quarter * 32 * EPI_PITCH) * 4;
        auto quarter_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory"); };
        const int p = ep.tile_p, zz = p * p;
        const bool pair_ok = (p % 2 == 0);
        const float alpha = ep.alpha * (ep.alpha_dev ? __ldg(ep.alpha_dev) : 1.0f);
        uint32_t tcount = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const int tn = (int)(tile % tiles_n);
            const int64_t r = tile / tiles_n;
            const int tm = (int)(r % tiles_m);
            const int batch = (int)(r / tiles_m);
            const int acc = tcount & 1;
            const int m0 = tm * TM, n0 = tn * TNP;
            const int gm = m0 + quarter * 32 + lane;                  // this thread's accumulator row
            mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + acc * 256 + ((uint32_t)(quarter * 32) << 16);
            float rs = alpha;
            if (ep.row_scale != nullptr && gm < ep.M) rs *= __ldg(ep.row_scale + gm);
            const float dcv = ep.dc_mode ? __ldg(ep.dc + batch) : 0.0f;

            if (ep.mode == 4) {
                // transposed hi/lo store straight from registers: out[b][n * ld + m]; for a fixed column
                // the 32 lanes of the warp hold 32 consecutive m -> one 64-byte store per plane
                const int64_t obase = (int64_t)batch * ep.batch_stride + gm;
                constexpr int kChunks = TNP / 32;
                const int last_c = ((kChunks - 1) & 1) == chalf ? kChunks - 1 : kChunks - 2;
#pragma unroll 1
                for (int c = chalf; c < kChunks; c += 2) {
                    uint32_t rr[32];
                    tmem_ld32(tmem_acc + c * 32, rr);
                    if (c == last_c) {                // this warp's share of the accumulator is read: hand it back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[acc])) : "memory");
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = n0 + c * 32 + j;
                        if (n < ep.N && gm < ep.M) {
                            float v = __uint_as_float(rr[j]) * rs;
                            if (ep.col_scale) v *= __ldg(ep.col_scale + n);
                            const __half h = __float2half_rn(v);
                            const __half l = __float2half_rn(v - __half2float(h));
                            ep.out_hi[obase + (int64_t)n * ep.ld] = h;
                            ep.out_lo[obase + (int64_t)n * ep.ld] = l;
                        }
                    }
                }
                continue;
            }

            // staged modes: two column halves of <= 128 through this warp's 32 x 132 staging slice
#pragma unroll 1
            for (int half = 0; half < (TNP + 127) / 128; ++half) {
                const int cbeg = half * 128;
                const int ccnt = (TNP - cbeg) < 128 ? (TNP - cbeg) : 128;      // 128 or 96
                for (int c = chalf; c < ccnt / 32; c += 2) {
                    uint32_t rr[32];
                    tmem_ld32(tmem_acc + cbeg + c * 32, rr);
                    const uint32_t dst = stage + (uint32_t)(lane * EPI_PITCH + c * 32) * 4;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + j * 4),
                                     "f"(__uint_as_float(rr[j]) * rs), "f"(__uint_as_float(rr[j + 1]) * rs),
                                     "f"(__uint_as_float(rr[j + 2]) * rs), "f"(__uint_as_float(rr[j + 3]) * rs) : "memory");
                }
                if (cbeg + ccnt >= TNP) {             // last half read: hand the accumulator back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[acc])) : "memory");
                }
                quarter_sync();                       // both warps' chunks are in the staging slice
                const int n = n0 + cbeg + lane * 4;
                const int n_valid = (lane * 4 < ccnt) ? ep.N - n : 0;
                int64_t col_off[4] = {0, 0, 0, 0};
                int64_t plane_off;
                if (ep.mode == 2) {
                    const int64_t img = batch / ep.channels;
                    const int ch = batch - (int)img * ep.channels;
                    plane_off = (img * ep.tiles_h * ep.tiles_w * ep.channels + ch) * (int64_t)zz;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int tw = (n + j) / p;
                        col_off[j] = (int64_t)tw * ep.channels * zz + ((n + j) - tw * p);
                    }
                } else {
                    plane_off = (int64_t)batch * ep.batch_stride + n;
                }
                float bias4[4] = {0.f, 0.f, 0.f, 0.f};
                if (ep.col_bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j < n_valid) bias4[j] = __ldg(ep.col_bias + n + j);
                }
                for (int r4 = chalf * 16; r4 < chalf * 16 + 16; r4 += 4) {
                  float4 v4[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u)
                      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v4[u].x), "=f"(v4[u].y), "=f"(v4[u].z), "=f"(v4[u].w)
                                   : "r"(stage + (uint32_t)((r4 + u) * EPI_PITCH + lane * 4) * 4));
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const int m = m0 + quarter * 32 + r4 + u;
                    if (m >= ep.M) break;
                    if (n_valid <= 0 && ep.mode != 5) continue;
                    float4 v = v4[u];
                    if (ep.dc_mode == 2) { v.x += dcv; v.y += dcv; v.z += dcv; v.w += dcv; }
                    else if (ep.dc_mode == 1 && m == 0 && n == 0) v.x += dcv;
                    v.x += bias4[0]; v.y += bias4[1]; v.z += bias4[2]; v.w += bias4[3];
                    if (ep.mode == 5) {
                        // LFQ sign + ballots of one 128-column half (a single N tile: n0 == 0), see gemm_split_kernel
                        const float a[4] = {v.x, v.y, v.z, v.w};
                        __half h[4];
                        uint32_t bal[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const bool pos = j < n_valid && a[j] > 0.0f;
                            h[j] = __float2half_rn(pos ? ep.lfq_scale : -ep.lfq_scale);
                            bal[j] = __ballot_sync(0xffffffffu, pos);
                        }
                        if (lane < 4 && cbeg < ep.N) ep.sign_bits[((int64_t)m * ((ep.N + 127) >> 7) + half) * 4 + lane] = bal[lane];
                        if (n_valid > 0) {
                            const int64_t o = plane_off + (int64_t)m * ep.ld;
                            if (n_valid >= 4) {
                                *reinterpret_cast<uint2*>(ep.out_hi + o) = *reinterpret_cast<const uint2*>(h);
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j) if (j < n_valid) ep.out_hi[o + j] = h[j];
                            }
                        }
                        continue;
                    }
                    if (ep.mode == 0) {
                        float* dst = ep.out_f32 + plane_off + (int64_t)m * ep.ld;
                        if (n_valid >= 4 && (ep.ld & 3) == 0) {
                            *reinterpret_cast<float4*>(dst) = v;
                        } else {
                            const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) if (j < n_valid) dst[j] = a[j];
                        }
                    } else if (ep.mode == 1) {
                        const float a[4] = {v.x, v.y, v.z, v.w};
                        __half h[4], l[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            h[j] = __float2half_rn(a[j]);
                            l[j] = __float2half_rn(a[j] - __half2float(h[j]));
                        }
                        const int64_t o = plane_off + (int64_t)m * ep.ld;
                        if (n_valid >= 4) {
                            *reinterpret_cast<uint2*>(ep.out_hi + o) = *reinterpret_cast<const uint2*>(h);
                            *reinterpret_cast<uint2*>(ep.out_lo + o) = *reinterpret_cast<const uint2*>(l);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) if (j < n_valid) { ep.out_hi[o + j] = h[j]; ep.out_lo[o + j] = l[j]; }
                        }
                    } else {
                        const int th = m / p, pi = m - th * p;
                        float* dst = ep.out_f32 + plane_off + (int64_t)th * ep.tiles_w * ep.channels * zz + pi * p;
                        if (pair_ok && n_valid >= 4) {
                            *reinterpret_cast<float2*>(dst + col_off[0]) = make_float2(v.x, v.y);
                            *reinterpret_cast<float2*>(dst + col_off[2]) = make_float2(v.z, v.w);
                        } else {
                            const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) if (j < n_valid) dst[col_off[j]] = a[j];
                        }
                    }
                  }
                }
                quarter_sync();                       // staging slice is reused by the next half / tile
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int TNP>
static int launch_gemm_persistent(const Operand& A, const Operand& B, int K, int64_t batch, const EpiArgs& ep, void* stream) {
    if (batch == 0 || ep.M == 0 || ep.N == 0) return DCTA_OK;
    if (batch >= (1ll << 31)) { set_error("gemm_split: too many batch items"); return DCTA_ERR_UNSUPPORTED; }
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    const int64_t ab = A.batch_stride ? batch : 1, bb = B.batch_stride ? batch : 1;
    if ((rc = make_map(&ma_hi, A.hi, A.rows, K, A.ld, ab, A.batch_stride, TM))) return rc;
    if ((rc = make_map(&ma_lo, A.lo, A.rows, K, A.ld, ab, A.batch_stride, TM))) return rc;
    if ((rc = make_map(&mb_hi, B.hi, B.rows, K, B.ld, bb, B.batch_stride, TNP))) return rc;
    if ((rc = make_map(&mb_lo, B.lo, B.rows, K, B.ld, bb, B.batch_stride, TNP))) return rc;
    cudaFuncSetAttribute(gemm_split_persistent_kernel<TNP>, cudaFuncAttributeMaxDynamicSharedMemorySize, PCfg<TNP>::kSmem);
    const int tiles_m = (int)ceil_div(ep.M, TM), tiles_n = (int)ceil_div(ep.N, TNP);
    const int64_t n_tiles = (int64_t)tiles_m * tiles_n * batch;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
    gemm_split_persistent_kernel<TNP><<<grid, P_THREADS, PCfg<TNP>::kSmem, as_stream(stream)>>>(
        ma_hi, ma_lo, mb_hi, mb_lo, A.batch_stride != 0, B.batch_stride != 0, (int)ceil_div(K, TK), tiles_m, tiles_n,
        (int)batch, ep);
    return check_launch("gemm_split_persistent");
}

// pick the N tile that wastes the least (224 tiles N = 448 exactly, 256 tiles N = 512 exactly)
static int launch_gemm_auto(const Operand& A, const Operand& B, int K, int64_t batch, const EpiArgs& ep, void* stream) {
    const int64_t w256 = ceil_div(ep.N, 256) * 256, w224 = ceil_div(ep.N, 224) * 224;
    if (w224 < w256) return launch_gemm_persistent<224>(A, B, K, batch, ep, stream);
    return launch_gemm_persistent<256>(A, B, K, batch, ep, stream);
}

// ------------------------------------------------------------------------------ split producers
// The tensor core accumulates in fp32 with truncation, so a large common-mode term (the image mean /
// the DC coefficient) would bias every partial sum the same way.  The producers therefore remove a
// per-plane constant before the split and hand it to the GEMM epilogue, which adds its exact
// contribution back (DCT of a constant is the DC coefficient only).
constexpr int kSumChunks = 32;   // deterministic two-level plane sums: [plane][chunk]

__device__ __forceinline__ void split16(float v, float scale, __half& h, __half& l) {
    const float s = v * scale;
    h = __float2half_rn(s);
    l = __float2half_rn(s - __half2float(h));
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = 0.f;
    if (wid == 0) {
        t = lane < 8 ? red[lane] : 0.f;
        t = warp_sum(t);
    }
    return t;   // valid in warp 0
}

__global__ void __launch_bounds__(256) split_f32_kernel(const float* __restrict__ x, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, int64_t n, float scale) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        split16(x[i], scale, hi[i], lo[i]);
}

// Sub-sampling pattern of the plane-mean estimate: runs of kRun consecutive quads (512 B, so DRAM
// sectors are fully used) out of every `stride` runs.  The k-th sampled quad of a chunk sits at
// (k / kRun) * kRun * stride + k % kRun.
constexpr int kRun = 32;
__host__ __device__ __forceinline__ int64_t sample_index(int64_t k, int stride) {
    return (k / kRun) * kRun * stride + (k % kRun);
}
__device__ __forceinline__ int64_t next_sample(int64_t cur, int step, int stride) {
    // cur is a sampled position; advance `step` samples
    const int64_t k = (cur / (kRun * stride)) * kRun + (cur % (kRun * stride));
    return sample_index(k + step, stride);
}

// sums[plane][chunk] = sum of every `stride`-th element-quad of the chunk (an ESTIMATE of the plane
// sum is enough: whatever constant is removed is added back exactly)
__global__ void __launch_bounds__(256) plane_sums_kernel(const float* __restrict__ x, float* __restrict__ sums,
                                                         int64_t plane4, int stride) {
    __shared__ float red[8];
    const int64_t pl = blockIdx.y;
    const int64_t per = (plane4 + kSumChunks - 1) / kSumChunks;
    const int64_t beg = blockIdx.x * per, end = min(plane4, beg + per);
    const float4* src = reinterpret_cast<const float4*>(x) + pl * plane4;
    float s = 0.f;
    for (int64_t i = beg + sample_index(threadIdx.x, stride); i < end; i = next_sample(i - beg, blockDim.x, stride) + beg) {
        const float4 v = ld_stream(src + i);
        s += (v.x + v.y) + (v.z + v.w);
    }
    s = block_sum_256(s, red);
    if (threadIdx.x == 0) sums[pl * kSumChunks + blockIdx.x] = s;
}

// number of quads plane_sums_kernel visits in a plane (so that mean = sum / (4 * count))
static int64_t sampled_quads(int64_t plane4, int stride) {
    const int64_t per = (plane4 + kSumChunks - 1) / kSumChunks;
    int64_t cnt = 0;
    for (int c = 0; c < kSumChunks; ++c) {
        const int64_t beg = c * per, end = plane4 < beg + per ? plane4 : beg + per;
        const int64_t len = end - beg;
        if (len <= 0) continue;
        const int64_t period = (int64_t)kRun * stride;
        const int64_t rem = len % period;
        cnt += (len / period) * kRun + (rem < kRun ? rem : kRun);
    }
    return cnt;
}

// mu[plane] = (sum of the chunk sums in fixed order) * inv_count ;  dc[plane] = mu * dc_factor
__global__ void finalize_means_kernel(const float* __restrict__ sums, float* __restrict__ mu, float* __restrict__ dc,
                                      int64_t n_planes, float inv_count, float dc_factor) {
    const int64_t pl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pl >= n_planes) return;
    float m = 0.f;
    for (int c = 0; c < kSumChunks; ++c) m += sums[pl * kSumChunks + c];
    m *= inv_count;
    mu[pl] = m;
    dc[pl] = m * dc_factor;
}

__global__ void __launch_bounds__(256) split_centered_kernel(const float* __restrict__ x, const float* __restrict__ mus,
                                                             __half* __restrict__ hi, __half* __restrict__ lo,
                                                             int64_t n_planes, int64_t plane4, float scale) {
    const int64_t total = n_planes * plane4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pl = i / plane4;
        const float mu = __ldg(mus + pl);
        const float4 v = ld_stream(reinterpret_cast<const float4*>(x) + i);
        __half oh[4], ol[4];
        split16(v.x - mu, scale, oh[0], ol[0]);
        split16(v.y - mu, scale, oh[1], ol[1]);
        split16(v.z - mu, scale, oh[2], ol[2]);
        split16(v.w - mu, scale, oh[3], ol[3]);
        reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<const uint2*>(oh);
        reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<const uint2*>(ol);
    }
}

__device__ __forceinline__ float signed_pow_tc(float v, float g) { return signed_pow(v, g); }

// the operand producer uses the accurate exp2 (common.cuh: signed_pow_fwd); the plane-mean ESTIMATE below keeps MUFU
__device__ __forceinline__ void rgb_px_to_ipt_acc(float r, float g, float b, const Mat3& A, const Mat3& B, float& o0,
                                                  float& o1, float& o2) {
    float l = fmaf(A.m[2], b, fmaf(A.m[1], g, A.m[0] * r));
    float m = fmaf(A.m[5], b, fmaf(A.m[4], g, A.m[3] * r));
    float s = fmaf(A.m[8], b, fmaf(A.m[7], g, A.m[6] * r));
    l = signed_pow_fwd(l, 0.43f);
    m = signed_pow_fwd(m, 0.43f);
    s = signed_pow_fwd(s, 0.43f);
    o0 = fmaf(B.m[2], s, fmaf(B.m[1], m, B.m[0] * l));
    o1 = fmaf(B.m[5], s, fmaf(B.m[4], m, B.m[3] * l));
    o2 = fmaf(B.m[8], s, fmaf(B.m[7], m, B.m[6] * l));
}

__device__ __forceinline__ void rgb_px_to_ipt(float r, float g, float b, const Mat3& A, const Mat3& B, float& o0,
                                              float& o1, float& o2) {
    float l = fmaf(A.m[2], b, fmaf(A.m[1], g, A.m[0] * r));
    float m = fmaf(A.m[5], b, fmaf(A.m[4], g, A.m[3] * r));
    float s = fmaf(A.m[8], b, fmaf(A.m[7], g, A.m[6] * r));
    l = signed_pow_tc(l, 0.43f);
    m = signed_pow_tc(m, 0.43f);
    s = signed_pow_tc(s, 0.43f);
    o0 = fmaf(B.m[2], s, fmaf(B.m[1], m, B.m[0] * l));
    o1 = fmaf(B.m[5], s, fmaf(B.m[4], m, B.m[3] * l));
    o2 = fmaf(B.m[8], s, fmaf(B.m[7], m, B.m[6] * l));
}

// IPT plane sums of a subsample of the pixels (every `stride`-th quad): sums[img*3 + c][chunk]
__global__ void __launch_bounds__(256) ipt_sums_kernel(const float* __restrict__ rgb, float* __restrict__ sums,
                                                       int64_t plane4, int stride, Mat3 A, Mat3 B) {
    __shared__ float red[8];
    const int64_t img = blockIdx.y;
    const int64_t per = (plane4 + kSumChunks - 1) / kSumChunks;
    const int64_t beg = blockIdx.x * per, end = min(plane4, beg + per);
    const float4* src = reinterpret_cast<const float4*>(rgb) + img * 3 * plane4;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int64_t i = beg + sample_index(threadIdx.x, stride); i < end; i = next_sample(i - beg, blockDim.x, stride) + beg) {
        const float4 c0 = ld_stream(src + i), c1 = ld_stream(src + plane4 + i), c2 = ld_stream(src + 2 * plane4 + i);
        const float r[4] = {c0.x, c0.y, c0.z, c0.w}, g[4] = {c1.x, c1.y, c1.z, c1.w}, b[4] = {c2.x, c2.y, c2.z, c2.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float o0, o1, o2;
            rgb_px_to_ipt(r[j], g[j], b[j], A, B, o0, o1, o2);
            s0 += o0; s1 += o1; s2 += o2;
        }
    }
    s0 = block_sum_256(s0, red);
    s1 = block_sum_256(s1, red);
    s2 = block_sum_256(s2, red);
    if (threadIdx.x == 0) {
        sums[(img * 3 + 0) * kSumChunks + blockIdx.x] = s0;
        sums[(img * 3 + 1) * kSumChunks + blockIdx.x] = s1;
        sums[(img * 3 + 2) * kSumChunks + blockIdx.x] = s2;
    }
}

// The same estimate in one launch for the folded path: one CTA per image visits the sampled quads of all chunks
// (warp w takes chunks w, w+8, ...), reduces the three plane sums and writes mu / dc itself.  8192 tiny CTAs plus a
// finalising launch cost more than the arithmetic.
template <typename TIn>
__global__ void __launch_bounds__(256) ipt_means_kernel(const TIn* __restrict__ rgb, float* __restrict__ mu,
                                                        float* __restrict__ dc, int64_t plane4, int stride, Mat3 A, Mat3 B,
                                                        float inv_count, float dc_factor) {
    __shared__ float red[8];
    const int64_t img = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = (plane4 + kSumChunks - 1) / kSumChunks;
    const TIn* src = rgb + img * 3 * plane4 * 4;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int c = warp; c < kSumChunks; c += 8) {
        const int64_t beg = c * per, len = min(plane4, beg + per) - beg;
        for (int64_t k = lane;; k += 32) {
            const int64_t i = sample_index(k, stride);
            if (i >= len) break;
            const float4 c0 = ld_px4(src, beg + i), c1 = ld_px4(src, plane4 + beg + i), c2 = ld_px4(src, 2 * plane4 + beg + i);
            const float r[4] = {c0.x, c0.y, c0.z, c0.w}, g[4] = {c1.x, c1.y, c1.z, c1.w}, b[4] = {c2.x, c2.y, c2.z, c2.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float o0, o1, o2;
                rgb_px_to_ipt(r[j], g[j], b[j], A, B, o0, o1, o2);
                s0 += o0; s1 += o1; s2 += o2;
            }
        }
    }
    s0 = block_sum_256(s0, red);
    s1 = block_sum_256(s1, red);
    s2 = block_sum_256(s2, red);
    if (threadIdx.x == 0) {
        const float m[3] = {s0 * inv_count, s1 * inv_count, s2 * inv_count};
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            mu[img * 3 + p] = m[p];
            dc[img * 3 + p] = m[p] * dc_factor;
        }
    }
}

// util.py:70-82 rgb_to_ipt, writing the centred, scaled fp16 hi/lo operand planes of the forward GEMM
__global__ void __launch_bounds__(256) rgb_to_ipt_split_kernel(const float* __restrict__ rgb, const float* __restrict__ mus,
                                                               __half* __restrict__ hi, __half* __restrict__ lo,
                                                               int64_t n_img, int64_t plane4, Mat3 A, Mat3 B,
                                                               float scale) {
    const int64_t total = n_img * plane4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t img = i / plane4, q = i - img * plane4;
        const float mu[3] = {__ldg(mus + img * 3), __ldg(mus + img * 3 + 1), __ldg(mus + img * 3 + 2)};
        const float4* src = reinterpret_cast<const float4*>(rgb) + img * 3 * plane4 + q;
        const float4 c0 = ld_stream(src), c1 = ld_stream(src + plane4), c2 = ld_stream(src + 2 * plane4);
        const float r[4] = {c0.x, c0.y, c0.z, c0.w}, g[4] = {c1.x, c1.y, c1.z, c1.w}, b[4] = {c2.x, c2.y, c2.z, c2.w};
        __half oh[3][4], ol[3][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float o0, o1, o2;
            rgb_px_to_ipt_acc(r[j], g[j], b[j], A, B, o0, o1, o2);
            split16(o0 - mu[0], scale, oh[0][j], ol[0][j]);
            split16(o1 - mu[1], scale, oh[1][j], ol[1][j]);
            split16(o2 - mu[2], scale, oh[2][j], ol[2][j]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int64_t o = (img * 3 + c) * plane4 + q;   // in units of 4 halves (8 bytes)
            reinterpret_cast<uint2*>(hi)[o] = *reinterpret_cast<const uint2*>(oh[c]);
            reinterpret_cast<uint2*>(lo)[o] = *reinterpret_cast<const uint2*>(ol[c]);
        }
    }
}

// feature_extraction_dct_autoencoder.py:635-653 un-patchify, writing scaled fp16 hi/lo planes with pitch `ld`.
// The DC coefficient goes to dc[plane] (already multiplied by dc_factor = 1/sqrt(h*w)) and is stored as 0.
__global__ void __launch_bounds__(128) unpatchify_split_kernel(const float* __restrict__ patches,
                                                               const int32_t* __restrict__ slot_map,
                                                               const int32_t* __restrict__ img_sel, int C, int th,
                                                               int tw, int p, int rows, int cols, int ld,
                                                               __half* __restrict__ hi, __half* __restrict__ lo,
                                                               float* __restrict__ dc, float dc_factor, float scale) {
    // one CTA per tile-row of one plane (p plane rows): block-uniform decomposition, each thread owns
    // 4 consecutive columns (at most two tiles) and walks the p rows
    const int z = p * p;
    const int tile_rows = rows / p;
    const unsigned id = blockIdx.x;
    const int ty = (int)(id % (unsigned)tile_rows);
    const unsigned t = id / (unsigned)tile_rows;
    const int c = (int)(t % (unsigned)C);
    const int sel = (int)(t / (unsigned)C);
    const int64_t img = img_sel ? img_sel[sel] : sel;
    const bool row_in = ty < th;
    const int32_t* smap = slot_map + ((img * C + c) * th + (row_in ? ty : 0)) * tw;
    const int64_t plane_row0 = ((int64_t)(sel * C + c) * rows + (int64_t)ty * p) * ld;
    for (int xv = threadIdx.x; xv < ld / 4; xv += blockDim.x) {
        const int x0 = xv * 4;
        const int tx0 = x0 / p, px0 = x0 - tx0 * p;
        const int n_first = min(4, p - px0);
        const int32_t slot0 = (row_in && tx0 < tw) ? __ldg(smap + tx0) : -1;
        const int32_t slot1 = (row_in && n_first < 4 && tx0 + 1 < tw) ? __ldg(smap + tx0 + 1) : -1;
        const float* s0 = slot0 >= 0 ? patches + (int64_t)slot0 * z + px0 : nullptr;
        const float* s1 = slot1 >= 0 ? patches + (int64_t)slot1 * z - n_first : nullptr;
        for (int py = 0; py < p; ++py) {
            __half oh[4], ol[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* src = j < n_first ? s0 : s1;
                float val = (src && x0 + j < cols) ? __ldg(src + py * p + j) : 0.0f;
                if (ty == 0 && py == 0 && x0 + j == 0) {
                    dc[sel * C + c] = val * dc_factor;
                    val = 0.0f;
                }
                split16(val, scale, oh[j], ol[j]);
            }
            const int64_t o = (plane_row0 + (int64_t)py * ld) / 4 + xv;
            reinterpret_cast<uint2*>(hi)[o] = *reinterpret_cast<const uint2*>(oh);
            reinterpret_cast<uint2*>(lo)[o] = *reinterpret_cast<const uint2*>(ol);
        }
    }
}

}  // namespace dcta

using namespace dcta;

// Scales of the split operands (powers of two, exact): images/IPT 2^8, forward intermediate 2^6,
// coefficient planes 2^4, inverse intermediate 2^6; basis 2^10 (folded into row_scale / alpha).
static const float kScaleX = 256.f, kScaleP = 64.f, kScaleY = 16.f, kScaleQ = 64.f, kScaleBasis = 1024.f;
static const int kSumStride = 64;  // plane-mean estimate from every 64th run of 32 quads: 4096 pixels of a 512^2 plane (any estimate is exact: it is added back)

extern "C" int dcta_gemm_split(const void* a_hi, const void* a_lo, int a_rows, int64_t a_ld, int64_t a_batch_stride,
                               const void* b_hi, const void* b_lo, int b_rows, int64_t b_ld, int64_t b_batch_stride,
                               int k, int64_t batch, const float* row_scale, float alpha, const float* col_bias, float* out,
                               int64_t out_ld, int64_t out_batch_stride, void* stream) {
    DCTA_REQUIRE(a_hi && b_hi && b_lo && out, "gemm_split: null pointer");
    DCTA_REQUIRE(a_rows > 0 && b_rows > 0 && k > 0 && batch >= 0, "gemm_split: bad sizes");
    Operand A{(const __half*)a_hi, (const __half*)(a_lo ? a_lo : a_hi), a_rows, a_ld, a_batch_stride};
    Operand B{(const __half*)b_hi, (const __half*)b_lo, b_rows, b_ld, b_batch_stride};
    EpiArgs ep{};
    ep.mode = 0; ep.out_f32 = out; ep.ld = out_ld; ep.batch_stride = out_batch_stride;
    ep.row_scale = row_scale; ep.alpha = alpha; ep.M = a_rows; ep.N = b_rows;
    ep.col_bias = col_bias; ep.a_lo_zero = a_lo == nullptr;
    // tall linear layer: the persistent kernel (epilogue of one tile behind the main loop of the next; a CTA keeps its
    // N tile when the tile count per row divides the grid, so the weights stay in L2 / shared memory)
    if (batch == 1 && a_rows >= 4096) return launch_gemm_auto(A, B, k, batch, ep, stream);
    return launch_gemm_split(A, B, k, batch, ep, stream);
}

// LFQ with projections in eval (lfq.py:136-227 with has_projections): project_in + bias + sign in the GEMM epilogue.
//   a_hi/a_lo (rows, a_ld): the split rows of the normalised tokens (dcta_split_rows_rowscale), row_scale their factors;
//   w_hi/w_lo (n, w_ld): project_in.weight split; bias (n) [nullable];
//   q_hi (rows, q_ld) fp16: +-codebook_scale, the operand of project_out (its lo plane is zero: see dcta_gemm_split);
//   sign_bits (rows, ceil(n / 128), 4) uint32: word j of a 128-column tile holds column 4*lane + j in bit `lane`.
extern "C" int dcta_lfq_project_sign(const void* a_hi, const void* a_lo, int64_t rows, int64_t a_ld, const void* w_hi,
                                     const void* w_lo, int n, int64_t w_ld, int k, const float* row_scale, const float* bias,
                                     float codebook_scale, void* q_hi, int64_t q_ld, uint32_t* sign_bits, void* stream) {
    DCTA_REQUIRE(a_hi && a_lo && w_hi && w_lo && q_hi && sign_bits, "lfq_project_sign: null pointer");
    DCTA_REQUIRE(rows >= 0 && rows <= 65535ll * TM && n > 0 && k > 0 && q_ld % 8 == 0 && q_ld >= n, "lfq_project_sign: bad sizes");
    if (rows == 0) return DCTA_OK;
    Operand A{(const __half*)a_hi, (const __half*)a_lo, (int)rows, a_ld, 0};
    Operand B{(const __half*)w_hi, (const __half*)w_lo, n, w_ld, 0};
    EpiArgs ep{};
    ep.mode = 5; ep.out_hi = (__half*)q_hi; ep.ld = q_ld; ep.batch_stride = 0;
    ep.row_scale = row_scale; ep.alpha = 1.0f; ep.M = (int)rows; ep.N = n;
    ep.col_bias = bias; ep.sign_bits = sign_bits; ep.lfq_scale = codebook_scale;
    // one N tile: the persistent kernel reads the tokens once and overlaps its epilogue with the next tile
    if (n <= 224) return launch_gemm_persistent<224>(A, B, k, 1, ep, stream);
    if (n <= 256) return launch_gemm_persistent<256>(A, B, k, 1, ep, stream);
    return launch_gemm_split(A, B, k, 1, ep, stream);
}

// sign_bits of dcta_lfq_project_sign -> indices (rows, c) int64, MSB first within a codebook (lfq.py:87, 187)
__global__ void __launch_bounds__(256) lfq_bits_to_codes_kernel(const uint32_t* __restrict__ bits, int64_t rows, int n_tiles,
                                                                int c, int d, int64_t* __restrict__ codes) {
    const int64_t total = rows * c;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / c;
        const int cb = (int)(i - m * c);
        const uint32_t* row = bits + m * n_tiles * 4;
        int64_t code = 0;
        for (int b = 0; b < d; ++b) {
            const int col = cb * d + b, t = col >> 7, r = col & 127;
            code = (code << 1) | ((__ldg(row + t * 4 + (r & 3)) >> (r >> 2)) & 1u);
        }
        codes[i] = code;
    }
}

extern "C" int dcta_lfq_bits_to_codes(const uint32_t* sign_bits, int64_t rows, int n, int c, int d, int64_t* codes, void* stream) {
    DCTA_REQUIRE(sign_bits && codes && c > 0 && d > 0 && d <= 62 && c * d == n && n <= 1024, "lfq_bits_to_codes: bad args");
    if (rows == 0) return DCTA_OK;
    lfq_bits_to_codes_kernel<<<grid_for(rows * c, 256), 256, 0, as_stream(stream)>>>(sign_bits, rows, (int)ceil_div(n, TN), c, d, codes);
    return check_launch("lfq_bits_to_codes");
}

extern "C" int dcta_split_f32(const float* x, void* hi, void* lo, int64_t n, float scale, void* stream) {
    DCTA_REQUIRE(x && hi && lo && n >= 0, "split_f32: bad args");
    if (n == 0) return DCTA_OK;
    split_f32_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x, (__half*)hi, (__half*)lo, n, scale);
    return check_launch("split_f32");
}

extern "C" int dcta_split_planes_centered(const float* x, void* hi, void* lo, float* dc, float* sums_scratch,
                                          int64_t n_planes, int h, int w, void* stream) {
    DCTA_REQUIRE(x && hi && lo && dc && sums_scratch, "split_planes_centered: null pointer");
    const int64_t plane = (int64_t)h * w;
    DCTA_REQUIRE(plane % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && n_planes <= 65535,
                 "split_planes_centered: plane %% 4 != 0, unaligned input or too many planes");
    if (n_planes == 0) return DCTA_OK;
    cudaStream_t st = as_stream(stream);
    plane_sums_kernel<<<dim3(kSumChunks, (unsigned)n_planes), 256, 0, st>>>(x, sums_scratch, plane / 4, kSumStride);
    const float inv_count = 1.0f / (4.0f * (float)sampled_quads(plane / 4, kSumStride));
    float* mus = sums_scratch + n_planes * kSumChunks;      // scratch holds (kSumChunks + 1) floats per plane
    finalize_means_kernel<<<(unsigned)ceil_div(n_planes, 128), 128, 0, st>>>(sums_scratch, mus, dc, n_planes, inv_count,
                                                                             sqrtf((float)h * (float)w));
    split_centered_kernel<<<grid_for(n_planes * (plane / 4), 256), 256, 0, st>>>(
        x, mus, (__half*)hi, (__half*)lo, n_planes, plane / 4, kScaleX);
    return check_launch("split_planes_centered");
}

extern "C" int dcta_rgb_to_ipt_split(const float* rgb, void* ipt_hi, void* ipt_lo, float* dc, float* sums_scratch,
                                     int64_t n_img, int h, int w, const float* m_rgb2lms_host,
                                     const float* m_ipt_host, void* stream) {
    DCTA_REQUIRE(rgb && ipt_hi && ipt_lo && dc && sums_scratch && m_rgb2lms_host && m_ipt_host,
                 "rgb_to_ipt_split: null pointer");
    const int64_t plane = (int64_t)h * w;
    DCTA_REQUIRE(plane % 4 == 0 && (reinterpret_cast<uintptr_t>(rgb) & 15) == 0 && n_img <= 65535,
                 "rgb_to_ipt_split: plane %% 4 != 0, unaligned input or too many images");
    if (n_img == 0 || plane == 0) return DCTA_OK;
    Mat3 A, B;
    for (int i = 0; i < 9; ++i) { A.m[i] = m_rgb2lms_host[i]; B.m[i] = m_ipt_host[i]; }
    cudaStream_t st = as_stream(stream);
    ipt_sums_kernel<<<dim3(kSumChunks, (unsigned)n_img), 256, 0, st>>>(rgb, sums_scratch, plane / 4, kSumStride, A, B);
    const float inv_count = 1.0f / (4.0f * (float)sampled_quads(plane / 4, kSumStride));
    float* mus = sums_scratch + n_img * 3 * kSumChunks;     // scratch holds (kSumChunks + 1) floats per plane
    finalize_means_kernel<<<(unsigned)ceil_div(n_img * 3, 128), 128, 0, st>>>(sums_scratch, mus, dc, n_img * 3, inv_count,
                                                                              sqrtf((float)h * (float)w));
    rgb_to_ipt_split_kernel<<<grid_for(n_img * (plane / 4), 256), 256, 0, st>>>(
        rgb, mus, (__half*)ipt_hi, (__half*)ipt_lo, n_img, plane / 4, A, B, kScaleX);
    return check_launch("rgb_to_ipt_split");
}

// plane means of the IPT image (sampled estimate) + the DC term they stand for; shared with dct_fold.cu.
// scratch holds (kSumChunks + 1) floats per plane; returns the device pointer of the means.
namespace dcta {
template <typename TIn>
static const float* launch_ipt_plane_means_t(const TIn* rgb, float* sums_scratch, float* dc, int64_t n_img, int h, int w,
                                             const Mat3& A, const Mat3& B, cudaStream_t st) {
    const int64_t plane = (int64_t)h * w;
    const float inv_count = 1.0f / (4.0f * (float)sampled_quads(plane / 4, kSumStride));
    float* mus = sums_scratch + n_img * 3 * kSumChunks;
    ipt_means_kernel<TIn><<<(unsigned)n_img, 256, 0, st>>>(rgb, mus, dc, plane / 4, kSumStride, A, B, inv_count,
                                                           sqrtf((float)h * (float)w));
    return mus;
}
const float* launch_ipt_plane_means(const float* rgb, float* sums_scratch, float* dc, int64_t n_img, int h, int w,
                                    const Mat3& A, const Mat3& B, cudaStream_t st) {
    return launch_ipt_plane_means_t(rgb, sums_scratch, dc, n_img, h, w, A, B, st);
}
const float* launch_ipt_plane_means_u8(const uint8_t* rgb, float* sums_scratch, float* dc, int64_t n_img, int h, int w,
                                       const Mat3& A, const Mat3& B, cudaStream_t st) {
    return launch_ipt_plane_means_t(rgb, sums_scratch, dc, n_img, h, w, A, B, st);
}
const float* launch_plane_means(const float* x, float* sums_scratch, float* dc, int64_t n_planes, int h, int w,
                                cudaStream_t st) {
    const int64_t plane = (int64_t)h * w;
    plane_sums_kernel<<<dim3(kSumChunks, (unsigned)n_planes), 256, 0, st>>>(x, sums_scratch, plane / 4, kSumStride);
    const float inv_count = 1.0f / (4.0f * (float)sampled_quads(plane / 4, kSumStride));
    float* mus = sums_scratch + n_planes * kSumChunks;
    finalize_means_kernel<<<(unsigned)ceil_div(n_planes, 128), 128, 0, st>>>(sums_scratch, mus, dc, n_planes, inv_count,
                                                                             sqrtf((float)h * (float)w));
    return mus;
}
}  // namespace dcta

extern "C" int dcta_unpatchify_split(const float* patches, const int32_t* slot_map, const int32_t* img_sel,
                                     int64_t n_img, int channels_n,
                                     int th, int tw, int p, int rows, int cols, int64_t ld, int out_h, int out_w,
                                     void* y_hi, void* y_lo, float* dc, void* stream) {
    DCTA_REQUIRE(patches && slot_map && y_hi && y_lo && dc, "unpatchify_split: null pointer");
    DCTA_REQUIRE(ld % 8 == 0 && ld >= cols && rows > 0 && cols > 0 && p > 0 && out_h > 0 && out_w > 0,
                 "unpatchify_split: bad sizes");
    if (n_img == 0) return DCTA_OK;
    DCTA_REQUIRE(rows % p == 0, "unpatchify_split: plane rows must be a multiple of the patch size");
    const int64_t n_rows_total = n_img * channels_n * (rows / p);
    DCTA_REQUIRE(n_rows_total < (1ll << 31) && ld < (1ll << 30), "unpatchify_split: too many plane rows for one launch");
    unpatchify_split_kernel<<<(unsigned)n_rows_total, 128, 0, as_stream(stream)>>>(
        patches, slot_map, img_sel, channels_n, th, tw, p, rows, cols, (int)ld, (__half*)y_hi, (__half*)y_lo, dc,
        1.0f / sqrtf((float)out_h * (float)out_w), kScaleY);
    return check_launch("unpatchify_split");
}

// forward: centred x planes (n_planes, h, w) as hi/lo (scale 2^8) + their removed DC -> token grid or planes (fp32)
//   basis_w = CW'[:kw] (kw x w, hi/lo, pitch w), basis_h = CH'[:kh] (kh x h, pitch ld_h); rs_w/rs_h their row scales
extern "C" int dcta_dct2_fwd_tc(const void* x_hi, const void* x_lo, const float* dc, const void* bw_hi,
                                const void* bw_lo, const float* rs_w, const void* bh_hi, const void* bh_lo,
                                const float* rs_h, void* work_hi, void* work_lo, float* y, int64_t n_planes, int h,
                                int w, int kh, int kw, int64_t ld_h, int tile_p, int channels, void* stream) {
    DCTA_REQUIRE(x_hi && x_lo && bw_hi && bw_lo && rs_w && bh_hi && bh_lo && rs_h && work_hi && work_lo && y,
                 "dct2_fwd_tc: null pointer");
    DCTA_REQUIRE(w % 8 == 0 && ld_h % 8 == 0 && ld_h >= h && kh <= h && kw <= w && kh > 0 && kw > 0,
                 "dct2_fwd_tc: needs w %% 8 == 0 and an 8-aligned pitch");
    if (tile_p > 0)
        DCTA_REQUIRE(channels > 0 && kh % tile_p == 0 && kw % tile_p == 0 && n_planes % channels == 0,
                     "dct2_fwd_tc: kh/kw must be multiples of the patch size");
    // pass 1: P[h, kw] = sum_w X'[h,w] * CW'[kw,w], written TRANSPOSED as hi/lo planes P^T (kw x ld_h), scale 2^6
    //         (M = h and N = kw tile without waste: 512 = 4 x 128, 448 = 2 x 224)
    Operand A1{(const __half*)x_hi, (const __half*)x_lo, h, w, (int64_t)h * w};
    Operand B1{(const __half*)bw_hi, (const __half*)bw_lo, kw, w, 0};
    EpiArgs e1{};
    e1.mode = 4; e1.out_hi = (__half*)work_hi; e1.out_lo = (__half*)work_lo;
    e1.ld = ld_h; e1.batch_stride = (int64_t)kw * ld_h;
    e1.col_scale = rs_w; e1.alpha = kScaleP / kScaleX; e1.M = h; e1.N = kw;
    int rc = launch_gemm_auto(A1, B1, w, n_planes, e1, stream);
    if (rc) return rc;
    // pass 2: Y[kh, kw] = sum_h CH'[kh,h] * P'^T[kw,h]   (+ the removed constant's DC at [0,0])
    Operand A2{(const __half*)bh_hi, (const __half*)bh_lo, kh, ld_h, 0};
    Operand B2{(const __half*)work_hi, (const __half*)work_lo, kw, ld_h, (int64_t)kw * ld_h};
    EpiArgs e2{};
    e2.out_f32 = y; e2.row_scale = rs_h; e2.alpha = 1.0f / kScaleP; e2.M = kh; e2.N = kw;
    e2.dc = dc; e2.dc_mode = dc ? 1 : 0;
    if (tile_p > 0) {
        e2.mode = 2; e2.tile_p = tile_p; e2.channels = channels; e2.tiles_h = kh / tile_p; e2.tiles_w = kw / tile_p;
    } else {
        e2.mode = 0; e2.ld = kw; e2.batch_stride = (int64_t)kh * kw;
    }
    return launch_gemm_auto(A2, B2, h, n_planes, e2, stream);
}

// inverse: y planes (n_planes, kh, ld_kw) hi/lo (scale 2^4, DC removed into dc[]) -> x (n_planes, h, w) fp32
//   bwt = CW'^T (w x kw, pitch ld_kw), bht = CH'^T (h x kh, pitch ld_kh); work planes (w x ld_kh)
extern "C" int dcta_dct2_inv_tc(const void* y_hi, const void* y_lo, const float* dc, const void* bwt_hi,
                                const void* bwt_lo, const void* bht_hi, const void* bht_lo, void* work_hi,
                                void* work_lo, float* x, int64_t n_planes, int h, int w, int kh, int kw,
                                int64_t ld_kh, int64_t ld_kw, void* stream) {
    DCTA_REQUIRE(y_hi && y_lo && bwt_hi && bwt_lo && bht_hi && bht_lo && work_hi && work_lo && x,
                 "dct2_inv_tc: null pointer");
    DCTA_REQUIRE(ld_kh % 8 == 0 && ld_kw % 8 == 0 && ld_kh >= kh && ld_kw >= kw && kh > 0 && kw > 0,
                 "dct2_inv_tc: pitches must be 8-aligned");
    // pass 1: Q^T[w, kh] = sum_kw CW'^T[w,kw] * Y'[kh,kw]  -> hi/lo planes (w x ld_kh), scale 2^6
    Operand A1{(const __half*)bwt_hi, (const __half*)bwt_lo, w, ld_kw, 0};
    Operand B1{(const __half*)y_hi, (const __half*)y_lo, kh, ld_kw, (int64_t)kh * ld_kw};
    EpiArgs e1{};
    e1.mode = 1; e1.out_hi = (__half*)work_hi; e1.out_lo = (__half*)work_lo;
    e1.ld = ld_kh; e1.batch_stride = (int64_t)w * ld_kh;
    e1.alpha = kScaleQ / (kScaleBasis * kScaleY); e1.M = w; e1.N = kh;
    int rc = launch_gemm_auto(A1, B1, kw, n_planes, e1, stream);
    if (rc) return rc;
    // pass 2: X[h, w] = sum_kh CH'^T[h,kh] * Q'^T[w,kh]   (+ the DC coefficient's constant everywhere)
    Operand A2{(const __half*)bht_hi, (const __half*)bht_lo, h, ld_kh, 0};
    Operand B2{(const __half*)work_hi, (const __half*)work_lo, w, ld_kh, (int64_t)w * ld_kh};
    EpiArgs e2{};
    e2.mode = 0; e2.out_f32 = x; e2.ld = w; e2.batch_stride = (int64_t)h * w;
    e2.alpha = 1.0f / (kScaleBasis * kScaleQ); e2.M = h; e2.N = w;
    e2.dc = dc; e2.dc_mode = dc ? 2 : 0;
    return launch_gemm_auto(A2, B2, kh, n_planes, e2, stream);
}

// fp32 coefficient planes (n_planes, kh, kw) -> split planes (n_planes, kh, ld_kw) with the DC moved to dc[]
namespace dcta {
__global__ void __launch_bounds__(256) split_coef_kernel(const float* __restrict__ y, __half* __restrict__ hi,
                                                         __half* __restrict__ lo, float* __restrict__ dc,
                                                         int64_t n_planes, int kh, int kw, int64_t ld, float dc_factor,
                                                         float scale) {
    const int64_t total = n_planes * kh * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % ld);
        const int64_t r = i / ld;
        const int yy = (int)(r % kh);
        const int64_t pl = r / kh;
        float v = x < kw ? y[(pl * kh + yy) * kw + x] : 0.f;
        if (x == 0 && yy == 0) { dc[pl] = v * dc_factor; v = 0.f; }
        split16(v, scale, hi[i], lo[i]);
    }
}
}  // namespace dcta

extern "C" int dcta_split_coef_planes(const float* y, void* hi, void* lo, float* dc, int64_t n_planes, int kh,
                                      int kw, int64_t ld, int out_h, int out_w, void* stream) {
    DCTA_REQUIRE(y && hi && lo && dc && ld % 8 == 0 && ld >= kw && kh > 0 && kw > 0, "split_coef_planes: bad args");
    if (n_planes == 0) return DCTA_OK;
    split_coef_kernel<<<grid_for(n_planes * kh * ld, 256), 256, 0, as_stream(stream)>>>(
        y, (__half*)hi, (__half*)lo, dc, n_planes, kh, kw, ld, 1.0f / sqrtf((float)out_h * (float)out_w), kScaleY);
    return check_launch("split_coef_planes");
}

// ------------------------------------------------------------------------------ VQ nearest code on tensor cores
// vector_quantize.py:29-33 / :467-469: approximate distances on tcgen05 (one fp16 MMA per product), four candidates per
// token, exact fp32 re-rank (vq_tc.cu).
extern "C" int dcta_vq_nearest_tc_masked(const float* x, const void* x_hi, const float* row_alpha, const float* embed,
                                         const void* e_hi, const float* e2, int32_t* cand, float* cand_val,
                                         const float* e2_max, const uint8_t* keep, int64_t* indices, float* quantized,
                                         int64_t n_tok, int n_codes, int d, int64_t ld, void* stream);

extern "C" int dcta_vq_nearest_tc(const float* x, const void* x_hi, const float* row_alpha, const float* embed,
                                  const void* e_hi, const float* e2, int32_t* cand, int64_t* indices, float* quantized,
                                  int64_t n_tok, int n_codes, int d, int64_t ld, void* stream) {
    return dcta_vq_nearest_tc_masked(x, x_hi, row_alpha, embed, e_hi, e2, cand, nullptr, nullptr, nullptr, indices, quantized,
                                     n_tok, n_codes, d, ld, stream);
}

extern "C" int dcta_vq_nearest_tc_masked(const float* x, const void* x_hi, const float* row_alpha, const float* embed,
                                         const void* e_hi, const float* e2, int32_t* cand, float* cand_val,
                                         const float* e2_max, const uint8_t* keep, int64_t* indices, float* quantized,
                                         int64_t n_tok, int n_codes, int d, int64_t ld, void* stream) {
    DCTA_REQUIRE(x && x_hi && row_alpha && embed && e_hi && e2 && cand && indices, "vq_nearest_tc: null pointer");
    DCTA_REQUIRE((cand_val == nullptr) == (e2_max == nullptr), "vq_nearest_tc: cand_val and e2_max go together");
    DCTA_REQUIRE(n_tok >= 0 && n_codes > 0 && d > 0 && ld >= d && ld % 8 == 0, "vq_nearest_tc: bad sizes");
    if (n_tok == 0) return DCTA_OK;
    int rc = launch_vq_pair(x_hi, e_hi, e2, row_alpha, cand, cand_val, n_tok, n_codes, d, ld, as_stream(stream));
    if (rc == DCTA_ERR_UNSUPPORTED) {
        set_error("vq_nearest_tc: a %d-wide token operand does not fit next to the code ring in shared memory "
                  "(use dcta_vq_nearest)", d);
        return rc;
    }
    if (rc) return rc;
    return launch_vq_rerank(x, embed, e2, cand, cand_val, e2_max, keep, n_tok, n_codes, d, indices, quantized, as_stream(stream));
}
