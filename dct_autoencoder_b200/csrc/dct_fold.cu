// Folded 2-D DCT / IDCT on tensor cores (sm_100a): half the multiply-adds of the plain basis GEMMs.
//
// The DCT-II basis is (anti)symmetric about the centre of the signal, C[k, N-1-n] = (-1)^k C[k, n], so
//   Y[2j]   = sum_{n < N/2} C[2j,   n] (x[n] + x[N-1-n])        Y[2j+1] = sum_{n < N/2} C[2j+1, n] (x[n] - x[N-1-n])
// and the 2-D transform of an H x W plane splits into four independent "quadrant" transforms of
// H/2 x W/2 inputs (row parity a of the coefficient, column parity b):
//   Y[2i+a, 2j+b] = sum_{h', w'} CH[2i+a, h'] Xq[b][a][h', w'] CW[2j+b, w']
// with Xq the four sign combinations of the mirrored pixels.  The inverse runs the same four transforms
// backwards and a final butterfly restores the mirrored pixels.  Reference: util.py:333-338 (dct2 / idct2),
// feature_extraction_dct_autoencoder.py:140, :149, :300-304, :360, :393.
//
// All four GEMM passes have the form  D[r, n] = sum_k Data[r, k] * Basis_g[n, k]  where the data rows of
// many planes are stacked (M tiles of 128 rows never waste lanes), the basis of a parity group g is tiny
// and every output element is written TRANSPOSED (the lanes of a warp hold consecutive r, so a store of
// one column n is one contiguous segment) -- exactly the K-major layout the next pass reads.
//
// Kernel: persistent CTA PAIRS (cta_group::2, M = 256 per tcgen05.mma).
//   * the basis slice (group, N tile) of a pair is loaded ONCE and stays in shared memory, split across the
//     two CTAs (N/2 rows each): L2 -> SM traffic is the data operand only;
//   * warp 0: TMA producer (ring of 16 KB stages: this CTA's 128 x 32 fp16 hi and lo tiles),
//     warp 1 of the leader CTA: tcgen05.mma issuer (3 MMAs per 16-wide k step: lo*hi, hi*lo, hi*hi),
//     warps 2..9: epilogue, two warps per TMEM lane quarter; the 512 TMEM columns hold two accumulators
//     so the epilogue of one tile overlaps the main loop of the next.
#include "lfq_norm.cuh"
#include <algorithm>
#include "tc_ptx.cuh"

namespace dcta {

constexpr int FK = 32;                    // k block: 32 fp16 = 64-byte rows (SWIZZLE_64B)
constexpr int F_ATILE = 128 * FK * 2;     // one 128-row operand tile (hi or lo): 8 KB
constexpr int F_STAGE = 2 * F_ATILE;      // a ring stage: this CTA's A_hi, A_lo
constexpr int F_THREADS = 320;
constexpr int F_MAX_STAGES = 8;
constexpr int F_SMEM_LIMIT = 227 * 1024 - 5120;   // dynamic shared memory we allow ourselves (static: ~2.3 KB)

struct FoldGemm {
    int n_seg;            // data segments (3rd tensor-map coordinate); basis group of a segment = seg & 1
    int rows_per_seg;     // stacked data rows in every segment
    int tiles_per_seg;    // pair tiles (256 rows) per segment
    int num_kb;           // ceil(K / 32)
    int n_tile;           // MMA N: basis rows of one slice (multiple of 16, <= 256)
    int n_ntiles;         // slices per group
    int n_valid;          // basis rows of a group that exist
    int stages;
    uint32_t basis_bytes; // one CTA's resident basis plane (hi or lo): num_kb * (n_tile/2) * 64
    int score_groups;     // > 0: tile rows of the token grid; a [score_groups][128] max|value| image follows the ring
    int chunk_w;          // accumulator columns an epilogue warp takes at a time (multiple of 4, <= 32; an even
                          // number of chunks covers n_tile so that the two warps of a lane quarter get equal shares)
    int lo_streamed;      // 1: only the hi plane of the basis is resident; this CTA's lo tile of the current k block
                          // travels in the ring behind the data (K too long for both planes, e.g. 1024-sample axes)
    uint32_t stage_bytes; // F_STAGE, + the lo tile rounded to 1 KB when lo_streamed
};

struct FoldEpi {
    int mode;                    // 0: fp16 hi/lo, 1: fp32, 2: fp32 token grid (forward pass 2),
                                 // 3: fp16 hi/lo with the output lines regrouped as [channel][token column][image][pj]
                                 //    (forward pass 1 feeding fold_codes_kernel; uses p, channels, tiles_w, batch)
    __half* out_hi;
    __half* out_lo;
    float* out_f32;
    int rows_per_item;           // stacked rows of one batch item; the row inside the item has output stride 1
    int64_t seg_stride;          // output elements between segments
    int64_t item_stride;         // output elements between items
    int col_mul, col_add;        // output line of basis row n of group g: n * col_mul + g * col_add
    int col_stride;              // output elements between lines
    float alpha;
    const float* basis_scale;    // optional (2, n_valid) per-basis-row factors
    const float* dc;             // optional per-item constant added to (group 0, basis row 0, row-in-item 0)
    int p, channels, tiles_h, tiles_w;   // modes 2, 3
    int batch;                           // mode 3: images
    float* maxabs;               // mode 2, optional: (n_img, tiles_h, tiles_w, channels) max |coefficient| of every token,
                                 // accumulated with atomicMax on the bit pattern (the buffer must start at zero)
};

// GEN variant of fold_gemm_kernel (inverse pass 1 straight from LFQ code words): the data operand is not loaded but
// GENERATED in shared memory by four producer warps.  A de-quantised, de-normalised coefficient is one of two values per
// position (median +- scale * sd, lfq.py:118-120 + patchnorm.py:177), so a data tile is a table of value pairs (shared by
// every image) selected by one bit per element.  A pair tile = 8 images x 32 coefficient rows i of one (a, channel):
// each 32-lane TMEM quarter is one image, so the transposed output stores stay 64-byte runs and one 8 KB table block
// serves four images.
struct FoldGen {
    const uint4* tab;       // [b][a][channel][i block][k block][4 parts][128 producer threads]: see decode_gen_tables_kernel
    const uint16_t* bv;     // [b][a][group of 4 images][channel][i block][k block][image][32 rows][4 chunks of 8 columns]:
                            // sign bits | valid bits << 8
    int n_img, C, kh2, n_iblk;
    int n_work;             // work items of a slice: ceil(n_img / 8) * 2 * C * n_iblk
    int64_t n_planes;
};
constexpr int F_GEN_THREADS = 128;     // groups of four producer warps taking alternate stages
constexpr int F_GEN_STAGES = 5;                 // ring of table blocks (8 KB) + the sign/valid words of 4 images (4 x 256 B)
constexpr int F_GEN_STAGE = 8192 + 4 * 256;

// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, M = 256 (pair), N = n_tile
__device__ __forceinline__ uint32_t fold_idesc(int n_tile) {
    return (1u << 4) | ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// Stores at base + 32-bit element offset: one IMAD.WIDE + one STG (the compiler would otherwise rebuild the
// 64-bit address from the kernel parameters for every column).
__device__ __forceinline__ void st_h16(uint64_t base, uint32_t off, __half v) {
    asm volatile("{\n.reg .u64 a;\nmad.wide.u32 a, %1, 2, %0;\nst.global.b16 [a], %2;\n}" ::"l"(base), "r"(off),
                 "h"(__half_as_ushort(v)) : "memory");
}
__device__ __forceinline__ void st_f32(uint64_t base, uint32_t off, float v) {
    asm volatile("{\n.reg .u64 a;\nmad.wide.u32 a, %1, 4, %0;\nst.global.f32 [a], %2;\n}" ::"l"(base), "r"(off), "f"(v)
                 : "memory");
}

struct ScoreCtx {
    const int32_t* col_grp;   // per column: tile-row index, bit 31 set where a run of columns of one tile row ends
    unsigned* smax;           // shared [tile row][128 accumulator rows] running max |value| bit patterns, + this row
};

__device__ __forceinline__ void st_h16_hint(uint64_t base, uint32_t off, __half v, uint64_t pol) {
    asm volatile("{\n.reg .u64 a;\nmad.wide.u32 a, %1, 2, %0;\nst.global.L2::cache_hint.b16 [a], %2, %3;\n}" ::"l"(base),
                 "r"(off), "h"(__half_as_ushort(v)), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_f32_hint(uint64_t base, uint32_t off, float v, uint64_t pol) {
    asm volatile("{\n.reg .u64 a;\nmad.wide.u32 a, %1, 4, %0;\nst.global.L2::cache_hint.f32 [a], %2, %3;\n}" ::"l"(base),
                 "r"(off), "f"(v), "l"(pol) : "memory");
}

// One 32-column chunk of an accumulator row: scale, (split,) store each column at its table offset.
// GUARD: offsets < 0 mark columns past the end of the basis (only the last chunk of a slice can have them).
// SCORE (fp32 token grid): also reduce max |value| per token (feature_extraction_dct_autoencoder.py:409).
// HINT: stores carry the L2 eviction policy `pol`.
template <int MODE, bool GUARD, bool SCORE, bool HINT = false>
__device__ __forceinline__ void store_chunk(const uint32_t (&rr)[32], const int32_t* col_off, const float* col_scale,
                                            uint64_t p_hi, uint64_t p_lo, uint64_t p_f32, float dcv, const ScoreCtx& sc_,
                                            int width, uint64_t pol = 0) {
    float m = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        if (j >= width) break;
        const int4 off = *reinterpret_cast<const int4*>(col_off + j);
        const float4 sc = *reinterpret_cast<const float4*>(col_scale + j);
        float v0 = __uint_as_float(rr[j]) * sc.x, v1 = __uint_as_float(rr[j + 1]) * sc.y;
        float v2 = __uint_as_float(rr[j + 2]) * sc.z, v3 = __uint_as_float(rr[j + 3]) * sc.w;
        if (j == 0) v0 += dcv;
        if (MODE == 0) {
            const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(v0 - f01.x, v1 - f01.y);
            const __half2 l23 = __floats2half2_rn(v2 - f23.x, v3 - f23.y);
            if (HINT) {
                if (!GUARD || off.x >= 0) { st_h16_hint(p_hi, off.x, __low2half(h01), pol); st_h16_hint(p_lo, off.x, __low2half(l01), pol); }
                if (!GUARD || off.y >= 0) { st_h16_hint(p_hi, off.y, __high2half(h01), pol); st_h16_hint(p_lo, off.y, __high2half(l01), pol); }
                if (!GUARD || off.z >= 0) { st_h16_hint(p_hi, off.z, __low2half(h23), pol); st_h16_hint(p_lo, off.z, __low2half(l23), pol); }
                if (!GUARD || off.w >= 0) { st_h16_hint(p_hi, off.w, __high2half(h23), pol); st_h16_hint(p_lo, off.w, __high2half(l23), pol); }
            } else {
                if (!GUARD || off.x >= 0) { st_h16(p_hi, off.x, __low2half(h01)); st_h16(p_lo, off.x, __low2half(l01)); }
                if (!GUARD || off.y >= 0) { st_h16(p_hi, off.y, __high2half(h01)); st_h16(p_lo, off.y, __high2half(l01)); }
                if (!GUARD || off.z >= 0) { st_h16(p_hi, off.z, __low2half(h23)); st_h16(p_lo, off.z, __low2half(l23)); }
                if (!GUARD || off.w >= 0) { st_h16(p_hi, off.w, __high2half(h23)); st_h16(p_lo, off.w, __high2half(l23)); }
            }
        } else if (HINT) {
            if (!GUARD || off.x >= 0) st_f32_hint(p_f32, off.x, v0, pol);
            if (!GUARD || off.y >= 0) st_f32_hint(p_f32, off.y, v1, pol);
            if (!GUARD || off.z >= 0) st_f32_hint(p_f32, off.z, v2, pol);
            if (!GUARD || off.w >= 0) st_f32_hint(p_f32, off.w, v3, pol);
        } else {
            if (!GUARD || off.x >= 0) st_f32(p_f32, off.x, v0);
            if (!GUARD || off.y >= 0) st_f32(p_f32, off.y, v1);
            if (!GUARD || off.z >= 0) st_f32(p_f32, off.z, v2);
            if (!GUARD || off.w >= 0) st_f32(p_f32, off.w, v3);
        }
        if (SCORE) {
            const int4 grp = *reinterpret_cast<const int4*>(sc_.col_grp + j);
            const float vv[4] = {v0, v1, v2, v3};
            const int gg[4] = {grp.x, grp.y, grp.z, grp.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                m = fmaxf(m, fabsf(vv[u]));
                if (gg[u] < 0) {          // warp-uniform: the run of this tile row ends here
                    atomicMax(sc_.smax + (gg[u] & 0x7fffffff) * 128, __float_as_uint(m));   // |v| >= 0 orders like its bits
                    m = 0.0f;
                }
            }
        }
    }
}

template <int MODE, bool GEN = false>   // MODE 0: fp16 hi/lo output, 1: fp32 output (plain or token grid)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F_THREADS + (GEN ? F_GEN_THREADS : 0), 1)
fold_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                 FoldGemm g, FoldEpi ep, FoldGen gen) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[F_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[F_MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full[2];
    __shared__ __align__(8) uint64_t tmem_empty[2];
    __shared__ __align__(8) uint64_t basis_bar;
    __shared__ __align__(8) uint64_t gen_full[F_GEN_STAGES];
    __shared__ __align__(8) uint64_t gen_empty[F_GEN_STAGES];
    __shared__ __align__(8) uint2 sel_lut[GEN ? 256 : 1];    // GEN: sign byte -> the four PRMT selectors of its bit pairs
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) int32_t col_off[288];     // n_tile + one chunk of slack (guarded reads past the end)
    __shared__ __align__(16) float col_scale[288];
    __shared__ __align__(16) int32_t col_grp[288];

    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* basis_hi = smem;
    uint8_t* basis_lo = smem + g.basis_bytes;                                    // unused when lo_streamed
    uint8_t* ring = smem + (g.lo_streamed ? ((g.basis_bytes + 1023u) & ~1023u) : 2 * g.basis_bytes);   // 1 KB aligned
    unsigned* smax = reinterpret_cast<unsigned*>(ring + g.stages * g.stage_bytes);
    uint8_t* gen_ring = ring + g.stages * g.stage_bytes + g.score_groups * 128 * 4;     // GEN only

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_slices = 2 * g.n_ntiles;
    const int slice = pair % n_slices;
    const int grp = slice & 1, nt = slice >> 1;
    const int pair_in_slice = pair / n_slices, pairs_per_slice = n_pairs / n_slices;
    const int n_work = GEN ? gen.n_work : (g.n_seg >> 1) * g.tiles_per_seg;        // work items of this slice
    const int n_lim = min(g.n_tile, g.n_valid - nt * g.n_tile);  // valid columns of this slice

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
        for (int s = 0; s < F_MAX_STAGES; ++s) {
            mbar_init(&full_bar[s], GEN ? 8 : 1);   // GEN: the four producer warps of a group, both CTAs
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 16);    // 8 epilogue warps in each CTA of the pair
        }
        mbar_init(&basis_bar, 1);
        for (int t = 0; t < F_GEN_STAGES; ++t) {
            mbar_init(&gen_full[t], 1);
            mbar_init(&gen_empty[t], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // per-slice output tables: where basis row n goes and what it is multiplied by
    for (int n = threadIdx.x; n < 288; n += blockDim.x) {
        int32_t off = -1;
        float sc = 0.f;
        if (n < n_lim) {
            const int ng = nt * g.n_tile + n;
            const int line = ng * ep.col_mul + grp * ep.col_add;
            if (ep.mode == 2) {
                const int th = line / ep.p, pi = line - th * ep.p;
                off = th * ep.tiles_w * ep.channels * ep.p * ep.p + pi * ep.p;
            } else if (ep.mode == 3) {
                const int tw = line / ep.p, pj = line - tw * ep.p;
                off = (tw * ep.batch * ep.p + pj) * ep.col_stride;
            } else {
                off = line * ep.col_stride;
            }
            sc = ep.alpha * (ep.basis_scale ? __ldg(ep.basis_scale + grp * g.n_valid + ng) : 1.0f);
        }
        col_off[n] = off;
        col_scale[n] = sc;
        int32_t grp_v = 0;
        if (ep.mode == 2 && n < n_lim) {
            const int line = (nt * g.n_tile + n) * ep.col_mul + grp * ep.col_add;
            const int th = line / ep.p;
            const bool ends = (n == n_lim - 1) || ((n % g.chunk_w) == g.chunk_w - 1) || ((line + ep.col_mul) / ep.p != th);
            grp_v = th | (ends ? (int32_t)0x80000000 : 0);
        }
        col_grp[n] = grp_v;
    }
    for (int i = threadIdx.x; i < g.score_groups * 128; i += blockDim.x) smax[i] = 0u;
    if (GEN) {
        // selector of a pair of elements: bytes 0-3 = the "bit is 1" word, 4-7 = the "bit is 0" word; low half <- first bit
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            uint32_t sel[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) sel[m] = (((((uint32_t)i >> (2 * m)) & 3u) * 0x2244u) & 0x4444u) ^ 0x7654u;
            sel_lut[i] = make_uint2(sel[0] | (sel[1] << 16), sel[2] | (sel[3] << 16));
        }
    }
    if (warp == 2) tmem_alloc_2sm(&tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer (both CTAs; completion bytes are credited to the leader's barriers)
        const uint32_t basis_bar_leader = mapa_u32(smem_u32(&basis_bar), 0);
        const int half_rows = g.n_tile >> 1;
        const uint32_t btile = (uint32_t)half_rows * 64;
        const int brow = nt * g.n_tile + (int)rank * half_rows;
        if (rank == 0) mbar_expect_tx(&basis_bar, (g.lo_streamed ? 2 : 4) * g.basis_bytes);
        for (int kb = 0; kb < g.num_kb; ++kb) {
            tma_load_3d_2sm(&map_b_hi, basis_bar_leader, basis_hi + kb * btile, kb * FK, brow, grp);
            if (!g.lo_streamed) tma_load_3d_2sm(&map_b_lo, basis_bar_leader, basis_lo + kb * btile, kb * FK, brow, grp);
        }
        const uint32_t full_leader0 = mapa_u32(smem_u32(&full_bar[0]), 0);
        int s = 0;
        uint32_t ph = 1;                                 // parity of a free slot
        if (GEN) {
            // table blocks and sign/valid words of this CTA's four images, one ring stage per k block
            int t = 0;
            uint32_t tph = 1;
            const uint32_t gen_s = smem_u32(gen_ring);
            const int64_t n_grp4 = (gen.n_img + 3) >> 2;
            for (int w = pair_in_slice; w < n_work; w += pairs_per_slice) {
                const int iblk = w % gen.n_iblk;
                int u = w / gen.n_iblk;
                const int ch = u % gen.C;
                u /= gen.C;
                const int a = u & 1, g4 = (u >> 1) * 2 + (int)rank;
                const uint4* tp = gen.tab + ((((int64_t)(grp * 2 + a) * gen.C + ch) * gen.n_iblk + iblk) * g.num_kb) * 512;
                const uint16_t* bp = gen.bv + (((((int64_t)(grp * 2 + a) * n_grp4 + g4) * gen.C + ch) * gen.n_iblk + iblk) * g.num_kb) * 512;
                const bool has_img = g4 < n_grp4;
                for (int kb = 0; kb < g.num_kb; ++kb, tp += 512, bp += 512) {
                    mbar_wait(&gen_empty[t], tph);
                    const uint32_t dst = gen_s + (uint32_t)t * F_GEN_STAGE;
                    mbar_expect_tx(&gen_full[t], has_img ? 8192u + 1024u : 8192u);
                    bulk_copy_g2s(dst, tp, 8192u, &gen_full[t]);
                    if (has_img) bulk_copy_g2s(dst + 8192u, bp, 1024u, &gen_full[t]);
                    if (++t == F_GEN_STAGES) { t = 0; tph ^= 1u; }
                }
            }
        }
        for (int w = pair_in_slice; !GEN && w < n_work; w += pairs_per_slice) {
            const int seg = (w / g.tiles_per_seg) * 2 + grp;
            const int row0 = (w % g.tiles_per_seg) * 256 + (int)rank * 128;
            for (int kb = 0; kb < g.num_kb; ++kb) {
                mbar_wait(&empty_bar[s], ph);
                uint8_t* st = ring + s * g.stage_bytes;
                const uint32_t full_leader = full_leader0 + 8u * s;
                if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * F_STAGE + (g.lo_streamed ? 2 * btile : 0u));
                tma_load_3d_2sm(&map_a_hi, full_leader, st, kb * FK, row0, seg);
                tma_load_3d_2sm(&map_a_lo, full_leader, st + F_ATILE, kb * FK, row0, seg);
                if (g.lo_streamed) tma_load_3d_2sm(&map_b_lo, full_leader, st + F_STAGE, kb * FK, brow, grp);
                if (++s == g.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ---------------- MMA issuer (leader CTA only): one thread, so the loop is kept to a wait, one add per
        // descriptor and the MMAs (no division for the ring slot, no descriptor rebuild)
        const uint32_t idesc = fold_idesc(g.n_tile);
        const uint32_t btile16 = ((uint32_t)(g.n_tile >> 1) * 64) >> 4;
        const uint32_t bh = smem_desc_lo(smem_u32(basis_hi)), bl = smem_desc_lo(smem_u32(basis_lo));
        const uint32_t ring16 = smem_desc_lo(smem_u32(ring)), stage16 = g.stage_bytes >> 4;
        mbar_wait_cluster(&basis_bar, 0);
        tc_fence_after();
        uint32_t tcount = 0, ph = 0;
        int s = 0;
        for (int w = pair_in_slice; w < n_work; w += pairs_per_slice, ++tcount) {
            const int acc = tcount & 1;
            mbar_wait_cluster(&tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + acc * 256;
            uint32_t b16 = 0;
            for (int kb = 0; kb < g.num_kb; ++kb, b16 += btile16) {
                mbar_wait_cluster(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a16 = ring16 + (uint32_t)s * stage16;
                const uint32_t bl16 = g.lo_streamed ? a16 + (F_STAGE >> 4) : bl + b16;
#pragma unroll
                for (int k = 0; k < FK / 16; ++k) {
                    const uint64_t a_hi = smem_desc_sw64_from_lo(a16 + 2 * k);
                    const uint64_t a_lo = smem_desc_sw64_from_lo(a16 + (F_ATILE >> 4) + 2 * k);
                    const uint64_t b_hi = smem_desc_sw64_from_lo(bh + b16 + 2 * k);
                    const uint64_t b_lo = smem_desc_sw64_from_lo(bl16 + 2 * k);
                    umma_f16_2sm(tmem_acc, a_lo, b_hi, idesc, (kb | k) ? 1u : 0u);   // small terms first
                    umma_f16_2sm(tmem_acc, a_hi, b_lo, idesc, 1u);
                    umma_f16_2sm(tmem_acc, a_hi, b_hi, idesc, 1u);
                }
                umma_commit_2sm(&empty_bar[s], 3);       // frees the stage in both CTAs
                if (++s == g.stages) { s = 0; ph ^= 1u; }
            }
            umma_commit_2sm(&tmem_full[acc], 3);         // accumulator complete in both CTAs
        }
    } else if (warp >= 2 && warp < F_THREADS / 32) {
        // ---------------- epilogue warps: TMEM lane quarter = warp & 3, 32-column chunks alternate between
        // the two warps of a quarter
        const int quarter = warp & 3, chalf = (warp - 2) >> 2;
        const int cwid = g.chunk_w;
        const int n_chunks = (n_lim + cwid - 1) / cwid;
        const uint32_t tmem_empty_leader0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
        const uint32_t tmem_empty_leader1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
        const bool dc_slice = (ep.dc != nullptr) && grp == 0 && nt == 0;
        uint32_t tcount = 0;
        for (int w = pair_in_slice; w < n_work; w += pairs_per_slice, ++tcount) {
            int seg, r, item, rin;
            bool row_ok;
            if (GEN) {
                // w = ((image group * 2 + a) * C + channel) * n_iblk + i block; TMEM quarter = image of the group
                const int iblk = w % gen.n_iblk;
                int t = w / gen.n_iblk;
                const int ch = t % gen.C;
                t /= gen.C;
                const int img = (t >> 1) * 8 + (int)rank * 4 + quarter;
                seg = grp;
                r = 0;
                rin = iblk * 32 + lane;
                item = (t & 1) * (int)gen.n_planes + img * gen.C + ch;
                row_ok = img < gen.n_img && rin < gen.kh2;
            } else {
                seg = (w / g.tiles_per_seg) * 2 + grp;
                r = (w % g.tiles_per_seg) * 256 + (int)rank * 128 + quarter * 32 + lane;   // stacked row
                row_ok = r < g.rows_per_seg;
                item = r / ep.rows_per_item;
                rin = r - item * ep.rows_per_item;
            }
            int64_t base;
            ScoreCtx sctx{};
            if (ep.mode == 2) {
                const int img = item / ep.channels, ch = item - img * ep.channels;
                const int tw = rin / ep.p, pj = rin - tw * ep.p;
                const int64_t tok0 = ((int64_t)img * ep.tiles_h * ep.tiles_w + tw) * ep.channels + ch;   // token at tile row 0
                base = tok0 * (ep.p * ep.p) + pj;
                sctx.smax = smax + quarter * 32 + lane;
            } else if (ep.mode == 3) {
                // item = a * planes + plane: the line block of (a, channel) starts at ((a*C + ch) * tiles_w * batch) * p,
                // inside it image img adds img * p lines (the token column adds tw * batch * p + pj: col_off)
                const int planes = ep.batch * ep.channels;
                const int a = item / planes, plane = item - a * planes;
                const int img = plane / ep.channels, ch = plane - img * ep.channels;
                base = (((int64_t)(a * ep.channels + ch) * ep.tiles_w * ep.batch + img) * ep.p) * ep.col_stride + rin;
            } else {
                base = (int64_t)seg * ep.seg_stride + (int64_t)item * ep.item_stride + rin;
            }
            const float dcv = (dc_slice && row_ok && rin == 0) ? __ldg(ep.dc + item) : 0.0f;
            const uint64_t p_hi = reinterpret_cast<uint64_t>(ep.out_hi + base);
            const uint64_t p_lo = reinterpret_cast<uint64_t>(ep.out_lo + base);
            const uint64_t p_f32 = reinterpret_cast<uint64_t>(ep.out_f32 + base);
            const int acc = tcount & 1;
            mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + acc * 256 + ((uint32_t)(quarter * 32) << 16);
            int last = n_chunks - 1;
            if ((last & 1) != chalf) --last;                 // this warp's last chunk (may be < first: none)
            if (last < chalf) {                              // nothing to read: release immediately
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(acc ? tmem_empty_leader1 : tmem_empty_leader0);
            }
            // software-pipelined read-out: the next chunk's tcgen05.ld is in flight while this one is stored
            auto process = [&](const uint32_t (&rr)[32], int c) {
                const float dcc = c == 0 ? dcv : 0.0f;
                if (MODE == 1 && g.score_groups > 0) {
                    if (row_ok) {
                        ScoreCtx sc2 = sctx;
                        sc2.col_grp = &col_grp[c * cwid];
                        store_chunk<MODE, true, true>(rr, &col_off[c * cwid], &col_scale[c * cwid], p_hi, p_lo, p_f32, dcc, sc2, cwid);
                    }
                    return;
                }
                if (!row_ok) return;
                if (c * cwid + cwid <= n_lim) store_chunk<MODE, false, false>(rr, &col_off[c * cwid], &col_scale[c * cwid], p_hi, p_lo, p_f32, dcc, sctx, cwid);
                else store_chunk<MODE, true, false>(rr, &col_off[c * cwid], &col_scale[c * cwid], p_hi, p_lo, p_f32, dcc, sctx, cwid);
            };
            auto hand_back = [&]() {                         // accumulator read out by this warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(acc ? tmem_empty_leader1 : tmem_empty_leader0);
            };
            uint32_t ra[32], rb[32];
            int c = chalf;
            if (c < n_chunks) tmem_ld32_nowait(tmem_acc + c * cwid, ra);
#pragma unroll 1
            while (c < n_chunks) {
                tmem_ld_wait();
                if (c + 2 < n_chunks) tmem_ld32_nowait(tmem_acc + (c + 2) * cwid, rb);
                if (c == last) hand_back();
                process(ra, c);
                c += 2;
                if (c >= n_chunks) break;
                tmem_ld_wait();
                if (c + 2 < n_chunks) tmem_ld32_nowait(tmem_acc + (c + 2) * cwid, ra);
                if (c == last) hand_back();
                process(rb, c);
                c += 2;
            }
            if (MODE == 1 && g.score_groups > 0) {
                // max |value| of every token touched by this CTA's 128 accumulator rows: the rows of a token are
                // p consecutive stacked rows; one thread per (tile row, token column)
                asm volatile("bar.sync 1, 256;" ::: "memory");          // the 8 epilogue warps
                const int R0 = (w % g.tiles_per_seg) * 256 + (int)rank * 128;
                const int Rend = min(R0 + 128, g.rows_per_seg);          // exclusive
                if (Rend > R0) {
                    const int s_first = R0 / ep.p, n_cols = (Rend - 1) / ep.p - s_first + 1;
                    for (int i = (int)threadIdx.x - 64; i < g.score_groups * n_cols; i += 256) {
                        const int th = i / n_cols, sc = s_first + (i - th * n_cols);
                        const int lo = max(R0, sc * ep.p) - R0, hi = min(Rend, sc * ep.p + ep.p) - R0;
                        unsigned m = 0;
                        for (int rr_ = lo; rr_ < hi; ++rr_) {
                            m = max(m, smax[th * 128 + rr_]);
                            smax[th * 128 + rr_] = 0u;
                        }
                        if (m != 0u) {
                            const int plane = sc / ep.tiles_w, tw = sc - plane * ep.tiles_w;
                            const int img = plane / ep.channels, ch = plane - img * ep.channels;
                            atomicMax(reinterpret_cast<unsigned*>(ep.maxabs) +
                                          (((int64_t)img * ep.tiles_h + th) * ep.tiles_w + tw) * ep.channels + ch, m);
                        }
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
        }
    } else if (GEN && warp >= F_THREADS / 32) {
        // ---------------- GEN: producer warps.  Thread pt = (row i of the 32-row block, 16-byte chunk c of the 64-byte
        // operand row) handles that (i, c) for the four images of this CTA: the table values are loaded once
        // (4 x 16 bytes: the hi / lo halves of the two candidate values of 8 columns, already paired for PRMT), and per
        // image one byte of sign bits picks the halves -- two PRMT per pair of elements, two 16-byte shared stores per
        // image row.  The table block and the sign words arrive through a small ring filled by warp 0 with bulk copies
        // (loads issued by these warps themselves would be drained by the release fence of every hand-over: measured
        // 2x slower).
        const int pt = (int)threadIdx.x - F_THREADS;
        const int il = pt >> 2, c = pt & 3;
        const uint32_t sw_off = (uint32_t)il * 64u + (uint32_t)((c ^ ((il >> 1) & 3)) << 4);     // SWIZZLE_64B
        const uint32_t ring_s = smem_u32(ring), gen_s = smem_u32(gen_ring);
        const uint32_t full_leader0 = mapa_u32(smem_u32(&full_bar[0]), 0);
        const int n_stages_total = ((n_work - pair_in_slice + pairs_per_slice - 1) / pairs_per_slice) * g.num_kb;
        int s = 0, t = 0;
        uint32_t ph = 1, tph = 0;
        // block of ring stage t -> registers (table words, sign/valid words, selectors), then the stage is handed back
        auto fetch = [&](int stage_no, uint4 (&T)[4], uint32_t (&bvv)[4], uint2 (&sls)[4]) {
            const int w = pair_in_slice + (stage_no / g.num_kb) * pairs_per_slice;
            const int img0 = ((w / (gen.n_iblk * gen.C)) >> 1) * 8 + (int)rank * 4;
            const int n_here = max(0, min(4, gen.n_img - img0));
            mbar_wait(&gen_full[t], tph);
            const uint32_t src = gen_s + (uint32_t)t * F_GEN_STAGE;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(T[k].x), "=r"(T[k].y), "=r"(T[k].z), "=r"(T[k].w)
                             : "r"(src + (uint32_t)k * 2048u + (uint32_t)pt * 16u));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint16_t v;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(src + 8192u + (uint32_t)q * 256u + (uint32_t)pt * 2u));
                bvv[q] = q < n_here ? (uint32_t)v : 0u;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) sls[q] = sel_lut[bvv[q] & 0xffu];
            __syncwarp();
            if (lane == 0) mbar_arrive_local(&gen_empty[t]);          // the block is in registers
            if (++t == F_GEN_STAGES) { t = 0; tph ^= 1u; }
        };
        auto emit = [&](const uint4 (&T)[4], const uint32_t (&bvv)[4], const uint2 (&sls)[4]) {
            mbar_wait(&empty_bar[s], ph);
            const uint32_t st_hi = ring_s + (uint32_t)s * F_STAGE + sw_off, st_lo = st_hi + F_ATILE;
            const uint32_t ph_[4] = {T[0].x, T[0].z, T[1].x, T[1].z}, nh_[4] = {T[0].y, T[0].w, T[1].y, T[1].w};
            const uint32_t pl_[4] = {T[2].x, T[2].z, T[3].x, T[3].z}, nl_[4] = {T[2].y, T[2].w, T[3].y, T[3].w};
            const bool all_valid = (bvv[0] & bvv[1] & bvv[2] & bvv[3]) >= 0xff00u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint2 sl = sls[q];                         // PRMT reads the low 16 bits of its selector
                const uint32_t sel[4] = {sl.x, sl.x >> 16, sl.y, sl.y >> 16};
                uint32_t wh[4], wl[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    wh[m] = __byte_perm(ph_[m], nh_[m], sel[m]);
                    wl[m] = __byte_perm(pl_[m], nl_[m], sel[m]);
                }
                if (!all_valid) {                        // tokens that were not kept decode to zero coefficients
                    const uint32_t valid = bvv[q] >> 8;
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const uint32_t vm = (((valid >> (2 * m)) & 1u) ? 0xffffu : 0u) | (((valid >> (2 * m + 1)) & 1u) ? 0xffff0000u : 0u);
                        wh[m] &= vm;
                        wl[m] &= vm;
                    }
                }
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_hi + (uint32_t)q * 2048u), "r"(wh[0]), "r"(wh[1]),
                             "r"(wh[2]), "r"(wh[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_lo + (uint32_t)q * 2048u), "r"(wl[0]), "r"(wl[1]),
                             "r"(wl[2]), "r"(wl[3]) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> the MMA's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive_remote_cta_release(full_leader0 + 8u * (uint32_t)s);
            if (++s == g.stages) { s = 0; ph ^= 1u; }
        };
        // software pipeline: the next stage's block is being read while this one is written
        uint4 Ta[4], Tb[4];
        uint32_t ba[4], bb[4];
        uint2 sa[4], sb[4];
        if (n_stages_total > 0) fetch(0, Ta, ba, sa);
#pragma unroll 1
        for (int n = 0; n < n_stages_total; n += 2) {
            if (n + 1 < n_stages_total) fetch(n + 1, Tb, bb, sb);
            emit(Ta, ba, sa);
            if (n + 1 >= n_stages_total) break;
            if (n + 2 < n_stages_total) fetch(n + 2, Ta, ba, sa);
            emit(Tb, bb, sb);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

// ------------------------------------------------------------------------------ forward pass 2 -> code words
// The forward pass 2 with the operand roles swapped: M = basis rows (the coefficient rows kh = 2i + a of one
// parity group, 128 per CTA of the pair, resident in shared memory), N = data rows (kw of a token-aligned run of
// `tokens_per_tile` token columns, streamed by TMA).  A thread then owns one coefficient row (tile row th, row in
// tile pi) and sees the p columns of a token as p CONSECUTIVE accumulator columns: the sign bits of
// clamp((Y - median) / (b*sqrt2 + eps)) (patchnorm.py:157-165 + lfq.py:175-187, one LFQ codebook per patch row)
// are packed into the code word in registers and stored once -- no ballots, no shared image, no token grid.
// The per-token max |Y| (feature_extraction_dct_autoencoder.py:409) is reduced over the p/2 lanes of a tile row
// with shuffles and one atomicMax per tile row, token and parity.
struct CodesArgs {
    int rows_per_seg;       // stacked data rows (planes * kw) in each of the 2 segments (= parity groups)
    int tiles_per_seg;      // pair tiles per segment
    int num_kb, stages;
    int n_valid;            // basis rows per group (kh / 2)
    int n_data;             // data rows per pair tile = tokens_per_tile * p  (MMA N, multiple of 16, <= 256)
    int tokens_per_tile;
    uint32_t basis_bytes;   // one CTA's resident basis plane (hi or lo): num_kb * 128 * 64
    uint32_t stage_bytes;   // one ring stage: this CTA's n_data/2 data rows, hi then lo (+ the basis lo tile, lo_streamed)
    int lo_streamed;        // 1: only the hi plane of the basis is resident, the lo tile of each k block rides in the ring
                            // at offset lo_off of the stage (K too long for both planes: 1024-sample axes)
    uint32_t lo_off;
    int p, channels, tiles_h, tiles_w;
    int batch;                  // images: the data rows are ordered [channel][token column tw][image][pj]
    float alpha;
    const float* basis_scale;   // (2, n_valid)
    const float* dc;            // per plane constant added to coefficient (0, 0)
    const float* med;           // PatchNorm tables (channels, stat_h, stat_w, p*p)
    const float* bstat;
    int stat_h, stat_w;
    float eps, clamp_lo, clamp_hi;
    const int32_t* tame;
    float* maxabs;              // (n_img, tiles_h, tiles_w, channels), zeroed
    int32_t* code_grid;         // (n_img, tiles_h, tiles_w, channels, p)
};

template <int N>
__device__ __forceinline__ float tree_max(float (&x)[N]) {
#pragma unroll
    for (int s = 1; s < N; s *= 2)
#pragma unroll
        for (int j = 0; j + s < N; j += 2 * s) x[j] = fmaxf(x[j], x[j + s]);
    return x[0];
}
template <int N>
__device__ __forceinline__ float tree_min(float (&x)[N]) {
#pragma unroll
    for (int s = 1; s < N; s *= 2)
#pragma unroll
        for (int j = 0; j + s < N; j += 2 * s) x[j] = fminf(x[j], x[j + s]);
    return x[0];
}

// Epilogue of fold_codes_kernel for tile width P (8 warps per CTA: lane quarter = warp & 3 picks the coefficient
// rows, the two warps of a quarter take half of the tile's token columns each).  The epilogue warps are only two
// per scheduler, so the code is written for instruction-level parallelism: two tokens are read out of TMEM and
// processed together, the per-token maxima / minima are reduction trees and the code word is assembled as two
// half chains, instead of three 14-step dependent chains per token.
template <int P>
__device__ __forceinline__ void codes_epilogue(const CodesArgs& g, uint32_t tmem_base, uint64_t* tmem_full,
                                               uint32_t tmem_empty_leader0, uint32_t tmem_empty_leader1, int warp,
                                               int lane, uint32_t rank, int grp, int pair_in_grp, int pairs_per_grp) {
    constexpr int Z = P * P, P2 = P / 2;
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int i = (int)rank * 128 + quarter * 32 + lane;         // basis row of this thread
    const bool i_ok = i < g.n_valid;
    const int kh = 2 * i + grp;
    const int th = kh / P, pi = kh - th * P;
    const float scale = i_ok ? g.alpha * (g.basis_scale ? __ldg(g.basis_scale + grp * g.n_valid + i) : 1.0f) : 0.0f;
    const bool tame = g.tame != nullptr && __ldg(g.tame) != 0 && g.eps > 1e-12f && g.clamp_lo < 0.0f && g.clamp_hi > 0.0f;
    const LfqNormParams q{nullptr, nullptr, 0, 0, 0, 0, g.eps, g.clamp_lo, g.clamp_hi, 0, 0, 0.f};
    // lanes of the same tile row th are contiguous: the first of each run publishes the row's maximum
    const int th_prev = __shfl_up_sync(0xffffffffu, th, 1);
    const bool th_leader = i_ok && (lane == 0 || th_prev != th);
    bool take[3];                                             // does lane + 1 / 2 / 4 belong to the same tile row?
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        const int oth = __shfl_down_sync(0xffffffffu, th, 1 << o);
        take[o] = (lane + (1 << o) < 32) && oth == th;
    }
    const int t_lo = half * ((g.tokens_per_tile + 1) >> 1);
    const int t_hi = half ? g.tokens_per_tile : ((g.tokens_per_tile + 1) >> 1);
    const int64_t stat_row = (int64_t)th * g.stat_w * Z + pi * P;        // + (ch * stat_h * stat_w + tw) * Z
    const int64_t stat_ch = (int64_t)g.stat_h * g.stat_w * Z;
    const int64_t tok_row = (int64_t)th * g.tiles_w * g.channels;        // token = (img*tiles_h*tiles_w + tw) * C + ch + tok_row
    const int64_t tok_img = (int64_t)g.tiles_h * g.tiles_w * g.channels; // token index step from one image to the next
    const bool dc_lane = kh == 0 && g.dc != nullptr;
    uint32_t tcount = 0;
    for (int w = pair_in_grp; w < g.tiles_per_seg; w += pairs_per_grp, ++tcount) {
        const int acc = tcount & 1;
        // token of this warp's first data column: the data rows run [channel][tw][image][pj]
        int tc = w * g.tokens_per_tile + t_lo;
        int img = tc % g.batch;
        int tw = (tc / g.batch) % g.tiles_w, ch = tc / (g.batch * g.tiles_w);
        mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc * 256 + ((uint32_t)(quarter * 32) << 16);
        // the medians of this coefficient row depend on (channel, tw) only: one load per run of images
        float mv[P];
        int m_tw = -1, m_ch = -1;
        int tk = t_lo;
        if (t_hi <= t_lo) {                                  // nothing to read: release immediately
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(acc ? tmem_empty_leader1 : tmem_empty_leader0);
        }
        while (tk < t_hi) {
            // a pair of tokens shares its medians unless the image index wraps between them (warp-uniform)
            const bool two = tk + 1 < t_hi && img + 1 < g.batch;
            uint32_t rr[2][16];
            tmem_ld16_nowait(tmem_acc + tk * P, rr[0]);
            if (two) tmem_ld16_nowait(tmem_acc + (tk + 1) * P, rr[1]);
            const bool ok[2] = {(int64_t)tc * P < g.rows_per_seg, two && (int64_t)(tc + 1) * P < g.rows_per_seg};
            if (ok[0] && (tw != m_tw || ch != m_ch)) {
                const float2* src = reinterpret_cast<const float2*>(g.med + ch * stat_ch + (int64_t)tw * Z + stat_row);
#pragma unroll
                for (int j = 0; j < P2; ++j) {
                    const float2 t = i_ok ? __ldg(src + j) : make_float2(0.f, 0.f);
                    mv[2 * j] = t.x;
                    mv[2 * j + 1] = t.y;
                }
                m_tw = tw;
                m_ch = ch;
            }
            float dcv[2] = {0.0f, 0.0f};
            if (dc_lane && tw == 0) {
                if (ok[0]) dcv[0] = __ldg(g.dc + img * g.channels + ch);
                if (ok[1]) dcv[1] = __ldg(g.dc + (img + 1) * g.channels + ch);
            }
            tmem_ld_wait();
            if (tk + (two ? 2 : 1) >= t_hi) {                  // accumulator read out by this warp: hand it back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(acc ? tmem_empty_leader1 : tmem_empty_leader0);
            }
            if (!two) {
#pragma unroll
                for (int j = 0; j < 16; ++j) rr[1][j] = rr[0][j];
            }
            const int64_t tok0 = ((int64_t)img * g.tiles_h * g.tiles_w + tw) * g.channels + ch + tok_row;
            unsigned word[2];
            float mx[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                // fast path: bit = sign of (median - Y) (Y > median <=> the difference is negative), valid when the
                // divisor is tame and no |difference| is tiny; funnel shifts build the word MSB first
                float v[P], nd[P], av[P], ad[P];
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    v[j] = __uint_as_float(rr[u][j]) * scale;
                    if (j == 0) v[j] += dcv[u];
                    nd[j] = __fsub_rn(mv[j], v[j]);
                    av[j] = fabsf(v[j]);
                    ad[j] = fabsf(nd[j]);
                }
                unsigned wa = 0, wb = 0;
#pragma unroll
                for (int j = 0; j < P2; ++j) {
                    wa = __funnelshift_l(__float_as_uint(nd[j]), wa, 1);
                    wb = __funnelshift_l(__float_as_uint(nd[P2 + j]), wb, 1);
                }
                word[u] = (wa << P2) | wb;
                const float dmin = tree_min<P>(ad);
                mx[u] = (ok[u] && i_ok) ? tree_max<P>(av) : 0.0f;
                if (ok[u] && i_ok && (!tame || !(dmin > 1e-20f))) {          // rare: exact sign of the clamped quotient
                    const float* bsrc = g.bstat + ch * stat_ch + (int64_t)tw * Z + stat_row;
                    unsigned wd = 0;
#pragma unroll
                    for (int j = 0; j < P; ++j) wd = (wd << 1) | norm_sign_bit(v[j], mv[j], __ldg(bsrc + j), q);
                    word[u] = wd;
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (ok[u] && i_ok) g.code_grid[(tok0 + u * tok_img) * P + pi] = (int32_t)word[u];
            // max over the lanes of one tile row (runs of P/2 <= 8 lanes)
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                const float o0 = __shfl_down_sync(0xffffffffu, mx[0], 1 << o);
                const float o1 = __shfl_down_sync(0xffffffffu, mx[1], 1 << o);
                if (take[o]) {
                    mx[0] = fmaxf(mx[0], o0);
                    mx[1] = fmaxf(mx[1], o1);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (ok[u] && th_leader && mx[u] > 0.0f)
                    atomicMax(reinterpret_cast<unsigned*>(g.maxabs) + tok0 + u * tok_img, __float_as_uint(mx[u]));
            const int n = two ? 2 : 1;
            tk += n;
            tc += n;
            img += n;
            if (img >= g.batch) { img = 0; if (++tw == g.tiles_w) { tw = 0; ++ch; } }
        }
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F_THREADS, 1)
fold_codes_kernel(const __grid_constant__ CUtensorMap map_d_hi, const __grid_constant__ CUtensorMap map_d_lo,
                  const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                  CodesArgs g) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[F_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[F_MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full[2];
    __shared__ __align__(8) uint64_t tmem_empty[2];
    __shared__ __align__(8) uint64_t basis_bar;
    __shared__ uint32_t tmem_base_slot;

    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* basis_hi = smem;
    uint8_t* basis_lo = smem + g.basis_bytes;                          // unused when lo_streamed
    uint8_t* ring = smem + (g.lo_streamed ? 1 : 2) * g.basis_bytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int grp = pair & 1;
    const int pair_in_grp = pair >> 1, pairs_per_grp = n_pairs >> 1;
    const int half_data = g.n_data >> 1;
    const uint32_t dtile = (uint32_t)half_data * 64;          // one data operand tile (hi or lo) of a stage

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_d_hi);
        tma_prefetch_desc(&map_d_lo);
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
        for (int s = 0; s < F_MAX_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 16);
        }
        mbar_init(&basis_bar, 1);
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
        const uint32_t basis_bar_leader = mapa_u32(smem_u32(&basis_bar), 0);
        if (rank == 0) mbar_expect_tx(&basis_bar, (g.lo_streamed ? 2 : 4) * g.basis_bytes);
        for (int kb = 0; kb < g.num_kb; ++kb) {          // this CTA's 128 basis rows (rows past n_valid are zero-filled)
            tma_load_3d_2sm(&map_b_hi, basis_bar_leader, basis_hi + kb * F_ATILE, kb * FK, (int)rank * 128, grp);
            if (!g.lo_streamed) tma_load_3d_2sm(&map_b_lo, basis_bar_leader, basis_lo + kb * F_ATILE, kb * FK, (int)rank * 128, grp);
        }
        const uint32_t full_leader0 = mapa_u32(smem_u32(&full_bar[0]), 0);
        int s = 0;
        uint32_t ph = 1;                                 // parity of a free slot
        for (int w = pair_in_grp; w < g.tiles_per_seg; w += pairs_per_grp) {
            const int row0 = w * g.n_data + (int)rank * half_data;
            for (int kb = 0; kb < g.num_kb; ++kb) {
                mbar_wait(&empty_bar[s], ph);
                uint8_t* st = ring + s * g.stage_bytes;
                const uint32_t full_leader = full_leader0 + 8u * s;
                if (rank == 0) mbar_expect_tx(&full_bar[s], 4 * dtile + (g.lo_streamed ? 2u * F_ATILE : 0u));
                tma_load_3d_2sm(&map_d_hi, full_leader, st, kb * FK, row0, grp);
                tma_load_3d_2sm(&map_d_lo, full_leader, st + dtile, kb * FK, row0, grp);
                if (g.lo_streamed) tma_load_3d_2sm(&map_b_lo, full_leader, st + g.lo_off, kb * FK, (int)rank * 128, grp);
                if (++s == g.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ---------------- MMA issuer: A = resident basis (M = 256 over the pair), B = data stage (N = n_data);
        // a single thread, kept to a wait, one add per descriptor and the MMAs (see fold_gemm_kernel)
        const uint32_t idesc = fold_idesc(g.n_data);
        const uint32_t bh = smem_desc_lo(smem_u32(basis_hi)), bl = smem_desc_lo(smem_u32(basis_lo));
        const uint32_t ring16 = smem_desc_lo(smem_u32(ring)), stage16 = g.stage_bytes >> 4, dtile16 = dtile >> 4;
        mbar_wait_cluster(&basis_bar, 0);
        tc_fence_after();
        uint32_t tcount = 0, ph = 0;
        int s = 0;
        for (int w = pair_in_grp; w < g.tiles_per_seg; w += pairs_per_grp, ++tcount) {
            const int acc = tcount & 1;
            mbar_wait_cluster(&tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + acc * 256;
            uint32_t a16 = 0;
            for (int kb = 0; kb < g.num_kb; ++kb, a16 += F_ATILE >> 4) {
                mbar_wait_cluster(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t d16 = ring16 + (uint32_t)s * stage16;
                const uint32_t al16 = g.lo_streamed ? d16 + (g.lo_off >> 4) : bl + a16;
#pragma unroll
                for (int k = 0; k < FK / 16; ++k) {
                    const uint64_t a_hi = smem_desc_sw64_from_lo(bh + a16 + 2 * k);
                    const uint64_t a_lo = smem_desc_sw64_from_lo(al16 + 2 * k);
                    const uint64_t b_hi = smem_desc_sw64_from_lo(d16 + 2 * k);
                    const uint64_t b_lo = smem_desc_sw64_from_lo(d16 + dtile16 + 2 * k);
                    // the same sequence of partial sums as fold_gemm_kernel (data_lo*basis_hi, data_hi*basis_lo,
                    // data_hi*basis_hi) so that both kernels produce bit-identical coefficients
                    umma_f16_2sm(tmem_acc, a_hi, b_lo, idesc, (kb | k) ? 1u : 0u);
                    umma_f16_2sm(tmem_acc, a_lo, b_hi, idesc, 1u);
                    umma_f16_2sm(tmem_acc, a_hi, b_hi, idesc, 1u);
                }
                umma_commit_2sm(&empty_bar[s], 3);
                if (++s == g.stages) { s = 0; ph ^= 1u; }
            }
            umma_commit_2sm(&tmem_full[acc], 3);
        }
    } else if (warp >= 2) {
        // ---------------- epilogue (codes_epilogue<P> above), one instantiation per supported tile width
        const uint32_t e0 = mapa_u32(smem_u32(&tmem_empty[0]), 0), e1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
        switch (g.p) {
            case 8: codes_epilogue<8>(g, tmem_base, tmem_full, e0, e1, warp, lane, rank, grp, pair_in_grp, pairs_per_grp); break;
            case 10: codes_epilogue<10>(g, tmem_base, tmem_full, e0, e1, warp, lane, rank, grp, pair_in_grp, pairs_per_grp); break;
            case 12: codes_epilogue<12>(g, tmem_base, tmem_full, e0, e1, warp, lane, rank, grp, pair_in_grp, pairs_per_grp); break;
            case 14: codes_epilogue<14>(g, tmem_base, tmem_full, e0, e1, warp, lane, rank, grp, pair_in_grp, pairs_per_grp); break;
            default: codes_epilogue<16>(g, tmem_base, tmem_full, e0, e1, warp, lane, rank, grp, pair_in_grp, pairs_per_grp); break;
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

// ------------------------------------------------------------------------------ host: launch
// 3-D fp16 tensor map (k, rows, segments) with a (32, box_rows, 1) box, SWIZZLE_64B
static int make_map3(CUtensorMap* map, const void* ptr, int64_t k, int64_t rows, int64_t segs, int64_t ld,
                     int64_t seg_stride, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DCTA_ERR_UNSUPPORTED; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8) || (segs > 1 && seg_stride % 8)) {
        set_error("fold_gemm: operand planes need a 16-byte aligned base, pitch and segment stride");
        return DCTA_ERR_INVALID_ARG;
    }
    cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)(segs > 0 ? segs : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(segs > 1 ? seg_stride : rows * ld) * 2};
    cuuint32_t box[3] = {FK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", (int)r); return DCTA_ERR_LAUNCH; }
    return DCTA_OK;
}

struct FoldOperand {
    const __half* hi;
    const __half* lo;
    int64_t ld;           // pitch (elements) of a row
    int64_t seg_stride;   // elements between segments (data) / groups (basis)
};

// slice geometry for a basis of n_valid rows and K columns: (n_tile, n_ntiles, stages), or false if nothing fits
static bool fold_geometry(int n_valid, int K, FoldGemm& g, int score_groups = 0, int extra_static = 0) {
    g.num_kb = (int)ceil_div(K, FK);
    g.n_valid = n_valid;
    g.score_groups = score_groups;
    const int64_t score_bytes = (int64_t)score_groups * 128 * 4;
    for (int nn = (int)ceil_div(n_valid, 256); nn <= 16; ++nn) {
        const int n_tile = (int)ceil_div(ceil_div(n_valid, nn), 16) * 16;
        const int64_t basis = (int64_t)g.num_kb * (n_tile / 2) * 64;
        int64_t stages = (F_SMEM_LIMIT - extra_static - 1024 - 2 * basis - score_bytes) / F_STAGE;
        g.lo_streamed = 0;
        g.stage_bytes = F_STAGE;
        if (n_tile <= 256 && stages < 3 && extra_static == 0) {
            // Both planes of the basis do not fit beside the ring.  Halving the slice (nn + 1) would stream every data
            // tile twice and run MMAs of half the width, which are bound by their shared-memory operand reads; keeping
            // the hi plane resident and sending the CTA's lo tile of each k block through the ring costs
            // n_tile/2 x 64 bytes per stage instead (it never leaves L2).
            const int64_t lo_tile = ceil_div((int64_t)(n_tile / 2) * 64, 1024) * 1024;
            const int64_t st2 = (F_SMEM_LIMIT - 1024 - ceil_div(basis, 1024) * 1024 - score_bytes) / (F_STAGE + lo_tile);
            if (st2 >= 3) {
                stages = st2;
                g.lo_streamed = 1;
                g.stage_bytes = (uint32_t)(F_STAGE + lo_tile);
            }
        }
        if (n_tile <= 256 && stages >= 3) {
            g.n_tile = n_tile;
            g.n_ntiles = nn;
            g.basis_bytes = (uint32_t)basis;
            g.stages = (int)(stages < F_MAX_STAGES ? stages : F_MAX_STAGES);
            const int pairs_of_chunks = (int)ceil_div(n_tile, 64);
            g.chunk_w = (int)ceil_div(ceil_div(n_tile, 2 * pairs_of_chunks), 4) * 4;
            return 4 * basis < (1 << 20);      // mbarrier transaction-count range
        }
    }
    return false;
}

static int launch_fold_gemm(const FoldOperand& A, int64_t rows_per_seg, int n_seg, const FoldOperand& Bas, int n_valid,
                            int K, FoldEpi ep, void* stream, const FoldGen* gen = nullptr) {
    if (rows_per_seg == 0 || n_seg == 0) return DCTA_OK;
    FoldGemm g{};
    const int score_groups = (ep.mode == 2 && ep.maxabs != nullptr) ? ep.tiles_h : 0;
    const int gen_bytes = gen ? F_GEN_STAGES * F_GEN_STAGE : 0;
    if (!fold_geometry(n_valid, K, g, score_groups, gen_bytes)) { set_error("fold_gemm: basis %d x %d does not fit in shared memory", n_valid, K); return DCTA_ERR_UNSUPPORTED; }
    if (rows_per_seg >= (1ll << 31) - 256 || (n_seg & 1)) { set_error("fold_gemm: bad segment geometry"); return DCTA_ERR_INVALID_ARG; }
    g.n_seg = n_seg;
    g.rows_per_seg = (int)rows_per_seg;
    g.tiles_per_seg = (int)ceil_div(rows_per_seg, 256);
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    if ((rc = make_map3(&mb_hi, Bas.hi, K, n_valid, 2, Bas.ld, Bas.seg_stride, g.n_tile / 2))) return rc;
    if ((rc = make_map3(&mb_lo, Bas.lo, K, n_valid, 2, Bas.ld, Bas.seg_stride, g.n_tile / 2))) return rc;
    if (gen) {                                    // the data operand is generated in shared memory: no tensor map
        ma_hi = mb_hi;
        ma_lo = mb_lo;
    } else {
        if ((rc = make_map3(&ma_hi, A.hi, K, rows_per_seg, n_seg, A.ld, A.seg_stride, 128))) return rc;
        if ((rc = make_map3(&ma_lo, A.lo, K, rows_per_seg, n_seg, A.ld, A.seg_stride, 128))) return rc;
    }
    const int smem_bytes = 1024 + (g.lo_streamed ? (int)(ceil_div((int64_t)g.basis_bytes, 1024) * 1024) : 2 * (int)g.basis_bytes) +
                           g.stages * (int)g.stage_bytes +
                           g.score_groups * 128 * 4 + gen_bytes;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_slices = 2 * g.n_ntiles;
    const int64_t work_per_slice = gen ? gen->n_work : (int64_t)(n_seg / 2) * g.tiles_per_seg;
    int64_t pps = (sms / 2) / n_slices;                      // pairs per slice
    if (pps < 1) pps = 1;
    if (pps > work_per_slice) pps = work_per_slice;
    const unsigned grid = (unsigned)(2 * pps * n_slices);
    const FoldGen no_gen{};
    cudaError_t e;
    if (gen) {
        if (ep.mode != 0) { set_error("fold_gemm: generated operands feed the fp16 hi/lo output only"); return DCTA_ERR_INVALID_ARG; }
        e = cudaFuncSetAttribute(fold_gemm_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e == cudaSuccess)
            fold_gemm_kernel<0, true><<<grid, F_THREADS + F_GEN_THREADS, smem_bytes, as_stream(stream)>>>(ma_hi, ma_lo, mb_hi, mb_lo, g, ep, *gen);
    } else if (ep.mode == 0 || ep.mode == 3) {          // fp16 hi/lo outputs
        e = cudaFuncSetAttribute(fold_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e == cudaSuccess)
            fold_gemm_kernel<0><<<grid, F_THREADS, smem_bytes, as_stream(stream)>>>(ma_hi, ma_lo, mb_hi, mb_lo, g, ep, no_gen);
    } else {
        e = cudaFuncSetAttribute(fold_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e == cudaSuccess)
            fold_gemm_kernel<1><<<grid, F_THREADS, smem_bytes, as_stream(stream)>>>(ma_hi, ma_lo, mb_hi, mb_lo, g, ep, no_gen);
    }
    if (e != cudaSuccess) { set_error("fold_gemm: %s", cudaGetErrorString(e)); return DCTA_ERR_LAUNCH; }
    return check_launch("fold_gemm");
}

// forward pass 2 straight to code words (fold_codes_kernel); DCTA_ERR_UNSUPPORTED when the geometry does not fit
static int launch_fold_codes(const FoldOperand& Data, int64_t rows_per_seg, const FoldOperand& Bas, int n_valid, int K,
                             CodesArgs g, void* stream) {
    if (rows_per_seg == 0) return DCTA_OK;
    const int p = g.p;
    if (n_valid > 256 || p < 8 || p > 16 || (p & 1) || rows_per_seg % p) return DCTA_ERR_UNSUPPORTED;
    // token-aligned data tiles: the largest multiple of p that is a multiple of 16 and at most 256 rows
    int tokens = 256 / p;
    while (tokens > 0 && (tokens * p) % 16) --tokens;
    if (tokens < 2) return DCTA_ERR_UNSUPPORTED;
    g.tokens_per_tile = tokens;
    g.n_data = tokens * p;
    g.rows_per_seg = (int)rows_per_seg;
    g.tiles_per_seg = (int)ceil_div(rows_per_seg, g.n_data);
    g.n_valid = n_valid;
    g.num_kb = (int)ceil_div(K, FK);
    g.basis_bytes = (uint32_t)g.num_kb * F_ATILE;
    g.stage_bytes = (uint32_t)ceil_div((int64_t)g.n_data * 64, 1024) * 1024;      // hi + lo halves of n_data/2 rows each
    int64_t stages = (F_SMEM_LIMIT - 1024 - 2 * (int64_t)g.basis_bytes) / g.stage_bytes;
    g.lo_streamed = 0;
    g.lo_off = 0;
    if (stages < 3) {             // long K: hi plane resident, lo tiles through the ring (see fold_geometry)
        g.lo_streamed = 1;
        g.lo_off = g.stage_bytes;
        g.stage_bytes += F_ATILE;
        stages = (F_SMEM_LIMIT - 1024 - (int64_t)g.basis_bytes) / g.stage_bytes;
    }
    if (stages < 3 || 4ll * g.basis_bytes >= (1 << 20) || rows_per_seg >= (1ll << 31) - 256) return DCTA_ERR_UNSUPPORTED;
    g.stages = (int)(stages < F_MAX_STAGES ? stages : F_MAX_STAGES);
    CUtensorMap md_hi, md_lo, mb_hi, mb_lo;
    int rc;
    if ((rc = make_map3(&md_hi, Data.hi, K, rows_per_seg, 2, Data.ld, Data.seg_stride, g.n_data / 2))) return rc;
    if ((rc = make_map3(&md_lo, Data.lo, K, rows_per_seg, 2, Data.ld, Data.seg_stride, g.n_data / 2))) return rc;
    if ((rc = make_map3(&mb_hi, Bas.hi, K, n_valid, 2, Bas.ld, Bas.seg_stride, 128))) return rc;
    if ((rc = make_map3(&mb_lo, Bas.lo, K, n_valid, 2, Bas.ld, Bas.seg_stride, 128))) return rc;
    const int smem_bytes = 1024 + (g.lo_streamed ? 1 : 2) * (int)g.basis_bytes + g.stages * (int)g.stage_bytes;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t ppg = (sms / 2) / 2;                     // pairs per parity group
    if (ppg < 1) ppg = 1;
    if (ppg > g.tiles_per_seg) ppg = g.tiles_per_seg;
    cudaError_t e = cudaFuncSetAttribute(fold_codes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { set_error("fold_codes: %s", cudaGetErrorString(e)); return DCTA_ERR_LAUNCH; }
    fold_codes_kernel<<<(unsigned)(4 * ppg), F_THREADS, smem_bytes, as_stream(stream)>>>(md_hi, md_lo, mb_hi, mb_lo, g);
    return check_launch("fold_codes");
}

// ------------------------------------------------------------------------------ fold / unfold kernels
__device__ __forceinline__ void split16x4(const float (&v)[4], float scale, uint2& hi, uint2& lo) {
    const float s0 = v[0] * scale, s1 = v[1] * scale, s2 = v[2] * scale, s3 = v[3] * scale;
    const __half2 h01 = __floats2half2_rn(s0, s1), h23 = __floats2half2_rn(s2, s3);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(s0 - f01.x, s1 - f01.y), l23 = __floats2half2_rn(s2 - f23.x, s3 - f23.y);
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

__device__ __forceinline__ void rgb_px_to_ipt_f(float r, float g, float b, const Mat3& A, const Mat3& B, float& o0,
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

// the four sign combinations of the mirrored samples p1 = x[h', w'], p2 = x[h', W-1-w'], p3 = x[H-1-h', w'],
// p4 = x[H-1-h', W-1-w']:  q[b*2 + a],  a = row parity of the coefficient, b = column parity
__device__ __forceinline__ void butterfly4(float p1, float p2, float p3, float p4, float (&q)[4]) {
    const float s12 = p1 + p2, d12 = p1 - p2, s34 = p3 + p4, d34 = p3 - p4;
    q[0] = s12 + s34;   // b = 0, a = 0
    q[1] = s12 - s34;   // b = 0, a = 1
    q[2] = d12 + d34;   // b = 1, a = 0
    q[3] = d12 - d34;   // b = 1, a = 1
}

// fp16 hi/lo halves of four values that already carry their scale
__device__ __forceinline__ void split16x4_scaled(const float (&v)[4], uint2& hi, uint2& lo) {
    const __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(v[0] - f01.x, v[1] - f01.y), l23 = __floats2half2_rn(v[2] - f23.x, v[3] - f23.y);
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

// rgb_px_to_ipt_f with the operand scale and the centring folded in: Bs = scale * B (a power of two: exact) and
// nmu[c] = -scale * mu[c] enter the last matrix product, so no instruction is spent on either.
__device__ __forceinline__ void rgb_px_to_ipt_centred(float r, float g, float b, const Mat3& A, const Mat3& Bs,
                                                      const float (&nmu)[3], float& o0, float& o1, float& o2) {
    float l = fmaf(A.m[2], b, fmaf(A.m[1], g, A.m[0] * r));
    float m = fmaf(A.m[5], b, fmaf(A.m[4], g, A.m[3] * r));
    float s = fmaf(A.m[8], b, fmaf(A.m[7], g, A.m[6] * r));
    l = signed_pow_fwd<false>(l, 0.43f);
    m = signed_pow_fwd<false>(m, 0.43f);
    s = signed_pow_fwd<false>(s, 0.43f);
    o0 = fmaf(Bs.m[2], s, fmaf(Bs.m[1], m, fmaf(Bs.m[0], l, nmu[0])));
    o1 = fmaf(Bs.m[5], s, fmaf(Bs.m[4], m, fmaf(Bs.m[3], l, nmu[1])));
    o2 = fmaf(Bs.m[8], s, fmaf(Bs.m[7], m, fmaf(Bs.m[6], l, nmu[2])));
}

// util.py:70-82 rgb_to_ipt fused with centring, the 2-D fold and the fp16 hi/lo split:
// xq[b][a][plane][h'][w'], plane = img * 3 + c.  One thread = 4 consecutive w' of one (img, h').
// The kernel is bound by instruction issue before HBM (48 pow per thread), so everything that is not colour math is
// kept off the per-pixel path: scale and centring live in the matrix (Bs, nmu), offsets inside an image are 32-bit.
template <typename TIn>       // float in [0, 1], or uint8 (read as u8 / 255)
__global__ void __launch_bounds__(256, 4) rgb_to_ipt_fold_kernel(const TIn* __restrict__ rgb, const float* __restrict__ mus,
                                                              __half* __restrict__ hi, __half* __restrict__ lo,
                                                              int64_t n_img, int h, int w, Mat3 A, Mat3 Bs, float scale) {
    const int h2 = h >> 1, w8 = w >> 3, w4 = w >> 2;    // w8: float4 groups of the half row
    const uint32_t plane4 = (uint32_t)h * w4, q4 = (uint32_t)h2 * w8;             // in float4 / uint2 units
    const int64_t quad = n_img * 3 * (int64_t)q4;                                 // one quadrant of all planes
    // grid: (blocks per image, images) -- one item per thread, 32-bit index arithmetic inside an image
    const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item < q4) {
        const int y = (int)(item / (uint32_t)w8);
        const int xv = (int)(item - (uint32_t)y * w8);
        const int64_t img = blockIdx.y;
        const TIn* src = rgb + img * 3 * (int64_t)plane4 * 4;
        const float nmu[3] = {-scale * __ldg(mus + img * 3), -scale * __ldg(mus + img * 3 + 1), -scale * __ldg(mus + img * 3 + 2)};
        const uint32_t top = (uint32_t)y * w4, bot = (uint32_t)(h - 1 - y) * w4;
        const uint32_t xl = xv, xr = w4 - 1 - xv;
        float ipt[4][3][4];          // [corner: top-left, top-right, bottom-left, bottom-right][channel][pixel]
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t o = ((k & 2) ? bot : top) + ((k & 1) ? xr : xl);
            const float4 c0 = ld_px4(src, o), c1 = ld_px4(src, plane4 + o), c2 = ld_px4(src, 2 * plane4 + o);
            const float r[4] = {c0.x, c0.y, c0.z, c0.w}, gg[4] = {c1.x, c1.y, c1.z, c1.w}, bb[4] = {c2.x, c2.y, c2.z, c2.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                rgb_px_to_ipt_centred(r[j], gg[j], bb[j], A, Bs, nmu, ipt[k][0][j], ipt[k][1][j], ipt[k][2][j]);
        }
        uint2* ph = reinterpret_cast<uint2*>(hi) + ((img * 3 * h2 + y) * (int64_t)w8 + xv);
        uint2* pl = reinterpret_cast<uint2*>(lo) + ((img * 3 * h2 + y) * (int64_t)w8 + xv);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float q[4][4];           // [quadrant][pixel]
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float qq[4];
                butterfly4(ipt[0][c][j], ipt[1][c][3 - j], ipt[2][c][j], ipt[3][c][3 - j], qq);
#pragma unroll
                for (int s = 0; s < 4; ++s) q[s][j] = qq[s];
            }
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                uint2 vh, vl;
                split16x4_scaled(q[s], vh, vl);
                ph[s * quad + c * q4] = vh;
                pl[s * quad + c * q4] = vl;
            }
        }
    }
}

// fp32 planes -> centred, folded, split quadrants (same layout, any number of planes)
__global__ void __launch_bounds__(256) fold_planes_kernel(const float* __restrict__ x, const float* __restrict__ mus,
                                                          __half* __restrict__ hi, __half* __restrict__ lo,
                                                          int64_t n_planes, int h, int w, float scale) {
    const int h2 = h >> 1, w8 = w >> 3;
    const int64_t total = n_planes * h2 * w8;
    const int64_t plane4 = (int64_t)h * w / 4, q4 = (int64_t)h2 * (w >> 1) / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xv = (int)(i % w8);
        const int64_t t = i / w8;
        const int y = (int)(t % h2);
        const int64_t pl = t / h2;
        const float4* src = reinterpret_cast<const float4*>(x) + pl * plane4;
        const int64_t top = (int64_t)y * (w >> 2), bot = (int64_t)(h - 1 - y) * (w >> 2);
        const int xr = (w >> 2) - 1 - xv;
        const float4 a4 = ld_stream(src + top + xv), b4 = ld_stream(src + top + xr);
        const float4 c4 = ld_stream(src + bot + xv), d4 = ld_stream(src + bot + xr);
        const float mu = __ldg(mus + pl);
        const float p1[4] = {a4.x, a4.y, a4.z, a4.w}, p2[4] = {b4.w, b4.z, b4.y, b4.x};
        const float p3[4] = {c4.x, c4.y, c4.z, c4.w}, p4[4] = {d4.w, d4.z, d4.y, d4.x};
        float q[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float qq[4];
            butterfly4(p1[j] - mu, p2[j] - mu, p3[j] - mu, p4[j] - mu, qq);
#pragma unroll
            for (int s = 0; s < 4; ++s) q[s][j] = qq[s];
        }
        const int64_t o = (pl * h2 + y) * (int64_t)w8 + xv;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            uint2 vh, vl;
            split16x4(q[s], scale, vh, vl);
            reinterpret_cast<uint2*>(hi)[s * n_planes * q4 + o] = vh;
            reinterpret_cast<uint2*>(lo)[s * n_planes * q4 + o] = vl;
        }
    }
}

__device__ __forceinline__ void ipt_px_to_rgb_f(float i0, float i1, float i2, const Mat3& A, const Mat3& B, float& o0,
                                                float& o1, float& o2) {
    float l = fmaf(A.m[2], i2, fmaf(A.m[1], i1, A.m[0] * i0));
    float m = fmaf(A.m[5], i2, fmaf(A.m[4], i1, A.m[3] * i0));
    float s = fmaf(A.m[8], i2, fmaf(A.m[7], i1, A.m[6] * i0));
    const float gamma = (float)(1.0 / 0.43);
    l = signed_pow(l, gamma);
    m = signed_pow(m, gamma);
    s = signed_pow(s, gamma);
    o0 = fmaf(B.m[2], s, fmaf(B.m[1], m, B.m[0] * l));
    o1 = fmaf(B.m[5], s, fmaf(B.m[4], m, B.m[3] * l));
    o2 = fmaf(B.m[8], s, fmaf(B.m[7], m, B.m[6] * l));
}

// inverse butterfly: the four mirrored samples from the quadrant transforms z[b*2 + a] (+ the constant of
// the DC coefficient)
__device__ __forceinline__ void unbutterfly4(const float (&z)[4], float dcv, float& p1, float& p2, float& p3, float& p4) {
    const float e = z[0] + z[1], f = z[0] - z[1];     // column-even part at rows h', H-1-h'
    const float o = z[2] + z[3], g = z[2] - z[3];     // column-odd part
    p1 = (e + o) + dcv;
    p2 = (e - o) + dcv;
    p3 = (f + g) + dcv;
    p4 = (f - g) + dcv;
}

// quadrant planes z[s][plane][h'][w'] -> un-folded IPT -> RGB (util.py:85-97); COLOR = false: plain planes out
template <bool COLOR, typename TOut = float>     // TOut = uint8_t: 8-bit pixels as torchvision's save_image stores them
__global__ void __launch_bounds__(256) unfold_kernel(const float* __restrict__ z, const float* __restrict__ dc,
                                                     TOut* __restrict__ out, int64_t n_items, int h, int w, Mat3 A, Mat3 B,
                                                     int64_t item0) {
    constexpr int CH = COLOR ? 3 : 1;
    const int h2 = h >> 1, w8 = w >> 3, w4 = w >> 2;
    const uint32_t plane4 = (uint32_t)h * w4, q4 = (uint32_t)h2 * w8;      // in float4 units
    const int64_t n_planes = n_items * CH;
    // grid: (blocks per item, items) -- one (item, h', 4 columns) per thread, 32-bit index arithmetic inside an item
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= q4) return;
    const int y = (int)(idx / (uint32_t)w8);
    const int xv = (int)(idx - (uint32_t)y * w8);
    const int64_t item = item0 + blockIdx.y;
    float px[4][CH][4];          // [corner][channel][pixel]
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int64_t pl = item * CH + c;
        const float dcv = dc ? __ldg(dc + pl) : 0.0f;
        const float4* zp = reinterpret_cast<const float4*>(z) + (pl * q4 + idx);
        float4 q[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) q[s] = ld_stream(zp + s * n_planes * (int64_t)q4);
        const float qa[4][4] = {{q[0].x, q[0].y, q[0].z, q[0].w}, {q[1].x, q[1].y, q[1].z, q[1].w},
                                {q[2].x, q[2].y, q[2].z, q[2].w}, {q[3].x, q[3].y, q[3].z, q[3].w}};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float zz[4] = {qa[0][j], qa[1][j], qa[2][j], qa[3][j]};
            unbutterfly4(zz, dcv, px[0][c][j], px[1][c][3 - j], px[2][c][j], px[3][c][3 - j]);
        }
    }
    TOut* dst = out + item * CH * (int64_t)plane4 * 4;
    const uint32_t top = (uint32_t)y * w4, bot = (uint32_t)(h - 1 - y) * w4;
    const uint32_t xr = w4 - 1 - xv;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (COLOR) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float r, gg, bb;
                ipt_px_to_rgb_f(px[k][0][j], px[k][CH > 1 ? 1 : 0][j], px[k][CH > 2 ? 2 : 0][j], A, B, r, gg, bb);
                px[k][0][j] = r; px[k][CH > 1 ? 1 : 0][j] = gg; px[k][CH > 2 ? 2 : 0][j] = bb;
            }
        }
        const uint32_t o = ((k & 2) ? bot : top) + ((k & 1) ? xr : (uint32_t)xv);
#pragma unroll
        for (int c = 0; c < CH; ++c)
            st_px4(dst, c * plane4 + o, make_float4(px[k][c][0], px[k][c][1], px[k][c][2], px[k][c][3]));
    }
}

// feature_extraction_dct_autoencoder.py:635-653 un-patchify (CODES = false) or its fusion with
// lfq.indices_to_codes + PatchNorm.inverse_norm (CODES = true; lfq.py:105-134, patchnorm.py:167-177),
// writing the folded coefficient quadrants yq[b][a][plane][i][j] = Y[2i+a, 2j+b] as scaled fp16 hi/lo.
// One CTA per (channel, tile row, group of images): p coefficient rows of up to `imgs_per_cta` planes.
// A thread owns 8 consecutive coefficient columns (4 per column parity: one 8-byte hi and lo store per parity)
// of one image at a time and walks the p rows of the tile row; the main loop has no block-wide barrier.
// Tiles are at least 8 columns wide, so the 8 columns of a thread touch at most two tokens.
// CODES: the PatchNorm statistics of the tile row (divisor b*sqrt2 + eps and median) are staged in shared
// memory ONCE per CTA as [row][column group][8] (16-byte halves swizzled so that 128-bit reads are
// conflict-free) and reused for every image.  FAST (one LFQ codebook per patch row, p <= 16): the 2 x p code
// words of a thread's two tokens are read in one burst and boiled down to one byte of sign bits per row.
// DENORM (with CODES = false): the patches are NORMALISED ones and PatchNorm.inverse_norm (patchnorm.py:167-177) is applied
// on the way, from the same staged statistics -- the de-normalised patches are never written.
template <bool CODES, bool FAST, bool DENORM = false>
__global__ void __launch_bounds__(256, 4) unpatchify_fold_kernel(const float* __restrict__ patches,
                                                                 const int64_t* __restrict__ codes,
                                                                 const int32_t* __restrict__ slot_map,
                                                                 const int32_t* __restrict__ img_sel, int64_t n_img,
                                                                 int imgs_per_cta, int C, int th, int tw, int p, int rows,
                                                                 int cols, int ldq, LfqNormParams q,
                                                                 __half* __restrict__ hi, __half* __restrict__ lo,
                                                                 float* __restrict__ dc, float dc_factor, float scale) {
    extern __shared__ __align__(16) float smem_f[];
    const int z = p * p;
    const int nxv = ldq >> 2;
    const int tile_rows = rows / p;
    const unsigned id = blockIdx.x;
    const int ty = (int)(id % (unsigned)tile_rows);
    const unsigned t = id / (unsigned)tile_rows;
    const int c = (int)(t % (unsigned)C);
    const int64_t sel0 = (int64_t)(t / (unsigned)C) * imgs_per_cta;
    const int n_here = (int)min((int64_t)imgs_per_cta, n_img - sel0);
    const bool row_in = ty < th;
    const int n_tx = row_in ? min(tw, cols / p) : 0;          // tiles of this row that can hold a token
    float4* sd_t = reinterpret_cast<float4*>(smem_f);         // [p][nxv][2] float4
    float4* med_t = sd_t + p * nxv * 2;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, n_thr = blockDim.x * blockDim.y;
    if (CODES) {
        const float* med_row = q.median + (((int64_t)c * q.H + ty) * q.W) * q.z;
        const float* b_row = q.b + (((int64_t)c * q.H + ty) * q.W) * q.z;
        float* sd_f = reinterpret_cast<float*>(sd_t);
        float* med_f = reinterpret_cast<float*>(med_t);
        // zero the column groups past the last tile, then scatter the tile-ordered statistics
        for (int i = tid; i < p * nxv * 8; i += n_thr) { sd_f[i] = 0.0f; med_f[i] = 0.0f; }
        __syncthreads();
        for (int e = tid; e < n_tx * z; e += n_thr) {
            const int tx = e / z, r = e - tx * z;
            const int py = r / p, px = r - py * p;
            const int x = tx * p + px, xv = x >> 3, k = x & 7;
            const int idx = ((py * nxv + xv) * 2 + ((k >> 2) ^ ((xv >> 2) & 1))) * 4 + (k & 3);
            sd_f[idx] = __fadd_rn(__fmul_rn(__ldg(b_row + e), kSqrt2f), q.eps);
            med_f[idx] = __ldg(med_row + e);
        }
        __syncthreads();
    }
    const int rows2 = rows >> 1;
    const int64_t n_planes = n_img * C;
    const int64_t quad = n_planes * rows2 * (int64_t)ldq;          // elements of one quadrant array
    const bool pairs = !CODES && (p & 1) == 0 && (reinterpret_cast<uintptr_t>(patches) & 7) == 0;
    for (int k = threadIdx.y; k < n_here; k += blockDim.y) {
        const int64_t img = img_sel ? img_sel[sel0 + k] : sel0 + k;
        const int64_t plane = (sel0 + k) * C + c;
        const int32_t* smap = slot_map + ((img * C + c) * th + (row_in ? ty : 0)) * tw;
        for (int xv = threadIdx.x; xv < nxv; xv += blockDim.x) {
            const int x0 = xv * 8;
            const int tx0 = x0 / p, px0 = x0 - tx0 * p;
            const int n_first = min(8, p - px0);                      // columns that belong to tile tx0
            const int32_t slot0 = tx0 < n_tx ? __ldg(smap + tx0) : -1;
            const int32_t slot1 = (n_first < 8 && tx0 + 1 < n_tx) ? __ldg(smap + tx0 + 1) : -1;
            const float* src0 = (!CODES && slot0 >= 0) ? patches + (int64_t)slot0 * z + px0 : nullptr;
            const float* src1 = (!CODES && slot1 >= 0) ? patches + (int64_t)slot1 * z - n_first : nullptr;
            const int64_t* cw0 = (CODES && slot0 >= 0) ? codes + (int64_t)slot0 * q.c : nullptr;
            const int64_t* cw1 = (CODES && slot1 >= 0) ? codes + (int64_t)slot1 * q.c : nullptr;
            // valid8: bit 7-j set where column j has a token
            const unsigned m_first = (0xffu << (8 - n_first)) & 0xffu;
            const unsigned valid8 = (slot0 >= 0 ? m_first : 0u) | (slot1 >= 0 ? (0xffu & ~m_first) : 0u);
            unsigned bpack[4] = {0u, 0u, 0u, 0u};                    // FAST: sign byte of row py in byte py
            if (CODES && FAST) {
                // column j < n_first takes bit (d-1 - (px0+j)) of the first token's word, the others bit (d-1 - (j-n_first))
                // of the second token's: two contiguous bit fields, MSB first
                const int sh0 = q.d - px0 - n_first, sh1 = q.d - (8 - n_first);
#pragma unroll
                for (int py = 0; py < 16; ++py) {
                    if (py < p) {
                        const unsigned w0 = cw0 ? (unsigned)__ldg(cw0 + py) : 0u;
                        const unsigned w1 = cw1 ? (unsigned)__ldg(cw1 + py) : 0u;
                        const unsigned f0 = (w0 >> sh0) << (8 - n_first);
                        const unsigned f1 = n_first < 8 ? (w1 >> sh1) & (0xffu >> n_first) : 0u;
                        bpack[py >> 2] |= ((f0 | f1) & 0xffu) << (8 * (py & 3));
                    }
                }
            }
#pragma unroll 2
            for (int py = 0; py < p; ++py) {
                float v[8];
                if (CODES) {
                    const int ti = (py * nxv + xv) * 2, sw = (xv >> 2) & 1;
                    const float4 s0 = sd_t[ti + sw], s1 = sd_t[ti + (sw ^ 1)];
                    const float4 m0 = med_t[ti + sw], m1 = med_t[ti + (sw ^ 1)];
                    const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                    const float mv[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
                    unsigned bits8;
                    if (FAST) {
                        bits8 = (bpack[(py >> 2) & 3] >> (8 * (py & 3))) & 0xffu;
                    } else {
                        bits8 = 0u;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const bool first = j < n_first;
                            const int64_t* cw = first ? cw0 : cw1;
                            const int e = py * p + (first ? px0 + j : j - n_first);
                            const int cb = e / q.d, bi = e - cb * q.d;
                            if (cw && (((unsigned long long)__ldg(cw + cb) >> (q.d - 1 - bi)) & 1ull)) bits8 |= 0x80u >> j;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float qv = (bits8 & (0x80u >> j)) ? q.scale : -q.scale;             // lfq.py:118-120
                        const float val = __fadd_rn(__fmul_rn(qv, sv[j]), mv[j]);                 // patchnorm.py:177
                        v[j] = (valid8 & (0x80u >> j)) ? val : 0.0f;
                    }
                } else if (pairs) {
                    // even tile width: px0 and n_first are even, so column pairs are 8-byte aligned in the token
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        const float* src = j < n_first ? src0 : src1;
                        const float2 t = src ? __ldg(reinterpret_cast<const float2*>(src + py * p + j)) : make_float2(0.0f, 0.0f);
                        v[j] = t.x;
                        v[j + 1] = t.y;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float* src = j < n_first ? src0 : src1;
                        v[j] = src ? __ldg(src + py * p + j) : 0.0f;
                    }
                }
                if (DENORM) {
                    // the statistics of the two tiles under these 8 columns, straight from the L2-resident tables (a
                    // shared-memory copy per CTA costs more than it saves: it is shared by 8 images only)
                    const int64_t st0 = (((int64_t)c * q.H + ty) * q.W + tx0) * q.z + py * p + px0;      // column j < n_first: + j
                    const int64_t st1 = st0 - px0 + q.z - n_first;                                       // the others: + j
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        const bool first = j < n_first;
                        const bool on = (valid8 & (0x80u >> j)) != 0u;          // p, px0, n_first even: pairs share a tile
                        float2 mv = make_float2(0.f, 0.f), bv = make_float2(0.f, 0.f);
                        if (on) {
                            const int64_t e = (first ? st0 : st1) + j;
                            mv = __ldg(reinterpret_cast<const float2*>(q.median + e));
                            bv = __ldg(reinterpret_cast<const float2*>(q.b + e));
                        }
                        const float s0 = __fadd_rn(__fmul_rn(bv.x, kSqrt2f), q.eps), s1 = __fadd_rn(__fmul_rn(bv.y, kSqrt2f), q.eps);
                        v[j] = on ? __fadd_rn(__fmul_rn(v[j], s0), mv.x) : 0.0f;                          // patchnorm.py:177
                        v[j + 1] = on ? __fadd_rn(__fmul_rn(v[j + 1], s1), mv.y) : 0.0f;
                    }
                }
                if (ty == 0 && py == 0 && xv == 0) {
                    dc[plane] = v[0] * dc_factor;
                    v[0] = 0.0f;
                }
                const int kh = ty * p + py;
                const int a = kh & 1, ii = kh >> 1;
                const int64_t o = ((a * n_planes + plane) * rows2 + ii) * (int64_t)ldq + xv * 4;   // + b * 2 * quad
                const float ev[4] = {v[0], v[2], v[4], v[6]}, od[4] = {v[1], v[3], v[5], v[7]};
                uint2 vh, vl;
                split16x4(ev, scale, vh, vl);
                *reinterpret_cast<uint2*>(hi + o) = vh;
                *reinterpret_cast<uint2*>(lo + o) = vl;
                split16x4(od, scale, vh, vl);
                *reinterpret_cast<uint2*>(hi + 2 * quad + o) = vh;
                *reinterpret_cast<uint2*>(lo + 2 * quad + o) = vl;
            }
        }
    }
}

// The same for one LFQ codebook per patch row (c == d == p <= 16), the configuration of the headline benchmark.
// A de-quantised, de-normalised coefficient can only be one of TWO values per position: median + s*sd or
// median - s*sd (lfq.py:118-120, patchnorm.py:177).  Both are computed once per CTA, already scaled and split into
// their fp16 hi/lo halves (packed in one 32-bit word: hi in the low half), and staged in shared memory as
// [row][column group][8] (16-byte halves swizzled: conflict-free 128-bit reads).  The main loop is then one select
// per element on the sign bit and one byte-permute per pair of elements to assemble the 8-byte hi and lo groups
// of each column parity: about a quarter of the instructions of the generic kernel, bit-identical results.
__device__ __forceinline__ uint32_t pack_split16(float v, float scale) {
    const float sv = v * scale;
    const __half h = __float2half_rn(sv);
    const __half l = __float2half_rn(sv - __half2float(h));
    return (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16);
}

// The two-value tables of EVERY (channel, tile row) of the PatchNorm statistics, in the shared-memory layout of
// decode_codes_rows_kernel ([channel][tile row][row][column group of the full statistics width][8], swizzled): a CTA
// then stages its rows with plain 128-bit copies.  tab: pos then neg, each C*H*p*nxv_tab*8 words.
__global__ void __launch_bounds__(256) decode_tables_kernel(LfqNormParams q, int p, int nxv_tab, float scale,
                                                            uint32_t* __restrict__ pos_tab, uint32_t* __restrict__ neg_tab) {
    const int z = p * p;
    const int64_t total = (int64_t)q.C * q.H * p * nxv_tab * 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k4 = (int)(i & 3);
        int64_t r = i >> 2;
        const int hs = (int)(r & 1);             // stored 16-byte half
        r >>= 1;
        const int xv = (int)(r % nxv_tab);
        r /= nxv_tab;
        const int py = (int)(r % p);
        r /= p;
        const int ty = (int)(r % q.H);
        const int c = (int)(r / q.H);
        const int k = ((hs ^ ((xv >> 2) & 1)) << 2) | k4;        // column inside the group of 8 (un-swizzled)
        const int x = xv * 8 + k;
        const int tx = x / p, px = x - tx * p;
        uint32_t vp = 0u, vn = 0u;
        if (tx < q.W && !(ty == 0 && tx == 0 && py == 0 && px == 0)) {
            const int64_t e = (((int64_t)c * q.H + ty) * q.W + tx) * z + py * p + px;
            const float sd = __fadd_rn(__fmul_rn(__ldg(q.b + e), kSqrt2f), q.eps);
            const float md = __ldg(q.median + e);
            vp = pack_split16(__fadd_rn(__fmul_rn(q.scale, sd), md), scale);
            vn = pack_split16(__fadd_rn(__fmul_rn(-q.scale, sd), md), scale);
        }
        pos_tab[i] = vp;
        neg_tab[i] = vn;
    }
}

template <bool PRE>      // PRE: the tables come from decode_tables_kernel (tab / nxv_tab), else they are built here
__global__ void __launch_bounds__(256, 3) decode_codes_rows_kernel(const int64_t* __restrict__ codes,
                                                                   const int32_t* __restrict__ slot_map,
                                                                   const int32_t* __restrict__ img_sel, int64_t n_img,
                                                                   int imgs_per_cta, int C, int th, int tw, int p, int rows,
                                                                   int cols, int ldq, LfqNormParams q,
                                                                   const uint32_t* __restrict__ pos_tab,
                                                                   const uint32_t* __restrict__ neg_tab, int nxv_tab,
                                                                   __half* __restrict__ hi, __half* __restrict__ lo,
                                                                   float* __restrict__ dc, float dc_factor, float scale) {
    extern __shared__ __align__(16) float smem_f[];
    const int z = p * p;
    const int nxv = ldq >> 2;
    const int tile_rows = rows / p;
    const unsigned id = blockIdx.x;
    const int ty = (int)(id % (unsigned)tile_rows);
    const unsigned t = id / (unsigned)tile_rows;
    const int c = (int)(t % (unsigned)C);
    const int64_t sel0 = (int64_t)(t / (unsigned)C) * imgs_per_cta;
    const int n_here = (int)min((int64_t)imgs_per_cta, n_img - sel0);
    const bool row_in = ty < th;
    const int n_tx = row_in ? min(tw, cols / p) : 0;
    uint4* pos_t = reinterpret_cast<uint4*>(smem_f);          // [p][nxv][2] words of 4: value when the bit is 1
    uint4* neg_t = pos_t + p * nxv * 2;                       //                            value when the bit is 0
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, n_thr = blockDim.x * blockDim.y;
    const float* med_row = q.median + (((int64_t)c * q.H + ty) * q.W) * q.z;
    const float* b_row = q.b + (((int64_t)c * q.H + ty) * q.W) * q.z;
    if (PRE) {
        const uint4* ps = reinterpret_cast<const uint4*>(pos_tab) + ((int64_t)c * q.H + ty) * p * nxv_tab * 2;
        const uint4* ns = reinterpret_cast<const uint4*>(neg_tab) + ((int64_t)c * q.H + ty) * p * nxv_tab * 2;
        for (int i = tid; i < p * nxv * 2; i += n_thr) {
            const int py = i / (nxv * 2), r = i - py * (nxv * 2);
            pos_t[i] = __ldg(ps + py * nxv_tab * 2 + r);
            neg_t[i] = __ldg(ns + py * nxv_tab * 2 + r);
        }
        __syncthreads();
    } else {
        uint32_t* pos_w = reinterpret_cast<uint32_t*>(pos_t);
        uint32_t* neg_w = reinterpret_cast<uint32_t*>(neg_t);
        for (int i = tid; i < p * nxv * 8; i += n_thr) { pos_w[i] = 0u; neg_w[i] = 0u; }
        __syncthreads();
        for (int e = tid; e < n_tx * z; e += n_thr) {
            const int tx = e / z, r = e - tx * z;
            const int py = r / p, px = r - py * p;
            const int x = tx * p + px, xv = x >> 3, k = x & 7;
            const int idx = ((py * nxv + xv) * 2 + ((k >> 2) ^ ((xv >> 2) & 1))) * 4 + (k & 3);
            const float sd = __fadd_rn(__fmul_rn(__ldg(b_row + e), kSqrt2f), q.eps);
            const float md = __ldg(med_row + e);
            float vp = __fadd_rn(__fmul_rn(q.scale, sd), md);                               // patchnorm.py:177
            float vn = __fadd_rn(__fmul_rn(-q.scale, sd), md);
            if (ty == 0 && e == 0) { vp = 0.0f; vn = 0.0f; }                                // the DC goes to dc[]
            pos_w[idx] = pack_split16(vp, scale);
            neg_w[idx] = pack_split16(vn, scale);
        }
        __syncthreads();
    }
    const int rows2 = rows >> 1;
    const int64_t n_planes = n_img * C;
    const int64_t quad = n_planes * rows2 * (int64_t)ldq;
    // a token's p code words are 16-byte aligned pairs when p is even (p int64 words per token)
    const bool vec_codes = (p & 1) == 0 && (reinterpret_cast<uintptr_t>(codes) & 15) == 0;
    for (int k = threadIdx.y; k < n_here; k += blockDim.y) {
        const int64_t img = img_sel ? img_sel[sel0 + k] : sel0 + k;
        const int64_t plane = (sel0 + k) * C + c;
        const int32_t* smap = slot_map + ((img * C + c) * th + (row_in ? ty : 0)) * tw;
        for (int xv = threadIdx.x; xv < nxv; xv += blockDim.x) {
            const int x0 = xv * 8;
            const int tx0 = x0 / p, px0 = x0 - tx0 * p;
            const int n_first = min(8, p - px0);
            const int32_t slot0 = tx0 < n_tx ? __ldg(smap + tx0) : -1;
            const int32_t slot1 = (n_first < 8 && tx0 + 1 < n_tx) ? __ldg(smap + tx0 + 1) : -1;
            const int64_t* cw0 = slot0 >= 0 ? codes + (int64_t)slot0 * q.c : nullptr;
            const int64_t* cw1 = slot1 >= 0 ? codes + (int64_t)slot1 * q.c : nullptr;
            const unsigned m_first = (0xffu << (8 - n_first)) & 0xffu;
            const unsigned valid8 = (slot0 >= 0 ? m_first : 0u) | (slot1 >= 0 ? (0xffu & ~m_first) : 0u);
            unsigned bpack[4] = {0u, 0u, 0u, 0u};                    // sign byte of row py in byte py
            const int sh0 = q.d - px0 - n_first, sh1 = q.d - (8 - n_first);
            // all code words of the two tokens first (independent 128-bit loads: two int64 words each), then the bits
            const longlong2* c0 = reinterpret_cast<const longlong2*>(cw0 ? cw0 : codes);
            const longlong2* c1 = reinterpret_cast<const longlong2*>(cw1 ? cw1 : codes);
            longlong2 w0v[8], w1v[8];
            if (vec_codes) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (2 * j < p) {
                        w0v[j] = __ldg(c0 + j);
                        w1v[j] = __ldg(c1 + j);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (2 * j < p) {
                        w0v[j].x = __ldg((cw0 ? cw0 : codes) + 2 * j);
                        w1v[j].x = __ldg((cw1 ? cw1 : codes) + 2 * j);
                        w0v[j].y = 2 * j + 1 < p ? __ldg((cw0 ? cw0 : codes) + 2 * j + 1) : 0;
                        w1v[j].y = 2 * j + 1 < p ? __ldg((cw1 ? cw1 : codes) + 2 * j + 1) : 0;
                    }
                }
            }
#pragma unroll
            for (int py = 0; py < 16; ++py) {
                if (py < p) {
                    const unsigned w0 = cw0 ? (unsigned)((py & 1) ? w0v[py >> 1].y : w0v[py >> 1].x) : 0u;
                    const unsigned w1 = cw1 ? (unsigned)((py & 1) ? w1v[py >> 1].y : w1v[py >> 1].x) : 0u;
                    const unsigned f0 = (w0 >> sh0) << (8 - n_first);
                    const unsigned f1 = n_first < 8 ? (w1 >> sh1) & (0xffu >> n_first) : 0u;
                    bpack[py >> 2] |= ((f0 | f1) & 0xffu) << (8 * (py & 3));
                }
            }
            if (ty == 0 && xv == 0) {
                // the DC coefficient (row 0, column 0 of tile row 0) is carried separately
                float dcval = 0.0f;
                if (slot0 >= 0) {
                    const float sd = __fadd_rn(__fmul_rn(__ldg(b_row), kSqrt2f), q.eps);
                    const float qv = (bpack[0] & 0x80u) ? q.scale : -q.scale;
                    dcval = __fadd_rn(__fmul_rn(qv, sd), __ldg(med_row)) * dc_factor;
                }
                dc[plane] = dcval;
            }
            const int sw = (xv >> 2) & 1;
            // output rows alternate between the two row parities a = kh & 1: one running offset per parity
            const int kh0 = ty * p, par0 = kh0 & 1;
            int64_t o_even = ((int64_t)plane * rows2 + ((kh0 + par0) >> 1)) * (int64_t)ldq + xv * 4;
            int64_t o_odd = ((n_planes + plane) * rows2 + ((kh0 + 1 - par0) >> 1)) * (int64_t)ldq + xv * 4;
            const uint4* pt_row = pos_t + xv * 2;
            const uint4* nt_row = neg_t + xv * 2;
#pragma unroll 2
            for (int py = 0; py < p; ++py, pt_row += nxv * 2, nt_row += nxv * 2) {
                const uint4 p0 = pt_row[sw], p1 = pt_row[sw ^ 1];
                const uint4 n0 = nt_row[sw], n1 = nt_row[sw ^ 1];
                const unsigned bits8 = ((bpack[(py >> 2) & 3] >> (8 * (py & 3))) & 0xffu) ;
                const uint32_t pw[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
                const uint32_t nw[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t v = (bits8 & (0x80u >> j)) ? pw[j] : nw[j];
                    w[j] = (valid8 & (0x80u >> j)) ? v : 0u;
                }
                // hi halves are the low 16 bits, lo halves the high 16 bits of each word; even columns -> parity b = 0
                const uint2 eh = make_uint2(__byte_perm(w[0], w[2], 0x5410), __byte_perm(w[4], w[6], 0x5410));
                const uint2 el = make_uint2(__byte_perm(w[0], w[2], 0x7632), __byte_perm(w[4], w[6], 0x7632));
                const uint2 oh = make_uint2(__byte_perm(w[1], w[3], 0x5410), __byte_perm(w[5], w[7], 0x5410));
                const uint2 ol = make_uint2(__byte_perm(w[1], w[3], 0x7632), __byte_perm(w[5], w[7], 0x7632));
                const bool odd = (kh0 + py) & 1;
                const int64_t o = odd ? o_odd : o_even;              // (a, plane, kh >> 1, xv * 4); + b * 2 * quad
                if (odd) o_odd += ldq; else o_even += ldq;
                *reinterpret_cast<uint2*>(hi + o) = eh;
                *reinterpret_cast<uint2*>(lo + o) = el;
                *reinterpret_cast<uint2*>(hi + 2 * quad + o) = oh;
                *reinterpret_cast<uint2*>(lo + 2 * quad + o) = ol;
            }
        }
    }
}

// ------------------------------------------------------------------------------ any plane size (scalar variants)
// One thread = one quadrant position (item, h', w') with h' < ceil(h/2), w' < ceil(w/2): the four mirrored samples of
// each channel, the self-paired middle row / column of odd sizes counted once.  Quadrant rows have pitch `ld`.
__device__ __forceinline__ float px_load(const float* p, int64_t i) { return __ldg(p + i); }
__device__ __forceinline__ float px_load(const uint8_t* p, int64_t i) { return u8_to_unit(__ldg(p + i)); }
__device__ __forceinline__ void px_store(float* p, int64_t i, float v) { p[i] = v; }
__device__ __forceinline__ void px_store(uint8_t* p, int64_t i, float v) { p[i] = (uint8_t)unit_to_u8(v); }

template <typename TIn, bool COLOR>
__global__ void __launch_bounds__(256) means_any_kernel(const TIn* __restrict__ x, float* __restrict__ mu, float* __restrict__ dc,
                                                        int64_t plane, Mat3 A, Mat3 B, float dc_factor) {
    constexpr int CH = COLOR ? 3 : 1;
    __shared__ float red[3][8];
    const int64_t item = blockIdx.x;
    const TIn* src = x + item * CH * plane;
    const int64_t step = plane > 8192 ? plane / 8192 : 1;       // any estimate is exact: what is removed is added back as DC
    const int64_t count = (plane + step - 1) / step;
    float s[3] = {0.f, 0.f, 0.f};
    for (int64_t k = threadIdx.x; k < count; k += blockDim.x) {
        const int64_t i = k * step;
        if (COLOR) {
            float o0, o1, o2;
            rgb_px_to_ipt_f(px_load(src, i), px_load(src, plane + i), px_load(src, 2 * plane + i), A, B, o0, o1, o2);
            s[0] += o0; s[1] += o1; s[2] += o2;
        } else {
            s[0] += px_load(src, i);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const float v = warp_sum(s[c]);
        if (lane == 0) red[c][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < CH) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += red[threadIdx.x][k];
        const float m = t / (float)count;
        mu[item * CH + threadIdx.x] = m;
        dc[item * CH + threadIdx.x] = m * dc_factor;
    }
}

template <typename TIn, bool COLOR>
__global__ void __launch_bounds__(256) fold_any_kernel(const TIn* __restrict__ x, const float* __restrict__ mus,
                                                       __half* __restrict__ hi, __half* __restrict__ lo, int64_t n_items, int h,
                                                       int w, int ld, Mat3 A, Mat3 B, float scale) {
    constexpr int CH = COLOR ? 3 : 1;
    const int h2 = (h + 1) >> 1, w2 = (w + 1) >> 1;
    const int64_t plane = (int64_t)h * w, n_planes = n_items * CH, quad = n_planes * h2 * (int64_t)ld;
    const int64_t total = n_items * h2 * w2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xv = (int)(i % w2);
        const int64_t t = i / w2;
        const int y = (int)(t % h2);
        const int64_t item = t / h2;
        const TIn* src = x + item * CH * plane;
        const int xr = w - 1 - xv, yb = h - 1 - y;
        const bool self_col = xr == xv, self_row = yb == y;
        const int64_t o[4] = {(int64_t)y * w + xv, (int64_t)y * w + xr, (int64_t)yb * w + xv, (int64_t)yb * w + xr};
        float v[4][3];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (COLOR) rgb_px_to_ipt_f(px_load(src, o[k]), px_load(src, plane + o[k]), px_load(src, 2 * plane + o[k]), A, B,
                                       v[k][0], v[k][1], v[k][2]);
            else v[k][0] = px_load(src, o[k]);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const float mu = __ldg(mus + item * CH + c);
            const float p1 = v[0][c] - mu, p2 = v[1][c] - mu, p3 = v[2][c] - mu, p4 = v[3][c] - mu;
            const float s_top = self_col ? p1 : p1 + p2, d_top = self_col ? 0.f : p1 - p2;
            const float s_bot = self_col ? p3 : p3 + p4, d_bot = self_col ? 0.f : p3 - p4;
            const float q[4] = {self_row ? s_top : s_top + s_bot, self_row ? 0.f : s_top - s_bot,
                                self_row ? d_top : d_top + d_bot, self_row ? 0.f : d_top - d_bot};   // [b*2 + a]
            const int64_t oq = ((item * CH + c) * h2 + y) * (int64_t)ld + xv;
#pragma unroll
            for (int sI = 0; sI < 4; ++sI) {
                const float sv = q[sI] * scale;
                const __half hh = __float2half_rn(sv);
                hi[sI * quad + oq] = hh;
                lo[sI * quad + oq] = __float2half_rn(sv - __half2float(hh));
            }
        }
    }
}

// z[s][plane][h'][w'] (dense, ceil(h/2) x ceil(w/2)) -> pixels
template <bool COLOR, typename TOut>
__global__ void __launch_bounds__(256) unfold_any_kernel(const float* __restrict__ z, const float* __restrict__ dc,
                                                         TOut* __restrict__ out, int64_t n_items, int h, int w, Mat3 A, Mat3 B) {
    constexpr int CH = COLOR ? 3 : 1;
    const int h2 = (h + 1) >> 1, w2 = (w + 1) >> 1;
    const int64_t plane = (int64_t)h * w, n_planes = n_items * CH, quad = n_planes * h2 * (int64_t)w2;
    const int64_t total = n_items * h2 * w2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xv = (int)(i % w2);
        const int64_t t = i / w2;
        const int y = (int)(t % h2);
        const int64_t item = t / h2;
        const int xr = w - 1 - xv, yb = h - 1 - y;
        const bool self_col = xr == xv, self_row = yb == y;
        float px[4][3];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int64_t oq = ((item * CH + c) * h2 + y) * (int64_t)w2 + xv;
            float zz[4] = {z[oq], z[quad + oq], z[2 * quad + oq], z[3 * quad + oq]};       // [b*2 + a]
            if (self_row) { zz[1] = 0.f; zz[3] = 0.f; }
            if (self_col) { zz[2] = 0.f; zz[3] = 0.f; }
            unbutterfly4(zz, dc ? __ldg(dc + item * CH + c) : 0.f, px[0][c], px[1][c], px[2][c], px[3][c]);
        }
        const int64_t o[4] = {(int64_t)y * w + xv, (int64_t)y * w + xr, (int64_t)yb * w + xv, (int64_t)yb * w + xr};
        const bool wr[4] = {true, !self_col, !self_row, !self_col && !self_row};
        TOut* dst = out + item * CH * plane;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!wr[k]) continue;
            if (COLOR) {
                float r, g, b;
                ipt_px_to_rgb_f(px[k][0], px[k][1], px[k][2], A, B, r, g, b);
                px_store(dst, o[k], r);
                px_store(dst, plane + o[k], g);
                px_store(dst, 2 * plane + o[k], b);
            } else {
                px_store(dst, o[k], px[k][0]);
            }
        }
    }
}

// ------------------------------------------------------------------------------ decode inside inverse pass 1 (GEN)
// The two-value tables of decode_tables_kernel in the order the producer warps of fold_gemm_kernel<0, true> read them:
// one 8 KB block per (b, a, channel, 32-row block of i, 32-column block of j) = [4 parts][128 threads pt = il*4 + c] uint4;
// a uint4 of part 0 / 1 holds {P, N, P, N} words of the pairs (0, 1) / (2, 3) of chunk c (columns j = kb*32 + c*8 + k)
// built from the fp16 HI halves, parts 2 / 3 the same from the LO halves: P = halves of the "bit is 1" values
// median + s*sd of the pair's two columns, N = of the "bit is 0" values median - s*sd.  Coefficient (kh, kw) =
// (2i + a, 2j + b); zero outside the plane and at the DC position (carried in dc[]).
__global__ void __launch_bounds__(256) decode_gen_tables_kernel(LfqNormParams q, int p, int rows, int cols, int n_iblk,
                                                                int num_kb, float scale, uint32_t* __restrict__ tab) {
    const int64_t total = (int64_t)4 * q.C * n_iblk * num_kb * 2048;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int wi = (int)(idx & 3), pt = (int)((idx >> 2) & 127), part = (int)((idx >> 9) & 3);
        int64_t r = idx >> 11;
        const int kb = (int)(r % num_kb);
        r /= num_kb;
        const int iblk = (int)(r % n_iblk);
        r /= n_iblk;
        const int ch = (int)(r % q.C);
        r /= q.C;
        const int a = (int)(r & 1), b = (int)(r >> 1);
        const int il = pt >> 2, c = pt & 3;
        const int kh = 2 * (iblk * 32 + il) + a;
        const int m = (part & 1) * 2 + (wi >> 1);
        const bool neg = wi & 1, lo_half = part >> 1;
        uint32_t word = 0u;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int kw = 2 * (kb * 32 + c * 8 + 2 * m + e) + b;
            uint32_t v = 0u;
            if (kh < rows && kw < cols && !(kh == 0 && kw == 0)) {
                const int ty = kh / p, py = kh - ty * p, tx = kw / p, px = kw - tx * p;
                if (ty < q.H && tx < q.W) {
                    const int64_t el = (((int64_t)ch * q.H + ty) * q.W + tx) * q.z + py * p + px;
                    const float sd = __fadd_rn(__fmul_rn(__ldg(q.b + el), kSqrt2f), q.eps);
                    const float md = __ldg(q.median + el);
                    v = pack_split16(__fadd_rn(__fmul_rn(neg ? -q.scale : q.scale, sd), md), scale);     // patchnorm.py:177
                }
            }
            word |= (lo_half ? (v >> 16) : (v & 0xffffu)) << (16 * e);
        }
        tab[idx] = word;
    }
}

// The sign bits of the code words (one LFQ codebook per patch row: c == d == p even, lfq.py:105-134) and the "token was
// kept" bits of a batch, re-ordered for the producer warps: per (b, a, plane, 32-row block of i, 32-column block of j)
// a 256-byte block [32 rows][4 chunks] of 16-bit words, low byte = sign bits of the 8 columns j = kb*32 + c*8 + k (bit k)
// of coefficient row kh = 2i + a and column parity b, high byte = valid bits.  One CTA = one (plane, i block): 64
// coefficient rows x groups of 8 tokens; a thread takes one row of 8 consecutive tokens, whose HW-bit fields of one
// parity concatenate to exactly HW bytes.  The words are staged in shared memory in their final order -- the k blocks
// of a (b, a, plane, i block) are contiguous in memory -- and written out with 16-byte stores.  Every token record
// (p code words) is read completely inside one CTA.  Also dc[plane] (the DC coefficient, scaled for the unfold kernel).
// bits 0, 2, 4, ... of the low half -> bits 0, 1, 2, ... of the low half, and the same inside the high half
__device__ __forceinline__ uint32_t compress_even_bits2x16(uint32_t x) {
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    return x;
}

template <int HW>      // HW = p / 2 bits of one column parity per token
__global__ void __launch_bounds__(256, 8) codes_bitplanes_kernel(const int64_t* __restrict__ codes, const int32_t* __restrict__ slot_map,
                                                                 const int32_t* __restrict__ img_sel, int64_t n_img, int C, int th,
                                                                 int tw, int rows, int cols, int n_iblk, int num_kb, int n_grp,
                                                                 LfqNormParams q, uint16_t* __restrict__ bv, float* __restrict__ dc,
                                                                 float dc_factor) {
    extern __shared__ __align__(16) uint16_t stage[];          // [b][a][kb][32 rows][4]
    constexpr int p = 2 * HW;
    const int64_t plane = blockIdx.x;
    const int iblk = blockIdx.y;
    const int blk_words = num_kb * 128;                        // words of one (b, a)
    if (iblk * 64 + 64 > rows)                                 // only the last block has rows that nobody writes
        for (int e = threadIdx.x; e < blk_words / 2; e += blockDim.x) reinterpret_cast<uint4*>(stage)[e] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    const int64_t k_img = plane / C;
    const int ch = (int)(plane - k_img * C);
    const int64_t img = img_sel ? img_sel[k_img] : k_img;
    const uint32_t* codes32 = reinterpret_cast<const uint32_t*>(codes);          // low words of the int64 code words
    const bool vec_slots = (tw & 7) == 0 && (reinterpret_cast<uintptr_t>(slot_map) & 15) == 0;
    for (int t = threadIdx.x; t < 64 * n_grp; t += blockDim.x) {
        const int rr = t / n_grp, G = t - rr * n_grp;           // rr = 2 * il + a
        const int kh = iblk * 64 + rr;
        if (kh >= rows) continue;
        const int ty = kh / p, py = kh - ty * p;
        const int n_tx = ty < th ? min(tw, cols / p) : 0;
        const int32_t* smap = slot_map + ((img * C + ch) * th + (ty < th ? ty : 0)) * tw;
        int32_t slots[8];
        if (vec_slots && G * 8 + 8 <= n_tx) {
            const int4 s0 = __ldg(reinterpret_cast<const int4*>(smap) + G * 2), s1 = __ldg(reinterpret_cast<const int4*>(smap) + G * 2 + 1);
            slots[0] = s0.x; slots[1] = s0.y; slots[2] = s0.z; slots[3] = s0.w;
            slots[4] = s1.x; slots[5] = s1.y; slots[6] = s1.z; slots[7] = s1.w;
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) slots[u] = G * 8 + u < n_tx ? __ldg(smap + G * 8 + u) : -1;
        }
        uint32_t words[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) words[u] = slots[u] >= 0 ? __ldg(codes32 + ((int64_t)slots[u] * q.c + py) * 2) : 0u;
        uint64_t s0 = 0, s1 = 0, sv = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (slots[u] >= 0) {
                const uint32_t r = __brev(words[u]) >> (32 - p);                  // bit px of r = sign of column px
                const uint32_t f = compress_even_bits2x16((r & 0xffffu) | ((r >> 1) << 16));   // even columns | odd columns << 16
                s0 |= (uint64_t)(f & 0xffffu) << (HW * u);
                s1 |= (uint64_t)(f >> 16) << (HW * u);
                sv |= (uint64_t)((1u << HW) - 1u) << (HW * u);
            }
        }
        if (kh == 0 && G == 0) {
            float dcval = 0.0f;
            if (sv & 1u) {
                const int64_t el = (((int64_t)ch * q.H) * q.W) * q.z;
                const float sd = __fadd_rn(__fmul_rn(__ldg(q.b + el), kSqrt2f), q.eps);
                const float qv = (s0 & 1u) ? q.scale : -q.scale;
                dcval = __fadd_rn(__fmul_rn(qv, sd), __ldg(q.median + el)) * dc_factor;
            }
            dc[plane] = dcval;
        }
        const int a = rr & 1, il = rr >> 1;
#pragma unroll
        for (int m = 0; m < HW; ++m) {
            const int jb = G * HW + m;                           // byte of the row's bit string: columns j = 8 jb .. 8 jb + 7
            if (jb < num_kb * 4) {
                const uint32_t vb = (uint32_t)(sv >> (8 * m)) & 0xffu;
                const int o = a * blk_words + (jb >> 2) * 128 + il * 4 + (jb & 3);
                stage[o] = (uint16_t)(((uint32_t)(s0 >> (8 * m)) & 0xffu) | (vb << 8));
                stage[o + 2 * blk_words] = (uint16_t)(((uint32_t)(s1 >> (8 * m)) & 0xffu) | (vb << 8));
            }
        }
    }
    __syncthreads();
    // 256-byte blocks [32 rows][4] at [b][a][group of 4 images][channel][i block][k block][image in group], 16-byte stores
    const int64_t n_grp4 = (n_img + 3) >> 2;
    for (int e = threadIdx.x; e < 4 * num_kb * 16; e += blockDim.x) {
        const int v = e & 15, kb = (e >> 4) % num_kb, ba = (e >> 4) / num_kb;
        uint4* dst = reinterpret_cast<uint4*>(bv + ((((((int64_t)ba * n_grp4 + (k_img >> 2)) * C + ch) * n_iblk + iblk) * num_kb + kb) * 4 +
                                                    (k_img & 3)) * 128) + v;
        *dst = reinterpret_cast<const uint4*>(stage)[e];
    }
}

// The same bit planes straight from the CODE GRID the forward pass wrote (code words in token-grid order,
// (n_img, th, tw, C, p) int32), for a round trip that kept every token: no slot map, no gather through the packed codes.
template <int HW>
__global__ void __launch_bounds__(256, 8) codes_bitplanes_grid_kernel(const int32_t* __restrict__ grid, int64_t n_img, int C, int th,
                                                                      int tw, int rows, int cols, int n_iblk, int num_kb, int n_grp,
                                                                      LfqNormParams q, uint16_t* __restrict__ bv,
                                                                      float* __restrict__ dc, float dc_factor) {
    extern __shared__ __align__(16) uint16_t stage[];          // [b][a][kb][32 rows][4]
    constexpr int p = 2 * HW;
    const int64_t plane = blockIdx.x;
    const int iblk = blockIdx.y;
    const int blk_words = num_kb * 128;
    if (iblk * 64 + 64 > rows)
        for (int e = threadIdx.x; e < blk_words / 2; e += blockDim.x) reinterpret_cast<uint4*>(stage)[e] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    const int64_t img = plane / C;
    const int ch = (int)(plane - img * C);
    const int n_tx = min(tw, cols / p);
    for (int t = threadIdx.x; t < 64 * n_grp; t += blockDim.x) {
        const int rr = t / n_grp, G = t - rr * n_grp;
        const int kh = iblk * 64 + rr;
        if (kh >= rows) continue;
        const int ty = kh / p, py = kh - ty * p;
        const int32_t* row = grid + (((img * th + ty) * tw) * C + ch) * (int64_t)p + py;      // + tx * C * p
        uint32_t words[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) words[u] = G * 8 + u < n_tx ? (uint32_t)__ldg(row + (int64_t)(G * 8 + u) * C * p) : 0u;
        uint64_t s0 = 0, s1 = 0, sv = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (G * 8 + u < n_tx) {
                const uint32_t r = __brev(words[u]) >> (32 - p);
                const uint32_t f = compress_even_bits2x16((r & 0xffffu) | ((r >> 1) << 16));
                s0 |= (uint64_t)(f & 0xffffu) << (HW * u);
                s1 |= (uint64_t)(f >> 16) << (HW * u);
                sv |= (uint64_t)((1u << HW) - 1u) << (HW * u);
            }
        }
        if (kh == 0 && G == 0) {
            const int64_t el = (((int64_t)ch * q.H) * q.W) * q.z;
            const float sd = __fadd_rn(__fmul_rn(__ldg(q.b + el), kSqrt2f), q.eps);
            const float qv = (s0 & 1u) ? q.scale : -q.scale;
            dc[plane] = __fadd_rn(__fmul_rn(qv, sd), __ldg(q.median + el)) * dc_factor;
        }
        const int a = rr & 1, il = rr >> 1;
#pragma unroll
        for (int m = 0; m < HW; ++m) {
            const int jb = G * HW + m;
            if (jb < num_kb * 4) {
                const uint32_t vb = (uint32_t)(sv >> (8 * m)) & 0xffu;
                const int o = a * blk_words + (jb >> 2) * 128 + il * 4 + (jb & 3);
                stage[o] = (uint16_t)(((uint32_t)(s0 >> (8 * m)) & 0xffu) | (vb << 8));
                stage[o + 2 * blk_words] = (uint16_t)(((uint32_t)(s1 >> (8 * m)) & 0xffu) | (vb << 8));
            }
        }
    }
    __syncthreads();
    const int64_t n_grp4 = (n_img + 3) >> 2;
    for (int e = threadIdx.x; e < 4 * num_kb * 16; e += blockDim.x) {
        const int v = e & 15, kb = (e >> 4) % num_kb, ba = (e >> 4) / num_kb;
        uint4* dst = reinterpret_cast<uint4*>(bv + ((((((int64_t)ba * n_grp4 + (img >> 2)) * C + ch) * n_iblk + iblk) * num_kb + kb) * 4 +
                                                    (img & 3)) * 128) + v;
        *dst = reinterpret_cast<const uint4*>(stage)[e];
    }
}

// fp32 coefficient planes (n_planes, kh, kw) -> folded split quadrants (DC moved to dc[])
__global__ void __launch_bounds__(256) fold_coef_kernel(const float* __restrict__ y, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, float* __restrict__ dc,
                                                        int64_t n_planes, int kh, int kw, int ldq, float dc_factor,
                                                        float scale) {
    const int kh2 = kh >> 1;
    const int64_t quad = n_planes * kh2 * (int64_t)ldq;
    const int64_t total = 4 * quad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % ldq);
        int64_t r = i / ldq;
        const int ii = (int)(r % kh2);
        r /= kh2;
        const int64_t pl = r % n_planes;
        const int s = (int)(r / n_planes);           // b * 2 + a
        const int a = s & 1, b = s >> 1;
        const int yy = 2 * ii + a, xx = 2 * j + b;
        float v = xx < kw ? y[(pl * kh + yy) * kw + xx] : 0.f;
        if (xx == 0 && yy == 0) { dc[pl] = v * dc_factor; v = 0.f; }
        const float sv = v * scale;
        const __half hh = __float2half_rn(sv);
        hi[i] = hh;
        lo[i] = __float2half_rn(sv - __half2float(hh));
    }
}

}  // namespace dcta

using namespace dcta;

// Scales of the split operands (powers of two, exact).  The fold sums up to four samples, so the image
// quadrants carry 2^6 where the plain path carries 2^8, and the forward intermediate 2^5 instead of 2^6.
static const float kFScaleX = 64.f, kFScaleP = 32.f, kFScaleY = 16.f, kFScaleQ = 64.f, kFScaleBasis = 1024.f;

// Any plane size: a quadrant holds ceil(n / 2) samples per axis (for odd n the middle sample pairs with itself: it
// enters the even rows once and the odd rows not at all, C[odd k, middle] = 0); rows are stored with an 8-element pitch.
static bool fold_dims_ok(int h, int w, int kh, int kw) {
    return h >= 2 && w >= 2 && kh > 0 && kw > 0 && kh % 2 == 0 && kw % 2 == 0 && kh <= h && kw <= w;
}
static inline int fold_half(int n) { return (n + 1) / 2; }
static inline int fold_pitch(int n) { return (int)(ceil_div(fold_half(n), 8) * 8); }
static inline bool fold_fast_dims(int h, int w) { return h % 16 == 0 && w % 16 == 0; }     // the vectorised colour kernels

extern "C" int dcta_fold_supported(int h, int w, int kh, int kw) {
    if (!fold_dims_ok(h, w, kh, kw)) return 0;
    FoldGemm g{};
    return fold_geometry(kw / 2, fold_half(w), g) && fold_geometry(kh / 2, fold_half(h), g, 64) &&
           fold_geometry(fold_half(w), kw / 2, g) && fold_geometry(fold_half(h), kh / 2, g);
}

template <typename TIn>
static int rgb_to_ipt_fold_any(const TIn* rgb, void* xq_hi, void* xq_lo, float* dc, float* sums_scratch, int64_t n_img, int h,
                               int w, const float* m_rgb2lms_host, const float* m_ipt_host, void* stream, const char* who) {
    DCTA_REQUIRE(rgb && xq_hi && xq_lo && dc && sums_scratch && m_rgb2lms_host && m_ipt_host, "%s: null pointer", who);
    DCTA_REQUIRE(h >= 2 && w >= 2 && n_img <= 65535, "%s: needs h, w >= 2 and at most 65535 images", who);
    if (n_img == 0) return DCTA_OK;
    Mat3 A, B;
    for (int i = 0; i < 9; ++i) { A.m[i] = m_rgb2lms_host[i]; B.m[i] = m_ipt_host[i]; }
    cudaStream_t st = as_stream(stream);
    if (fold_fast_dims(h, w) && (reinterpret_cast<uintptr_t>(rgb) & 15) == 0) {
        const float* mus;
        if constexpr (sizeof(TIn) == 1) mus = launch_ipt_plane_means_u8(rgb, sums_scratch, dc, n_img, h, w, A, B, st);
        else mus = launch_ipt_plane_means(rgb, sums_scratch, dc, n_img, h, w, A, B, st);
        Mat3 Bs;                                       // the operand scale (a power of two) folded into the matrix
        for (int i = 0; i < 9; ++i) Bs.m[i] = B.m[i] * kFScaleX;
        rgb_to_ipt_fold_kernel<TIn><<<dim3((unsigned)ceil_div((int64_t)(h / 2) * (w / 8), 256), (unsigned)n_img), 256, 0, st>>>(
            rgb, mus, (__half*)xq_hi, (__half*)xq_lo, n_img, h, w, A, Bs, kFScaleX);
    } else {            // any size: scalar kernels, quadrant rows padded to 8 elements
        float* mus = sums_scratch;
        means_any_kernel<TIn, true><<<(unsigned)n_img, 256, 0, st>>>(rgb, mus, dc, (int64_t)h * w, A, B, sqrtf((float)h * (float)w));
        fold_any_kernel<TIn, true><<<grid_for(n_img * fold_half(h) * fold_half(w), 256), 256, 0, st>>>(
            rgb, mus, (__half*)xq_hi, (__half*)xq_lo, n_img, h, w, fold_pitch(w), A, B, kFScaleX);
    }
    return check_launch(who);
}

extern "C" int dcta_rgb_to_ipt_fold(const float* rgb, void* xq_hi, void* xq_lo, float* dc, float* sums_scratch,
                                    int64_t n_img, int h, int w, const float* m_rgb2lms_host, const float* m_ipt_host,
                                    void* stream) {
    return rgb_to_ipt_fold_any(rgb, xq_hi, xq_lo, dc, sums_scratch, n_img, h, w, m_rgb2lms_host, m_ipt_host, stream,
                               "rgb_to_ipt_fold");
}

extern "C" int dcta_rgb_u8_to_ipt_fold(const uint8_t* rgb, void* xq_hi, void* xq_lo, float* dc, float* sums_scratch,
                                       int64_t n_img, int h, int w, const float* m_rgb2lms_host, const float* m_ipt_host,
                                       void* stream) {
    return rgb_to_ipt_fold_any(rgb, xq_hi, xq_lo, dc, sums_scratch, n_img, h, w, m_rgb2lms_host, m_ipt_host, stream,
                               "rgb_u8_to_ipt_fold");
}

extern "C" int dcta_fold_planes(const float* x, void* xq_hi, void* xq_lo, float* dc, float* sums_scratch,
                                int64_t n_planes, int h, int w, void* stream) {
    DCTA_REQUIRE(x && xq_hi && xq_lo && dc && sums_scratch, "fold_planes: null pointer");
    DCTA_REQUIRE(h >= 2 && w >= 2 && n_planes <= 65535 * 3, "fold_planes: needs h, w >= 2");
    if (n_planes == 0) return DCTA_OK;
    cudaStream_t st = as_stream(stream);
    Mat3 A{}, B{};
    if (fold_fast_dims(h, w) && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const float* mus = launch_plane_means(x, sums_scratch, dc, n_planes, h, w, st);
        fold_planes_kernel<<<grid_for(n_planes * (h / 2) * (w / 8), 256), 256, 0, st>>>(x, mus, (__half*)xq_hi, (__half*)xq_lo,
                                                                                      n_planes, h, w, kFScaleX);
    } else {
        float* mus = sums_scratch;
        means_any_kernel<float, false><<<(unsigned)n_planes, 256, 0, st>>>(x, mus, dc, (int64_t)h * w, A, B, sqrtf((float)h * (float)w));
        fold_any_kernel<float, false><<<grid_for(n_planes * fold_half(h) * fold_half(w), 256), 256, 0, st>>>(
            x, mus, (__half*)xq_hi, (__half*)xq_lo, n_planes, h, w, fold_pitch(w), A, B, kFScaleX);
    }
    return check_launch("fold_planes");
}

// forward: xq[b][a][plane][H2][ldw] (scale 2^6; H2 = ceil(h/2), W2 = ceil(w/2), ldw = round8(W2)) + removed DC
//   -> token grid (tile_p > 0) or planes (n_planes, kh, kw)
//   bw: (2, kw/2, ldw) folded basis of the width transform, group = column parity b; rs_w (2, kw/2) its row factors
//   bh: (2, kh/2, ldh) for the height transform, group = row parity a; work: (2, n_planes, kw, ldh) hi/lo
extern "C" int dcta_dct2_fwd_fold(const void* xq_hi, const void* xq_lo, const float* dc, const void* bw_hi,
                                  const void* bw_lo, const float* rs_w, const void* bh_hi, const void* bh_lo,
                                  const float* rs_h, void* work_hi, void* work_lo, float* y, float* maxabs,
                                  int64_t n_planes, int h, int w, int kh, int kw, int tile_p,
                                  int channels, void* stream) {
    DCTA_REQUIRE(xq_hi && xq_lo && bw_hi && bw_lo && rs_w && bh_hi && bh_lo && rs_h && work_hi && work_lo && y,
                 "dct2_fwd_fold: null pointer");
    DCTA_REQUIRE(maxabs == nullptr || tile_p > 0, "dct2_fwd_fold: maxabs needs the token-grid output");
    DCTA_REQUIRE(fold_dims_ok(h, w, kh, kw), "dct2_fwd_fold: needs h, w >= 2 and even kh <= h, kw <= w");
    if (tile_p > 0)
        DCTA_REQUIRE(channels > 0 && kh % tile_p == 0 && kw % tile_p == 0 && n_planes % channels == 0,
                     "dct2_fwd_fold: kh/kw must be multiples of the patch size");
    if (n_planes == 0) return DCTA_OK;
    const int h2 = fold_half(h), w2 = fold_half(w), ldw = fold_pitch(w), ldh = fold_pitch(h);
    // pass 1: P[(a, plane, h'), j] = sum_w' xq[b][a][plane][h', w'] CW[2j+b, w'], stored as P^T[a][plane][kw = 2j+b][h']
    FoldOperand A1{(const __half*)xq_hi, (const __half*)xq_lo, ldw, 2 * n_planes * (int64_t)h2 * ldw};
    FoldOperand B1{(const __half*)bw_hi, (const __half*)bw_lo, ldw, (int64_t)(kw / 2) * ldw};
    FoldEpi e1{};
    e1.mode = 0; e1.out_hi = (__half*)work_hi; e1.out_lo = (__half*)work_lo;
    e1.rows_per_item = h2; e1.seg_stride = 0; e1.item_stride = (int64_t)kw * ldh;
    e1.col_mul = 2; e1.col_add = 1; e1.col_stride = ldh;
    e1.alpha = kFScaleP / kFScaleX; e1.basis_scale = rs_w;
    // pass 2: Y[2i+a, kw] = sum_h' CH[2i+a, h'] P^T[a][plane][kw][h']   (+ the removed constant's DC at [0,0])
    FoldOperand A2{(const __half*)work_hi, (const __half*)work_lo, ldh, n_planes * (int64_t)kw * ldh};
    FoldOperand B2{(const __half*)bh_hi, (const __half*)bh_lo, ldh, (int64_t)(kh / 2) * ldh};
    FoldEpi e2{};
    e2.out_f32 = y; e2.rows_per_item = kw; e2.col_mul = 2; e2.col_add = 1;
    e2.alpha = 1.0f / kFScaleP; e2.basis_scale = rs_h; e2.dc = dc;
    if (tile_p > 0) {
        e2.mode = 2; e2.p = tile_p; e2.channels = channels; e2.tiles_h = kh / tile_p; e2.tiles_w = kw / tile_p;
        e2.maxabs = maxabs;
        if (maxabs)
            cudaMemsetAsync(maxabs, 0, sizeof(float) * (n_planes / channels) * e2.tiles_h * e2.tiles_w * channels,
                            as_stream(stream));
    } else {
        e2.mode = 1; e2.seg_stride = 0; e2.item_stride = (int64_t)kh * kw; e2.col_stride = kw;
    }
    int rc = launch_fold_gemm(A1, 2 * n_planes * (int64_t)h2, 2, B1, kw / 2, w2, e1, stream);
    if (rc) return rc;
    return launch_fold_gemm(A2, n_planes * (int64_t)kw, 2, B2, kh / 2, h2, e2, stream);
}

extern "C" int dcta_fold_codes_supported(int h, int w, int kh, int kw, int tile_p) {
    if (!dcta_fold_supported(h, w, kh, kw)) return 0;
    if (tile_p < 8 || tile_p > 16 || (tile_p & 1) || kh % tile_p || kw % tile_p || kh / 2 > 256) return 0;
    int tokens = 256 / tile_p;
    while (tokens > 0 && (tokens * tile_p) % 16) --tokens;
    if (tokens < 2) return 0;
    const int64_t basis = ceil_div(fold_half(h), FK) * (int64_t)F_ATILE;
    const int64_t stage = ceil_div((int64_t)tokens * tile_p * 64, 1024) * 1024;
    return ((F_SMEM_LIMIT - 1024 - 2 * basis) / stage >= 3 || (F_SMEM_LIMIT - 1024 - basis) / (stage + F_ATILE) >= 3) &&
           4 * basis < (1 << 20);
}

// forward straight to LFQ code words (one codebook per patch row: c == d == tile_p): the pass-2 epilogue forms
// sign(clamp((Y - median) / (b*sqrt2 + eps))) of every coefficient and never writes the token grid.
//   code_grid (n_planes/channels, kh/p, kw/p, channels, p) int32, maxabs as in dcta_dct2_fwd_fold.
extern "C" int dcta_dct2_fwd_fold_codes(const void* xq_hi, const void* xq_lo, const float* dc, const void* bw_hi,
                                        const void* bw_lo, const float* rs_w, const void* bh_hi, const void* bh_lo,
                                        const float* rs_h, void* work_hi, void* work_lo, float* maxabs,
                                        int32_t* code_grid, const float* median, const float* b, int H, int W,
                                        float eps, float lo, float hi, int32_t* tame_scratch, int tame_known,
                                        int64_t n_planes, int h, int w, int kh, int kw, int tile_p, int channels,
                                        void* stream) {
    DCTA_REQUIRE(xq_hi && xq_lo && bw_hi && bw_lo && rs_w && bh_hi && bh_lo && rs_h && work_hi && work_lo && maxabs &&
                 code_grid && median && b && tame_scratch, "dct2_fwd_fold_codes: null pointer");
    DCTA_REQUIRE(fold_dims_ok(h, w, kh, kw), "dct2_fwd_fold_codes: needs h, w >= 2 and even kh <= h, kw <= w");
    DCTA_REQUIRE(dcta_fold_codes_supported(h, w, kh, kw, tile_p), "dct2_fwd_fold_codes: geometry not supported");
    DCTA_REQUIRE(channels > 0 && n_planes % channels == 0 && kh / tile_p <= H && kw / tile_p <= W,
                 "dct2_fwd_fold_codes: needs a token grid inside the PatchNorm tables");
    if (n_planes == 0) return DCTA_OK;
    const int h2 = fold_half(h), w2 = fold_half(w), ldw = fold_pitch(w), ldh = fold_pitch(h);
    cudaStream_t st = as_stream(stream);
    int rc = tame_known ? DCTA_OK : launch_b_tame(b, (int64_t)channels * H * W * tile_p * tile_p, tame_scratch, st);
    if (rc) return rc;
    FoldOperand A1{(const __half*)xq_hi, (const __half*)xq_lo, ldw, 2 * n_planes * (int64_t)h2 * ldw};
    FoldOperand B1{(const __half*)bw_hi, (const __half*)bw_lo, ldw, (int64_t)(kw / 2) * ldw};
    FoldEpi e1{};
    // pass 1 writes P^T with its lines regrouped as [a][channel][token column][image][pj]: the 16 token columns of a
    // fold_codes_kernel tile are then 16 images at the SAME (channel, tw), i.e. they share their PatchNorm medians
    e1.mode = 3; e1.out_hi = (__half*)work_hi; e1.out_lo = (__half*)work_lo;
    e1.rows_per_item = h2; e1.col_mul = 2; e1.col_add = 1; e1.col_stride = ldh;
    e1.p = tile_p; e1.channels = channels; e1.tiles_w = kw / tile_p; e1.batch = (int)(n_planes / channels);
    e1.alpha = kFScaleP / kFScaleX; e1.basis_scale = rs_w;
    rc = launch_fold_gemm(A1, 2 * n_planes * (int64_t)h2, 2, B1, kw / 2, w2, e1, stream);
    if (rc) return rc;
    FoldOperand A2{(const __half*)work_hi, (const __half*)work_lo, ldh, n_planes * (int64_t)kw * ldh};
    FoldOperand B2{(const __half*)bh_hi, (const __half*)bh_lo, ldh, (int64_t)(kh / 2) * ldh};
    const int64_t n_tok = (n_planes / channels) * (kh / tile_p) * (kw / tile_p) * channels;
    cudaMemsetAsync(maxabs, 0, sizeof(float) * n_tok, st);
    CodesArgs cg{};
    cg.p = tile_p; cg.channels = channels; cg.tiles_h = kh / tile_p; cg.tiles_w = kw / tile_p;
    cg.batch = (int)(n_planes / channels);
    cg.alpha = 1.0f / kFScaleP; cg.basis_scale = rs_h; cg.dc = dc; cg.med = median; cg.bstat = b;
    cg.stat_h = H; cg.stat_w = W; cg.eps = eps; cg.clamp_lo = lo; cg.clamp_hi = hi; cg.tame = tame_scratch;
    cg.maxabs = maxabs; cg.code_grid = code_grid;
    rc = launch_fold_codes(A2, n_planes * (int64_t)kw, B2, kh / 2, h2, cg, stream);
    if (rc == DCTA_ERR_UNSUPPORTED) set_error("dct2_fwd_fold_codes: geometry not supported (see dcta_fold_codes_supported)");
    return rc;
}

// inverse: yq[b][a][plane][kh/2][ldq] (scale 2^4, DC removed) -> z[b*2+a][plane][H2][W2] fp32 quadrant transforms
//   (H2 = ceil(h/2), W2 = ceil(w/2));  bwt: (2, W2, ldq) = CW[2j+b, w']^T, bht: (2, H2, ldi) = CH[2i+a, h']^T;
//   work: (2, 2, n_planes, W2, ldi) hi/lo
extern "C" int dcta_dct2_inv_fold(const void* yq_hi, const void* yq_lo, const void* bwt_hi, const void* bwt_lo,
                                  const void* bht_hi, const void* bht_lo, void* work_hi, void* work_lo, float* z,
                                  int64_t n_planes, int h, int w, int kh, int kw, void* stream) {
    DCTA_REQUIRE(yq_hi && yq_lo && bwt_hi && bwt_lo && bht_hi && bht_lo && work_hi && work_lo && z,
                 "dct2_inv_fold: null pointer");
    DCTA_REQUIRE(fold_dims_ok(h, w, kh, kw), "dct2_inv_fold: needs h, w >= 2 and even kh <= h, kw <= w");
    if (n_planes == 0) return DCTA_OK;
    const int h2 = fold_half(h), w2 = fold_half(w), kh2 = kh / 2, kw2 = kw / 2;
    const int64_t ldq = ceil_div(kw2, 8) * 8, ldi = ceil_div(kh2, 8) * 8;
    // pass 1: Q[(a, plane, i), w'] = sum_j yq[b][a][plane][i, j] CW[2j+b, w'], stored as Q^T[b][a][plane][w'][i]
    FoldOperand A1{(const __half*)yq_hi, (const __half*)yq_lo, ldq, 2 * n_planes * (int64_t)kh2 * ldq};
    FoldOperand B1{(const __half*)bwt_hi, (const __half*)bwt_lo, ldq, (int64_t)w2 * ldq};
    FoldEpi e1{};
    e1.mode = 0; e1.out_hi = (__half*)work_hi; e1.out_lo = (__half*)work_lo;
    e1.rows_per_item = kh2; e1.seg_stride = 2 * n_planes * (int64_t)w2 * ldi; e1.item_stride = (int64_t)w2 * ldi;
    e1.col_mul = 1; e1.col_add = 0; e1.col_stride = (int)ldi;
    e1.alpha = kFScaleQ / (kFScaleBasis * kFScaleY);
    // pass 2: Z[(b, a)][plane][h', w'] = sum_i CH[2i+a, h'] Q^T[b][a][plane][w'][i]
    FoldOperand A2{(const __half*)work_hi, (const __half*)work_lo, ldi, n_planes * (int64_t)w2 * ldi};
    FoldOperand B2{(const __half*)bht_hi, (const __half*)bht_lo, ldi, (int64_t)h2 * ldi};
    FoldEpi e2{};
    e2.mode = 1; e2.out_f32 = z; e2.rows_per_item = w2;
    e2.seg_stride = n_planes * (int64_t)h2 * w2; e2.item_stride = (int64_t)h2 * w2;
    e2.col_mul = 1; e2.col_add = 0; e2.col_stride = w2;
    e2.alpha = 1.0f / (kFScaleBasis * kFScaleQ);
    int rc = launch_fold_gemm(A1, 2 * n_planes * (int64_t)kh2, 2, B1, w2, kw2, e1, stream);
    if (rc) return rc;
    return launch_fold_gemm(A2, n_planes * (int64_t)w2, 4, B2, h2, kh2, e2, stream);
}

// ---- decode from LFQ codes INSIDE inverse pass 1: the coefficient quadrants yq are never materialised
static void decode_inv_sizes(int64_t n_img, int C, int kh, int kw, int& n_iblk, int& num_kb, int64_t& tab_bytes, int64_t& bv_bytes) {
    n_iblk = (int)ceil_div(kh / 2, 32);
    num_kb = (int)ceil_div(kw / 2, FK);
    tab_bytes = (int64_t)4 * C * n_iblk * num_kb * 8192;
    bv_bytes = (int64_t)4 * ((n_img + 3) / 4 * 4) * C * n_iblk * num_kb * 256;      // 32 rows x 4 chunks x 2 bytes per (i block, k block)
}

extern "C" int64_t dcta_decode_gen_tables_bytes(int channels_n, int kh, int kw) {
    int n_iblk, num_kb;
    int64_t tab_bytes, bv_bytes;
    decode_inv_sizes(0, channels_n, kh, kw, n_iblk, num_kb, tab_bytes, bv_bytes);
    return tab_bytes;
}

extern "C" int64_t dcta_decode_codes_inv_fold_scratch_bytes(int64_t n_img, int channels_n, int kh, int kw) {
    int n_iblk, num_kb;
    int64_t tab_bytes, bv_bytes;
    decode_inv_sizes(n_img, channels_n, kh, kw, n_iblk, num_kb, tab_bytes, bv_bytes);
    return bv_bytes;
}

extern "C" int dcta_decode_gen_tables(const float* median, const float* b, int channels_n, int H, int W, float eps, int p,
                                      int kh, int kw, float scale, void* tab, void* stream) {
    DCTA_REQUIRE(median && b && tab, "decode_gen_tables: null pointer");
    DCTA_REQUIRE(p >= 8 && p <= 16 && !(p & 1) && kh > 0 && kw > 0 && kh % p == 0 && kw % p == 0 && kh / p <= H && kw / p <= W,
                 "decode_gen_tables: needs an even patch size in 8..16 and a token grid inside the PatchNorm tables");
    int n_iblk, num_kb;
    int64_t tab_bytes, bv_bytes;
    decode_inv_sizes(0, channels_n, kh, kw, n_iblk, num_kb, tab_bytes, bv_bytes);
    LfqNormParams q{median, b, channels_n, H, W, p * p, eps, 0.f, 0.f, p, p, scale};
    decode_gen_tables_kernel<<<grid_for(tab_bytes / 4, 256), 256, 0, as_stream(stream)>>>(q, p, kh, kw, n_iblk, num_kb, kFScaleY,
                                                                                          reinterpret_cast<uint32_t*>(tab));
    return check_launch("decode_gen_tables");
}

extern "C" int dcta_decode_codes_inv_fold_supported(int h, int w, int kh, int kw, int p, int c, int d) {
    if (!fold_dims_ok(h, w, kh, kw) || p < 8 || p > 16 || (p & 1) || c != p || d != p || kh % p || kw % p) return 0;
    FoldGemm g{};
    return fold_geometry(fold_half(w), kw / 2, g, 0, F_GEN_STAGES * F_GEN_STAGE) && fold_geometry(fold_half(h), kh / 2, g);
}

static int decode_inv_fold_impl(const int64_t* codes, const int32_t* slot_map, const int32_t* img_sel, const int32_t* code_grid,
                                int64_t n_img, int channels_n, int th, int tw, int p, int kh, int kw, int h, int w,
                                const float* median, const float* b, int H, int W, float eps, int c, int d, float scale,
                                const void* bwt_hi, const void* bwt_lo, const void* bht_hi, const void* bht_lo, void* work_hi,
                                void* work_lo, float* z, float* dc, const void* tab_in, void* scratch, void* stream) {
    DCTA_REQUIRE((code_grid || (codes && slot_map)) && median && b && bwt_hi && bwt_lo && bht_hi && bht_lo && work_hi && work_lo &&
                 z && dc && tab_in && scratch, "decode_codes_inv_fold: null pointer");
    DCTA_REQUIRE(dcta_decode_codes_inv_fold_supported(h, w, kh, kw, p, c, d),
                 "decode_codes_inv_fold: needs one LFQ codebook per patch row (c == d == p, p even in 8..16), even kh <= h, kw <= w");
    DCTA_REQUIRE(th <= H && tw <= W && kh / p <= H && kw / p <= W && n_img * channels_n < (1 << 24),
                 "decode_codes_inv_fold: token grid outside the PatchNorm tables");
    if (n_img == 0) return DCTA_OK;
    const int64_t n_planes = n_img * channels_n;
    const int h2 = fold_half(h), w2 = fold_half(w), kh2 = kh / 2, kw2 = kw / 2;
    const int64_t ldq = ceil_div(kw2, 8) * 8, ldi = ceil_div(kh2, 8) * 8;
    int n_iblk, num_kb;
    int64_t tab_bytes, bv_bytes;
    decode_inv_sizes(n_img, channels_n, kh, kw, n_iblk, num_kb, tab_bytes, bv_bytes);
    const uint32_t* tab = reinterpret_cast<const uint32_t*>(tab_in);
    uint16_t* bv = reinterpret_cast<uint16_t*>(scratch);
    LfqNormParams q{median, b, channels_n, H, W, p * p, eps, 0.f, 0.f, c, d, scale};
    {
        const int n_grp = (int)ceil_div(kw / p, 8);
        DCTA_REQUIRE(n_iblk <= 65535, "decode_codes_inv_fold: plane too tall for the bit-plane kernel");
        const dim3 grid((unsigned)n_planes, (unsigned)n_iblk);
        int threads = 64 * n_grp;
        threads = threads > 512 ? 512 : (threads + 31) / 32 * 32;
        const int stage_bytes = 4 * num_kb * 256;
        const float dcf = 1.0f / sqrtf((float)h * (float)w);
#define DCTA_BITPLANES(HW)                                                                                                   \
    codes_bitplanes_kernel<HW><<<grid, threads, stage_bytes, as_stream(stream)>>>(codes, slot_map, img_sel, n_img, channels_n, th, \
                                                                                  tw, kh, kw, n_iblk, num_kb, n_grp, q, bv, dc, dcf)
#define DCTA_BITPLANES_GRID(HW)                                                                                              \
    codes_bitplanes_grid_kernel<HW><<<grid, threads, stage_bytes, as_stream(stream)>>>(code_grid, n_img, channels_n, th, tw, kh, \
                                                                                       kw, n_iblk, num_kb, n_grp, q, bv, dc, dcf)
        if (code_grid) {
            switch (p / 2) {
                case 4: DCTA_BITPLANES_GRID(4); break;
                case 5: DCTA_BITPLANES_GRID(5); break;
                case 6: DCTA_BITPLANES_GRID(6); break;
                case 7: DCTA_BITPLANES_GRID(7); break;
                default: DCTA_BITPLANES_GRID(8); break;
            }
        } else {
            switch (p / 2) {
                case 4: DCTA_BITPLANES(4); break;
                case 5: DCTA_BITPLANES(5); break;
                case 6: DCTA_BITPLANES(6); break;
                case 7: DCTA_BITPLANES(7); break;
                default: DCTA_BITPLANES(8); break;
            }
        }
#undef DCTA_BITPLANES_GRID
#undef DCTA_BITPLANES
    }
    int rc = check_launch("decode_codes_inv_fold (bit planes)");
    if (rc) return rc;
    // pass 1 with the generated operand: Q^T[b][a][plane][w'][i] = sum_j Y[2i+a, 2j+b] CW[2j+b, w']
    FoldGen gen{};
    gen.tab = reinterpret_cast<const uint4*>(tab);
    gen.bv = bv;
    gen.n_img = (int)n_img; gen.C = channels_n; gen.kh2 = kh2; gen.n_iblk = n_iblk; gen.n_planes = n_planes;
    gen.n_work = (int)(ceil_div(n_img, 8) * 2 * channels_n * n_iblk);
    FoldOperand A1{nullptr, nullptr, ldq, 0};
    FoldOperand B1{(const __half*)bwt_hi, (const __half*)bwt_lo, ldq, (int64_t)w2 * ldq};
    FoldEpi e1{};
    e1.mode = 0; e1.out_hi = (__half*)work_hi; e1.out_lo = (__half*)work_lo;
    e1.rows_per_item = kh2; e1.seg_stride = 2 * n_planes * (int64_t)w2 * ldi; e1.item_stride = (int64_t)w2 * ldi;
    e1.col_mul = 1; e1.col_add = 0; e1.col_stride = (int)ldi;
    e1.alpha = kFScaleQ / (kFScaleBasis * kFScaleY);
    rc = launch_fold_gemm(A1, 2 * n_planes * (int64_t)kh2, 2, B1, w2, kw2, e1, stream, &gen);
    if (rc) return rc;
    // pass 2 as in dcta_dct2_inv_fold
    FoldOperand A2{(const __half*)work_hi, (const __half*)work_lo, ldi, n_planes * (int64_t)w2 * ldi};
    FoldOperand B2{(const __half*)bht_hi, (const __half*)bht_lo, ldi, (int64_t)h2 * ldi};
    FoldEpi e2{};
    e2.mode = 1; e2.out_f32 = z; e2.rows_per_item = w2;
    e2.seg_stride = n_planes * (int64_t)h2 * w2; e2.item_stride = (int64_t)h2 * w2;
    e2.col_mul = 1; e2.col_add = 0; e2.col_stride = w2;
    e2.alpha = 1.0f / (kFScaleBasis * kFScaleQ);
    return launch_fold_gemm(A2, n_planes * (int64_t)w2, 4, B2, h2, kh2, e2, stream);
}

extern "C" int dcta_decode_codes_inv_fold(const int64_t* codes, const int32_t* slot_map, const int32_t* img_sel, int64_t n_img,
                                          int channels_n, int th, int tw, int p, int kh, int kw, int h, int w,
                                          const float* median, const float* b, int H, int W, float eps, int c, int d,
                                          float scale, const void* bwt_hi, const void* bwt_lo, const void* bht_hi,
                                          const void* bht_lo, void* work_hi, void* work_lo, float* z, float* dc,
                                          const void* tab_in, void* scratch, void* stream) {
    DCTA_REQUIRE(codes && slot_map, "decode_codes_inv_fold: null pointer");
    return decode_inv_fold_impl(codes, slot_map, img_sel, nullptr, n_img, channels_n, th, tw, p, kh, kw, h, w, median, b, H, W, eps,
                                c, d, scale, bwt_hi, bwt_lo, bht_hi, bht_lo, work_hi, work_lo, z, dc, tab_in, scratch, stream);
}

// The same from the code grid of dcta_dct2_fwd_fold_codes (every token kept: a round trip without a top-k cut)
extern "C" int dcta_decode_grid_inv_fold(const int32_t* code_grid, int64_t n_img, int channels_n, int p, int kh, int kw, int h,
                                         int w, const float* median, const float* b, int H, int W, float eps, float scale,
                                         const void* bwt_hi, const void* bwt_lo, const void* bht_hi, const void* bht_lo,
                                         void* work_hi, void* work_lo, float* z, float* dc, const void* tab_in, void* scratch,
                                         void* stream) {
    DCTA_REQUIRE(code_grid && p > 0 && kh % p == 0 && kw % p == 0, "decode_grid_inv_fold: bad arguments");
    return decode_inv_fold_impl(nullptr, nullptr, nullptr, code_grid, n_img, channels_n, kh / p, kw / p, p, kh, kw, h, w, median, b,
                                H, W, eps, p, p, scale, bwt_hi, bwt_lo, bht_hi, bht_lo, work_hi, work_lo, z, dc, tab_in, scratch,
                                stream);
}

template <typename TOut>
static int unfold_ipt_to_rgb_any(const float* z, const float* dc, TOut* rgb, int64_t n_img, int h, int w,
                                 const float* m_ipt_inv_host, const float* m_lms2rgb_host, void* stream, const char* who) {
    DCTA_REQUIRE(z && rgb && m_ipt_inv_host && m_lms2rgb_host, "%s: null pointer", who);
    DCTA_REQUIRE(h >= 2 && w >= 2, "%s: bad sizes", who);
    if (n_img == 0) return DCTA_OK;
    Mat3 A, B;
    for (int i = 0; i < 9; ++i) { A.m[i] = m_ipt_inv_host[i]; B.m[i] = m_lms2rgb_host[i]; }
    if (fold_fast_dims(h, w) && (reinterpret_cast<uintptr_t>(rgb) & 15) == 0)
        for (int64_t i0 = 0; i0 < n_img; i0 += 65535)          // grid.y limit
            unfold_kernel<true, TOut><<<dim3((unsigned)ceil_div((int64_t)(h / 2) * (w / 8), 256), (unsigned)std::min<int64_t>(65535, n_img - i0)),
                                        256, 0, as_stream(stream)>>>(z, dc, rgb, n_img, h, w, A, B, i0);
    else
        unfold_any_kernel<true, TOut><<<grid_for(n_img * fold_half(h) * fold_half(w), 256), 256, 0, as_stream(stream)>>>(
            z, dc, rgb, n_img, h, w, A, B);
    return check_launch(who);
}

extern "C" int dcta_unfold_ipt_to_rgb(const float* z, const float* dc, float* rgb, int64_t n_img, int h, int w,
                                      const float* m_ipt_inv_host, const float* m_lms2rgb_host, void* stream) {
    return unfold_ipt_to_rgb_any(z, dc, rgb, n_img, h, w, m_ipt_inv_host, m_lms2rgb_host, stream, "unfold_ipt_to_rgb");
}

extern "C" int dcta_unfold_ipt_to_rgb_u8(const float* z, const float* dc, uint8_t* rgb, int64_t n_img, int h, int w,
                                         const float* m_ipt_inv_host, const float* m_lms2rgb_host, void* stream) {
    return unfold_ipt_to_rgb_any(z, dc, rgb, n_img, h, w, m_ipt_inv_host, m_lms2rgb_host, stream, "unfold_ipt_to_rgb_u8");
}

extern "C" int dcta_unfold_planes(const float* z, const float* dc, float* x, int64_t n_planes, int h, int w, void* stream) {
    DCTA_REQUIRE(z && x, "unfold_planes: null pointer");
    DCTA_REQUIRE(h >= 2 && w >= 2, "unfold_planes: bad sizes");
    if (n_planes == 0) return DCTA_OK;
    Mat3 A{}, B{};
    if (fold_fast_dims(h, w) && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
        for (int64_t i0 = 0; i0 < n_planes; i0 += 65535)       // grid.y limit
            unfold_kernel<false><<<dim3((unsigned)ceil_div((int64_t)(h / 2) * (w / 8), 256), (unsigned)std::min<int64_t>(65535, n_planes - i0)),
                                   256, 0, as_stream(stream)>>>(z, dc, x, n_planes, h, w, A, B, i0);
    else
        unfold_any_kernel<false, float><<<grid_for(n_planes * fold_half(h) * fold_half(w), 256), 256, 0, as_stream(stream)>>>(
            z, dc, x, n_planes, h, w, A, B);
    return check_launch("unfold_planes");
}

extern "C" int dcta_fold_coef_planes(const float* y, void* yq_hi, void* yq_lo, float* dc, int64_t n_planes, int kh,
                                     int kw, int out_h, int out_w, void* stream) {
    DCTA_REQUIRE(y && yq_hi && yq_lo && dc && kh > 0 && kw > 0 && kh % 2 == 0 && kw % 2 == 0, "fold_coef_planes: bad args");
    if (n_planes == 0) return DCTA_OK;
    const int ldq = (int)(ceil_div(kw / 2, 8) * 8);
    fold_coef_kernel<<<grid_for(4 * n_planes * (kh / 2) * ldq, 256), 256, 0, as_stream(stream)>>>(
        y, (__half*)yq_hi, (__half*)yq_lo, dc, n_planes, kh, kw, ldq, 1.0f / sqrtf((float)out_h * (float)out_w), kFScaleY);
    return check_launch("fold_coef_planes");
}

static int launch_unpatchify_fold(bool with_codes, const float* patches, const int64_t* codes, const int32_t* slot_map,
                                  const int32_t* img_sel, int64_t n_img, int C, int th, int tw, int p, int rows, int cols,
                                  int out_h, int out_w, const LfqNormParams& q, void* yq_hi, void* yq_lo, float* dc,
                                  void* stream, const char* who, uint32_t* tab_scratch = nullptr, bool denorm = false) {
    DCTA_REQUIRE(slot_map && yq_hi && yq_lo && dc, "%s: null pointer", who);
    DCTA_REQUIRE(rows > 0 && cols > 0 && p > 0 && out_h > 0 && out_w > 0 && rows % 2 == 0 && cols % 2 == 0 && rows % p == 0,
                 "%s: needs even plane sizes and rows %% patch == 0", who);
    if (n_img == 0) return DCTA_OK;
    DCTA_REQUIRE(p >= 8, "%s: tiles must be at least 8 columns wide", who);
    const int ldq = (int)(ceil_div(cols / 2, 8) * 8);
    const int nxv = ldq / 4;
    // threads: x = groups of 8 columns, y = images handled side by side
    const int bx = nxv < 256 ? nxv : 256;
    int by = 256 / bx;
    if (by < 1) by = 1;
    if (by > n_img) by = (int)n_img;
    const int imgs_per_cta = (int)(n_img < 2 * by ? n_img : 2 * by);
    const int64_t n_ctas = ceil_div(n_img, imgs_per_cta) * C * (rows / p);
    DCTA_REQUIRE(n_ctas < (1ll << 31), "%s: too many plane rows for one launch", who);
    const float dcf = 1.0f / sqrtf((float)out_h * (float)out_w);
    const int smem_bytes = with_codes ? 2 * p * 8 * nxv * (int)sizeof(float) : 0;
    DCTA_REQUIRE(smem_bytes <= 200 * 1024, "%s: tile row of %d bytes does not fit in shared memory", who, smem_bytes);
    const bool fast = with_codes && q.d == p && p <= 16;      // one codebook per patch row
    const dim3 block((unsigned)bx, (unsigned)by);
    cudaError_t e = cudaSuccess;
    if (with_codes && fast) {
        const int nxv_tab = (int)ceil_div((int64_t)q.W * p, 8);
        if (tab_scratch != nullptr && nxv <= nxv_tab) {
            const int64_t n_tab = (int64_t)q.C * q.H * p * nxv_tab * 8;
            decode_tables_kernel<<<grid_for(n_tab, 256), 256, 0, as_stream(stream)>>>(q, p, nxv_tab, kFScaleY, tab_scratch,
                                                                                      tab_scratch + n_tab);
            e = cudaFuncSetAttribute(decode_codes_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
            if (e == cudaSuccess)
                decode_codes_rows_kernel<true><<<(unsigned)n_ctas, block, smem_bytes, as_stream(stream)>>>(
                    codes, slot_map, img_sel, n_img, imgs_per_cta, C, th, tw, p, rows, cols, ldq, q, tab_scratch,
                    tab_scratch + n_tab, nxv_tab, (__half*)yq_hi, (__half*)yq_lo, dc, dcf, kFScaleY);
        } else {
            e = cudaFuncSetAttribute(decode_codes_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
            if (e == cudaSuccess)
                decode_codes_rows_kernel<false><<<(unsigned)n_ctas, block, smem_bytes, as_stream(stream)>>>(
                    codes, slot_map, img_sel, n_img, imgs_per_cta, C, th, tw, p, rows, cols, ldq, q, nullptr, nullptr, 0,
                    (__half*)yq_hi, (__half*)yq_lo, dc, dcf, kFScaleY);
        }
    } else if (with_codes) {
        e = cudaFuncSetAttribute(unpatchify_fold_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e == cudaSuccess)
            unpatchify_fold_kernel<true, false><<<(unsigned)n_ctas, block, smem_bytes, as_stream(stream)>>>(
                patches, codes, slot_map, img_sel, n_img, imgs_per_cta, C, th, tw, p, rows, cols, ldq, q, (__half*)yq_hi,
                (__half*)yq_lo, dc, dcf, kFScaleY);
    } else if (denorm) {
        unpatchify_fold_kernel<false, false, true><<<(unsigned)n_ctas, block, 0, as_stream(stream)>>>(
            patches, codes, slot_map, img_sel, n_img, imgs_per_cta, C, th, tw, p, rows, cols, ldq, q, (__half*)yq_hi,
            (__half*)yq_lo, dc, dcf, kFScaleY);
    } else {
        unpatchify_fold_kernel<false, false><<<(unsigned)n_ctas, block, 0, as_stream(stream)>>>(
            patches, codes, slot_map, img_sel, n_img, imgs_per_cta, C, th, tw, p, rows, cols, ldq, q, (__half*)yq_hi,
            (__half*)yq_lo, dc, dcf, kFScaleY);
    }
    if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return DCTA_ERR_LAUNCH; }
    return check_launch(who);
}

extern "C" int dcta_unpatchify_fold(const float* patches, const int32_t* slot_map, const int32_t* img_sel, int64_t n_img,
                                    int channels_n, int th, int tw, int p, int rows, int cols, int out_h, int out_w,
                                    void* yq_hi, void* yq_lo, float* dc, void* stream) {
    DCTA_REQUIRE(patches, "unpatchify_fold: null pointer");
    LfqNormParams q{};
    return launch_unpatchify_fold(false, patches, nullptr, slot_map, img_sel, n_img, channels_n, th, tw, p, rows, cols,
                                  out_h, out_w, q, yq_hi, yq_lo, dc, stream, "unpatchify_fold");
}

extern "C" int dcta_unpatchify_denorm_fold(const float* patches, const int32_t* slot_map, const int32_t* img_sel, int64_t n_img,
                                           int channels_n, int th, int tw, int p, int rows, int cols, int out_h, int out_w,
                                           const float* median, const float* b, int H, int W, float eps, void* yq_hi,
                                           void* yq_lo, float* dc, void* stream) {
    DCTA_REQUIRE(patches && median && b, "unpatchify_denorm_fold: null pointer");
    DCTA_REQUIRE(th <= H && tw <= W && rows / p <= H, "unpatchify_denorm_fold: the token grid must lie inside the PatchNorm tables");
    DCTA_REQUIRE(p % 2 == 0 && ((reinterpret_cast<uintptr_t>(median) | reinterpret_cast<uintptr_t>(b)) & 7) == 0,
                 "unpatchify_denorm_fold: needs an even patch size and 8-byte aligned tables");
    LfqNormParams q{median, b, channels_n, H, W, p * p, eps, 0.f, 0.f, 1, p * p, 1.0f};
    return launch_unpatchify_fold(false, patches, nullptr, slot_map, img_sel, n_img, channels_n, th, tw, p, rows, cols,
                                  out_h, out_w, q, yq_hi, yq_lo, dc, stream, "unpatchify_denorm_fold", nullptr, true);
}

extern "C" int dcta_decode_codes_fold(const int64_t* codes, const int32_t* slot_map, const int32_t* img_sel,
                                      int64_t n_img, int channels_n, int th, int tw, int p, int rows, int cols, int out_h,
                                      int out_w, const float* median, const float* b, int H, int W, float eps, int c,
                                      int d, float scale, void* yq_hi, void* yq_lo, float* dc, void* tab_scratch,
                                      void* stream) {
    DCTA_REQUIRE(codes && median && b, "decode_codes_fold: null pointer");
    DCTA_REQUIRE(th <= H && tw <= W && c > 0 && d > 0 && d <= 62 && c * d == p * p && rows / p <= H,
                 "decode_codes_fold: needs a projection-free LFQ (c*d == p*p) and a token grid inside the PatchNorm tables");
    LfqNormParams q{median, b, channels_n, H, W, p * p, eps, 0.f, 0.f, c, d, scale};
    return launch_unpatchify_fold(true, nullptr, codes, slot_map, img_sel, n_img, channels_n, th, tw, p, rows, cols,
                                  out_h, out_w, q, yq_hi, yq_lo, dc, stream, "decode_codes_fold", (uint32_t*)tab_scratch);
}
