// Truncated orthonormal 2-D DCT-II / DCT-III as two batched basis-matrix GEMMs per direction
// (reference: util.py:333-338 via torch_dct, feature_extraction_dct_autoencoder.py:140,149).
//
//   forward : P = X . CW[:kw]^T   (h x kw),   Y = CH[:kh] . P        (kh x kw)
//   inverse : Q = CH[:kh]^T . Y   (h x kw),   X = Q . CW[:kw]        (h x w)
//
// This file is the exact-fp32 path: a register-tiled FFMA SGEMM (128x128x8 CTA tile, 8x8 per
// thread, double-buffered shared memory).  It bounds at the fp32 FMA pipe, not at HBM; the
// tensor-core split-precision path replaces it where its tolerance is accepted.
#include "common.cuh"

namespace dcta {

constexpr int BM = 128, BN = 128, BK = 8, PAD = 4;

struct GemmArgs {
    const float* A;
    const float* B;
    float* C;
    int M, N, K;
    int64_t sAm, sAk, sAb;  // element strides of A[m,k] and its batch stride
    int64_t sBk, sBn, sBb;  // element strides of B[k,n] and its batch stride
    int64_t ldc, sCb;       // MODE 0: C[b][m*ldc + n]
    int tile_p, channels, tiles_h, tiles_w;  // MODE 1: token-grid layout (FE:374-380)
    int vec_store;          // MODE 0: 128-bit stores allowed
    int64_t batch0;         // first batch index of this launch (blockIdx.z is 16-bit)
};

// kAK: A is K-contiguous (sAk == 1) else M-contiguous (sAm == 1)
// kBN: B is N-contiguous (sBn == 1) else K-contiguous (sBk == 1)
template <bool kAK, bool kBN, int kMode>
__global__ void __launch_bounds__(256, 2) sgemm_tile_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int64_t batch = g.batch0 + blockIdx.z;
    const float* __restrict__ A = g.A + batch * g.sAb;
    const float* __restrict__ B = g.B + batch * g.sBb;

    // global -> register staging (4 elements of A and 4 of B per thread per k-block)
    float ra[4], rb[4];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int m, k;
            if (kAK) { m = tid >> 1; k = (tid & 1) * 4 + j; }
            else     { m = tid & 127; k = (tid >> 7) + 2 * j; }
            const int gm = m0 + m, gk = k0 + k;
            ra[j] = (gm < g.M && gk < g.K) ? __ldg(A + gm * g.sAm + gk * g.sAk) : 0.0f;
            int n, kb;
            if (kBN) { n = tid & 127; kb = (tid >> 7) + 2 * j; }
            else     { n = tid >> 1; kb = (tid & 1) * 4 + j; }
            const int gn = n0 + n, gkb = k0 + kb;
            rb[j] = (gn < g.N && gkb < g.K) ? __ldg(B + gkb * g.sBk + gn * g.sBn) : 0.0f;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int m, k;
            if (kAK) { m = tid >> 1; k = (tid & 1) * 4 + j; }
            else     { m = tid & 127; k = (tid >> 7) + 2 * j; }
            As[buf][k][m] = ra[j];
            int n, kb;
            if (kBN) { n = tid & 127; kb = (tid >> 7) + 2 * j; }
            else     { n = tid >> 1; kb = (tid & 1) * 4 + j; }
            Bs[buf][kb][n] = rb[j];
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    const int nk = (g.K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kb = 0; kb < nk; ++kb) {
        const int cur = kb & 1;
        if (kb + 1 < nk) load_tiles((kb + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][kk][64 + tx * 4]);
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

    // epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int n = n0 + half * 64 + tx * 4;
            if (n >= g.N) continue;
            const float* v = &acc[i][half * 4];
            if (kMode == 0) {
                float* dst = g.C + batch * g.sCb + (int64_t)m * g.ldc + n;
                if (g.vec_store && n + 3 < g.N) {
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < g.N) dst[j] = v[j];
                }
            } else {
                // token-grid layout: (img, th, tw, c, p*p)
                const int p = g.tile_p;
                const int64_t img = batch / g.channels;
                const int c = (int)(batch - img * g.channels);
                const int th = m / p, pi = m - th * p;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int nn = n + j;
                    if (nn >= g.N) break;
                    const int tw = nn / p, pj = nn - tw * p;
                    const int64_t tok = ((img * g.tiles_h + th) * g.tiles_w + tw) * g.channels + c;
                    g.C[tok * (p * p) + pi * p + pj] = v[j];
                }
            }
        }
    }
}

template <bool kAK, bool kBN, int kMode>
int launch_gemm(const GemmArgs& g, int64_t batch, void* stream) {
    if (g.M == 0 || g.N == 0 || batch == 0) return DCTA_OK;
    const int64_t max_z = 65535;
    // blockIdx.z is limited to 65535: walk the batch in slabs
    for (int64_t b0 = 0; b0 < batch; b0 += max_z) {
        GemmArgs s = g;
        s.batch0 = b0;
        const int64_t nb = (batch - b0 < max_z) ? batch - b0 : max_z;
        dim3 grid((unsigned)ceil_div(g.N, BN), (unsigned)ceil_div(g.M, BM), (unsigned)nb);
        sgemm_tile_kernel<kAK, kBN, kMode><<<grid, 256, 0, as_stream(stream)>>>(s);
    }
    return check_launch("dct gemm");
}

static inline int aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace dcta

using namespace dcta;

extern "C" int dcta_dct2_fwd(const float* x, const float* ch, const float* cw, float* work,
                             float* y, int64_t n_planes, int h, int w, int kh, int kw, int tile_p,
                             int channels, void* stream) {
    DCTA_REQUIRE(x && ch && cw && work && y, "dct2_fwd: null pointer");
    DCTA_REQUIRE(n_planes >= 0 && h > 0 && w > 0 && kh > 0 && kw > 0 && kh <= h && kw <= w,
                 "dct2_fwd: bad sizes h=%d w=%d kh=%d kw=%d", h, w, kh, kw);
    if (tile_p > 0) {
        DCTA_REQUIRE(channels > 0 && kh % tile_p == 0 && kw % tile_p == 0 && n_planes % channels == 0,
                     "dct2_fwd: kh/kw must be multiples of the patch size and n_planes of channels");
    }
    // pass 1: P[b] (h x kw) = X[b] (h x w) . CW^T   -- A = X (k-contig), B[k=w][n=kw] = CW[kw][w] (k-contig)
    GemmArgs g1{};
    g1.A = x; g1.B = cw; g1.C = work;
    g1.M = h; g1.N = kw; g1.K = w;
    g1.sAm = w; g1.sAk = 1; g1.sAb = (int64_t)h * w;
    g1.sBk = 1; g1.sBn = w; g1.sBb = 0;
    g1.ldc = kw; g1.sCb = (int64_t)h * kw;
    g1.vec_store = (kw % 4 == 0) && aligned16(work);
    int rc = launch_gemm<true, false, 0>(g1, n_planes, stream);
    if (rc) return rc;
    // pass 2: Y[b] (kh x kw) = CH (kh x h) . P[b] (h x kw) -- A = CH (k-contig, shared), B = P (n-contig)
    GemmArgs g2{};
    g2.A = ch; g2.B = work; g2.C = y;
    g2.M = kh; g2.N = kw; g2.K = h;
    g2.sAm = h; g2.sAk = 1; g2.sAb = 0;
    g2.sBk = kw; g2.sBn = 1; g2.sBb = (int64_t)h * kw;
    g2.ldc = kw; g2.sCb = (int64_t)kh * kw;
    g2.vec_store = (kw % 4 == 0) && aligned16(y);
    if (tile_p > 0) {
        g2.tile_p = tile_p; g2.channels = channels;
        g2.tiles_h = kh / tile_p; g2.tiles_w = kw / tile_p;
        return launch_gemm<true, true, 1>(g2, n_planes, stream);
    }
    return launch_gemm<true, true, 0>(g2, n_planes, stream);
}

extern "C" int dcta_dct2_inv(const float* y, const float* ch, const float* cw, float* work,
                             float* x, int64_t n_planes, int h, int w, int kh, int kw,
                             void* stream) {
    DCTA_REQUIRE(x && ch && cw && work && y, "dct2_inv: null pointer");
    DCTA_REQUIRE(n_planes >= 0 && h > 0 && w > 0 && kh > 0 && kw > 0 && kh <= h && kw <= w,
                 "dct2_inv: bad sizes h=%d w=%d kh=%d kw=%d", h, w, kh, kw);
    // pass 1: Q[b] (h x kw) = CH^T (h x kh) . Y[b] (kh x kw) -- A[m=h][k=kh] = CH[kh][h] (m-contig), B = Y (n-contig)
    GemmArgs g1{};
    g1.A = ch; g1.B = y; g1.C = work;
    g1.M = h; g1.N = kw; g1.K = kh;
    g1.sAm = 1; g1.sAk = h; g1.sAb = 0;
    g1.sBk = kw; g1.sBn = 1; g1.sBb = (int64_t)kh * kw;
    g1.ldc = kw; g1.sCb = (int64_t)h * kw;
    g1.vec_store = (kw % 4 == 0) && aligned16(work);
    int rc = launch_gemm<false, true, 0>(g1, n_planes, stream);
    if (rc) return rc;
    // pass 2: X[b] (h x w) = Q[b] (h x kw) . CW (kw x w) -- A = Q (k-contig), B = CW (n-contig, shared)
    GemmArgs g2{};
    g2.A = work; g2.B = cw; g2.C = x;
    g2.M = h; g2.N = w; g2.K = kw;
    g2.sAm = kw; g2.sAk = 1; g2.sAb = (int64_t)h * kw;
    g2.sBk = w; g2.sBn = 1; g2.sBb = 0;
    g2.ldc = w; g2.sCb = (int64_t)h * w;
    g2.vec_store = (w % 4 == 0) && aligned16(x);
    return launch_gemm<true, true, 0>(g2, n_planes, stream);
}
