"""GPU tests of the tensor-core (tcgen05, split-precision fp16x3) DCT path against float64.

Stated tolerance for this path: max|dY| <= TC_RTOL * max|Y| with TC_RTOL = 4e-7, the same bound the
exact-fp32 FFMA path is held to (tests/test_gpu_parity.py), i.e. fp32-class accuracy.
"""
import numpy as np
import pytest
import torch

import dcta_oracle as O

pytestmark = pytest.mark.gpu
TC_RTOL = 4e-7


@pytest.fixture(scope="module")
def D():
    import dct_autoencoder_b200 as d
    d._lib.load()
    return d


def _gemm_split(D, a, b, row_scale=None, alpha=1.0):
    """a (B?, M, K), b (B?, N, K) float64 numpy -> fp32 (batch, M, N) through dcta_gemm_split."""
    from dct_autoencoder_b200 import _lib
    from dct_autoencoder_b200.util import _split_host, _round8
    a_shared, b_shared = a.ndim == 2, b.ndim == 2
    batch = 1 if (a_shared and b_shared) else (b.shape[0] if a_shared else a.shape[0])
    K = a.shape[-1]
    ld = _round8(K)

    def prep(m):
        p = np.zeros(m.shape[:-1] + (ld,), np.float64)
        p[..., :K] = m
        hi, lo = _split_host(p)
        return torch.from_numpy(hi).cuda(), torch.from_numpy(lo).cuda()

    ah, al = prep(a)
    bh, bl = prep(b)
    M, N = a.shape[-2], b.shape[-2]
    out = torch.full((batch, M, N), float("nan"), dtype=torch.float32, device="cuda")
    rs = None if row_scale is None else torch.from_numpy(row_scale.astype(np.float32)).cuda()
    _lib.call("dcta_gemm_split", _lib.ptr(ah), _lib.ptr(al), M, ld, 0 if a_shared else M * ld,
              _lib.ptr(bh), _lib.ptr(bl), N, ld, 0 if b_shared else N * ld, K, batch, _lib.ptr(rs), float(alpha),
              None, _lib.ptr(out), N, M * N, _lib.stream_ptr())
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("M,N,K,batch", [(128, 128, 32, 1), (128, 128, 64, 1), (128, 128, 512, 2),
                                          (448, 512, 512, 3), (100, 70, 40, 2), (256, 130, 100, 1)])
def test_gemm_split_matches_float64(D, M, N, K, batch):
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((M, K))                 # shared A (the basis role)
    b = rng.standard_normal((batch, N, K))
    rs = rng.random(M) + 0.5
    got = _gemm_split(D, a, b, row_scale=rs, alpha=0.25)
    ref = 0.25 * rs[None, :, None] * np.einsum("mk,bnk->bmn", a, b)
    assert not np.isnan(got).any()
    err = np.abs(got - ref).max()
    assert err <= 2e-6 * np.sqrt(K) * 4, (err, np.abs(ref).max())


def test_gemm_split_batched_a(D):
    rng = np.random.default_rng(9)
    a = rng.standard_normal((3, 130, 72))
    b = rng.standard_normal((3, 200, 72))
    got = _gemm_split(D, a, b)
    ref = np.einsum("bmk,bnk->bmn", a, b)
    assert np.abs(got - ref).max() <= 1e-4


@pytest.mark.parametrize("h,w,kh,kw", [(64, 64, 64, 64), (128, 96, 112, 84), (252, 256, 252, 252),
                                        (512, 512, 448, 448), (300, 456, 294, 448)])
def test_tc_dct_matches_float64_definition(D, h, w, kh, kw):
    rng = np.random.default_rng(h * 7 + w)
    x = rng.random((2, 3, h, w), dtype=np.float32) * 2 - 0.5
    y = D.util.dct2_truncated_tc(torch.from_numpy(x).cuda(), kh, kw).cpu().numpy()
    y64 = O.dct2(x.astype(np.float64))[..., :kh, :kw]
    assert np.abs(y - y64).max() <= TC_RTOL * np.abs(y64).max()
    back = D.util.idct2_truncated_tc(torch.from_numpy(y64.astype(np.float32)).cuda(), h, w).cpu().numpy()
    pad = np.zeros((2, 3, h, w))
    pad[..., :kh, :kw] = y64
    ref = O.idct2(pad)
    assert np.abs(back - ref).max() <= 2e-5


def test_tc_token_grid_matches_fp32_path(D):
    torch.manual_seed(0)
    x = torch.rand(4, 3, 512, 512, device="cuda")
    ipt = D.util.rgb_to_ipt(x)
    ref = D.util.dct2_truncated(ipt, 448, 448, tile_p=14, channels=3)
    hi, lo, dc = D.util.rgb_to_ipt_split(x)
    got = D.util.dct2_fwd_tc(hi, lo, dc, 448, 448, tile_p=14, channels=3)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 2 * TC_RTOL * float(ref.abs().max())
    # against float64: the tensor-core path is held to the same bound as the FFMA path
    y64 = O.dct2(ipt.double().cpu().numpy())[..., :448, :448]
    planes = D.util.dct2_fwd_tc(hi, lo, dc, 448, 448).cpu().numpy()
    print("tc  max err vs f64:", np.abs(planes - y64).max(), " fp32-ffma:",
          np.abs(D.util.dct2_truncated(ipt, 448, 448).cpu().numpy() - y64).max(), " max|Y|:", np.abs(y64).max())
    assert np.abs(planes - y64).max() <= TC_RTOL * np.abs(y64).max()


def test_tc_natural_image_statistics(D):
    """1/f-like images (large low-frequency content, arbitrary mean): same bound."""
    rng = np.random.default_rng(5)
    h = w = 256
    fy, fx = np.meshgrid(np.fft.fftfreq(h), np.fft.fftfreq(w), indexing="ij")
    amp = 1.0 / np.maximum(np.hypot(fy, fx), 1.0 / h)
    x = np.stack([np.real(np.fft.ifft2(amp * np.exp(2j * np.pi * rng.random((h, w))))) for _ in range(6)])
    x = (x - x.min()) / (x.max() - x.min())
    x = x.reshape(2, 3, h, w).astype(np.float32)
    y64 = O.dct2(x.astype(np.float64))[..., :252, :252]
    y = D.util.dct2_truncated_tc(torch.from_numpy(x).cuda(), 252, 252).cpu().numpy()
    yf = D.util.dct2_truncated(torch.from_numpy(x).cuda(), 252, 252).cpu().numpy()
    print("natural: tc", np.abs(y - y64).max(), "ffma", np.abs(yf - y64).max(), "max|Y|", np.abs(y64).max())
    assert np.abs(y - y64).max() <= TC_RTOL * np.abs(y64).max()


@pytest.mark.parametrize("T,C,d", [(1000, 512, 64), (3000, 8192, 256), (129, 100, 7), (257, 130, 20)])
def test_vq_tc_matches_oracle(D, T, C, d):
    """tensor-core nearest-code search vs the numpy oracle (reference formula); indices may differ
    only where the two best squared distances are within 1e-4 relative."""
    from dct_autoencoder_b200.vector_quantize import nearest_code
    rng = np.random.default_rng(T + C)
    x = (rng.standard_normal((T, d)) * 3).astype(np.float32)
    e = rng.standard_normal((C, d)).astype(np.float32)
    e[C // 2] = e[3]                      # exact duplicate code: the first index must win
    x[0] = e[3]
    xt, et = torch.from_numpy(x).cuda(), torch.from_numpy(e).cuda()
    idx, q = nearest_code(xt, et, impl="tc")
    idx32, _ = nearest_code(xt, et, impl="fp32")
    oi, best, second = O.vq_nearest(x, e)
    got = idx.cpu().numpy()
    diff = got != oi
    if diff.any():
        d2 = ((x[diff] - e[got[diff]]) ** 2).sum(-1)
        assert np.all(np.abs(d2 - best[diff]) <= 1e-4 * np.abs(best[diff]) + 1e-6)
    assert diff.mean() < 0.01
    assert int(idx[0]) == 3
    assert np.array_equal(q.cpu().numpy(), e[got])
    assert float((idx != idx32).float().mean()) < 0.01


def test_vq_module_uses_tc_by_default(D):
    torch.manual_seed(0)
    v = D.VectorQuantize(64, 256).cuda().eval()
    v32 = D.VectorQuantize(64, 256, vq_impl="fp32").cuda().eval()
    v32.codebook = v.codebook
    x = torch.randn(2, 500, 64, device="cuda")
    q, ind, _ = v(x)
    q2, ind2, _ = v32(x)
    assert float((ind != ind2).float().mean()) < 0.005
    same = ind == ind2
    assert torch.equal(q[same], q2[same])


def test_vq_padding_mask_is_applied_in_the_gather(D):
    """vector_quantize.py:1043-1048: torch.where(mask, quantize, orig_input).  For the projection-free layer in eval the
    re-rank kernel's gather writes the input row of a masked token itself (no extra pass over the (b, n, d) tensors)."""
    torch.manual_seed(0)
    for dim in (64, 30):                        # 128-bit and scalar variants of the re-rank kernel
        vq = D.VectorQuantize(dim=dim, codebook_size=512).cuda().eval()
        x = torch.randn(3, 200, dim).cuda()
        mask = torch.rand(3, 200).cuda() > 0.3
        q, ind, _ = vq(x, mask=mask)
        q_all, ind_all, _ = vq(x, mask=torch.ones_like(mask))
        assert torch.equal(ind, ind_all)
        assert torch.equal(q, torch.where(mask[..., None], q_all, x))
        assert torch.equal(q_all, vq._codebook.embed[0][ind_all])


def test_vq_rerank_shortcut_returns_the_exact_rerank_indices(D):
    """The second pass skips the exact re-scoring of a token whose best candidate leads the runner-up by more than twice
    the rigorous error bound of the first pass: indices and codes equal those of the plain re-rank (dcta_vq_nearest_tc),
    on well separated codes (shortcut taken) and on a codebook of near-duplicates (shortcut refused)."""
    from dct_autoencoder_b200 import _lib
    from dct_autoencoder_b200.vector_quantize import _codebook_operand, nearest_code
    torch.manual_seed(3)
    T, C, d = 5000, 2048, 128
    x = torch.randn(T, d, device="cuda")
    base = torch.randn(C // 2, d, device="cuda")
    for embed in (torch.randn(C, d, device="cuda"),
                  torch.cat([base, base + 1e-4 * torch.randn_like(base)]).contiguous()):
        idx, q = nearest_code(x, embed)
        e_hi, e2, s_e, _ = _codebook_operand(embed)
        x_hi = torch.empty((T, d), dtype=torch.float16, device="cuda")
        row_alpha = torch.empty(T, dtype=torch.float32, device="cuda")
        st = _lib.stream_ptr(x.device)
        _lib.call("dcta_split_rows_rowscale", _lib.ptr(x), None, None, 0.0, _lib.ptr(x_hi), None, _lib.ptr(row_alpha),
                  -2.0 / s_e, T, d, d, st)
        cand = torch.empty((T, 4), dtype=torch.int32, device="cuda")
        idx0 = torch.empty(T, dtype=torch.int64, device="cuda")
        q0 = torch.empty_like(x)
        _lib.call("dcta_vq_nearest_tc", _lib.ptr(x), _lib.ptr(x_hi), _lib.ptr(row_alpha), _lib.ptr(embed), _lib.ptr(e_hi),
                  _lib.ptr(e2), _lib.ptr(cand), _lib.ptr(idx0), _lib.ptr(q0), T, C, d, d, st)
        assert torch.equal(idx, idx0) and torch.equal(q, q0)
        # and both are the true nearest code up to the stated EPS_VQ
        d2 = torch.cdist(x.double(), embed.double()) ** 2
        best = d2.min(dim=1).values
        got = d2.gather(1, idx[:, None])[:, 0]
        assert bool(((got - best) <= 1e-4 * best).all())
