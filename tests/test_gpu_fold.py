"""GPU tests of the FOLDED tensor-core DCT path (csrc/dct_fold.cu: even/odd symmetry of the DCT-II basis,
persistent CTA pairs with a shared-memory-resident basis) against float64.

Stated tolerance: max|dY| <= 4e-7 * max|Y|, the bound every DCT implementation of this repo is held to.
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


def test_fold_supported_predicate(D):
    ok = D.util.fold_ok
    assert ok(512, 512, 448, 448) and ok(1024, 1024, 448, 448) and ok(256, 256, 252, 252) and ok(64, 64, 64, 64)
    assert ok(300, 456, 294, 448) and ok(300, 451, 294, 448) and ok(37, 53, 28, 42)      # any plane size
    assert not ok(512, 512, 441, 448)        # odd coefficient count


@pytest.mark.parametrize("h,w,kh,kw,n", [(64, 64, 64, 64, 2), (128, 96, 112, 84, 2), (256, 256, 252, 252, 2),
                                          (512, 512, 448, 448, 1), (304, 464, 294, 448, 1), (1024, 1024, 448, 448, 1),
                                          (512, 512, 512, 512, 1), (32, 48, 28, 42, 5),
                                          # sizes that are not multiples of 16: padded pitches; odd sizes: the middle
                                          # row / column pairs with itself
                                          (300, 452, 294, 448, 2), (300, 451, 294, 448, 2), (301, 450, 294, 448, 1),
                                          (37, 53, 28, 42, 3), (511, 513, 448, 448, 1), (18, 18, 14, 14, 2),
                                          # K > 256: the hi plane of the basis resident, its lo tiles streamed (and the
                                          # sliced fallback where even that does not fit)
                                          (1024, 768, 448, 448, 2), (1000, 1016, 448, 434, 1), (801, 1023, 448, 448, 1),
                                          (1280, 720, 448, 448, 1), (1600, 1200, 448, 448, 1)])
def test_fold_dct_matches_float64_definition(D, h, w, kh, kw, n):
    rng = np.random.default_rng(h * 7 + w)
    x = rng.random((n, 3, h, w), dtype=np.float32) * 2 - 0.5
    assert D.util.fold_ok(h, w, kh, kw)
    y = D.util.dct2_truncated_fold(torch.from_numpy(x).cuda(), kh, kw).cpu().numpy()
    y64 = O.dct2(x.astype(np.float64))[..., :kh, :kw]
    assert y.shape == y64.shape
    assert not np.isnan(y).any()
    err = np.abs(y - y64).max()
    print(f"fold fwd {h}x{w}->{kh}x{kw}: err {err:.3e} = {err / np.abs(y64).max():.2e} max|Y|")
    assert err <= TC_RTOL * np.abs(y64).max()
    back = D.util.idct2_truncated_fold(torch.from_numpy(y64.astype(np.float32)).cuda(), h, w).cpu().numpy()
    pad = np.zeros((n, 3, h, w))
    pad[..., :kh, :kw] = y64
    ref = O.idct2(pad)
    assert not np.isnan(back).any()
    print(f"fold inv err {np.abs(back - ref).max():.3e}")
    assert np.abs(back - ref).max() <= 2e-5


def test_fold_token_grid_matches_unfolded_path(D):
    torch.manual_seed(0)
    x = torch.rand(5, 3, 512, 512, device="cuda")
    hi, lo, dc = D.util.rgb_to_ipt_split(x)
    ref = D.util.dct2_fwd_tc(hi, lo, dc, 448, 448, tile_p=14, channels=3)
    qhi, qlo, qdc = D.util.rgb_to_ipt_fold(x)
    got = D.util.dct2_fwd_fold(qhi, qlo, qdc, 448, 448, tile_p=14, channels=3)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 2 * TC_RTOL * float(ref.abs().max())
    ipt = D.util.rgb_to_ipt(x)
    y64 = O.dct2(ipt.double().cpu().numpy())[..., :448, :448]
    planes = D.util.dct2_fwd_fold(qhi, qlo, qdc, 448, 448, out_shape=(5, 3)).cpu().numpy()
    print("fold max err vs f64:", np.abs(planes - y64).max(), " max|Y|:", np.abs(y64).max())
    assert np.abs(planes - y64).max() <= TC_RTOL * np.abs(y64).max()


def test_fold_natural_image_statistics(D):
    rng = np.random.default_rng(5)
    h = w = 256
    fy, fx = np.meshgrid(np.fft.fftfreq(h), np.fft.fftfreq(w), indexing="ij")
    amp = 1.0 / np.maximum(np.hypot(fy, fx), 1.0 / h)
    x = np.stack([np.real(np.fft.ifft2(amp * np.exp(2j * np.pi * rng.random((h, w))))) for _ in range(6)])
    x = (x - x.min()) / (x.max() - x.min())
    x = x.reshape(2, 3, h, w).astype(np.float32)
    y64 = O.dct2(x.astype(np.float64))[..., :252, :252]
    y = D.util.dct2_truncated_fold(torch.from_numpy(x).cuda(), 252, 252).cpu().numpy()
    print("natural: fold", np.abs(y - y64).max(), "max|Y|", np.abs(y64).max())
    assert np.abs(y - y64).max() <= TC_RTOL * np.abs(y64).max()


def test_fold_many_planes_exercises_the_persistent_schedule(D):
    """More pair tiles than CTA pairs, odd plane count (tail tile of the stacked rows)."""
    torch.manual_seed(3)
    x = torch.rand(37, 3, 128, 128, device="cuda")
    y = D.util.dct2_truncated_fold(x, 112, 112)
    ref = D.util.dct2_truncated(x, 112, 112)
    assert float((y - ref).abs().max()) <= 2 * TC_RTOL * float(ref.abs().max())
    back = D.util.idct2_truncated_fold(ref, 128, 128)
    ref_back = D.util.idct2_truncated(ref, 128, 128)
    assert float((back - ref_back).abs().max()) <= 2e-5


def test_fold_rgb_round_trip(D):
    torch.manual_seed(1)
    x = torch.rand(3, 3, 256, 256, device="cuda")
    qhi, qlo, qdc = D.util.rgb_to_ipt_fold(x)
    y = D.util.dct2_fwd_fold(qhi, qlo, qdc, 256, 256, out_shape=(3, 3))       # full spectrum: lossless
    from dct_autoencoder_b200 import _lib
    ldq = 128
    yh = torch.empty((2, 2, 9, 128, ldq), dtype=torch.float16, device="cuda")
    yl = torch.empty_like(yh)
    dc = torch.empty(9, dtype=torch.float32, device="cuda")
    _lib.call("dcta_fold_coef_planes", _lib.ptr(y), _lib.ptr(yh), _lib.ptr(yl), _lib.ptr(dc), 9, 256, 256, 256, 256,
              _lib.stream_ptr())
    z = D.util.dct2_inv_fold(yh, yl, 256, 256, 256, 256)
    rgb = D.util.unfold_ipt_to_rgb(z, dc, 256, 256)
    assert float((rgb - x).abs().max()) <= 2e-5


def test_fold_maxabs_matches_tile_scores(D):
    """amax|tile| reduced in the GEMM epilogue == a pass over the token grid (bit-exact), hence same order."""
    torch.manual_seed(4)
    for shape in [(5, 3, 512, 512), (3, 3, 256, 320), (2, 3, 1024, 1024)]:
        x = torch.rand(*shape, device="cuda")
        fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
        tiles, maxabs = fe._token_grid(x, want_maxabs=True)
        assert maxabs is not None
        ref = tiles.abs().amax(dim=-1)
        assert torch.equal(maxabs, ref)
        assert torch.equal(fe._sorted_order(tiles, maxabs), fe._sorted_order(tiles))


@pytest.mark.parametrize("n_tok", [100, 128, 500, 1024, 1025, 2000, 3072, 4096, 5000])
def test_sort_tokens_matches_torch_sort(D, n_tok):
    """Per-image descending sort (ties: ascending index), both kernels (register/shuffle variant up to 4096 keys)."""
    from dct_autoencoder_b200 import _lib
    torch.manual_seed(n_tok)
    scores = torch.randn(7, n_tok, device="cuda")
    idx = torch.arange(0, n_tok - 1, 5, device="cuda")
    scores[:, idx] = scores[:, idx + 1]                                    # exact ties
    order = torch.empty((7, n_tok), dtype=torch.int32, device="cuda")
    _lib.call("dcta_sort_tokens", _lib.ptr(scores), _lib.ptr(order), 7, n_tok, _lib.stream_ptr())
    ref = torch.sort(scores, dim=1, descending=True, stable=True).indices
    assert torch.equal(order.long(), ref)
