"""CPU tests: pin oracle/dcta_oracle.py against the fixtures the reference produced
(tests/golden/make_golden.py) and against the float64 definition of the DCT."""
import os
import random

import numpy as np
import pytest
import scipy.fft

import dcta_oracle as O


def batch_from(g, prefix):
    return O.Patches(g[prefix + "patches"], g[prefix + "key_pad_mask"], g[prefix + "image_ids"],
                     g[prefix + "channels"], g[prefix + "positions"],
                     [tuple(x) for x in g[prefix + "patch_sizes"]],
                     [tuple(x) for x in g[prefix + "original_sizes"]])


def test_constants_match_reference(golden):
    g = golden("constants")
    assert np.array_equal(g["Trgb2lms"], O.TRGB2LMS)
    assert np.array_equal(g["Tlms2rgb"], O.TLMS2RGB)
    assert np.array_equal(g["Mipt"], O.MIPT)
    assert np.array_equal(g["MiptInv"], O.MIPT_INV)


def test_colorspace(golden):
    g = golden("colorspace")
    ipt = O.rgb_to_ipt(g["rgb"].copy())
    np.testing.assert_allclose(ipt, g["ipt"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(O.ipt_to_rgb(g["ipt"].copy()), g["rgb_back"], rtol=0, atol=2e-6)
    # UT round trip (test_colorspaces.py:36-38 intent)
    np.testing.assert_allclose(g["rgb_back"], g["rgb"], atol=1e-5)


def test_standin_matches_float64_definition():
    """The fp32 FFT stand-in for torch_dct (used to make the fixtures) agrees with the
    float64 orthonormal DCT-II/III definition."""
    import torch
    import torch_dct_standin as S
    rng = np.random.default_rng(0)
    for h, w in [(64, 64), (45, 70), (252, 252), (300, 451)]:
        x = rng.random((3, h, w), dtype=np.float32)
        y = S.dct_2d(torch.from_numpy(x), "ortho").numpy()
        y64 = scipy.fft.dctn(x.astype(np.float64), type=2, norm="ortho", axes=(-2, -1))
        assert np.abs(y - y64).max() <= 2e-7 * np.abs(y64).max()
        back = S.idct_2d(torch.from_numpy(y), "ortho").numpy()
        assert np.abs(back - x).max() <= 5e-6
    # the matrix form used by the CUDA kernels is the same transform
    c = O.dct_basis(45)
    x = rng.random((45,))
    np.testing.assert_allclose(c @ x, scipy.fft.dct(x, type=2, norm="ortho"), atol=1e-12)
    np.testing.assert_allclose(c.T @ c, np.eye(45), atol=1e-12)


def test_transform(golden):
    g = golden("transform")
    for i in range(3):
        co = O.transform_image_in(g[f"im{i}"])
        assert np.abs(co - g[f"coef{i}"]).max() <= 4e-7 * np.abs(g[f"coef{i}"]).max()
        back = O.transform_image_out(g[f"coef{i}"])
        np.testing.assert_allclose(back, g[f"back{i}"], atol=2e-5)


PRE = {
    "a": dict(channels=3, patch_size=14, sample_patches_beta=0.0, max_patch_h=32, max_patch_w=32, max_seq_len=3072),
    "b": dict(channels=3, patch_size=4, sample_patches_beta=0.0, max_patch_h=5, max_patch_w=4, max_seq_len=40),
    "c": dict(channels=3, patch_size=8, sample_patches_beta=0.05, max_patch_h=6, max_patch_w=6, max_seq_len=64),
    "d": dict(channels=1, patch_size=2, sample_patches_beta=0.0, max_patch_h=8, max_patch_w=8, max_seq_len=64,
              channel_importances=(8.0,)),
}


def test_preprocess_selection_is_bit_exact_on_reference_coefficients(golden):
    """Stage-isolated: feed the reference's own coefficient planes -> identical tokens, positions,
    channels (integer outputs bit-exact; patches are copies so also exact)."""
    g = golden("preprocess")
    random.seed(42)   # same RNG stream as make_golden.py (one draw, case c)
    for name, kw in PRE.items():
        fe = O.FeatureExtractor(**kw)
        fe._transform_image_in = lambda x, c=g[name + "_coef"]: c   # already cropped
        fe._crop_image = lambda x: x
        out = fe._patch_image(g[name + "_coef"])
        assert np.array_equal(out[1], g[name + "_positions"]), name
        assert np.array_equal(out[2], g[name + "_channels"]), name
        assert np.array_equal(out[0], g[name + "_patches"]), name


def test_preprocess_end_to_end(golden):
    g = golden("preprocess")
    random.seed(42)
    for name, kw in PRE.items():
        fe = O.FeatureExtractor(**kw)
        if kw["channels"] != 3:
            fe._transform_image_in = lambda x: x
        out = fe.preprocess(g[name + "_im"])
        assert tuple(out["original_sizes"]) == tuple(g[name + "_original_size"])
        assert tuple(out["patch_sizes"]) == tuple(g[name + "_patch_size"])
        assert out["patches"].shape == g[name + "_patches"].shape
        # same set of tokens; order may differ only among near-tied scores
        ours = set(zip(out["channels"].tolist(), *out["positions"].T.tolist()))
        ref = set(zip(g[name + "_channels"].tolist(), *g[name + "_positions"].T.tolist()))
        if out["patches"].shape[0] == len(ref | ours):
            assert ours == ref
        same = (out["channels"] == g[name + "_channels"]) & (out["positions"] == g[name + "_positions"]).all(-1)
        assert same.mean() > 0.98
        tol = 4e-7 * np.abs(g[name + "_coef"]).max() + 1e-7
        assert np.abs(out["patches"][same] - g[name + "_patches"][same]).max() <= tol


PACK = dict(channels=3, patch_size=8, sample_patches_beta=0.0, max_patch_h=4, max_patch_w=4, max_seq_len=80)


def _pack_items(g, fe):
    items = []
    for i in range(9):
        it = fe.preprocess(g[f"im{i}"])
        assert it["patches"].shape[0] == int(g[f"k{i}"])
        items.append(it)
    return items


def _check_batch(b, g, prefix, coef_tol):
    assert np.array_equal(b.key_pad_mask, g[prefix + "key_pad_mask"])
    assert np.array_equal(b.batched_image_ids, g[prefix + "image_ids"])
    assert np.array_equal(b.attn_mask, g[prefix + "attn_mask"])
    assert [tuple(x) for x in g[prefix + "patch_sizes"]] == [tuple(x) for x in b.patch_sizes]
    assert [tuple(x) for x in g[prefix + "original_sizes"]] == [tuple(x) for x in b.original_sizes]
    same = (b.patch_channels == g[prefix + "channels"]) & (b.patch_positions == g[prefix + "positions"]).all(-1)
    assert same.mean() > 0.98
    assert np.abs(b.patches[same] - g[prefix + "patches"][same]).max() <= coef_tol


def test_packing_and_postprocess(golden):
    g = golden("packing")
    fe = O.FeatureExtractor(**PACK)
    items = _pack_items(g, fe)
    collate = lambda its: {k: [it[k] for it in its] for k in its[0]}
    b = next(fe.iter_batches(iter([collate(items)]), None))
    _check_batch(b, g, "none_", 1e-5)
    # stage-isolated decode: reference batch in -> planes and images out
    rb = batch_from(g, "none_")
    planes = fe.revert_patching(rb)
    recs = fe.postprocess(rb)
    assert len(planes) == 9
    for i in range(9):
        assert np.array_equal(planes[i], g[f"none_plane{i}"])
        np.testing.assert_allclose(recs[i], g[f"none_rec{i}"], atol=3e-5)
    # streaming mode, batch_size=2: tail dropped exactly as the reference does
    loader = iter([collate(items[i:i + 3]) for i in range(0, 9, 3)])
    got = list(fe.iter_batches(loader, 2))
    assert len(got) == int(g["bs2_num_batches"])
    for j, bb in enumerate(got):
        _check_batch(bb, g, f"bs2_{j}_", 1e-5)


def test_identity_transform_round_trip_is_lossless():
    """Intent of the reference's testpatching.py:12-71: with the transform replaced by the
    identity and nothing dropped, postprocess(batch(preprocess(x))) == cropped x."""
    random.seed(42)
    for channels in (1, 3, 4):
        for p in (2, 8, 16):
            n = 4
            fe = O.FeatureExtractor(channels, p, 0.0, n, n, channels * n * n,
                                    channel_importances=(8.0, 1.0, 1.0, 1.0)[:channels])
            fe._transform_image_in = lambda x: x
            fe._transform_image_out = lambda x: x
            xs = []
            for _ in range(6):
                h, w = random.randint(p, p * n), random.randint(p, p * n)
                xs.append(np.arange(channels * h * w, dtype=np.float32).reshape(channels, h, w))
            items = [fe.preprocess(x) for x in xs]
            b = next(fe.iter_batches(iter([{k: [it[k] for it in items] for k in items[0]}]), None))
            for x, r in zip(xs, fe.postprocess(b)):
                ch, cw = fe._get_crop_dims(*x.shape[1:])
                assert np.array_equal(x[:, :ch, :cw], r[:, :ch, :cw])


def test_patchnorm(golden):
    g = golden("patchnorm")
    pn = O.PatchNorm(3, 3, 4, 3)
    for step in range(2):
        b = batch_from(g, f"s{step}_")
        out = pn.forward(b)
        assert np.array_equal(out, g[f"s{step}_out"])
        assert np.array_equal(pn.n, g[f"s{step}_n"])
        np.testing.assert_allclose(pn.median, g[f"s{step}_median"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(pn.b, g[f"s{step}_b"], rtol=2e-6, atol=1e-7)
    pn.frozen = True
    fwd = pn.forward(b)
    np.testing.assert_allclose(fwd, g["fwd"], rtol=1e-5, atol=1e-5)
    b.patches = g["fwd"]
    np.testing.assert_allclose(pn.inverse_norm(b), g["inv"], rtol=1e-5, atol=1e-6)


def test_patchnorm_ignores_padding_outliers():
    """Intent of the reference's testnorm.py:18-55: masked (padding) tokens, however large,
    must not pollute the statistics."""
    rng = np.random.default_rng(0)
    z = 16
    x = rng.normal(np.arange(z), 1.0, size=(8, 10, z)).astype(np.float32)
    pad = np.zeros((8, 10), bool)
    pad[:, 7:] = True
    x[pad] *= 1000
    pos = np.zeros((8, 10, 2), np.int64)
    dp = O.Patches(x, pad, np.zeros((8, 10), np.int64), np.zeros((8, 10), np.int64), pos, [], [])
    pn = O.PatchNorm(2, 2, 4, 1)
    pn.forward(dp)
    assert pn.n[0, 0, 0] == 56
    assert np.abs(pn.median[0, 0, 0] - np.arange(z)).max() < 0.6
    assert pn.b[0, 0, 0].max() < 1.5


def test_lfq(golden):
    g = golden("lfq")
    l = O.LFQ(codebook_size=16, num_codebooks=3)
    q, idx, commit, dist = l.forward(g["a_x"], g["a_mask"])
    assert np.array_equal(q, g["a_q"]) and np.array_equal(idx, g["a_idx"]) and idx.dtype == np.int64
    assert np.array_equal(l.indices_to_codes(idx), g["a_codes"])
    l.training = True
    q, idx, commit, dist = l.forward(g["a_x"], g["a_mask"])
    assert np.array_equal(idx, g["a_train_idx"])
    np.testing.assert_allclose(q, g["a_train_q"], atol=1e-6)
    np.testing.assert_allclose(commit, g["a_commit"], rtol=1e-5)
    np.testing.assert_allclose(dist, g["a_dist"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(O.compute_entropy_loss(g["a_dist"], g["a_mask"]), g["a_entropy"], rtol=1e-4, atol=1e-5)
    l2 = O.LFQ(dim=10, codebook_size=8, num_codebooks=4, w_in=g["b_w_in"], b_in=g["b_b_in"],
               w_out=g["b_w_out"], b_out=g["b_b_out"])
    q, idx, _, _ = l2.forward(g["b_x"], np.ones(g["b_x"].shape[:2], bool))
    assert np.array_equal(idx, g["b_idx"])
    np.testing.assert_allclose(q, g["b_q"], atol=1e-5)
    np.testing.assert_allclose(l2.indices_to_codes(idx), g["b_codes"], atol=1e-5)
    l3 = O.LFQ(codebook_size=2 ** 14, num_codebooks=14)
    q, idx, _, _ = l3.forward(g["c_x"], np.ones((1, 9), bool))
    assert np.array_equal(idx, g["c_idx"]) and np.array_equal(q, g["c_q"])
    with pytest.raises(NotImplementedError):
        l3.forward(g["c_x"], None)
    np.testing.assert_allclose(O.calculate_perplexity(g["p_codes"], 16), g["p_perplexity"], rtol=1e-5)


def test_vq(golden):
    g = golden("vq")
    v = O.VectorQuantize(32, 64, embed=g["a_embed"])
    q, ind, loss = v.forward(g["a_x"], g["a_mask"])
    assert np.array_equal(ind, g["a_ind"]) and ind.dtype == np.int64
    np.testing.assert_allclose(q, g["a_q"], atol=1e-6)
    assert loss.shape == (1,) and loss[0] == 0
    v2 = O.VectorQuantize(24, 32, codebook_dim=8, heads=4, embed=g["b_embed"], w_in=g["b_w_in"],
                          b_in=g["b_b_in"], w_out=g["b_w_out"], b_out=g["b_b_out"])
    q, ind, _ = v2.forward(g["b_x"], np.ones((2, 9), bool))
    assert ind.shape == g["b_ind"].shape
    assert (ind == g["b_ind"]).mean() > 0.97     # near-ties under different fp32 GEMM order
    np.testing.assert_allclose(q[(ind == g["b_ind"]).all(-1)], g["b_q"][(ind == g["b_ind"]).all(-1)], atol=1e-5)


def test_pipeline(golden):
    g = golden("pipeline")
    fe = O.FeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    pn = O.PatchNorm(32, 32, 14, 3)
    fit = O.Patches(g["fit_patches"], g["fit_key_pad_mask"], np.zeros_like(g["fit_channels"]),
                    g["fit_channels"], g["fit_positions"], [], [])
    pn.forward(fit)
    assert int((pn.n > 0).sum()) == int(g["n_used"])
    assert np.array_equal(pn.n[:, :7, :7], g["n"])
    np.testing.assert_allclose(pn.median[:, :7, :7], g["median"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(pn.b[:, :7, :7], g["b"], rtol=2e-6, atol=1e-7)
    pn.frozen = True
    lfq = O.LFQ(codebook_size=2 ** 14, num_codebooks=14)
    rec, codes = O.run_pipeline(g["ims"], fe, pn, lfq)
    same = (codes == g["codes"]).all(-1)
    assert same.mean() > 0.97
    np.testing.assert_allclose(np.stack(rec), g["rec"], atol=5e-3)
    psnr = -10 * np.log10(np.mean((np.stack(rec) - g["rec"]) ** 2) + 1e-20)
    assert psnr > 60


def test_vq_training_oracle_matches_reference_golden(golden):
    """k-means (VQ:180-220) and the EMA codebook update + commitment loss (VQ:479-500, 837-1050) against the
    unmodified reference run in training mode (tests/golden/make_golden_vq_train.py)."""
    g = golden("vq_train")
    means, bins = O.vq_kmeans(g["km_samples"][0], g["km_means0"][0], 10)
    assert np.array_equal(bins, g["km_bins"][0])
    np.testing.assert_allclose(means, g["km_means"][0], rtol=2e-6, atol=2e-7)
    embed, cs, avg = g["ema_embed0"][0], np.zeros(64, np.float32), g["ema_embed0"][0].copy()
    for step in range(2):
        q, ind, loss, embed, cs, avg = O.vq_train_step(g[f"ema_x{step}"], g["ema_mask"], embed, cs, avg, decay=0.8,
                                                       eps=1e-5, commitment_weight=0.7)
        assert np.array_equal(ind, g[f"ema_ind{step}"])
        np.testing.assert_allclose(q, g[f"ema_q{step}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(loss, g[f"ema_loss{step}"], rtol=2e-6)
        np.testing.assert_allclose(cs, g[f"ema_cluster_size{step + 1}"][0], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(avg, g[f"ema_embed_avg{step + 1}"][0], rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(embed, g[f"ema_embed{step + 1}"][0], rtol=5e-6, atol=1e-7)
