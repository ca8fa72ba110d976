"""GPU parity tests: the CUDA path (through the C ABI) against the golden fixtures produced by the
reference and against the numpy oracle on seeded inputs.

Tolerances (defined once in tests/parity_rules.py, used below):
  * integer / index / code outputs: bit-exact when the inputs are identical (stage-isolated tests);
    end-to-end, selection order may differ only among tokens whose scores differ by < EPS_SCORE,
    LFQ bits only where |Y - median| < EPS_LFQ * max|Y| (the coefficient-scale statement of the
    sign-boundary rule), VQ indices only where the two best squared distances differ by < EPS_VQ
    (relative).
  * DCT coefficients: max|dY| <= COEF_RTOL * max|Y|  (fp32; the reference's own fp32 FFT path
    differs from the float64 definition by ~8e-8 * max|Y|).
  * normalised patches / reconstructions: abs 1e-4 outside clamped entries / 5e-5 for images.
"""
import random

import numpy as np
import pytest
import torch

import dcta_oracle as O

pytestmark = pytest.mark.gpu

from parity_rules import COEF_RTOL, EPS_LFQ, EPS_SCORE, EPS_VQ, lfq_bit_exempt, std_at_tokens  # noqa: E402


@pytest.fixture(scope="module")
def D():
    import dct_autoencoder_b200 as d
    d._lib.load()
    return d


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def npy(t):
    return t.detach().cpu().numpy()


def patches_from(D, g, prefix):
    return D.DCTPatches(
        patches=cu(g[prefix + "patches"]), key_pad_mask=cu(g[prefix + "key_pad_mask"]),
        batched_image_ids=cu(g[prefix + "image_ids"]), patch_channels=cu(g[prefix + "channels"]),
        patch_positions=cu(g[prefix + "positions"]),
        patch_sizes=[tuple(int(v) for v in x) for x in g[prefix + "patch_sizes"]] if prefix + "patch_sizes" in g else [],
        original_sizes=[tuple(int(v) for v in x) for x in g[prefix + "original_sizes"]] if prefix + "original_sizes" in g else [])


# ------------------------------------------------------------------------------ colour space
def test_colorspace(D, golden):
    g = golden("colorspace")
    ipt = D.util.rgb_to_ipt(cu(g["rgb"]))
    np.testing.assert_allclose(npy(ipt), g["ipt"], rtol=0, atol=2e-6)
    back = D.util.ipt_to_rgb(cu(g["ipt"]))
    np.testing.assert_allclose(npy(back), g["rgb_back"], rtol=0, atol=4e-6)
    # odd plane size -> scalar kernel variant
    x = np.random.default_rng(0).random((2, 3, 7, 9), dtype=np.float32)
    np.testing.assert_allclose(npy(D.util.rgb_to_ipt(cu(x))), O.rgb_to_ipt(x.copy()), atol=2e-6)
    np.testing.assert_allclose(npy(D.util.ipt_to_rgb(D.util.rgb_to_ipt(cu(x)))), x, atol=2e-5)


# ------------------------------------------------------------------------------ DCT
@pytest.mark.parametrize("h,w", [(64, 64), (45, 70), (128, 96), (252, 252), (300, 451), (512, 512)])
def test_dct_matches_float64_definition(D, h, w):
    rng = np.random.default_rng(h * 1000 + w)
    x = rng.random((3, h, w), dtype=np.float32)
    y = npy(D.util.dct2(cu(x), "ortho"))
    y64 = O.dct2(x.astype(np.float64))
    assert np.abs(y - y64).max() <= COEF_RTOL * np.abs(y64).max()
    back = npy(D.util.idct2(cu(y), "ortho"))
    assert np.abs(back - x).max() <= 2e-5
    with pytest.raises(NotImplementedError):
        D.util.dct2(cu(x))


def test_truncated_dct_equals_crop_of_full(D):
    x = np.random.default_rng(1).random((2, 3, 100, 75), dtype=np.float32)
    full = O.dct2(x.astype(np.float64))
    t = npy(D.util.dct2_truncated(cu(x), 56, 42))
    assert np.abs(t - full[..., :56, :42]).max() <= COEF_RTOL * np.abs(full).max()
    # truncated inverse == inverse of the zero-padded plane (FE:300-304)
    pad = np.zeros_like(full)
    pad[..., :56, :42] = full[..., :56, :42]
    inv = npy(D.util.idct2_truncated(cu(full[..., :56, :42].astype(np.float32)), 100, 75))
    np.testing.assert_allclose(inv, O.idct2(pad), atol=2e-5)


def test_transform_golden(D, golden):
    g = golden("transform")
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    for i in range(3):
        co = npy(fe._transform_image_in(cu(g[f"im{i}"])))
        assert np.abs(co - g[f"coef{i}"]).max() <= COEF_RTOL * np.abs(g[f"coef{i}"]).max()
        back = npy(fe._transform_image_out(cu(g[f"coef{i}"])))
        np.testing.assert_allclose(back, g[f"back{i}"], atol=5e-5)


# ------------------------------------------------------------------------------ selection
PRE = {
    "a": dict(channels=3, patch_size=14, sample_patches_beta=0.0, max_patch_h=32, max_patch_w=32, max_seq_len=3072),
    "b": dict(channels=3, patch_size=4, sample_patches_beta=0.0, max_patch_h=5, max_patch_w=4, max_seq_len=40),
    "c": dict(channels=3, patch_size=8, sample_patches_beta=0.05, max_patch_h=6, max_patch_w=6, max_seq_len=64),
    "d": dict(channels=1, patch_size=2, sample_patches_beta=0.0, max_patch_h=8, max_patch_w=8, max_seq_len=64,
              channel_importances=(8.0,)),
}


def test_selection_bit_exact_on_reference_coefficients(D, golden):
    """Stage-isolated: the reference's own coefficient planes in -> identical tokens out."""
    g = golden("preprocess")
    random.seed(42)
    for name, kw in PRE.items():
        fe = D.DCTAutoencoderFeatureExtractor(**kw)
        coef = cu(g[name + "_coef"])
        fe._transform_image_in = lambda x, c=coef: c
        out = fe.preprocess(cu(g[name + "_im"]))
        assert np.array_equal(npy(out["positions"]), g[name + "_positions"]), name
        assert np.array_equal(npy(out["channels"]), g[name + "_channels"]), name
        assert np.array_equal(npy(out["patches"]), g[name + "_patches"]), name
        assert out["positions"].dtype == torch.int64 and out["channels"].dtype == torch.int64


def _order_ok(scores_ref, ours_keys, ref_keys):
    """same sequence except permutations inside runs of near-tied reference scores"""
    if ours_keys == ref_keys:
        return True
    score_of = dict(zip(ref_keys, scores_ref))
    for a, b in zip(ours_keys, ref_keys):
        if a != b and (a not in score_of or abs(score_of[a] - score_of[b]) >= EPS_SCORE):
            return False
    return True


def test_preprocess_end_to_end(D, golden):
    g = golden("preprocess")
    random.seed(42)
    for name, kw in PRE.items():
        if kw["channels"] != 3:
            continue
        fe = D.DCTAutoencoderFeatureExtractor(**kw)
        out = fe.preprocess(cu(g[name + "_im"]))
        assert tuple(out["original_sizes"]) == tuple(g[name + "_original_size"])
        assert tuple(out["patch_sizes"]) == tuple(g[name + "_patch_size"])
        assert out["patches"].shape == g[name + "_patches"].shape
        ofe = O.FeatureExtractor(**kw)
        _, hi, wi, sc = ofe.importance_scores(g[name + "_coef"])
        c = kw["channels"]
        keys_all = [(ch, int(h), int(w)) for h, w in zip(hi, wi) for ch in range(c)]
        score_of = dict(zip(keys_all, sc.reshape(-1)))
        ref_keys = list(zip(g[name + "_channels"].tolist(), *g[name + "_positions"].T.tolist()))
        ours_keys = list(zip(npy(out["channels"]).tolist(), *npy(out["positions"]).T.tolist()))
        assert _order_ok([score_of[k] for k in ref_keys], ours_keys, ref_keys), name
        same = np.array([a == b for a, b in zip(ours_keys, ref_keys)])
        tol = COEF_RTOL * np.abs(g[name + "_coef"]).max()
        assert np.abs(npy(out["patches"])[same] - g[name + "_patches"][same]).max() <= tol


# ------------------------------------------------------------------------------ packing / decode
PACK = dict(channels=3, patch_size=8, sample_patches_beta=0.0, max_patch_h=4, max_patch_w=4, max_seq_len=80)


def _oracle_items(g):
    fe = O.FeatureExtractor(**PACK)
    return fe, [fe.preprocess(g[f"im{i}"]) for i in range(9)]


def _to_cuda_items(items):
    return [dict(patches=cu(it["patches"]), positions=cu(it["positions"]), channels=cu(it["channels"]),
                 original_sizes=it["original_sizes"], patch_sizes=it["patch_sizes"], tag=i)
            for i, it in enumerate(items)]


def _assert_batch_equal(b, ob):
    assert np.array_equal(npy(b.patches), ob.patches)
    assert np.array_equal(npy(b.key_pad_mask), ob.key_pad_mask)
    assert np.array_equal(npy(b.batched_image_ids), ob.batched_image_ids)
    assert np.array_equal(npy(b.patch_channels), ob.patch_channels)
    assert np.array_equal(npy(b.patch_positions), ob.patch_positions)
    assert np.array_equal(npy(b.attn_mask), ob.attn_mask)
    assert [tuple(x) for x in b.patch_sizes] == [tuple(x) for x in ob.patch_sizes]
    assert [tuple(x) for x in b.original_sizes] == [tuple(x) for x in ob.original_sizes]


def test_iter_batches_stage_isolated(D, golden):
    """identical per-image token lists in -> bit-identical DCTPatches, both modes, tail dropped."""
    g = golden("packing")
    ofe, items = _oracle_items(g)
    fe = D.DCTAutoencoderFeatureExtractor(**PACK)
    citems = _to_cuda_items(items)
    collate = lambda its: {k: [it[k] for it in its] for k in its[0]}
    b = next(fe.iter_batches(iter([D.dict_collate(citems)]), None))
    ob = next(ofe.iter_batches(iter([collate(items)]), None))
    _assert_batch_equal(b, ob)
    assert b._data["tag"] == list(range(9))
    assert b.patches.dtype == torch.float32 and b.key_pad_mask.dtype == torch.bool
    got = list(fe.iter_batches(iter([D.dict_collate(citems[i:i + 3]) for i in range(0, 9, 3)]), 2))
    ogot = list(ofe.iter_batches(iter([collate(items[i:i + 3]) for i in range(0, 9, 3)]), 2))
    assert len(got) == len(ogot) == int(g["bs2_num_batches"])
    for x, y in zip(got, ogot):
        _assert_batch_equal(x, y)
        assert x.patches.shape[0] == 2


def test_packing_golden_end_to_end(D, golden):
    g = golden("packing")
    fe = D.DCTAutoencoderFeatureExtractor(**PACK)
    items = [fe.preprocess(cu(g[f"im{i}"])) for i in range(9)]
    for i, it in enumerate(items):
        assert it["patches"].shape[0] == int(g[f"k{i}"])
    b = next(fe.iter_batches(iter([D.dict_collate(items)]), None))
    assert np.array_equal(npy(b.key_pad_mask), g["none_key_pad_mask"])
    assert np.array_equal(npy(b.batched_image_ids), g["none_image_ids"])
    assert np.array_equal(npy(b.attn_mask), g["none_attn_mask"])
    same = (npy(b.patch_channels) == g["none_channels"]) & (npy(b.patch_positions) == g["none_positions"]).all(-1)
    assert same.mean() > 0.98
    assert np.abs(npy(b.patches)[same] - g["none_patches"][same]).max() <= 1e-5


def test_revert_and_postprocess_golden(D, golden):
    g = golden("packing")
    fe = D.DCTAutoencoderFeatureExtractor(**PACK)
    rb = patches_from(D, g, "none_")
    planes = fe.revert_patching(rb)
    recs = fe.postprocess(rb)
    assert len(planes) == len(recs) == 9
    for i in range(9):
        assert np.array_equal(npy(planes[i]), g[f"none_plane{i}"])
        np.testing.assert_allclose(npy(recs[i]), g[f"none_rec{i}"], atol=5e-5)


def test_identity_transform_round_trip_is_lossless(D):
    """Intent of the reference's testpatching.py:12-71."""
    random.seed(42)
    for channels in (1, 3, 4):
        for p in (2, 8, 16):
            n = 4
            fe = D.DCTAutoencoderFeatureExtractor(channels, p, 0.0, n, n, channels * n * n,
                                                  channel_importances=(8.0, 1.0, 1.0, 1.0)[:channels])
            fe._transform_image_in = lambda x: x
            fe._transform_image_out = lambda x: x
            xs = []
            for _ in range(6):
                h, w = random.randint(p, p * n), random.randint(p, p * n)
                xs.append(torch.arange(channels * h * w, dtype=torch.float32).reshape(channels, h, w).cuda())
            items = [fe.preprocess(x) for x in xs]
            b = next(fe.iter_batches(iter([D.dict_collate(items)]), None))
            for x, r in zip(xs, fe.postprocess(b)):
                ch, cw = fe._get_crop_dims(*x.shape[1:])
                assert torch.equal(x[:, :ch, :cw], r[:, :ch, :cw])


def test_process_batch_equals_per_image_path(D):
    torch.manual_seed(0)
    random.seed(7)
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.01, 32, 32, 256)
    x = torch.rand(5, 3, 100, 130, device="cuda")
    state = random.getstate()
    b1 = fe.process_batch(x)
    random.setstate(state)
    items = [fe.preprocess(im) for im in x]
    b2 = next(fe.iter_batches(iter([D.dict_collate(items)]), None))
    for f in ("patches", "key_pad_mask", "batched_image_ids", "patch_channels", "patch_positions"):
        assert torch.equal(getattr(b1, f), getattr(b2, f)), f
    assert b1.patch_sizes == b2.patch_sizes and b1.original_sizes == b2.original_sizes
    assert torch.equal(b1.attn_mask, b2.attn_mask)
    r1 = fe.postprocess_batch(b1)
    r2 = torch.stack(fe.postprocess(b2))
    assert torch.equal(r1, r2)


# ------------------------------------------------------------------------------ PatchNorm
def test_patchnorm_golden(D, golden):
    g = golden("patchnorm")
    pn = D.PatchNorm(3, 3, 4, 3).cuda()
    pn.train()
    for step in range(2):
        b = patches_from(D, g, f"s{step}_")
        out = pn(b)
        assert np.array_equal(npy(out), g[f"s{step}_out"])
        assert np.array_equal(npy(pn.n), g[f"s{step}_n"])
        np.testing.assert_allclose(npy(pn.median), g[f"s{step}_median"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(npy(pn.b), g[f"s{step}_b"], rtol=2e-6, atol=1e-7)
    pn.frozen = True
    fwd = pn(b)
    np.testing.assert_allclose(npy(fwd), g["fwd"], rtol=1e-5, atol=1e-5)
    b.patches = cu(g["fwd"])
    np.testing.assert_allclose(npy(pn.inverse_norm(b)), g["inv"], rtol=1e-5, atol=1e-6)
    assert set(pn.state_dict().keys()) == {"n", "median", "b"}


def test_patchnorm_matches_oracle_bitwise(D):
    """fit (2 steps), normalise and denormalise on random packed rows: bit-exact vs the oracle's
    fp32 arithmetic (same operation order, no fma contraction)."""
    rng = np.random.default_rng(3)
    C, H, W, p = 3, 5, 4, 4
    z = p * p
    on = O.PatchNorm(H, W, p, C)
    pn = D.PatchNorm(H, W, p, C).cuda().train()
    for step in range(2):
        b, s = 7, 70
        x = (rng.standard_normal((b, s, z)) * (1 + step)).astype(np.float32)
        ch = rng.integers(0, C, (b, s))
        pos = np.stack([rng.integers(0, H, (b, s)), rng.integers(0, W - 1, (b, s))], -1)
        pad = np.arange(s)[None, :] >= rng.integers(20, s + 1, (b, 1))
        x[pad] = 0
        ch[pad] = 0
        pos[pad] = 0
        odp = O.Patches(x, pad, np.zeros((b, s), np.int64), ch, pos, [], [])
        ddp = D.DCTPatches(patches=cu(x), key_pad_mask=cu(pad), batched_image_ids=cu(np.zeros((b, s), np.int64)),
                           patch_channels=cu(ch), patch_positions=cu(pos), patch_sizes=[], original_sizes=[])
        assert np.array_equal(npy(pn(ddp)), on.forward(odp))
        assert np.array_equal(npy(pn.n), on.n)
        assert np.array_equal(npy(pn.median), on.median)
        np.testing.assert_allclose(npy(pn.b), on.b, rtol=1e-6, atol=1e-7)   # np.add.at vs ordered sum
    pn.frozen = True
    on.frozen = True
    on.b = npy(pn.b).copy()
    y = pn(ddp)
    assert np.array_equal(npy(y), on.forward(odp))
    ddp.patches = y
    odp.patches = npy(y)
    assert np.array_equal(npy(pn.inverse_norm(ddp)), on.inverse_norm(odp))
    # eval mode takes the normalise branch too (PN:101)
    pn.frozen = False
    pn.eval()
    assert np.array_equal(npy(pn(ddp)), on.forward(odp))


@pytest.mark.parametrize("H,W,b,s,p", [(1, 2, 40, 40, 14), (3, 3, 9, 64, 14), (2, 2, 6, 50, 6), (4, 4, 12, 60, 14), (6, 6, 10, 70, 6)])
def test_patchnorm_median_long_and_short_lists(D, H, W, b, s, p):
    """batch_median_kernel: lists beyond its shared-memory staging (800 tokens per position: counted from global memory), lists of up to 64 tokens (the bit-sliced kernel, one and two token blocks),
    lists inside it, coefficient counts that are not a multiple of its 32-lane chunk, even list lengths (lower median,
    PN:129) and duplicated values: n, median bit-exact vs the oracle."""
    rng = np.random.default_rng(11)
    z = p * p
    on = O.PatchNorm(H, W, p, 1)
    pn = D.PatchNorm(H, W, p, 1).cuda().train()
    for step in range(2):
        x = np.round(rng.standard_normal((b, s, z)) * 8).astype(np.float32) / 4 - step   # many exact ties
        ch = np.zeros((b, s), np.int64)
        pos = np.stack([rng.integers(0, H, (b, s)), rng.integers(0, W, (b, s))], -1)
        pad = np.arange(s)[None, :] >= rng.integers(s - 3, s + 1, (b, 1))
        x[pad] = 0
        pos[pad] = 0
        odp = O.Patches(x, pad, np.zeros((b, s), np.int64), ch, pos, [], [])
        ddp = D.DCTPatches(patches=cu(x), key_pad_mask=cu(pad), batched_image_ids=cu(np.zeros((b, s), np.int64)),
                           patch_channels=cu(ch), patch_positions=cu(pos), patch_sizes=[], original_sizes=[])
        assert np.array_equal(npy(pn(ddp)), on.forward(odp))
        assert np.array_equal(npy(pn.n), on.n)
        assert np.array_equal(npy(pn.median), on.median)
        np.testing.assert_allclose(npy(pn.b), on.b, rtol=1e-6, atol=1e-7)


def test_patchnorm_ignores_padding_outliers(D):
    """Intent of the reference's testnorm.py:18-55."""
    rng = np.random.default_rng(0)
    z = 16
    x = rng.normal(np.arange(z), 1.0, size=(8, 10, z)).astype(np.float32)
    pad = np.zeros((8, 10), bool)
    pad[:, 7:] = True
    x[pad] *= 1000
    dp = D.DCTPatches(patches=cu(x), key_pad_mask=cu(pad), batched_image_ids=cu(np.zeros((8, 10), np.int64)),
                      patch_channels=cu(np.zeros((8, 10), np.int64)), patch_positions=cu(np.zeros((8, 10, 2), np.int64)),
                      patch_sizes=[], original_sizes=[])
    pn = D.PatchNorm(2, 2, 4, 1).cuda().train()
    out = pn(dp)
    assert float(pn.n[0, 0, 0]) == 56
    assert np.abs(npy(pn.median)[0, 0, 0] - np.arange(z)).max() < 0.6
    assert npy(pn.b)[0, 0, 0].max() < 1.5
    assert float(out[:, 7:].abs().max()) == 0.0


# ------------------------------------------------------------------------------ LFQ
def test_lfq_golden(D, golden):
    g = golden("lfq")
    l = D.LFQ(codebook_size=16, num_codebooks=3).cuda().eval()
    q, idx, commit, dist = l(cu(g["a_x"]), cu(g["a_mask"]))
    assert np.array_equal(npy(q), g["a_q"]) and np.array_equal(npy(idx), g["a_idx"])
    assert idx.dtype == torch.int64 and float(commit) == 0.0 and float(dist) == 0.0
    assert np.array_equal(npy(l.indices_to_codes(idx)), g["a_codes"])
    l.train()
    q, idx, commit, dist = l(cu(g["a_x"]), cu(g["a_mask"]))
    assert np.array_equal(npy(idx), g["a_train_idx"])
    np.testing.assert_allclose(npy(q), g["a_train_q"], atol=1e-6)
    np.testing.assert_allclose(float(commit), g["a_commit"], rtol=1e-5)
    np.testing.assert_allclose(npy(dist), g["a_dist"], rtol=1e-5, atol=1e-5)
    ent = D.util.compute_entropy_loss(dist, cu(g["a_mask"]))
    np.testing.assert_allclose(float(ent), g["a_entropy"], rtol=1e-4, atol=1e-5)
    l2 = D.LFQ(dim=10, codebook_size=8, num_codebooks=4).cuda().eval()
    with torch.no_grad():
        l2.project_in.weight.copy_(cu(g["b_w_in"])); l2.project_in.bias.copy_(cu(g["b_b_in"]))
        l2.project_out.weight.copy_(cu(g["b_w_out"])); l2.project_out.bias.copy_(cu(g["b_b_out"]))
        q, idx, _, _ = l2(cu(g["b_x"]), torch.ones(3, 5, dtype=torch.bool, device="cuda"))
    assert np.array_equal(npy(idx), g["b_idx"])
    np.testing.assert_allclose(npy(q), g["b_q"], atol=1e-5)
    np.testing.assert_allclose(npy(l2.indices_to_codes(idx)), g["b_codes"], atol=1e-5)
    l3 = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).cuda().eval()
    q, idx, _, _ = l3(cu(g["c_x"]), torch.ones(1, 9, dtype=torch.bool, device="cuda"))
    assert np.array_equal(npy(idx), g["c_idx"]) and np.array_equal(npy(q), g["c_q"])
    with pytest.raises(NotImplementedError):
        l3(cu(g["c_x"]), None)
    with pytest.raises(AssertionError):
        l3(cu(g["c_x"])[..., :100], torch.ones(1, 9, dtype=torch.bool, device="cuda"))
    assert set(l3.state_dict().keys()) == {"mask"}
    ppl = D.util.calculate_perplexity(cu(g["p_codes"]), 16)
    np.testing.assert_allclose(float(ppl), g["p_perplexity"], rtol=1e-5)


@pytest.mark.parametrize("c,d,n_tok", [(14, 14, 3000), (16, 13, 777), (1, 10, 100), (3, 5, 65)])
def test_lfq_matches_oracle_bitwise(D, c, d, n_tok):
    rng = np.random.default_rng(c * 100 + d)
    x = rng.standard_normal((2, n_tok, c * d)).astype(np.float32)
    x[0, :3] = 0.0
    x[1, 5, :7] = np.nan
    l = D.LFQ(codebook_size=2 ** d, num_codebooks=c).cuda().eval()
    ol = O.LFQ(codebook_size=2 ** d, num_codebooks=c)
    m = np.ones((2, n_tok), bool)
    q, idx, _, _ = l(cu(x), cu(m))
    oq, oidx, _, _ = ol.forward(x, m)
    assert np.array_equal(npy(q), oq) and np.array_equal(npy(idx), oidx)
    # decode of codes == quantised vector (lfq.py:105-134)
    assert torch.equal(l.indices_to_codes(idx), q)


def test_lfq_train_terms_match_oracle(D):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, 40, 24)).astype(np.float32)
    m = rng.random((3, 40)) > 0.3
    l = D.LFQ(codebook_size=2 ** 8, num_codebooks=3).cuda().train()
    ol = O.LFQ(codebook_size=2 ** 8, num_codebooks=3)
    ol.training = True
    xt = cu(x).requires_grad_(True)
    q, idx, commit, dist = l(xt, cu(m))
    oq, oidx, ocommit, odist = ol.forward(x, m)
    assert np.array_equal(npy(idx), oidx)
    np.testing.assert_allclose(float(commit), ocommit, rtol=2e-5)
    np.testing.assert_allclose(npy(dist), odist, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(float(D.util.compute_entropy_loss(dist.detach(), cu(m))),
                               O.compute_entropy_loss(odist, m), rtol=2e-4, atol=1e-5)
    # straight-through gradient and commitment gradient
    (q.sum() + commit).backward()
    qn = np.where(x > 0, 1.0, -1.0)
    expect = 1.0 + 2 * (x - qn) * m[..., None] / (m.sum() * 24)
    np.testing.assert_allclose(npy(xt.grad), expect, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------ VQ
def _vq_check(x, embed, ind):
    oi, best, second = O.vq_nearest(x, embed)
    diff = ind != oi
    if diff.any():
        # exempt near-ties (different fp32 GEMM summation order)
        d2_ours = ((x[diff] - embed[ind[diff]]) ** 2).sum(-1)
        assert np.all(np.abs(d2_ours - best[diff]) <= EPS_VQ * np.abs(best[diff]) + 1e-6)
    return diff.mean()


def test_vq_golden(D, golden):
    g = golden("vq")
    v = D.VectorQuantize(32, 64).cuda().eval()
    v.codebook = cu(g["a_embed"])
    q, ind, loss = v(cu(g["a_x"]), mask=cu(g["a_mask"]))
    assert np.array_equal(npy(ind), g["a_ind"]) and ind.dtype == torch.int64
    np.testing.assert_allclose(npy(q), g["a_q"], atol=1e-6)
    assert tuple(loss.shape) == (1,) and float(loss) == 0.0
    v2 = D.VectorQuantize(24, 32, heads=4, codebook_dim=8).cuda().eval()
    with torch.no_grad():
        v2.codebook = cu(g["b_embed"])
        v2.project_in.weight.copy_(cu(g["b_w_in"])); v2.project_in.bias.copy_(cu(g["b_b_in"]))
        v2.project_out.weight.copy_(cu(g["b_w_out"])); v2.project_out.bias.copy_(cu(g["b_b_out"]))
        q, ind, _ = v2(cu(g["b_x"]), mask=torch.ones(2, 9, dtype=torch.bool, device="cuda"))
    assert ind.shape == g["b_ind"].shape
    ok = (npy(ind) == g["b_ind"])
    assert ok.mean() > 0.97
    rows = ok.all(-1)
    np.testing.assert_allclose(npy(q)[rows], g["b_q"][rows], atol=1e-5)
    assert "_codebook.embed" in v.state_dict()
    # training of the EMA codebook is covered by tests/test_gpu_vq_train.py; the variants that are not implemented refuse
    with pytest.raises(NotImplementedError):
        D.VectorQuantize(dim=8, codebook_size=16, affine_param=True)
    with pytest.raises(NotImplementedError):
        D.VectorQuantize(dim=8, codebook_size=16, learnable_codebook=True, ema_update=False)


@pytest.mark.parametrize("T,C,d", [(1000, 512, 64), (300, 8192, 256), (129, 100, 7)])
def test_vq_matches_oracle(D, T, C, d):
    rng = np.random.default_rng(T)
    x = rng.standard_normal((T, d)).astype(np.float32)
    e = rng.standard_normal((C, d)).astype(np.float32)
    e[C // 2] = e[3]                     # exact duplicate code -> first index must win
    x[0] = e[3]
    from dct_autoencoder_b200.vector_quantize import nearest_code
    idx, q = nearest_code(cu(x), cu(e))
    frac = _vq_check(x, e, npy(idx))
    assert frac < 0.01
    assert int(idx[0]) == 3
    assert np.array_equal(npy(q), e[npy(idx)])


# ------------------------------------------------------------------------------ whole path
def test_pipeline_golden(D, golden):
    g = golden("pipeline")
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    pn = D.PatchNorm(32, 32, 14, 3).cuda().train()
    fit = D.DCTPatches(patches=cu(g["fit_patches"]), key_pad_mask=cu(g["fit_key_pad_mask"]),
                       batched_image_ids=cu(np.zeros_like(g["fit_channels"])), patch_channels=cu(g["fit_channels"]),
                       patch_positions=cu(g["fit_positions"]), patch_sizes=[], original_sizes=[])
    pn(fit)
    assert int((pn.n > 0).sum()) == int(g["n_used"])
    assert np.array_equal(npy(pn.n)[:, :7, :7], g["n"])
    np.testing.assert_allclose(npy(pn.median)[:, :7, :7], g["median"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(npy(pn.b)[:, :7, :7], g["b"], rtol=2e-6, atol=1e-7)
    pn.frozen = True
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).cuda().eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    batch, q, codes = pipe.encode(cu(g["ims"]))
    assert np.array_equal(npy(batch.key_pad_mask), g["key_pad_mask"])
    assert np.array_equal(npy(batch.batched_image_ids), g["image_ids"])
    same_tok = (npy(batch.patch_channels) == g["channels"]) & (npy(batch.patch_positions) == g["positions"]).all(-1)
    assert same_tok.mean() > 0.98
    # code bits against the reference's, on identical tokens: differences only under the EPS_LFQ rule
    opn = O.PatchNorm(32, 32, 14, 3)
    opn.median[:, :7, :7], opn.b[:, :7, :7] = g["median"], g["b"]
    opn.frozen = True
    ob = O.Patches(g["patches"], g["key_pad_mask"], g["image_ids"], g["channels"], g["positions"], [], [])
    onormed = opn.forward(ob)
    shifts = np.arange(13, -1, -1)
    bits = lambda a: ((a[..., None] >> shifts) & 1).reshape(a.shape[:-1] + (196,)).astype(bool)
    diff = (bits(npy(codes)) != bits(g["codes"])) & same_tok[..., None]
    exempt = lfq_bit_exempt(onormed, std_at_tokens(opn.b, ob.patch_channels, ob.patch_positions, opn.eps),
                            float(np.abs(g["patches"]).max()))
    assert not (diff & ~exempt).any()
    assert diff.mean() < 2e-4
    rec = pipe.decode(batch, q)
    psnr = -10 * np.log10(np.mean((npy(rec) - g["rec"]) ** 2) + 1e-20)
    assert psnr > 60
    # decoding the REFERENCE's codes on its own tokens reproduces its images to the stated 5e-5
    gb = D.DCTPatches(patches=None, key_pad_mask=cu(g["key_pad_mask"]), batched_image_ids=cu(g["image_ids"]),
                      patch_channels=cu(g["channels"]), patch_positions=cu(g["positions"]),
                      patch_sizes=batch.patch_sizes, original_sizes=batch.original_sizes)
    rec_g = pipe.decode_codes(gb, cu(g["codes"]))
    np.testing.assert_allclose(npy(rec_g), g["rec"], atol=5e-5)


def test_pipeline_matches_oracle_with_bit_level_exemptions(D):
    """encode on seeded images vs the oracle: LFQ bits equal except where |Y - median| < EPS_LFQ * max|Y|."""
    torch.manual_seed(3)
    x = torch.rand(3, 3, 112, 84)
    x_fit = torch.rand(8, 3, 112, 84)     # statistics from OTHER images (else 1/3 of the values sit on the median)
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    ofe = O.FeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    pn = D.PatchNorm(32, 32, 14, 3).cuda()
    opn = O.PatchNorm(32, 32, 14, 3)
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).cuda().eval()
    olfq = O.LFQ(codebook_size=2 ** 14, num_codebooks=14)
    pipe = D.TransformPipeline(fe, pn, lfq)
    pipe.fit_norm(x_fit.cuda())
    collate = lambda its: {k: [it[k] for it in its] for k in its[0]}
    fit_items = [ofe.preprocess(im.numpy()) for im in x_fit]
    opn.forward(next(ofe.iter_batches(iter([collate(fit_items)]), None)))
    opn.frozen = True
    np.testing.assert_allclose(npy(pn.median), opn.median, atol=2e-5)
    np.testing.assert_allclose(npy(pn.b), opn.b, atol=2e-5)
    items = [ofe.preprocess(im.numpy()) for im in x]
    ob = next(ofe.iter_batches(iter([collate(items)]), None))
    batch, q, codes = pipe.encode(x.cuda())
    ob.patches = opn.forward(ob)
    oq, ocodes, _, _ = olfq.forward(ob.patches, ~ob.key_pad_mask)
    tok_same = (npy(batch.patch_channels) == ob.patch_channels) & (npy(batch.patch_positions) == ob.patch_positions).all(-1)
    assert tok_same.mean() > 0.98
    bits_diff = (npy(q) != oq) & tok_same[..., None]
    ymax = max(float(np.abs(it["patches"]).max()) for it in items)
    exempt = lfq_bit_exempt(ob.patches, std_at_tokens(opn.b, ob.patch_channels, ob.patch_positions, opn.eps), ymax)
    assert not (bits_diff & ~exempt).any()                # differences only next to the sign boundary
    assert bits_diff.mean() < 2e-4
    rec = pipe.decode(batch, q)
    ob.patches = oq
    ob.patches = opn.inverse_norm(ob)
    orec = np.stack(ofe.postprocess(ob))
    psnr = -10 * np.log10(np.mean((npy(rec) - orec) ** 2) + 1e-20)
    assert psnr > 50


# ------------------------------------------------------------------------------ full-size properties
def test_config2_shape_properties(D):
    """512^2, patch 14 (BASELINE config 2 geometry, smaller batch): size-independent properties."""
    torch.manual_seed(0)
    x = torch.rand(4, 3, 512, 512, device="cuda")
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    b = fe.process_batch(x)
    assert tuple(b.patches.shape) == (4, 3072, 196) and b.patch_sizes[0] == (36, 36)
    assert not bool(b.key_pad_mask.any())
    # every (c, h, w) appears exactly once per image; positions in range
    key = (b.patch_channels * 32 + b.patch_positions[..., 0]) * 32 + b.patch_positions[..., 1]
    assert all(len(torch.unique(k)) == 3072 for k in key)
    # Parseval on the kept block: energy(tokens) == energy of the truncated coefficient plane
    ipt = D.util.rgb_to_ipt(x)
    coef = D.util.dct2_truncated(ipt, 448, 448)
    np.testing.assert_allclose(float((b.patches.double() ** 2).sum()), float((coef.double() ** 2).sum()), rtol=5e-6)   # coef via the FFMA path, tokens via tensor cores
    full = D.util.dct2(ipt, "ortho")
    np.testing.assert_allclose(float((full.double() ** 2).sum()), float((ipt.double() ** 2).sum()), rtol=1e-5)
    # linearity of the transform
    y1 = D.util.dct2_truncated(ipt[:1], 448, 448)
    y2 = D.util.dct2_truncated(ipt[1:2], 448, 448)
    y12 = D.util.dct2_truncated(ipt[:1] + ipt[1:2], 448, 448)
    # three independently rounded results: 3x the single-transform tolerance
    assert float((y12 - y1 - y2).abs().max()) <= 3 * COEF_RTOL * float(y12.abs().max())
    # decode(encode) with nothing quantised == low-pass of the image: IDCT of the kept block
    rec = fe.postprocess_batch(b)
    ref = D.util.ipt_to_rgb(D.util.idct2_truncated(coef, 512, 512))
    assert float((rec - ref).abs().max()) <= 2e-5     # tensor-core decode vs exact-fp32 decode
    # scores are sorted: descending per image
    tiles = fe._token_grid(x)
    order = fe._sorted_order(tiles).long()
    mags = tiles.abs().amax(-1).reshape(4, -1) * 0.1
    t = torch.arange(3072, device="cuda")
    hw = (t // 3) // 32 + (t // 3) % 32
    imp = torch.tensor([8.0, 1.0, 1.0], device="cuda")[t % 3]
    sc = mags + (-hw).float() / imp
    srt = torch.gather(sc, 1, order)
    assert bool((srt[:, :-1] >= srt[:, 1:] - 1e-6).all())


def test_topk_cap_config3a_geometry(D):
    """1024^2, max_seq_len=1024: pure top-k cap, one image per row (BASELINE config 3a)."""
    torch.manual_seed(0)
    x = torch.rand(2, 3, 1024, 1024, device="cuda")
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 1024)
    b = fe.process_batch(x)
    assert tuple(b.patches.shape) == (2, 1024, 196) and b.patch_sizes[0] == (73, 73)
    ofe = O.FeatureExtractor(3, 14, 0.0, 32, 32, 1024)
    oi = ofe.preprocess(x[0].cpu().numpy())
    ours = set(zip(npy(b.patch_channels[0]).tolist(), *npy(b.patch_positions[0]).T.tolist()))
    ref = set(zip(oi["channels"].tolist(), *oi["positions"].T.tolist()))
    assert len(ours & ref) >= 1020      # the cut may move tokens within EPS_SCORE of the k-th score
    rec = fe.postprocess_batch(b)
    assert tuple(rec.shape) == (2, 3, 1024, 1024)


def test_variable_k_packing_config3b(D):
    """beta > 0: variable k from Python's RNG in the reference's draw order, several images per row."""
    torch.manual_seed(1)
    x = torch.rand(12, 3, 256, 256, device="cuda")
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.004, 32, 32, 1024)
    ofe = O.FeatureExtractor(3, 14, 0.004, 32, 32, 1024)
    random.seed(42)
    b = fe.process_batch(x)
    random.seed(42)
    ks = [ofe._choose_k(972) for _ in range(12)]
    st = ofe.group_by_max_seq_len(ks)
    rows = st["groups"] + [st["group"]]
    assert b.row_num_images() == [len(r) for r in rows]
    lengths = (~b.key_pad_mask).sum(1).tolist()
    assert lengths == [sum(ks[i] for i in r) for r in rows]
    rec = fe.postprocess(b)
    assert len(rec) == 12 and all(tuple(r.shape) == (3, 256, 256) for r in rec)
    # foreign batch (no host bookkeeping): same decode through a device read of the ids
    b2 = b.shallow_copy()
    b2._row_num_images = None
    rec2 = fe.postprocess(b2)
    assert all(torch.equal(a, c) for a, c in zip(rec, rec2))
