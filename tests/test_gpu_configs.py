"""BASELINE.json configs pinned DIRECTLY to the oracle / to fixtures of the unmodified reference, at the
benchmarked geometry ("next" 1): config 1 (the reference's own images at 256^2), config 2
(512^2, patch 14, fused LFQ path), config 3a / 3b (1024^2 with the max_seq_len cap / variable k + packing).
Tolerances: tests/parity_rules.py."""
import random

import numpy as np
import pytest
import torch

import dcta_oracle as O
from parity_rules import (COEF_RTOL, COEF_RTOL_NATURAL, EPS_LFQ, EPS_SCORE, lfq_bit_exempt, order_equal_up_to_score_ties,
                          std_at_tokens)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import dct_autoencoder_b200 as d
    d._lib.load()
    return d


def npy(t):
    return t.detach().cpu().numpy()


def collate(items):
    return {k: [it[k] for it in items] for k in items[0]}


def oracle_order(ofe, im):
    """Full descending token order of one image + the sorted scores + max|Y| (the oracle's own arithmetic)."""
    coef = ofe._crop_image(ofe._transform_image_in(im))
    _, h_idx, w_idx, scores = ofe.importance_scores(coef)
    c = coef.shape[0]
    flat = scores.reshape(-1)
    order = np.argsort(-flat.astype(np.float64), kind="stable")
    hs = np.repeat(h_idx[:, None], c, 1).reshape(-1)[order]
    ws = np.repeat(w_idx[:, None], c, 1).reshape(-1)[order]
    cs = np.repeat(np.arange(c, dtype=np.int64)[None, :], len(h_idx), 0).reshape(-1)[order]
    return (cs * 64 + hs) * 64 + ws, flat[order], float(np.abs(coef).max())


def token_keys(channels, positions):
    return (channels * 64 + positions[..., 0]) * 64 + positions[..., 1]


def check_tokens_and_bits(batch, codes, ob, ocodes, onormed, opn, ofe, images, lfq_bits, rows_of_image):
    """Shared body: selection order per image under EPS_SCORE, then code bits under EPS_LFQ on tokens that sit at
    the same slot in both.  Returns (fraction of slots holding the same token, number of differing bits)."""
    ch, pos = npy(batch.patch_channels), npy(batch.patch_positions)
    assert np.array_equal(npy(batch.key_pad_mask), ob.key_pad_mask)
    assert np.array_equal(npy(batch.batched_image_ids), ob.batched_image_ids)
    ours_keys = token_keys(ch, pos)
    ymax = 0.0
    for i, (r, lo, k) in enumerate(rows_of_image):
        full_keys, full_scores, ym = oracle_order(ofe, images[i])
        ymax = max(ymax, ym)
        assert np.array_equal(token_keys(ob.patch_channels, ob.patch_positions)[r, lo:lo + k], full_keys[:k])
        ok, moved = order_equal_up_to_score_ties(ours_keys[r, lo:lo + k], full_keys, full_scores)
        assert ok, f"image {i}: selection order differs outside score ties ({moved} tokens moved)"
    same_tok = (ours_keys == token_keys(ob.patch_channels, ob.patch_positions)) & ~ob.key_pad_mask
    # code words -> bits, MSB first (LFQ:87): compare bit by bit on identical tokens
    c = codes.shape[-1]
    shifts = np.arange(lfq_bits - 1, -1, -1)
    bits = lambda a: ((a[..., None] >> shifts) & 1).reshape(a.shape[:-1] + (c * lfq_bits,)).astype(bool)
    diff = (bits(npy(codes)) != bits(ocodes)) & same_tok[..., None]
    std = std_at_tokens(opn.b, ob.patch_channels, ob.patch_positions, opn.eps)
    exempt = lfq_bit_exempt(onormed, std, ymax)
    assert not (diff & ~exempt).any(), "a sign bit differs where |Y - median| >= EPS_LFQ * max|Y|"
    return same_tok.sum() / max(1, (~ob.key_pad_mask).sum()), int(diff.sum())


# ------------------------------------------------------------------------------------------------ config 2
def test_config2_fused_path_matches_the_oracle_at_512(D):
    """BASELINE config 2 at its own geometry (512^2, patch 14, K = 448, 3072 tokens, LFQ 14 x 14 bit) on 8 images:
    the FUSED roundtrip (fold_codes_kernel + decode from code words, what bench.py times) against
    O.run_pipeline -- token order, positions, channels, ids, masks, code bits, reconstructions."""
    torch.manual_seed(0)
    x = torch.rand(8, 3, 512, 512)
    torch.manual_seed(1)
    x_fit = torch.rand(8, 3, 512, 512)
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    pn = D.PatchNorm(32, 32, 14, 3).cuda()
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).cuda().eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    pipe.fit_norm(x_fit.cuda())
    assert pipe.fusable()
    assert bool(D._lib.load().dcta_fold_codes_supported(512, 512, 448, 448, 14))
    rec, codes = pipe.roundtrip(x.cuda())
    batch, codes2 = pipe.encode_codes(x.cuda())
    assert torch.equal(codes, codes2) and batch.patches is None

    ofe = O.FeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    opn = O.PatchNorm(32, 32, 14, 3)
    opn.forward(next(ofe.iter_batches(iter([collate([ofe.preprocess(im.numpy()) for im in x_fit])]), None)))
    opn.frozen = True
    # the fitted statistics themselves: counts exact, medians / b to the coefficient tolerance
    assert np.array_equal(npy(pn.n), opn.n)
    ymax_fit = float(np.abs(npy(pn.median)).max())
    assert np.abs(npy(pn.median) - opn.median).max() <= COEF_RTOL * max(ymax_fit, 256.0)
    np.testing.assert_allclose(npy(pn.b), opn.b, rtol=1e-4, atol=COEF_RTOL * 256.0)

    olfq = O.LFQ(codebook_size=2 ** 14, num_codebooks=14)
    items = [ofe.preprocess(im.numpy()) for im in x]
    ob = next(ofe.iter_batches(iter([collate(items)]), None))
    onormed = opn.forward(ob)
    oq, ocodes, _, _ = olfq.forward(onormed, ~ob.key_pad_mask)
    rows = [(i, 0, 3072) for i in range(8)]
    frac, nbits = check_tokens_and_bits(batch, codes, ob, ocodes, onormed, opn, ofe, x.numpy(), 14, rows)
    assert frac > 0.999
    assert nbits < 2e-4 * ocodes.size * 14
    # reconstruction: decode OUR codes with the oracle (isolates the decode side), then the whole round trip
    ob2 = ob
    ob2.patches = olfq.indices_to_codes(npy(codes))
    ob2.patch_channels, ob2.patch_positions = npy(batch.patch_channels), npy(batch.patch_positions)
    ob2.patches = opn.inverse_norm(ob2)
    orec = np.stack(ofe.postprocess(ob2))
    assert np.abs(npy(rec) - orec).max() < 1e-4                      # same codes, their tables vs ours
    orec_full, _ = O.run_pipeline(x.numpy(), ofe, opn, olfq)
    mse = np.mean((npy(rec) - np.stack(orec_full)) ** 2)
    assert -10 * np.log10(mse + 1e-20) > 60                           # a few bits at the sign boundary flipped
    p_ours = -10 * np.log10(np.mean((npy(rec) - x.numpy()) ** 2))
    p_ref = -10 * np.log10(np.mean((np.stack(orec_full) - x.numpy()) ** 2))
    assert abs(p_ours - p_ref) < 0.01                                 # round-trip PSNR within 0.01 dB


# ------------------------------------------------------------------------------------------------ config 3
def test_config3a_topk_cap_matches_the_oracle_at_1024(D):
    """1024^2, max_seq_len 1024, beta 0: the 3072 in-bounds candidates are cut to the best 1024 (SURVEY 8d config
    3a).  Token order vs the oracle under EPS_SCORE, patches to the coefficient tolerance, decode vs oracle."""
    torch.manual_seed(0)
    x = torch.rand(3, 3, 1024, 1024)
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 1024)
    ofe = O.FeatureExtractor(3, 14, 0.0, 32, 32, 1024)
    b = fe.process_batch(x.cuda())
    assert tuple(b.patches.shape) == (3, 1024, 196) and b.patch_sizes[0] == (73, 73)
    keys = token_keys(npy(b.patch_channels), npy(b.patch_positions))
    for i in range(3):
        full_keys, full_scores, ymax = oracle_order(ofe, x[i].numpy())
        ok, moved = order_equal_up_to_score_ties(keys[i], full_keys, full_scores)
        assert ok, (i, moved)
        it = ofe.preprocess(x[i].numpy())
        same = keys[i] == full_keys[:1024]
        assert same.mean() > 0.99
        assert np.abs(npy(b.patches[i])[same] - it["patches"][same]).max() <= COEF_RTOL * ymax
    # channel split of the cut (SURVEY 8d [measured] 968 / 28 / 28 on U[0,1) noise): luminance dominates
    counts = np.bincount(npy(b.patch_channels[0]), minlength=3)
    assert counts[0] > 900 and counts.sum() == 1024
    ob = next(ofe.iter_batches(iter([collate([ofe.preprocess(im.numpy()) for im in x])]), None))
    ob.patches, ob.patch_channels, ob.patch_positions = npy(b.patches), npy(b.patch_channels), npy(b.patch_positions)
    orec = np.stack(ofe.postprocess(ob))
    rec = fe.postprocess_batch(b)
    assert np.abs(npy(rec) - orec).max() < 5e-5


def test_config3b_variable_k_packing_matches_the_oracle_at_1024(D):
    """1024^2, beta 0.004, max_seq_len 1024, random.seed(42): variable k drawn in the reference's RNG order,
    several images per row, padding, multi-image revert_patching -- per-image token sets and
    order, packed rows, ids, masks, the fused codes and the decoded images against the oracle."""
    n = 10
    torch.manual_seed(2)
    x = torch.rand(n, 3, 1024, 1024)
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.004, 32, 32, 1024)
    ofe = O.FeatureExtractor(3, 14, 0.004, 32, 32, 1024)
    pn = D.PatchNorm(32, 32, 14, 3).cuda()
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).cuda().eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    random.seed(7)
    pipe.fit_norm(x[:4].flip(-1).cuda())
    random.seed(42)
    b = fe.process_batch(x.cuda())
    random.seed(42)
    items = [ofe.preprocess(im.numpy()) for im in x]
    ks = [it["patches"].shape[0] for it in items]
    assert len(set(ks)) > 3, ks                                 # really variable
    ob = next(ofe.iter_batches(iter([collate(items)]), None))
    assert b.row_num_images() == [int(ob.batched_image_ids[r][~ob.key_pad_mask[r]].max()) + 1 for r in range(ob.patches.shape[0])]
    assert max(b.row_num_images()) > 1                          # rows really are shared
    assert [tuple(s) for s in b.patch_sizes] == [tuple(s) for s in ob.patch_sizes]
    # (row, offset, k) of every image, from the oracle's packing
    rows, r, off = [], 0, 0
    for k in ks:
        if off + k > 1024:
            r, off = r + 1, 0
        rows.append((r, off, k))
        off += k
    # PatchNorm tables: use OURS on both sides (the fit is pinned by its own tests), so bits isolate the encode
    opn = O.PatchNorm(32, 32, 14, 3)
    opn.n, opn.median, opn.b = npy(pn.n), npy(pn.median), npy(pn.b)
    opn.frozen = True
    olfq = O.LFQ(codebook_size=2 ** 14, num_codebooks=14)
    onormed = opn.forward(ob)
    oq, ocodes, _, _ = olfq.forward(onormed, ~ob.key_pad_mask)
    random.seed(42)
    batch, codes = pipe.encode_codes(x.cuda())
    frac, nbits = check_tokens_and_bits(batch, codes, ob, ocodes, onormed, opn, ofe, x.numpy(), 14, rows)
    assert frac > 0.99
    # staged patches agree with the oracle's on identical tokens
    same = (token_keys(npy(b.patch_channels), npy(b.patch_positions)) == token_keys(ob.patch_channels, ob.patch_positions))
    same &= ~ob.key_pad_mask
    assert np.abs(npy(b.patches)[same] - ob.patches[same]).max() <= COEF_RTOL * 512 * 1.01
    assert not npy(b.patches)[ob.key_pad_mask].any()            # padding rows are zeros (UT:149-164)
    # decode our codes with both implementations
    rec = pipe.decode_codes(batch, codes)
    ob.patches = olfq.indices_to_codes(npy(codes))
    ob.patch_channels, ob.patch_positions = npy(batch.patch_channels), npy(batch.patch_positions)
    ob.patches = opn.inverse_norm(ob)
    orec = np.stack(ofe.postprocess(ob))
    assert np.abs(npy(rec) - orec).max() < 1e-4


# ------------------------------------------------------------------------------------------------ config 1
def test_config1_real_images_match_the_reference(D, golden):
    """BASELINE config 1: the reference's own images/*.jpg at 256^2, B = 16, conf/patch14-l.json geometry;
    preprocess -> iter_batches(None) -> postprocess.  The fixture holds the UNMODIFIED reference's outputs
    (tests/golden/make_golden_configs.py)."""
    g = golden("config1")
    ims = torch.from_numpy(g["images"]).float() / 255
    x = torch.stack([ims[i % 13] for i in range(16)])
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    # the reference's calling sequence, image by image
    items = [fe.preprocess(im.cuda()) for im in x]
    assert all(it["patches"].shape == (972, 196) for it in items)          # 18 x 18 x 3 tokens
    b = next(fe.iter_batches(iter([D.dict_collate(items)]), None))
    assert tuple(b.patches.shape) == (6, 3072, 196)                          # 3 images per row, 6 rows
    assert np.array_equal(npy(b.key_pad_mask), g["key_pad_mask"])
    assert np.array_equal(npy(b.batched_image_ids), g["image_ids"].astype(np.int64))
    assert [tuple(s) for s in b.patch_sizes] == [tuple(s) for s in g["patch_sizes"].tolist()]
    assert [tuple(s) for s in b.original_sizes] == [tuple(s) for s in g["original_sizes"].tolist()]
    ofe = O.FeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    keys = token_keys(npy(b.patch_channels), npy(b.patch_positions))
    gkeys = token_keys(g["channels"].astype(np.int64), g["positions"].astype(np.int64))
    n_moved = 0
    for i in range(16):
        r, lo = i // 3, (i % 3) * 972
        full_keys, full_scores, _ = oracle_order(ofe, x[i].numpy())
        ok, moved = order_equal_up_to_score_ties(keys[r, lo:lo + 972], gkeys[r, lo:lo + 972], full_scores)
        assert ok, (i, moved)
        n_moved += moved
    # (natural images: the high-frequency tiles of the two chroma channels are ~0, so tokens on the same
    #  anti-diagonal h + w have scores within EPS_SCORE of each other and may swap -- about 7 % of the tokens do)
    assert n_moved <= 16 * 972 * 0.15
    ymax = float(np.abs(g["patches0"]).max())
    same0 = keys[0, :256] == gkeys[0, :256]
    # coefficients: the bound is stated against the float64 definition (the oracle); the fixture comes from the
    # reference's fp32 FFT path, which carries its own error of up to 2e-7 * max|Y| against that definition
    # (tests/test_oracle.py::test_standin_matches_float64_definition), hence + 2e-7 against the fixture
    ours0 = npy(b.patches[0, :256])
    it0 = ofe.preprocess(x[0].numpy())
    okeys0 = token_keys(it0["channels"], it0["positions"])[:256]
    same_o = keys[0, :256] == okeys0
    err_o = np.abs(ours0[same_o] - it0["patches"][:256][same_o]).max()
    assert err_o <= COEF_RTOL_NATURAL * ymax, (err_o, ymax, err_o / ymax)
    err0 = np.abs(ours0[same0] - g["patches0"][same0]).max()
    assert err0 <= (COEF_RTOL_NATURAL + 2e-7) * ymax, (err0, ymax, err0 / ymax)
    rec = fe.postprocess(b)
    assert len(rec) == 16
    assert np.abs(npy(rec[11]) - g["rec_f32_11"]).max() < 5e-5
    rec_u8 = torch.stack([(r.clamp(0, 1) * 255).round().to(torch.uint8) for r in rec[:4]])
    d8 = np.abs(npy(rec_u8).astype(np.int16) - g["rec_u8"].astype(np.int16))
    assert d8.max() <= 1 and (d8 > 0).mean() < 1e-3                         # 8-bit output: rounding ties only
    psnr = np.array([float(-10 * torch.log10(((r.cpu() - im) ** 2).mean())) for r, im in zip(rec, x)])
    assert np.abs(psnr - g["psnr"]).max() < 0.01                             # 42.9 .. 54.8 dB, mean 49.3
    # the batched path gives the same images
    rb = fe.postprocess_batch(fe.process_batch(x.cuda()))
    assert all(torch.equal(rb[i], rec[i]) for i in range(16))


def test_config1_model_quantiser_lfq_8192x16_matches_the_reference(D, golden):
    """The conf/patch14-l.json quantiser (LFQ dim 196 -> 16 codebooks x 13 bits with 196 -> 208 -> 196 projections)
    on PatchNorm-normalised tokens of a real image, against the unmodified reference's outputs."""
    g = golden("config1")
    lfq = D.LFQ(dim=196, codebook_size=8192, num_codebooks=16).cuda().eval()
    with torch.no_grad():
        lfq.project_in.weight.copy_(torch.from_numpy(g["lfq16_w_in"]))
        lfq.project_in.bias.copy_(torch.from_numpy(g["lfq16_b_in"]))
        lfq.project_out.weight.copy_(torch.from_numpy(g["lfq16_w_out"]))
        lfq.project_out.bias.copy_(torch.from_numpy(g["lfq16_b_out"]))
        x = torch.from_numpy(g["lfq16_in"]).cuda()[None]
        q, idx, _, _ = lfq(x, torch.ones(1, x.shape[1], dtype=torch.bool, device="cuda"))
    pre = g["lfq16_pre"]                                         # the reference's projected values
    shifts = np.arange(12, -1, -1)
    bits = lambda a: ((a[..., None].astype(np.int64) >> shifts) & 1).reshape(a.shape[0], -1)
    diff = bits(npy(idx[0])) != bits(g["lfq16_idx"])
    assert np.all(np.abs(pre[diff]) < 1e-5)                      # fp32 GEMM rounding next to zero only
    assert diff.mean() < 1e-3
    same_tok = ~diff.any(-1)
    assert np.abs(npy(q[0])[same_tok] - g["lfq16_q"][same_tok]).max() < 1e-5
    assert np.abs(npy(lfq.indices_to_codes(idx))[0][same_tok] - g["lfq16_q"][same_tok]).max() < 1e-5
