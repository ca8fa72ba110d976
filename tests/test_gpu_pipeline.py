"""GPU tests of the assembled pipeline helpers (host streaming, both DCT implementations)."""
import numpy as np
import pytest
import torch

import dcta_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import dct_autoencoder_b200 as d
    d._lib.load()
    return d


def _pipe(D, impl, max_seq_len=3072):
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, max_seq_len, dct_impl=impl)
    pn = D.PatchNorm(32, 32, 14, 3).cuda()
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).cuda().eval()
    return D.TransformPipeline(fe, pn, lfq)


def test_roundtrip_host_equals_device_roundtrip(D):
    torch.manual_seed(0)
    x = torch.rand(10, 3, 128, 160)
    pipe = _pipe(D, "tc", max_seq_len=9 * 11 * 3)      # one image per row, so rows == images
    pipe.fit_norm(torch.rand(6, 3, 128, 160).cuda())
    rec, codes = pipe.roundtrip(x.cuda())
    hx = x.pin_memory()
    out_images, out_codes = pipe.roundtrip_host(hx, chunk=4)      # 3 chunks, last one ragged
    torch.cuda.synchronize()
    assert torch.equal(out_images, rec.cpu()) and torch.equal(out_codes, codes.cpu())
    o2 = torch.empty_like(out_images).pin_memory()
    c2 = torch.empty_like(out_codes).pin_memory()
    pipe.roundtrip_host(hx, o2, c2, chunk=16)
    torch.cuda.synchronize()
    assert torch.equal(o2, rec.cpu()) and torch.equal(c2, codes.cpu())


def test_tc_and_fp32_pipelines_agree(D):
    """The tensor-core and exact-fp32 DCT implementations give the same codes (outside the sign
    boundary) and the same images to 2e-5."""
    torch.manual_seed(1)
    x = torch.rand(4, 3, 256, 256).cuda()
    fit = torch.rand(8, 3, 256, 256).cuda()
    a, b = _pipe(D, "tc"), _pipe(D, "fp32")
    a.fit_norm(fit)
    b.fit_norm(fit)
    assert torch.allclose(a.norm.median, b.norm.median, rtol=1e-6, atol=2e-5)
    ba, qa, ca = a.encode(x)
    bb, qb, cb = b.encode(x)
    same_tok = (ba.patch_channels == bb.patch_channels) & (ba.patch_positions == bb.patch_positions).all(-1)
    assert float(same_tok.float().mean()) > 0.99
    diff = (qa != qb) & same_tok[..., None]
    assert float(diff.float().mean()) < 2e-4
    assert float(bb.patches[diff].abs().max() if diff.any() else 0.0) < 1e-3
    ra, rb = a.decode(ba, qa), b.decode(bb, qb)
    # decode the SAME quantised tokens with both implementations
    rb2 = b.decode(ba, qa)
    assert float((ra - rb2).abs().max()) < 2e-5


@pytest.mark.parametrize("h,w", [(90, 101), (300, 451), (255, 320), (128, 112)])
def test_any_image_size_runs_on_the_folded_tensor_core_path(D, h, w):
    """Sizes that are not multiples of 16 (or are odd) stay on the folded tcgen05 path: quadrants of ceil(n/2) samples
    with padded pitches, the middle row / column of an odd size paired with itself (csrc/dct_fold.cu fold_any_kernel).
    Same tokens as the exact-fp32 FFMA extractor and the oracle within the DCT tolerance, same fused codes round trip."""
    torch.manual_seed(2)
    x = torch.rand(3, 3, h, w).cuda()
    fe_tc = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072, dct_impl="tc")
    fe_32 = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072, dct_impl="fp32")
    _, _, th, tw = fe_tc._geometry(h, w)
    assert D.util.fold_ok(h, w, th * 14, tw * 14)
    l0 = D._lib.launch_count
    g_tc = fe_tc._token_grid(x)
    assert D._lib.launch_count - l0 == 4          # plane means + fold kernel + two fold_gemm launches: no FFMA kernel
    g_32 = fe_32._token_grid(x)
    scale = float(g_32.abs().max())
    assert float((g_tc - g_32).abs().max()) <= 8e-7 * scale
    ofe = O.FeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    it = ofe.preprocess(x[0].cpu().numpy())
    b1 = fe_tc.process_batch(x)
    k = it["patches"].shape[0]
    # same token set and order wherever the scores are not within the tie rule (random images: no near ties expected)
    same = (b1.patch_positions[0, :k].cpu().numpy() == it["positions"]).all(-1) & (b1.patch_channels[0, :k].cpu().numpy() == it["channels"])
    assert same.mean() > 0.99
    assert np.abs(b1.patches[0, :k].cpu().numpy()[same] - it["patches"][same]).max() < 1e-4
    r1, r2 = fe_tc.postprocess_batch(b1), fe_32.postprocess_batch(fe_32.process_batch(x))
    assert float((r1 - r2).abs().max()) < 2e-5
    # fused codes path (forward pass 2 -> code words, decode inside inverse pass 1) against the staged modules
    pn = D.PatchNorm(32, 32, 14, 3).cuda()
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).cuda().eval()
    pipe = D.TransformPipeline(fe_tc, pn, lfq)
    pipe.fit_norm(torch.rand(4, 3, h, w).cuda())
    rec_s, codes_s = pipe.roundtrip(x, fused=False)
    rec_f, codes_f = pipe.roundtrip(x)
    assert torch.equal(codes_f, codes_s)
    assert torch.equal(rec_f, rec_s)


@pytest.mark.parametrize("beta,max_seq_len,size,patch,cb", [
    (0.0, 3072, (256, 256), 14, (14, 14)), (0.0, 972, (256, 256), 14, (14, 14)),
    (0.01, 512, (128, 160), 14, (14, 14)), (0.0, 200, (128, 160), 14, (14, 14)),
    (0.0, 3072, (128, 128), 4, (4, 4)),        # vector kernel, several codebooks per 32-bit word
    (0.02, 300, (96, 120), 4, (2, 8)),         # codebook != patch row
    (0.0, 3072, (96, 96), 3, (3, 3)),          # z % 4 != 0: scalar kernel
    # 1024-sample axes (K = 512): only the hi plane of the basis is resident in both forward kernels (BASELINE config 3)
    (0.0, 1024, (1024, 1024), 14, (14, 14)), (0.0, 3072, (1024, 720), 14, (14, 14)),
])
def test_fused_roundtrip_is_bit_identical_to_staged(D, beta, max_seq_len, size, patch, cb):
    """encode_codes / decode_codes (PatchNorm + LFQ inside the pack / un-patchify kernels) give the
    same codes, metadata and images, bit for bit, as the module-by-module path -- including padding
    slots, several images per row, top-k cuts and variable k."""
    import random
    torch.manual_seed(4)
    h, w = size
    x = torch.rand(9, 3, h, w).cuda()
    fe = D.DCTAutoencoderFeatureExtractor(3, patch, beta, 32, 32, max_seq_len)
    pn = D.PatchNorm(32, 32, patch, 3).cuda()
    lfq = D.LFQ(codebook_size=2 ** cb[1], num_codebooks=cb[0]).cuda().eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    random.seed(3)
    pipe.fit_norm(torch.rand(8, 3, h, w).cuda())
    assert pipe.fusable()
    random.seed(11)
    rec_s, codes_s = pipe.roundtrip(x, fused=False)
    random.seed(11)
    batch_s, q_s, _ = pipe.encode(x)
    random.seed(11)
    batch_f, codes_f = pipe.encode_codes(x)
    assert batch_f.patches is None
    for f in ("key_pad_mask", "batched_image_ids", "patch_channels", "patch_positions"):
        assert torch.equal(getattr(batch_f, f), getattr(batch_s, f)), f
    assert torch.equal(codes_f, codes_s)
    rec_f = pipe.decode_codes(batch_f, codes_f)
    assert torch.equal(rec_f, rec_s)
    random.seed(11)
    rec_auto, codes_auto = pipe.roundtrip(x)
    assert torch.equal(rec_auto, rec_s) and torch.equal(codes_auto, codes_s)
    # decode_codes from codes alone == the reference's decode_from_codes chain
    b2 = batch_s.shallow_copy()
    b2.patches = lfq.indices_to_codes(codes_s)
    b2.patches = pn.inverse_norm(b2)
    assert torch.equal(fe.postprocess_batch(b2), rec_f)
    # not fusable -> staged path is used transparently
    lfq2 = D.LFQ(dim=patch * patch, codebook_size=8192, num_codebooks=16).cuda().eval()
    assert not D.TransformPipeline(fe, pn, lfq2).fusable()


def test_fused_codes_match_staged_when_values_tie_with_the_median(D):
    """PatchNorm fitted on the very images that are encoded: a quarter of the coefficients EQUAL their median, so any
    last-bit difference between the kernel that wrote the fitted coefficients (token grid) and the one that forms the
    code words in its epilogue would flip sign bits.  Both must produce bit-identical coefficients."""
    torch.manual_seed(0)
    for shape in [(4, 3, 128, 112), (8, 3, 512, 512)]:
        x = torch.rand(*shape).cuda()
        fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
        pn = D.PatchNorm(32, 32, 14, 3).cuda()
        lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).cuda().eval()
        pipe = D.TransformPipeline(fe, pn, lfq)
        pipe.fit_norm(x)
        assert pipe.fusable()
        rec_s, codes_s = pipe.roundtrip(x, fused=False)
        rec_f, codes_f = pipe.roundtrip(x)
        assert torch.equal(codes_f, codes_s)
        assert torch.equal(rec_f, rec_s)


@pytest.mark.parametrize("patch,size,max_seq_len,beta", [
    (16, (256, 256), 1024, 0.0),       # 16-column tiles: 16 token columns per 256-row data tile
    (8, (128, 192), 3072, 0.0),        # 8-column tiles: 32 token columns per tile
    (14, (512, 512), 3072, 0.0),       # config 2
    (14, (1024, 1024), 1024, 0.004),   # config 3b: top-k cap, variable k, several images per row (K = 512: only the hi
                                       # plane of the basis is resident, its lo tiles ride in the ring)
    (12, (192, 240), 700, 0.0),        # 12-column tiles: 20 token columns = 240 rows per tile
])
def test_codes_in_the_dct_epilogue_for_other_tile_sizes(D, patch, size, max_seq_len, beta):
    """fold_codes_kernel (forward pass 2 straight to code words) for every tile size it accepts: same codes,
    metadata and reconstructions, bit for bit, as the staged modules."""
    import random
    from dct_autoencoder_b200 import _lib
    torch.manual_seed(patch)
    h, w = size
    n = 3 if h >= 1024 else 6
    x = torch.rand(n, 3, h, w).cuda()
    fe = D.DCTAutoencoderFeatureExtractor(3, patch, beta, 32, 32, max_seq_len)
    _, _, th, tw = fe._geometry(h, w)
    assert bool(_lib.load().dcta_fold_codes_supported(h, w, th * patch, tw * patch, patch))
    pn = D.PatchNorm(32, 32, patch, 3).cuda()
    lfq = D.LFQ(codebook_size=2 ** patch, num_codebooks=patch).cuda().eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    random.seed(1)
    pipe.fit_norm(torch.rand(4, 3, h, w).cuda())
    assert pipe.fusable()
    random.seed(5)
    rec_s, codes_s = pipe.roundtrip(x, fused=False)
    random.seed(5)
    batch_s, _, _ = pipe.encode(x)
    random.seed(5)
    batch_f, codes_f = pipe.encode_codes(x)
    for f in ("key_pad_mask", "batched_image_ids", "patch_channels", "patch_positions"):
        assert torch.equal(getattr(batch_f, f), getattr(batch_s, f)), f
    assert torch.equal(codes_f, codes_s)
    assert torch.equal(pipe.decode_codes(batch_f, codes_f), rec_s)


def test_pack_on_the_side_stream_changes_nothing(D):
    """The fused round trip launches the sort and the gather into the packed tensors on a second stream when every token
    was kept (the decode only needs the code grid): results, metadata and repeated steps are those of one stream."""
    torch.manual_seed(11)
    pipe = _pipe(D, "tc")
    assert pipe.overlap_pack
    pipe.fit_norm(torch.rand(4, 3, 512, 512).cuda())
    xs = [torch.rand(12, 3, 512, 512).cuda() for _ in range(3)]
    pipe.overlap_pack = False
    want = []
    for x in xs:
        batch, codes, rec = pipe._fused_roundtrip(x, None, torch.float32)
        want.append((rec.clone(), codes.clone(), batch.patch_positions.clone(), batch.patch_channels.clone(),
                     batch.batched_image_ids.clone(), batch.key_pad_mask.clone()))
    pipe.overlap_pack = True
    for rep in range(3):               # back-to-back steps recycle the allocator's blocks: a missing join shows up here
        got = [pipe._fused_roundtrip(x, None, torch.float32) for x in xs]
        torch.cuda.synchronize()
        for (batch, codes, rec), w in zip(got, want):
            assert torch.equal(rec, w[0]) and torch.equal(codes, w[1])
            assert torch.equal(batch.patch_positions, w[2]) and torch.equal(batch.patch_channels, w[3])
            assert torch.equal(batch.batched_image_ids, w[4]) and torch.equal(batch.key_pad_mask, w[5])
    assert getattr(pipe.extractor, "_pack_pending", None) is None
    # a top-k cut needs the packed codes for the decode: nothing is moved to the side stream, same results either way
    ks = [2000] * 12
    r1, c1 = pipe.roundtrip(xs[0], ks)
    pipe.overlap_pack = False
    r2, c2 = pipe.roundtrip(xs[0], ks)
    assert torch.equal(r1, r2) and torch.equal(c1, c2)


@pytest.mark.parametrize("beta,max_seq_len,size", [(0.0, 972, 256), (0.01, 512, 256), (0.0, 700, 300)])
def test_projection_lfq_roundtrip_is_bit_identical_to_staged(D, beta, max_seq_len, size):
    """conf/patch14-l.json's quantiser (LFQ with 196 -> 208 -> 196 projections): PatchNorm.forward runs inside the operand
    split of project_in and inverse_norm inside the un-patchify kernel; codes and pixels are those of the staged modules
    (padding rows, several images per row, top-k cuts included)."""
    import random
    torch.manual_seed(21)
    random.seed(5)
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, beta, 32, 32, max_seq_len)
    pn = D.PatchNorm(32, 32, 14, 3).cuda()
    lfq = D.LFQ(dim=196, codebook_size=8192, num_codebooks=16).cuda().eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    pipe.fit_norm(torch.rand(5, 3, size, size + 16).cuda(), ks=[min(600, max_seq_len)] * 5)
    x = torch.rand(7, 3, size, size + 16).cuda()
    assert not pipe.fusable() and pipe.proj_fusable()
    n_tok = (size // 14) * ((size + 16) // 14) * 3
    ks = [max(1, min(n_tok, max_seq_len, int(v))) for v in torch.randint(50, 1200, (7,))] if beta > 0 else None
    want_rec, want_codes = pipe.roundtrip_staged(x, ks)
    rec, codes = pipe.roundtrip(x, ks)
    assert torch.equal(codes, want_codes)
    assert torch.equal(rec, want_rec)
    # the two fused entry points on their own
    batch = fe.process_batch(x, ks)
    from dct_autoencoder_b200.linear import _split_rows
    raw = batch.patches.reshape(-1, 196)
    a = _split_rows(raw, 0.5, patchnorm=(pn, batch.patch_channels, batch.patch_positions))
    b = _split_rows(pn(batch).reshape(-1, 196), 0.5)
    assert all(torch.equal(u, v) for u, v in zip(a, b))
    q = batch.shallow_copy()
    q.patches = torch.randn_like(batch.patches)
    fused = fe.postprocess_batch(q, denorm=pn)
    q.patches = pn.inverse_norm(q)
    assert torch.equal(fused, fe.postprocess_batch(q))


def test_graphed_roundtrip_replays_the_eager_step(D):
    torch.manual_seed(4)
    pipe = _pipe(D, "tc")
    x = torch.rand(6, 3, 256, 256).cuda()
    pipe.fit_norm(torch.rand(4, 3, 256, 256).cuda())
    rec, codes = pipe.roundtrip(x)
    g = pipe.graphed(x)
    assert g.launches > 0
    for _ in range(3):
        r2, c2 = g()
        torch.cuda.synchronize()
        assert torch.equal(r2, rec) and torch.equal(c2, codes)
    y = torch.rand(6, 3, 256, 256).cuda()
    want_rec, want_codes = pipe.roundtrip(y)
    r3, c3 = g(y)                                    # another tensor of the same shape is copied in first
    assert torch.equal(r3, want_rec) and torch.equal(c3, want_codes)
    # staged modules under capture as well (x now holds y's values: the graph reads its input in place)
    assert torch.equal(x, y)
    gs = pipe.graphed(x, fused=False)
    r4, c4 = gs()
    assert torch.equal(r4, want_rec) and torch.equal(c4, want_codes)
    # the captured kernels read tables derived from the statistics: after a re-fit a replay must refuse, a new graph works
    pipe.fit_norm(torch.rand(4, 3, 256, 256).cuda())
    with pytest.raises(RuntimeError):
        g()
    g2 = pipe.graphed(x)
    r5, c5 = g2()
    w5, wc5 = pipe.roundtrip(x)
    assert torch.equal(r5, w5) and torch.equal(c5, wc5) and not torch.equal(r5, want_rec)


def test_config2_full_batch_properties(D):
    """BASELINE config 2 at its full size (256 x 512^2, patch 14, 14 x 14-bit LFQ), checked through properties
    that do not need the CPU oracle: every image's codes and reconstruction are independent of the batch it is
    processed in (bit for bit, full batch vs chunks of 32 vs the staged modules), every (channel, h, w) token
    appears exactly once per image, code words are 14-bit, and the reconstruction error is that of the
    quantiser, not of the transform."""
    torch.manual_seed(0)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(256, 3, 512, 512, device="cuda", generator=g)
    pipe = _pipe(D, "tc")
    pipe.fit_norm(torch.rand(32, 3, 512, 512, device="cuda", generator=g))
    batch, codes = pipe.encode_codes(x)
    rec = pipe.decode_codes(batch, codes)
    assert tuple(codes.shape) == (256, 3072, 14) and codes.dtype == torch.int64
    assert int(codes.min()) >= 0 and int(codes.max()) < 2 ** 14
    assert not bool(batch.key_pad_mask.any())
    key = (batch.patch_channels * 32 + batch.patch_positions[..., 0]) * 32 + batch.patch_positions[..., 1]
    assert bool((torch.sort(key, dim=1).values == torch.arange(3072, device="cuda")).all())
    # checksum of checksums: chunks of 32 images reproduce the full batch exactly
    for lo in range(0, 256, 32):
        r2, c2 = pipe.roundtrip(x[lo:lo + 32])
        assert torch.equal(c2, codes[lo:lo + 32]) and torch.equal(r2, rec[lo:lo + 32]), lo
    # ... and so do the staged drop-in modules (on a slice, they materialise every intermediate)
    r3, c3 = pipe.roundtrip_staged(x[64:96])
    assert torch.equal(c3, codes[64:96]) and torch.equal(r3, rec[64:96])
    # un-quantised round trip = low-pass of the image; quantised one is bounded by the clamp of PatchNorm
    plain = pipe.extractor.postprocess_batch(pipe.extractor.process_batch(x[:8]))
    assert float((plain - x[:8]).abs().mean()) < float((rec[:8] - x[:8]).abs().mean())
    assert bool(torch.isfinite(rec).all())


def test_uint8_pixels_in_and_out(D):
    """8-bit pixels on both sides (what the reference's callers hold: read_image(path) / 255 in, save_image out):
    uint8 input == the fp32 input x / 255 bit for bit on every path; uint8 output == save_image's quantisation of the
    fp32 output; the compact host round trip == the fp32 one after those two conversions, codes as wire records."""
    torch.manual_seed(5)
    # all 256 byte values: u8 -> float is torch's IEEE division, float -> u8 is floor(clamp(x)*255 + 0.5)
    b = torch.arange(256, dtype=torch.uint8).repeat(3).cuda()[:767]
    # (the CPU is the reference: torch's CUDA kernel multiplies by fl(1/255) instead, which is 1 ulp off for 126 values)
    assert torch.equal(D.util.u8_to_unit(b).cpu(), b.cpu().float() / 255)
    v = torch.cat([torch.rand(1001, device="cuda") * 1.2 - 0.1, torch.tensor([0.0, 1.0, 0.5 / 255, 1.5 / 255, float("nan")], device="cuda")])
    want = v.nan_to_num(0.0).clamp(0, 1).mul(255).add_(0.5).clamp_(0, 255).to(torch.uint8)
    assert torch.equal(D.util.unit_to_u8(v), want)

    x8 = torch.randint(0, 256, (6, 3, 128, 160), dtype=torch.uint8)
    xf = x8.float() / 255
    pipe = _pipe(D, "tc", max_seq_len=9 * 11 * 3)
    pipe.fit_norm(torch.rand(6, 3, 128, 160).cuda())
    rec_f, codes_f = pipe.roundtrip(xf.cuda())
    rec_8, codes_8 = pipe.roundtrip(x8.cuda(), out_dtype=torch.uint8)
    assert torch.equal(codes_8, codes_f)
    assert rec_8.dtype == torch.uint8 and torch.equal(rec_8, D.util.unit_to_u8(rec_f))
    # staged modules and the per-image API take uint8 too
    b8, bf = pipe.extractor.process_batch(x8.cuda()), pipe.extractor.process_batch(xf.cuda())
    assert b8.patches.dtype == torch.float32 and torch.equal(b8.patches, bf.patches)
    i8, i_f = pipe.extractor.preprocess(x8[0]), pipe.extractor.preprocess(xf[0])
    assert torch.equal(i8["patches"], i_f["patches"]) and torch.equal(i8["positions"], i_f["positions"])
    assert torch.equal(pipe.extractor.postprocess_batch(bf, out_dtype=torch.uint8), D.util.unit_to_u8(pipe.extractor.postprocess_batch(bf)))
    # host to host, compact: uint8 in, uint8 + wire records out
    out_img, records, counts = pipe.roundtrip_host(x8.pin_memory(), chunk=4, compact=True)
    torch.cuda.synchronize()
    assert torch.equal(out_img, rec_8.cpu()) and counts.tolist() == [9 * 11 * 3] * 6
    batch, _ = pipe.encode_codes(xf.cuda())
    blobs = D.to_bytes(batch, codes_f, 2 ** 14)
    for i, blob in enumerate(blobs):
        assert blob[D.dct_patches.WIRE_HEADER_BYTES:] == records[i, :int(counts[i])].numpy().tobytes()
        dp, c2 = D.from_bytes(blob)
        assert torch.equal(c2, codes_f[i])


@pytest.mark.parametrize("size,max_seq_len,n,beta", [
    ((512, 512), 3072, 5, 0.0),        # the benchmark geometry: every token kept, 5 images (a partial group of 8)
    ((256, 256), 700, 9, 0.0),         # top-k cut: dropped tokens decode to zero; 126 coefficient rows per parity
    ((128, 160), 3072, 8, 0.01),       # variable k, odd row count per parity (63), K padded to a 32-column block
    ((64, 48), 3072, 3, 0.0),          # one 32-row block, mostly padding
])
def test_decode_inside_inverse_pass1_is_bit_identical(D, size, max_seq_len, n, beta):
    """dcta_decode_codes_inv_fold (the operand of inverse pass 1 generated in shared memory from the code bits,
    csrc/dct_fold.cu fold_gemm_kernel<0, true>) against the separate decode kernel + plain inverse: same pixels,
    bit for bit (FE:607-656, PN:167-177, LFQ:105-134, FE:289-310)."""
    import random
    from dct_autoencoder_b200.util import decode_codes_inv_fold_ok
    torch.manual_seed(7)
    h, w = size
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, beta, 32, 32, max_seq_len)
    pn = D.PatchNorm(32, 32, 14, 3).cuda()
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).cuda().eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    random.seed(5)
    pipe.fit_norm(torch.rand(6, 3, h, w).cuda())
    kh, kw = min(h // 14, 32) * 14, min(w // 14, 32) * 14
    assert decode_codes_inv_fold_ok(h, w, kh, kw, 14, 14, 14)
    x = torch.rand(n, 3, h, w).cuda()
    batch, codes = pipe.encode_codes(x)
    assert fe.decode_in_gemm
    rec_gen = pipe.decode_codes(batch, codes)
    rec_gen_u8 = pipe.decode_codes(batch, codes, out_dtype=torch.uint8)
    fe.decode_in_gemm = False
    rec_sep = pipe.decode_codes(batch, codes)
    rec_sep_u8 = pipe.decode_codes(batch, codes, out_dtype=torch.uint8)
    assert torch.isfinite(rec_gen).all()
    assert torch.equal(rec_gen, rec_sep)
    assert torch.equal(rec_gen_u8, rec_sep_u8)
    # roundtrip(): when every token was kept the decode reads the forward pass's code grid (dcta_decode_grid_inv_fold)
    fe.decode_in_gemm = True
    random.seed(5)
    l0 = D._lib.launch_count
    rec_rt, codes_rt = pipe.roundtrip(x)
    n_launch = D._lib.launch_count - l0
    random.seed(5)
    batch2, codes2 = pipe.encode_codes(x)
    assert torch.equal(codes_rt, codes2) and torch.equal(rec_rt, pipe.decode_codes(batch2, codes2))
    kept_all = beta == 0.0 and max_seq_len >= (min(h // 14, 32) * min(w // 14, 32) * 3)
    assert n_launch == (11 if kept_all else 12), n_launch       # no slot-map kernel on the grid path


def test_decode_refuses_statistics_beyond_the_fp16_operand_range(D):
    """The tensor-core decode carries |AC coefficient| < 2^11 in its fp16 hi/lo planes; statistics that put
    median +- scale*(b*sqrt(2)+eps) beyond that are refused where the two-value table is built instead of overflowing
    silently (the reference's fp32 path, patchnorm.py:167-177, has no such limit: dct_impl="fp32" keeps it)."""
    torch.manual_seed(0)
    x = torch.rand(2, 3, 128, 112).cuda()
    pipe = _pipe(D, "tc")
    pipe.fit_norm(x)
    rec, codes = pipe.roundtrip(x)
    assert torch.isfinite(rec).all()
    with torch.no_grad():
        pipe.norm.b[1, 2, 3, 5] = 4000.0
    with pytest.raises(ValueError, match="2\\^11"):
        pipe.roundtrip(x)
    with torch.no_grad():
        pipe.norm.b[1, 2, 3, 5] = 1.0
        pipe.norm.median[2, 0, 0, 0] = 1.0e5       # a DC term: carried in fp32, no limit
    rec, _ = pipe.roundtrip(x)
    assert torch.isfinite(rec).all()
