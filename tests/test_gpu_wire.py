"""GPU tests of the compact code wire format: to_bytes / from_bytes against
to_dict / from_dict (dct_patches.py:54-122) and a pure-Python statement of the byte layout."""
import struct

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import dct_autoencoder_b200 as d
    d._lib.load()
    return d


def _pack_ref(obj, c, d):
    """The layout of include/dcta.h dcta_wire_pack, from the reference-format dict of one image."""
    out = bytearray(struct.pack("<4sBBBBHHIII", b"DCTW", 1, c, d, 0, *obj["size"], *obj["original_size"],
                                len(obj["codes"])))
    for e in obj["codes"]:
        out += struct.pack("<H", e["c"] << 12 | e["h"] << 6 | e["w"])
        bits = "".join(format(v, f"0{d}b") for v in e["data"])
        bits += "0" * (-len(bits) % 8)
        out += int(bits, 2).to_bytes(len(bits) // 8, "big")
    return bytes(out)


def _batch(D, ks, c, d, s, seed=0):
    """A packed batch with the bookkeeping process_batch produces and random code words."""
    fe = D.DCTAutoencoderFeatureExtractor(3, 8, 0.0, 8, 8, s)
    torch.manual_seed(seed)
    x = fe.process_batch(torch.rand(len(ks), 3, 64, 72).cuda(), ks)
    g = torch.Generator(device="cuda").manual_seed(seed)
    hi = torch.randint(0, 2 ** min(d, 31), x.key_pad_mask.shape + (c,), device="cuda", generator=g)
    if d == 32:
        hi = hi * 2 + torch.randint(0, 2, hi.shape, device="cuda", generator=g)
    return x, hi


@pytest.mark.parametrize("c,d", [(14, 14), (16, 13), (3, 5), (1, 32), (7, 1), (5, 8)])
def test_to_bytes_matches_the_documented_layout(D, c, d):
    ks = [40, 150, 1, 192, 64, 17]                       # several images per row, ragged rows, one full row
    x, codes = _batch(D, ks, c, d, 192)
    blobs = D.to_bytes(x, codes, 2 ** d)
    objs = D.to_dict(x, codes)
    assert len(blobs) == len(objs) == len(ks)
    assert [len(o["codes"]) for o in objs] == ks
    for blob, obj in zip(blobs, objs):
        assert blob == _pack_ref(obj, c, d)
        assert len(blob) == 24 + len(obj["codes"]) * D.dct_patches.wire_record_bytes(c, d)


def test_from_bytes_inverts_to_bytes_like_from_dict(D):
    x, codes = _batch(D, [100, 33, 192], 14, 14, 192, seed=3)
    for blob, obj in zip(D.to_bytes(x, codes, 2 ** 14), D.to_dict(x, codes)):
        dp_b, codes_b = D.from_bytes(blob)
        dp_d, codes_d = D.from_dict(obj, device="cuda")
        assert torch.equal(codes_b, codes_d)
        for f in ("key_pad_mask", "batched_image_ids", "patch_channels", "patch_positions", "attn_mask"):
            assert torch.equal(getattr(dp_b, f), getattr(dp_d, f)), f
        assert dp_b.patch_sizes == dp_d.patch_sizes and dp_b.original_sizes == dp_d.original_sizes


def test_wire_round_trip_decodes_to_the_same_image(D):
    """codes -> bytes -> codes -> decode equals decoding the original codes."""
    torch.manual_seed(1)
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 8, 8, 192)
    pn = D.PatchNorm(8, 8, 14, 3).cuda()
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).cuda().eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    imgs = torch.rand(3, 3, 112, 112).cuda()
    pipe.fit_norm(imgs)
    batch, codes = pipe.encode_codes(imgs)
    want = pipe.decode_codes(batch, codes)
    for i, blob in enumerate(D.to_bytes(batch, codes, 2 ** 14)):
        dp, c1 = D.from_bytes(blob)
        got = pipe.decode_codes(dp, c1[None])
        assert torch.equal(got[0], want[i])


def test_wire_rejects_what_it_cannot_hold(D):
    x, codes = _batch(D, [10], 14, 14, 16)
    x.patch_positions[0, 0, 0] = 64
    with pytest.raises(ValueError):
        D.to_bytes(x, codes, 2 ** 14)
    with pytest.raises(ValueError):
        D.from_bytes(b"nope" + bytes(20))
    x, codes = _batch(D, [10], 14, 14, 16)
    blob = D.to_bytes(x, codes, 2 ** 14)[0]
    with pytest.raises(ValueError):
        D.from_bytes(blob[:-1])
    with pytest.raises(D._lib.DctaError):                 # 40 x 32 bits do not fit the staging buffer
        D.to_bytes(*_batch(D, [10], 40, 32, 16), 2 ** 32)
