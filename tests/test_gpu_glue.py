"""Model glue on the device: to_patch_embedding + LayerNorm + position embeddings, the quantiser's
projections, proj_out and decode_from_codes, against the UNMODIFIED reference model run with identity transformer
stacks (tests/golden/glue.npz, tests/golden/make_golden_configs.py) and against float64 PyTorch on the same inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import dct_autoencoder_b200 as d
    d._lib.load()
    return d


def npy(t):
    return t.detach().cpu().numpy()


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("t,k,n", [(300, 196, 1024), (1000, 1024, 196), (77, 208, 256), (129, 16, 24)])
def test_linear_rows_matches_float64(D, t, k, n):
    """x @ W^T on the split-precision tensor-core GEMM with per-row scaling: fp32-class accuracy whatever the range
    of a row (rows scaled by 1e-3 .. 1e3)."""
    from dct_autoencoder_b200.linear import linear_rows
    torch.manual_seed(t)
    x = torch.randn(t, k, device="cuda") * torch.logspace(-3, 3, t, device="cuda")[:, None]
    w = torch.randn(n, k, device="cuda") / k ** 0.5
    y = linear_rows(x, w)
    ref = x.double() @ w.double().T
    scale = (x.double().abs() @ w.double().abs().T)                    # forward error bound of an fp32 dot product
    assert float(((y.double() - ref).abs() / scale).max()) < 2e-6      # (a sequential fp32 dot product: up to k * 2^-24; the tensor-core accumulator truncates)
    ln = torch.nn.LayerNorm(k, eps=1e-4).cuda()
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.normal_(0, 0.1)
    y2 = linear_rows(x, w, ln=ln)
    ref2 = torch.nn.functional.layer_norm(x.double(), (k,), ln.weight.double(), ln.bias.double(), 1e-4) @ w.double().T
    assert float((y2.double() - ref2).abs().max()) < 2e-5 * float(ref2.abs().max())


@pytest.mark.parametrize("f", [1024, 256, 208, 20])
def test_ln_pos_rows_matches_torch(D, f):
    from dct_autoencoder_b200.linear import ln_pos_rows
    torch.manual_seed(f)
    x = torch.randn(3, 50, f, device="cuda") * 3 + 1
    ln = torch.nn.LayerNorm(f, eps=1e-4).cuda()
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.normal_(0, 0.1)
    pc, ph, pw = torch.randn(3, f, device="cuda"), torch.randn(7, f, device="cuda"), torch.randn(5, f, device="cuda")
    ch = torch.randint(0, 3, (3, 50), device="cuda")
    pos = torch.stack([torch.randint(0, 7, (3, 50), device="cuda"), torch.randint(0, 5, (3, 50), device="cuda")], -1)
    bias = torch.randn(f, device="cuda")
    got = ln_pos_rows(x, ln=ln, pos=(pc, ph, pw), channels=ch, positions=pos)
    with torch.no_grad():
        want = ln(x) + ph[pos[..., 0]] + pw[pos[..., 1]] + pc[ch]
    assert float((got - want).abs().max()) < 2e-5
    assert torch.equal(ln_pos_rows(x, pos=(pc, ph, pw), channels=ch, positions=pos), x + ph[pos[..., 0]] + pw[pos[..., 1]] + pc[ch])
    assert torch.equal(ln_pos_rows(x, bias=bias), x + bias)


def _glue_from_golden(D, g):
    m = D.DCTAutoencoderGlue(image_channels=3, max_patch_h=6, max_patch_w=6, patch_size=14, feature_dim=256,
                             vq_type="lfq", vq_codebook_size=8192, vq_num_codebooks=16)
    sd = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w:")}
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and not res.missing_keys, res          # the reference's parameter names, all of them
    m = m.cuda().eval()
    m.patchnorm.frozen = True
    return m


def _batch(D, g, patches):
    return D.DCTPatches(patches=patches, key_pad_mask=cu(g["key_pad_mask"]), batched_image_ids=cu(g["image_ids"]),
                        patch_channels=cu(g["channels"]), patch_positions=cu(g["positions"]),
                        patch_sizes=[tuple(x) for x in g["patch_sizes"].tolist()],
                        original_sizes=[tuple(x) for x in g["original_sizes"].tolist()])


def test_glue_encode_matches_the_reference_model(D, golden):
    g = golden("glue")
    m = _glue_from_golden(D, g)
    b = _batch(D, g, cu(g["in_patches"]))
    emb = m.embed(m.normalize_(b.shallow_copy()))
    # the reference's `embedded` is before the position embeddings: add them with the reference's own tables
    want = torch.from_numpy(g["embedded"]).cuda() + m.encoder_pos_embed_height[b.h_indices] + \
        m.encoder_pos_embed_width[b.w_indices] + m.encoder_pos_embed_channel[b.patch_channels]
    assert float((emb.patches - want).abs().max()) < 2e-5
    out, codes, commit, dist = m.encode(_batch(D, g, cu(g["in_patches"])), do_normalize=True)
    assert codes.dtype == torch.int64 and tuple(codes.shape) == (2, 120, 16)
    shifts = np.arange(12, -1, -1)
    bits = lambda a: ((a[..., None].astype(np.int64) >> shifts) & 1).reshape(a.shape[:-1] + (208,))
    diff = bits(npy(codes)) != bits(g["codes"])
    assert np.all(np.abs(g["lfq_pre"][diff]) < 2e-5)          # sign bits differ only where the projected value is ~0
    assert diff.mean() < 1e-3
    same = ~diff.any(-1)
    assert np.abs(npy(out.patches)[same] - g["enc_patches"][same]).max() < 2e-5


def test_glue_decode_from_codes_matches_the_reference_model(D, golden):
    g = golden("glue")
    m = _glue_from_golden(D, g)
    dec = m.decode_from_codes(cu(g["codes"].astype(np.int64)), do_inv_norm=True, key_pad_mask=cu(g["key_pad_mask"]),
                              batched_image_ids=cu(g["image_ids"]), patch_channels=cu(g["channels"]),
                              patch_positions=cu(g["positions"]),
                              patch_sizes=[tuple(x) for x in g["patch_sizes"].tolist()],
                              original_sizes=[tuple(x) for x in g["original_sizes"].tolist()])
    want = g["dec_patches"]
    assert np.abs(npy(dec.patches) - want).max() < 2e-5 * max(1.0, float(np.abs(want).max()))
    # forward = encode + decode; the pixels come out of the extractor as usual
    res = m(_batch(D, g, cu(g["in_patches"])), do_normalize=True)
    assert set(res) == {"dct_patches", "commit_loss", "codes", "distances"}
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 6, 6, 120)
    res["dct_patches"].patches = m.patchnorm.inverse_norm(res["dct_patches"])
    ims = fe.postprocess(res["dct_patches"])
    assert len(ims) == 3 and tuple(ims[2].shape) == (3, 84, 84) and all(bool(torch.isfinite(i).all()) for i in ims)
