"""VectorQuantize codebook learning on the GPU: EMA update, commitment loss, dead-code expiry and the k-means
steps against the UNMODIFIED reference run in training mode (tests/golden/vq_train.npz, make_golden_vq_train.py) and
against the oracle's restatement (oracle/dcta_oracle.py vq_train_step / vq_kmeans; VQ:180-220, :417-434, :479-500)."""
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


def test_ema_training_steps_match_the_reference(D, golden):
    g = golden("vq_train")
    vq = D.VectorQuantize(dim=32, codebook_size=64, decay=0.8, commitment_weight=0.7).cuda()
    vq.train()
    with torch.no_grad():
        vq._codebook.embed.copy_(torch.from_numpy(g["ema_embed0"]))
        vq._codebook.embed_avg.copy_(torch.from_numpy(g["ema_embed0"]))
    mask = torch.from_numpy(g["ema_mask"]).cuda()
    for step in range(2):
        x = torch.from_numpy(g[f"ema_x{step}"]).cuda().requires_grad_(True)
        q, ind, loss = vq(x, mask=mask)
        assert np.array_equal(ind.cpu().numpy(), g[f"ema_ind{step}"])
        np.testing.assert_allclose(q.detach().cpu().numpy(), g[f"ema_q{step}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(loss.detach().cpu().numpy(), g[f"ema_loss{step}"], rtol=5e-6)
        cb = vq._codebook
        np.testing.assert_allclose(cb.cluster_size.cpu().numpy(), g[f"ema_cluster_size{step + 1}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(cb.embed_avg.cpu().numpy(), g[f"ema_embed_avg{step + 1}"], rtol=5e-6, atol=2e-7)
        np.testing.assert_allclose(cb.embed.cpu().numpy(), g[f"ema_embed{step + 1}"], rtol=1e-5, atol=2e-7)
        # straight-through estimator + commitment loss: d(loss + sum(q)) / dx on unmasked tokens (VQ:944-952, 986-994)
        (loss.sum() + q.sum()).backward()
        n_valid = float(mask.sum()) * 32
        want = 1.0 + 0.7 * 2.0 * (g[f"ema_x{step}"] - g[f"ema_q{step}"]) / n_valid
        got = x.grad.cpu().numpy()
        m = g["ema_mask"]
        np.testing.assert_allclose(got[m], want[m], rtol=1e-4, atol=1e-6)
        assert np.all(got[~m] == 1.0)          # masked tokens pass through unchanged (VQ:1043-1048)
    # eval afterwards uses the learned codebook and leaves it alone
    vq.eval()
    before = vq._codebook.embed.clone()
    x = torch.from_numpy(g["ema_x0"]).cuda()
    q, ind, loss = vq(x, mask=mask)
    oi, _, _ = O.vq_nearest(g["ema_x0"].reshape(-1, 32), before[0].cpu().numpy())
    assert np.array_equal(ind.cpu().numpy().reshape(-1), oi) and torch.equal(before, vq._codebook.embed)
    assert float(loss) == 0.0


def test_dead_codes_are_replaced_like_the_reference(D, golden):
    g = golden("vq_train")
    torch.manual_seed(0)
    vq = D.VectorQuantize(dim=16, codebook_size=32, decay=0.5, threshold_ema_dead_code=2).cuda()
    vq.train()
    with torch.no_grad():
        vq._codebook.embed.copy_(torch.from_numpy(g["dead_embed0"]))
        vq._codebook.embed_avg.copy_(torch.from_numpy(g["dead_embed0"]))
    x = torch.from_numpy(g["dead_x"]).cuda()
    q, ind, loss = vq(x)
    assert np.array_equal(ind.cpu().numpy(), g["dead_ind"])
    cs, ref_cs = vq._codebook.cluster_size.cpu().numpy()[0], g["dead_cluster_size"][0]
    np.testing.assert_allclose(cs, ref_cs, rtol=1e-6)
    # the codes the reference expired: its new embedding is one of the batch vectors (a surviving code can also sit at
    # cluster size 2.0, the reset value = threshold, VQ:268)
    flat = g["dead_x"].reshape(-1, 16)
    expired = np.array([np.abs(flat - g["dead_embed"][0][c]).max(-1).min() == 0.0 for c in range(32)])
    assert expired.any() and (~expired).any() and np.all(ref_cs[expired] == 2.0)
    emb, avg = vq._codebook.embed.cpu().numpy()[0], vq._codebook.embed_avg.cpu().numpy()[0]
    np.testing.assert_allclose(emb[~expired], g["dead_embed"][0][~expired], rtol=1e-5, atol=2e-7)
    # replaced codes are batch vectors (which ones is random), with embed_avg = vector * reset_cluster_size
    for c in np.nonzero(expired)[0]:
        assert np.abs(flat - emb[c]).max(-1).min() == 0.0
        np.testing.assert_allclose(avg[c], emb[c] * 2.0, rtol=1e-6)


def test_kmeans_steps_match_the_reference(D, golden):
    from dct_autoencoder_b200 import _lib
    from dct_autoencoder_b200.vector_quantize import invalidate_codebook_cache, nearest_code
    g = golden("vq_train")
    x = torch.from_numpy(g["km_samples"][0]).cuda()
    means = torch.from_numpy(g["km_means0"][0]).cuda().contiguous()
    cb = D.vector_quantize.EuclideanCodebook(8, 12).cuda()
    for _ in range(10):
        buckets, _ = nearest_code(x, means, return_quantized=False, impl="fp32")
        counts, sums = cb.cluster_stats(0, x, buckets, None)
        _lib.call("dcta_vq_kmeans_means", _lib.ptr(means), _lib.ptr(counts), _lib.ptr(sums), 12, 8, _lib.stream_ptr(x.device))
        invalidate_codebook_cache()
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), g["km_bins"][0])
    np.testing.assert_allclose(means.cpu().numpy(), g["km_means"][0], rtol=5e-6, atol=5e-7)
    om, ob = O.vq_kmeans(g["km_samples"][0], g["km_means0"][0], 10)
    np.testing.assert_allclose(means.cpu().numpy(), om, rtol=5e-6, atol=5e-7)


def test_kmeans_init_and_ema_at_scale(D):
    """k-means initialisation on the first batch, then EMA steps, at a size where the statistics kernel matters
    (200 k tokens x 1024 codes x 64): the codebook must track the data (quantisation error far below the data variance
    of 1 and not drifting under the EMA steps on fresh batches) and the running cluster sizes must keep summing to the
    token count."""
    torch.manual_seed(1)
    centres = torch.randn(1024, 64, device="cuda")
    vq = D.VectorQuantize(dim=64, codebook_size=1024, kmeans_init=True, kmeans_iters=5, decay=0.9).cuda()
    vq.train()
    errs = []
    for step in range(4):
        x = centres[torch.randint(0, 1024, (200000,), device="cuda")] + 0.05 * torch.randn(200000, 64, device="cuda")
        q, ind, loss = vq(x[None])
        errs.append(float(((q - x[None]) ** 2).mean()))
    assert bool(vq._codebook.initted)
    assert errs[-1] <= 1.05 * errs[0] and errs[-1] < 0.3, errs
    cs = vq._codebook.cluster_size.sum()
    assert abs(float(cs) - 200000.0) < 1.0          # k-means sets it to the bin counts; every EMA step keeps the total
