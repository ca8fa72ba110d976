"""Training-time quantiser terms at scale: the factorised LFQ entropy loss and the commitment loss,
forward and backward, against the oracle's dense restatement of util.py:355-387 / lfq.py:187-200 at sizes where the
dense (T, c, 2^d) tensor exists, against float64 autograd of the dense formula, and at the benchmark shape
(786 432 tokens x 14 codebooks x 2^14 codes: 721 GB dense) where only the factorised form can run."""
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


def dense_loss_f64(x, mask, scale, temperature=0.01, eps=1e-9):
    """util.py:355-387 on lfq.py:191's distance, float64 torch ops (differentiable)."""
    b, n, c, d = x.shape
    codes = torch.arange(2 ** d, device=x.device)
    bits = ((codes[:, None] >> torch.arange(d - 1, -1, -1, device=x.device)) & 1).double()
    cb = bits * 2 * scale - scale
    aff = -2 * torch.einsum("bncd,jd->bncj", x.double(), cb)
    m = mask.reshape(b * n).double()
    aff = aff.reshape(b * n, c, -1)
    logits = aff / temperature + eps
    probs, logp = logits.softmax(-1), logits.log_softmax(-1)
    avg = (probs * m[:, None, None] / m.sum()).sum(0).mean(0)
    avg_ent = -(avg * (avg + eps).log()).sum()
    sample = -(((probs * logp).sum(-1) * m[:, None]) / m.sum()).sum()
    return sample - avg_ent


@pytest.mark.parametrize("c,d,scale_x", [(3, 4, 0.004), (2, 8, 0.01), (4, 7, 0.003), (1, 1, 0.01), (2, 11, 0.002), (3, 12, 0.002), (16, 13, 0.002), (14, 14, 0.0015)])
def test_factorized_entropy_matches_the_dense_formula(D, c, d, scale_x):
    from dct_autoencoder_b200.util import FactorizedDistance, compute_entropy_loss
    torch.manual_seed(d)
    b, n = 3, 41
    x = (torch.randn(b, n, c, d, device="cuda") * scale_x).requires_grad_(True)      # soft distributions at T = 0.01
    mask = torch.rand(b, n, device="cuda") > 0.25
    loss = compute_entropy_loss(FactorizedDistance(x, 1.0), mask)
    loss.backward()
    x64 = x.detach().double().requires_grad_(True)
    ref = dense_loss_f64(x64, mask, 1.0)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 2e-5 * max(1.0, abs(float(ref)))
    gmax = float(x64.grad.abs().max())
    assert float((x.grad.double() - x64.grad).abs().max()) < 2e-4 * gmax
    assert not bool(x.grad[~mask].any())
    if d <= 8:   # the oracle's numpy restatement of the reference (float32, dense)
        xn = x.detach().cpu().numpy()
        olfq = O.LFQ(codebook_size=2 ** d, num_codebooks=c)
        dist = -2.0 * np.einsum("bncd,jd->bncj", xn, olfq.codebook.astype(np.float32))
        want = O.compute_entropy_loss(dist.astype(np.float32), mask.cpu().numpy())
        assert abs(float(loss) - float(want)) < 2e-4 * max(1.0, abs(float(want)))
        # and our own dense kernel from the materialised distance
        dense = compute_entropy_loss(FactorizedDistance(x.detach(), 1.0).dense(), mask)
        assert abs(float(loss) - float(dense)) < 2e-4 * max(1.0, abs(float(dense)))


def test_lfq_training_forward_switches_to_the_factorized_distance(D):
    """LFQ.forward in training mode: dense distance while it is small, FactorizedDistance beyond dense_distance_limit;
    both give the same entropy loss, commit loss and gradients."""
    from dct_autoencoder_b200.util import FactorizedDistance, compute_entropy_loss
    torch.manual_seed(0)
    x = (torch.randn(2, 30, 24, device="cuda") * 0.01).requires_grad_(True)
    mask = torch.rand(2, 30, device="cuda") > 0.2
    grads = []
    for limit in (2 ** 28, 0):
        lfq = D.LFQ(codebook_size=2 ** 8, num_codebooks=3, dense_distance_limit=limit).cuda().train()
        out, idx, commit, dist = lfq(x, mask)
        assert isinstance(dist, FactorizedDistance) == (limit == 0)
        loss = compute_entropy_loss(dist, mask) * 0.1 + commit * 0.25
        (g,) = torch.autograd.grad(loss, x)
        grads.append((float(loss), g))
    assert abs(grads[0][0] - grads[1][0]) < 1e-5 * max(1.0, abs(grads[0][0]))
    assert float((grads[0][1] - grads[1][1]).abs().max()) < 2e-4 * float(grads[0][1].abs().max())
    # commit loss gradient against autograd of the formula (lfq.py:195-200)
    x2 = x.detach().double().requires_grad_(True)
    q = torch.where(x2 > 0, 1.0, -1.0)
    m = mask.double()[..., None]
    ref = (((x2 - q) ** 2) * m / mask.sum()).sum(0).sum(0).mean()
    lfq = D.LFQ(codebook_size=2 ** 8, num_codebooks=3).cuda().train()
    _, _, commit, _ = lfq(x, mask)
    (gc,) = torch.autograd.grad(commit, x)
    (gr,) = torch.autograd.grad(ref, x2)
    assert abs(float(commit) - float(ref)) < 1e-5 * float(ref)
    assert float((gc.double() - gr).abs().max()) < 1e-5 * float(gr.abs().max())


def test_factorized_entropy_at_the_benchmark_shape(D):
    """786 432 tokens x 14 codebooks x 2^14 codes (config 2's token count): the dense tensor would be 721 GB.
    Checked through properties: equals the loss of a 1/16 subsample within sampling error when the tokens are i.i.d.,
    hard (saturated) inputs give sample entropy ~ 0 and avg entropy = log-count entropy of the code histogram,
    gradients are finite and vanish on masked tokens."""
    import time
    from dct_autoencoder_b200.util import FactorizedDistance, compute_entropy_loss
    torch.manual_seed(1)
    T, c, d = 256 * 3072, 14, 14
    x = (torch.randn(256, 3072, c, d, device="cuda") * 0.002).requires_grad_(True)
    mask = torch.rand(256, 3072, device="cuda") > 0.1
    torch.cuda.synchronize()
    t0 = time.time()
    loss = compute_entropy_loss(FactorizedDistance(x, 1.0), mask)
    loss.backward()
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"factorised entropy fwd+bwd at {T} x {c} x 2^{d}: {dt * 1e3:.1f} ms")
    sub = compute_entropy_loss(FactorizedDistance(x.detach()[:16], 1.0), mask[:16])
    assert abs(float(loss) - float(sub)) < 5e-2 * abs(float(sub))
    assert bool(torch.isfinite(x.grad).all()) and not bool(x.grad[~mask].any())
    # saturated inputs: every softmax is one-hot, the loss is minus the entropy of the code histogram
    xs = torch.randn(8, 3072, c, d, device="cuda")
    xs = torch.sign(xs) * (0.5 + xs.abs())                  # |x| >= 0.5: |u| >= 200, every Bernoulli is saturated
    ms = torch.ones(8, 3072, dtype=torch.bool, device="cuda")
    ls = compute_entropy_loss(FactorizedDistance(xs, 1.0), ms)
    idx = ((xs > 0).long() * (2 ** torch.arange(d - 1, -1, -1, device="cuda"))).sum(-1).reshape(-1)
    hist = torch.bincount(idx, minlength=2 ** d).double() / idx.numel()
    want = (hist * (hist + 1e-9).log()).sum()
    assert abs(float(ls) - float(want)) < 2e-3 * abs(float(want))


def test_tensor_core_forward_equals_the_fma_forward(D):
    """d = 14 and d = 13 (the conf/patch14-l.json quantiser): the forward contraction on tcgen05 (fp16 hi / lo operands, TMEM accumulators flushed every 1024 pairs)
    against the fp32 FMA kernel on the same input -- loss, entropies and the gradient that is computed from its tables;
    token counts that leave a ragged last stage, several flush rounds per CTA, masked tokens, saturated and soft inputs."""
    from dct_autoencoder_b200 import _lib
    from dct_autoencoder_b200.util import FactorizedDistance, compute_entropy_loss
    lib = _lib.load()
    torch.manual_seed(5)
    for (b, n, c, scale_x, d) in [(1, 3, 1, 0.002, 14), (7, 501, 14, 0.002, 14), (64, 3000, 14, 0.004, 14), (3, 700, 5, 0.05, 14),
                                  (5, 333, 16, 0.002, 13), (32, 3000, 16, 0.003, 13), (2, 90, 3, 0.08, 13)]:
        x = (torch.randn(b, n, c, d, device="cuda") * scale_x)
        mask = torch.rand(b, n, device="cuda") > 0.2
        mask[0, 0] = True
        out = []
        for on in (1, 0):
            lib.dcta_lfq_entropy_use_tensor_cores(on)
            try:
                xr = x.clone().requires_grad_(True)
                loss = compute_entropy_loss(FactorizedDistance(xr, 1.0), mask)
                loss.backward()
                out.append((float(loss), xr.grad.clone()))
            finally:
                lib.dcta_lfq_entropy_use_tensor_cores(1)
        (l_tc, g_tc), (l_fma, g_fma) = out
        assert abs(l_tc - l_fma) < 1e-5 * max(1.0, abs(l_fma)), (b, n, c, l_tc, l_fma)
        gmax = float(g_fma.abs().max())
        assert float((g_tc - g_fma).abs().max()) < 2e-5 * gmax
