"""Training-time quantiser terms at config-2 scale: device times with CUDA events + the library's own
per-launch events.

  * factorised LFQ entropy loss (util.py:355-387 on lfq.py:187-204's distance), forward + backward, at
    786 432 tokens x 14 codebooks x 2^14 codes -- the dense (T, c, 2^d) tensor would be 721 GB;
  * LFQ.forward in training mode (quantise + commit loss + entropy loss) + backward at the same shape;
  * VectorQuantize.forward in training mode (nearest code, EMA codebook update, commit loss, straight-through) at
    393 216 tokens x 8192 codes x 256.

    python tools/bench_train_terms.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dct_autoencoder_b200 as D  # noqa: E402
from dct_autoencoder_b200 import _lib  # noqa: E402
from dct_autoencoder_b200.util import FactorizedDistance, compute_entropy_loss  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, fn, tokens, dev):
    ms = timed(fn)
    print("%-72s %8.3f ms  %7.1f M tokens/s" % (name, ms, tokens / ms / 1e3))
    with _lib.profile(dev) as p:
        fn()
    for g, m in p.groups:
        print("      %9.1f us  %s" % (m * 1e3, g))


def main():
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    b, n, c, d = 256, 3072, 14, 14
    T = b * n
    x = (torch.randn(b, n, c, d, device=dev) * 0.002).requires_grad_(True)
    mask = torch.rand(b, n, device=dev) > 0.1

    def entropy():
        x.grad = None
        compute_entropy_loss(FactorizedDistance(x, 1.0), mask).backward()
    report("factorised LFQ entropy loss fwd + bwd, %d x %d x 2^%d" % (T, c, d), entropy, T, dev)

    x13 = (torch.randn(b, n, 16, 13, device=dev) * 0.002).requires_grad_(True)     # conf/patch14-l.json: 16 codebooks of 2^13

    def entropy13():
        x13.grad = None
        compute_entropy_loss(FactorizedDistance(x13, 1.0), mask).backward()
    report("factorised LFQ entropy loss fwd + bwd, %d x 16 x 2^13 (conf/patch14-l.json)" % T, entropy13, T, dev)
    del x13

    lfq = D.LFQ(codebook_size=2 ** d, num_codebooks=c).to(dev).train()
    xf = x.detach().reshape(b, n, c * d).requires_grad_(True)

    def lfq_step():
        xf.grad = None
        out, idx, commit, dist = lfq(xf, mask)
        (compute_entropy_loss(dist, mask) * 0.1 + commit * 0.25 + out.sum() * 1e-9).backward()
    report("LFQ.forward (train: quantise + commit + entropy) + backward", lfq_step, T, dev)

    tv = 128 * 3072
    vq = D.VectorQuantize(dim=256, codebook_size=8192, decay=0.8, commitment_weight=1.0).to(dev).train()
    xv = torch.randn(128, 3072, 256, device=dev, requires_grad=True)
    mv = torch.rand(128, 3072, device=dev) > 0.1

    def vq_step():
        xv.grad = None
        q, idx, loss = vq(xv, mask=mv)
        (loss.sum() + q.sum() * 1e-9).backward()
    report("VectorQuantize.forward (train: nearest + EMA update + commit) + backward, %d x 8192 x 256" % tv, vq_step, tv, dev)


if __name__ == "__main__":
    main()
