"""Generates vq_train.npz by running the UNMODIFIED reference VectorQuantize / kmeans (imported from /root/reference
through oracle/ref_shim.py) in TRAINING mode on CPU.  Runs only in the build container.

    python tests/golden/make_golden_vq_train.py

vq_train.npz
  ema_*     VectorQuantize(dim=32, codebook_size=64, decay=0.8, commitment_weight=0.7) (EMA Euclidean codebook,
            vector_quantize.py:436-507 + :837-1050), two training steps on seeded inputs with a mask:
            x0/x1 (2, 40, 32), mask (2, 40), embed0, and after each step embed / cluster_size / embed_avg, the returned
            quantize, indices and loss.
  dead_*    the same module with threshold_ema_dead_code=2 after one step: which codes were expired (the replacement
            vectors are random samples of the batch: only the set of expired codes and the reset statistics are pinned).
  km_*      kmeans() (vector_quantize.py:180-220) from fixed initial means: samples (1, 300, 8), means0 (1, 12, 8),
            10 iterations -> means, bins.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

ref_shim.import_reference()
from dct_autoencoder.vector_quantize import VectorQuantize, kmeans  # noqa: E402


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


out = {}
torch.manual_seed(11)
vq = VectorQuantize(dim=32, codebook_size=64, decay=0.8, commitment_weight=0.7)
vq.train()
out["ema_embed0"] = npy(vq._codebook.embed).copy()
mask = torch.ones(2, 40, dtype=torch.bool)
mask[0, 33:] = False
mask[1, 25:] = False
out["ema_mask"] = npy(mask)
for step in range(2):
    x = torch.randn(2, 40, 32) * 0.3
    q, ind, loss = vq(x, mask=mask)
    out[f"ema_x{step}"] = npy(x)
    out[f"ema_q{step}"] = npy(q)
    out[f"ema_ind{step}"] = npy(ind)
    out[f"ema_loss{step}"] = npy(loss)
    out[f"ema_embed{step + 1}"] = npy(vq._codebook.embed).copy()
    out[f"ema_cluster_size{step + 1}"] = npy(vq._codebook.cluster_size).copy()
    out[f"ema_embed_avg{step + 1}"] = npy(vq._codebook.embed_avg).copy()

torch.manual_seed(12)
vq2 = VectorQuantize(dim=16, codebook_size=32, decay=0.5, threshold_ema_dead_code=2)
vq2.train()
out["dead_embed0"] = npy(vq2._codebook.embed).copy()
x = torch.randn(3, 20, 16) * 0.2
out["dead_x"] = npy(x)
q, ind, loss = vq2(x)
out["dead_ind"] = npy(ind)
out["dead_cluster_size"] = npy(vq2._codebook.cluster_size).copy()
out["dead_embed"] = npy(vq2._codebook.embed).copy()
out["dead_embed_avg"] = npy(vq2._codebook.embed_avg).copy()

torch.manual_seed(13)
samples = torch.randn(1, 300, 8)
means0 = samples[:, torch.randperm(300)[:12]].clone()
means, bins = kmeans(samples, 12, 10, sample_fn=lambda s, n: means0.clone())
out.update(km_samples=npy(samples), km_means0=npy(means0), km_means=npy(means), km_bins=npy(bins))

path = os.path.join(HERE, "vq_train.npz")
np.savez_compressed(path, **out)
print(f"vq_train: {os.path.getsize(path) / 1024:.1f} KiB, keys={sorted(out)}")
