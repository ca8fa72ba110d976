"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference through oracle/ref_shim.py) on seeded inputs.  Runs only in the build
container; the committed .npz files are what travels to the GPU box.

    python tests/golden/make_golden.py

Fixture keys are documented next to each block.  ``torch_dct`` is the stand-in described in
oracle/torch_dct_standin.py (parity of the DCT arithmetic is therefore "unpinned" at the bit
level).
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

ref_shim.import_reference()
from dct_autoencoder import util as U  # noqa: E402
from dct_autoencoder.dataset import dict_collate  # noqa: E402
from dct_autoencoder.dct_patches import DCTPatches  # noqa: E402
from dct_autoencoder.feature_extraction_dct_autoencoder import DCTAutoencoderFeatureExtractor  # noqa: E402
from dct_autoencoder.lfq import LFQ  # noqa: E402
from dct_autoencoder.patchnorm import PatchNorm  # noqa: E402
from dct_autoencoder.vector_quantize import VectorQuantize  # noqa: E402


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: npy(v) for k, v in kw.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, keys={list(kw)}")


def batch_fields(prefix, b: DCTPatches):
    return {
        prefix + "patches": b.patches, prefix + "key_pad_mask": b.key_pad_mask,
        prefix + "attn_mask": b.attn_mask, prefix + "image_ids": b.batched_image_ids,
        prefix + "channels": b.patch_channels, prefix + "positions": b.patch_positions,
        prefix + "patch_sizes": np.asarray(b.patch_sizes, dtype=np.int64),
        prefix + "original_sizes": np.asarray(b.original_sizes, dtype=np.int64),
    }


# ---------------------------------------------------------------------------- constants
save("constants", Trgb2lms=U.Trgb2lms, Tlms2rgb=U.Tlms2rgb, Mipt=U.Mipt, MiptInv=U.Mipt.inverse())

# ---------------------------------------------------------------------------- colourspace
torch.manual_seed(10)
x = torch.rand(2, 3, 12, 20)
x[0, :, :2] = -x[0, :, :2]          # negative lobes exercise the sign branch (UT:76-78)
x[1, :, 0, :4] = 0.0
ipt = U.rgb_to_ipt(x.clone())
save("colorspace", rgb=x, ipt=ipt, rgb_back=U.ipt_to_rgb(ipt.clone()))

# ---------------------------------------------------------------------------- transform in/out (DCT)
fe = DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
torch.manual_seed(11)
tr = {}
for i, (h, w) in enumerate([(64, 64), (45, 70), (128, 96)]):
    im = torch.rand(3, h, w)
    co = fe._transform_image_in(im)
    tr[f"im{i}"] = im
    tr[f"coef{i}"] = co
    tr[f"back{i}"] = fe._transform_image_out(co)
save("transform", **tr)

# ---------------------------------------------------------------------------- preprocess (select / top-k)
pp = {}
cases = [
    # name, ctor kwargs, image size, seed
    ("a", dict(channels=3, patch_size=14, sample_patches_beta=0.0, max_patch_h=32, max_patch_w=32, max_seq_len=3072), (3, 72, 60)),
    ("b", dict(channels=3, patch_size=4, sample_patches_beta=0.0, max_patch_h=5, max_patch_w=4, max_seq_len=40), (3, 33, 26)),
    ("c", dict(channels=3, patch_size=8, sample_patches_beta=0.05, max_patch_h=6, max_patch_w=6, max_seq_len=64), (3, 50, 64)),
    ("d", dict(channels=1, patch_size=2, sample_patches_beta=0.0, max_patch_h=8, max_patch_w=8, max_seq_len=64,
               channel_importances=(8.0,)), (1, 9, 17)),
]
torch.manual_seed(12)
random.seed(42)
for name, kw, shape in cases:
    fx = DCTAutoencoderFeatureExtractor(**kw)
    if kw["channels"] != 3:
        # the colour transform is 3-channel only; as in the reference's own testpatching.py:42-43
        fx._transform_image_in = lambda x: x
    im = torch.rand(*shape)
    coef = fx._crop_image(fx._transform_image_in(im))
    out = fx.preprocess(im)
    pp[name + "_im"] = im
    pp[name + "_coef"] = coef
    pp[name + "_patches"] = out["patches"]
    pp[name + "_positions"] = out["positions"]
    pp[name + "_channels"] = out["channels"]
    pp[name + "_original_size"] = np.asarray(out["original_sizes"])
    pp[name + "_patch_size"] = np.asarray(out["patch_sizes"])
save("preprocess", **pp)

# ---------------------------------------------------------------------------- iter_batches / packing / postprocess
kw = dict(channels=3, patch_size=8, sample_patches_beta=0.0, max_patch_h=4, max_patch_w=4, max_seq_len=80)
fx = DCTAutoencoderFeatureExtractor(**kw)
torch.manual_seed(13)
sizes = [(3, 32, 32), (3, 17, 40), (3, 24, 24), (3, 40, 9), (3, 33, 33), (3, 16, 16), (3, 32, 20), (3, 8, 8), (3, 30, 30)]
ims = [torch.rand(*s) for s in sizes]
items = [fx.preprocess(im) for im in ims]
pk = {f"im{i}": im for i, im in enumerate(ims)}
for i, it in enumerate(items):
    pk[f"k{i}"] = np.asarray(it["patches"].shape[0])
# batch_size=None: one shot over everything
b_none = next(fx.iter_batches(iter([dict_collate(items)]), None))
pk.update(batch_fields("none_", b_none))
rec = fx.postprocess(b_none)
for i, r in enumerate(rec):
    pk[f"none_rec{i}"] = r
planes = fx.revert_patching(b_none)
for i, r in enumerate(planes):
    pk[f"none_plane{i}"] = r
# batch_size=2, loader delivering 3 images at a time
loader = iter([dict_collate(items[i:i + 3]) for i in range(0, 9, 3)])
got = list(fx.iter_batches(loader, 2))
pk["bs2_num_batches"] = np.asarray(len(got))
for j, b in enumerate(got):
    pk.update(batch_fields(f"bs2_{j}_", b))
save("packing", **pk)

# ---------------------------------------------------------------------------- PatchNorm
torch.manual_seed(14)
kw = dict(channels=3, patch_size=4, sample_patches_beta=0.0, max_patch_h=3, max_patch_w=3, max_seq_len=27)
fx = DCTAutoencoderFeatureExtractor(**kw)
pn = PatchNorm(3, 3, 4, 3)
pnk = {}
pn.train()
for step in range(2):
    ims = [torch.rand(3, 12 + 4 * (i % 2), 12) * (1 + step) for i in range(5 + step)]
    items = [fx.preprocess(im) for im in ims]
    b = next(fx.iter_batches(iter([dict_collate(items)]), None))
    pnk.update(batch_fields(f"s{step}_", b))
    out = pn(b)
    pnk[f"s{step}_out"] = out
    pnk[f"s{step}_n"] = pn.n.data.clone()
    pnk[f"s{step}_median"] = pn.median.data.clone()
    pnk[f"s{step}_b"] = pn.b.data.clone()
pn.frozen = True
normed = pn(b)
pnk["fwd"] = normed
b2 = b.shallow_copy()
b2.patches = normed
pnk["inv"] = pn.inverse_norm(b2)
save("patchnorm", **pnk)

# ---------------------------------------------------------------------------- LFQ
torch.manual_seed(15)
lq = {}
x = torch.randn(2, 7, 12)
x[0, 0, :3] = 0.0
mask = torch.ones(2, 7, dtype=torch.bool)
mask[1, 4:] = False
lfq = LFQ(codebook_size=16, num_codebooks=3)          # dim = 12, no projections
lfq.eval()
q, idx, commit, dist = lfq(x, mask)
lq.update(a_x=x, a_mask=mask, a_q=q, a_idx=idx, a_codes=lfq.indices_to_codes(idx))
lfq.train()
q, idx, commit, dist = lfq(x, mask)
lq.update(a_train_q=q, a_train_idx=idx, a_commit=commit, a_dist=dist,
          a_entropy=U.compute_entropy_loss(dist, mask))
lfq2 = LFQ(dim=10, codebook_size=8, num_codebooks=4)  # 10 -> 12 -> 10 projections
lfq2.eval()
x2 = torch.randn(3, 5, 10)
m2 = torch.ones(3, 5, dtype=torch.bool)
with torch.no_grad():
    q, idx, _, _ = lfq2(x2, m2)
    lq.update(b_x=x2, b_q=q, b_idx=idx, b_codes=lfq2.indices_to_codes(idx),
              b_w_in=lfq2.project_in.weight, b_b_in=lfq2.project_in.bias,
              b_w_out=lfq2.project_out.weight, b_b_out=lfq2.project_out.bias)
lfq3 = LFQ(codebook_size=2 ** 14, num_codebooks=14)   # config-2 quantizer, 196 sign bits
lfq3.eval()
x3 = torch.randn(1, 9, 196)
q, idx, _, _ = lfq3(x3, torch.ones(1, 9, dtype=torch.bool))
lq.update(c_x=x3, c_q=q, c_idx=idx)
codes = torch.randint(0, 16, (4, 50))
codes[0, :5] = -1
lq.update(p_codes=codes, p_perplexity=U.calculate_perplexity(codes, 16))
save("lfq", **lq)

# ---------------------------------------------------------------------------- VectorQuantize (eval)
torch.manual_seed(16)
vk = {}
vq = VectorQuantize(dim=32, codebook_size=64)
vq.eval()
x = torch.randn(3, 11, 32)
mask = torch.ones(3, 11, dtype=torch.bool)
mask[2, 6:] = False
with torch.no_grad():
    q, ind, loss = vq(x, mask=mask)
vk.update(a_x=x, a_mask=mask, a_embed=vq.codebook, a_q=q, a_ind=ind, a_loss=loss)
vq2 = VectorQuantize(dim=24, codebook_size=32, heads=4, codebook_dim=8)   # model variant: shared codebook, projections
vq2.eval()
x = torch.randn(2, 9, 24)
with torch.no_grad():
    q, ind, loss = vq2(x, mask=torch.ones(2, 9, dtype=torch.bool))
vk.update(b_x=x, b_embed=vq2.codebook, b_q=q, b_ind=ind,
          b_w_in=vq2.project_in.weight, b_b_in=vq2.project_in.bias,
          b_w_out=vq2.project_out.weight, b_b_out=vq2.project_out.bias)
save("vq", **vk)

# ---------------------------------------------------------------------------- whole path, patch 14
torch.manual_seed(17)
fx = DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
ims = [torch.rand(3, 96, 96) for _ in range(4)]
items = [fx.preprocess(im) for im in ims]
fit = next(fx.iter_batches(iter([dict_collate([fx.preprocess(torch.rand(3, 96, 96)) for _ in range(6)])]), None))
pn = PatchNorm(32, 32, 14, 3)
pn.train()
pn(fit)
pn.frozen = True
b = next(fx.iter_batches(iter([dict_collate(items)]), None))
lfq = LFQ(codebook_size=2 ** 14, num_codebooks=14).eval()
nb = b.shallow_copy()
nb.patches = pn(b)
q, codes, _, _ = lfq(nb.patches, ~nb.key_pad_mask)
nb.patches = q
nb.patches = pn.inverse_norm(nb)
rec = fx.postprocess(nb)
used = (pn.n.data > 0)
save("pipeline", ims=torch.stack(ims), fit_patches=fit.patches, fit_channels=fit.patch_channels,
     fit_positions=fit.patch_positions, fit_key_pad_mask=fit.key_pad_mask,
     n=pn.n.data[:, :7, :7], median=pn.median.data[:, :7, :7], b=pn.b.data[:, :7, :7],
     n_used=np.asarray(int(used.sum())),
     patches=b.patches, positions=b.patch_positions, channels=b.patch_channels,
     image_ids=b.batched_image_ids, key_pad_mask=b.key_pad_mask,
     codes=codes, rec=torch.stack(rec))
print("done")
