"""Generates the BASELINE-config fixtures by running the UNMODIFIED reference (imported from
/root/reference through oracle/ref_shim.py).  Runs only in the build container.

    python tests/golden/make_golden_configs.py

config1.npz   BASELINE config 1: the reference's own images/*.jpg (13 files; `books.jpeg` is not matched
              by *.jpg), decoded as RGB (ImageReadMode.RGB: the 1-channel lennon.jpg is replicated), resized
              to 256x256 with antialiased bilinear interpolation (what torchvision's
              Resize((256, 256), antialias=True) runs on tensors: F.interpolate(mode="bilinear",
              antialias=True)), rounded to uint8 so that the fixture's input is exact, cycled to B = 16;
              conf/patch14-l.json geometry (patch 14, max 32x32 tiles), beta = 0, max_seq_len 3072;
              preprocess -> iter_batches(batch_size=None) -> postprocess (FE:155-177, FE:180-287, FE:289-310).
              Keys: images u8 (13,3,256,256); positions / channels / image_ids / key_pad_mask of the batch;
              patches0 (first 256 tokens of image 0); rec_u8 (4,3,256,256) = the reference's first four
              reconstructions clamped and rounded to 8 bits (what its callers do, testpipe.py:74); rec_f32_11
              one full-precision reconstruction; psnr (16,) of each reconstruction against its input; lfq16_* = the
              conf/patch14-l.json quantiser (LFQ dim 196, 8192 x 16 codebooks, 196->208->196 projections,
              manual_seed(2)) applied to the first 128 PatchNorm-normalised tokens of image 0.
glue.npz      the reference's DCTAutoencoder (modeling_dct_autoencoder.py) with its two CLIPEncoder stacks replaced by
              the identity: encode() (to_patch_embedding + encoder position embeddings + LFQ 256 -> 16 x 13 bits ->
              256) and decode_from_codes() (indices_to_codes + decoder position embeddings + proj_out + inverse
              PatchNorm) on a packed 3-image batch; every parameter is stored so the test loads the same weights.
to_dict.json  the reference's to_dict (DP:54-87) on a packed 3-image batch with 3-codebook codes, and
to_dict.npz   the batch it was computed from + from_dict's (DP:90-122) outputs for object 1.
"""
import glob
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

ref_shim.import_reference()
from dct_autoencoder.dataset import dict_collate  # noqa: E402
from dct_autoencoder.dct_patches import from_dict, to_dict  # noqa: E402
from dct_autoencoder.feature_extraction_dct_autoencoder import DCTAutoencoderFeatureExtractor  # noqa: E402
from dct_autoencoder.lfq import LFQ  # noqa: E402
from dct_autoencoder.patchnorm import PatchNorm  # noqa: E402


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: npy(v) for k, v in kw.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, keys={list(kw)}")


# ---------------------------------------------------------------------------- config 1: real images
from PIL import Image  # noqa: E402

files = sorted(glob.glob(os.path.join(ref_shim.REFERENCE_ROOT, "images", "*.jpg")))
assert len(files) == 13, files
ims_u8 = []
for f in files:
    a = np.asarray(Image.open(f).convert("RGB"))                      # (H, W, 3) u8
    t = torch.from_numpy(a.copy()).permute(2, 0, 1)[None].float() / 255
    r = F.interpolate(t, size=(256, 256), mode="bilinear", antialias=True, align_corners=False)[0]
    ims_u8.append((r.clamp(0, 1) * 255).round().to(torch.uint8))
ims_u8 = torch.stack(ims_u8)
batch_ims = [ims_u8[i % 13].float() / 255 for i in range(16)]

fe = DCTAutoencoderFeatureExtractor(channels=3, patch_size=14, sample_patches_beta=0.0, max_patch_h=32,
                                    max_patch_w=32, max_seq_len=3072)
items = [fe.preprocess(im) for im in batch_ims]
b = next(fe.iter_batches(iter([dict_collate(items)]), None))
rec = fe.postprocess(b)
psnr = [float(-10 * torch.log10(((r - im) ** 2).mean())) for r, im in zip(rec, batch_ims)]
print("config 1 round-trip PSNR: mean %.2f dB, min %.2f, max %.2f" % (np.mean(psnr), min(psnr), max(psnr)))

# the model's quantiser (conf/patch14-l.json: vq_type lfq, 8192 x 16) on normalised tokens of image 0
pn = PatchNorm(32, 32, 14, 3)
pn.train()
pn(b)
pn.frozen = True
normed = pn(b)
torch.manual_seed(2)
lfq16 = LFQ(dim=196, codebook_size=8192, num_codebooks=16).eval()
k0 = 128          # the first 128 tokens of image 0 keep the fixture small
with torch.no_grad():
    q16, idx16, _, _ = lfq16(normed[:1, :k0], torch.ones(1, k0, dtype=torch.bool))
    pre16 = lfq16.project_in(normed[:1, :k0])

save("config1",
     images=ims_u8,
     positions=b.patch_positions.to(torch.int16), channels=b.patch_channels.to(torch.int8),
     image_ids=b.batched_image_ids.to(torch.int8), key_pad_mask=b.key_pad_mask,
     patch_sizes=np.asarray(b.patch_sizes, dtype=np.int64), original_sizes=np.asarray(b.original_sizes, dtype=np.int64),
     patches0=items[0]["patches"][:256],
     rec_u8=torch.stack([(r.clamp(0, 1) * 255).round().to(torch.uint8) for r in rec[:4]]),
     rec_f32_11=rec[11], psnr=np.asarray(psnr),
     pn_median0=pn.median.data[:, :18, :18], pn_b0=pn.b.data[:, :18, :18],
     lfq16_in=normed[0, :k0], lfq16_pre=pre16[0], lfq16_q=q16[0], lfq16_idx=idx16[0].to(torch.int16),
     lfq16_w_in=lfq16.project_in.weight, lfq16_b_in=lfq16.project_in.bias,
     lfq16_w_out=lfq16.project_out.weight, lfq16_b_out=lfq16.project_out.bias)

# ---------------------------------------------------------------------------- to_dict / from_dict
torch.manual_seed(21)
fx = DCTAutoencoderFeatureExtractor(channels=3, patch_size=4, sample_patches_beta=0.0, max_patch_h=4, max_patch_w=4,
                                    max_seq_len=60)
ims = [torch.rand(3, 12, 12), torch.rand(3, 9, 17), torch.rand(3, 16, 8)]
its = [fx.preprocess(im) for im in ims]
bt = next(fx.iter_batches(iter([dict_collate(its)]), None))
codes = torch.randint(0, 2 ** 10, (bt.patches.shape[0], bt.patches.shape[1], 3))
objs = to_dict(bt, codes)
with open(os.path.join(HERE, "to_dict.json"), "w") as f:
    json.dump(objs, f)
dp1, codes1 = from_dict(objs[1])
save("to_dict", patches_shape=np.asarray(bt.patches.shape), key_pad_mask=bt.key_pad_mask, image_ids=bt.batched_image_ids,
     channels=bt.patch_channels, positions=bt.patch_positions, codes=codes,
     patch_sizes=np.asarray(bt.patch_sizes, dtype=np.int64), original_sizes=np.asarray(bt.original_sizes, dtype=np.int64),
     fd_channels=dp1.patch_channels, fd_positions=dp1.patch_positions, fd_key_pad_mask=dp1.key_pad_mask,
     fd_image_ids=dp1.batched_image_ids, fd_attn_mask=dp1.attn_mask, fd_codes=codes1,
     fd_patches=dp1.patches)

# ---------------------------------------------------------------------------- model glue (identity transformer stacks)
import types  # noqa: E402

from dct_autoencoder.configuration_dct_autoencoder import DCTAutoencoderConfig  # noqa: E402
from dct_autoencoder.modeling_dct_autoencoder import DCTAutoencoder  # noqa: E402

torch.manual_seed(31)
tiny = dict(hidden_size=256, intermediate_size=64, num_attention_heads=2, num_hidden_layers=1)
cfg = DCTAutoencoderConfig(image_channels=3, patch_size=14, max_patch_h=6, max_patch_w=6, vq_codebook_size=8192,
                           vq_num_codebooks=16, vq_type="lfq", encoder_config=tiny, decoder_config=tiny)
model = DCTAutoencoder(cfg).eval()


class _IdentityStack(torch.nn.Module):
    def forward(self, hidden, attention_mask=None):
        return types.SimpleNamespace(last_hidden_state=hidden)


model.encoder, model.decoder = _IdentityStack(), _IdentityStack()
with torch.no_grad():
    model.to_patch_embedding[1].weight.uniform_(0.5, 1.5)
    model.to_patch_embedding[1].bias.normal_(0, 0.1)
    model.proj_out[0].weight.uniform_(0.5, 1.5)
    model.proj_out[0].bias.normal_(0, 0.1)
    model.patchnorm.median.normal_(0, 0.2)
    model.patchnorm.b.uniform_(0.2, 1.0)
model.patchnorm.frozen = True
fx = DCTAutoencoderFeatureExtractor(channels=3, patch_size=14, sample_patches_beta=0.0, max_patch_h=6, max_patch_w=6,
                                    max_seq_len=120)
ims = [torch.rand(3, 70, 56), torch.rand(3, 42, 84), torch.rand(3, 84, 84)]
bt = next(fx.iter_batches(iter([dict_collate([fx.preprocess(im) for im in ims])]), None))
gl = dict(in_patches=bt.patches.clone(), key_pad_mask=bt.key_pad_mask, image_ids=bt.batched_image_ids,
          channels=bt.patch_channels, positions=bt.patch_positions,
          patch_sizes=np.asarray(bt.patch_sizes, dtype=np.int64), original_sizes=np.asarray(bt.original_sizes, dtype=np.int64))
with torch.no_grad():
    emb = model.to_patch_embedding(model.patchnorm(bt))
    enc, codes, commit, dist = model.encode(bt.shallow_copy(), do_normalize=True)
    gl.update(embedded=emb, enc_patches=enc.patches, codes=codes.to(torch.int16))
    dec = model.decode_from_codes(codes, do_inv_norm=True, key_pad_mask=bt.key_pad_mask, attn_mask=bt.attn_mask,
                                  batched_image_ids=bt.batched_image_ids, patch_channels=bt.patch_channels,
                                  patch_positions=bt.patch_positions, patch_sizes=bt.patch_sizes,
                                  original_sizes=bt.original_sizes)
    gl.update(dec_patches=dec.patches)
    pre = model.vq_model.project_in(model.to_patch_embedding(model.patchnorm(bt)) +
                                    model.encoder_pos_embed_height[bt.h_indices] + model.encoder_pos_embed_width[bt.w_indices] +
                                    model.encoder_pos_embed_channel[bt.patch_channels])
    gl.update(lfq_pre=pre)
for k, v in model.state_dict().items():
    gl["w:" + k] = v
save("glue", **gl)
print("done")
