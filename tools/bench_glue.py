"""Model glue on the device (modeling_dct_autoencoder.py:41-64, 85-112, 129-178) at config-2 token counts:
patch embedding (Linear 196 -> 1024 + LayerNorm + three position-embedding gathers), the quantiser of conf/patch14-l.json
on 1024 features, decoder position embeddings + LayerNorm + proj_out (1024 -> 196).  Per-launch device times from the
library's own events, next to the same modules run by eager PyTorch (cuBLAS fp32 + separate LayerNorm / gather passes).

    python tools/bench_glue.py [--images 64]
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dct_autoencoder_b200 as D  # noqa: E402
from dct_autoencoder_b200 import _lib  # noqa: E402
from dct_autoencoder_b200.modeling_dct_autoencoder import DCTAutoencoderGlue  # noqa: E402


def timed(fn, reps=10, warm=3):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    glue = DCTAutoencoderGlue().to(dev).eval()
    x = torch.rand(a.images, 3, 512, 512, device=dev)
    batch = fe.process_batch(x)
    tokens = batch.patches.shape[0] * batch.patches.shape[1]
    print(f"{a.images} images of 512^2 -> {tokens} tokens of 196, feature_dim 1024, LFQ 16 x 13 bit (208)")
    src = batch.patches.clone()

    def fresh():
        b = batch.shallow_copy()
        b.patches = src
        return b

    def eager_embed(b):
        lin, ln = glue.to_patch_embedding[0], glue.to_patch_embedding[1]
        y = F.layer_norm(F.linear(b.patches, lin.weight), (1024,), ln.weight, ln.bias, ln.eps)
        ch, pos = b.patch_channels, b.patch_positions
        return y + glue.encoder_pos_embed_channel[ch] + glue.encoder_pos_embed_height[pos[..., 0]] \
            + glue.encoder_pos_embed_width[pos[..., 1]]

    def eager_decode(h, b):
        ch, pos = b.patch_channels, b.patch_positions
        h = h + glue.decoder_pos_embed_channel[ch] + glue.decoder_pos_embed_height[pos[..., 0]] \
            + glue.decoder_pos_embed_width[pos[..., 1]]
        ln, lin = glue.proj_out[0], glue.proj_out[1]
        return F.linear(F.layer_norm(h, (1024,), ln.weight, ln.bias, ln.eps), lin.weight)

    with torch.no_grad():
        emb = glue.embed(fresh()).patches
        ref = eager_embed(fresh())
        print("embed: max |ours - eager| = %.3g (|eager| max %.3g)" % (float((emb - ref).abs().max()), float(ref.abs().max())))
        hid = emb.clone()

        def dec():
            b = fresh()
            b.patches = hid
            return glue.decode(b).patches
        out, oref = dec(), eager_decode(hid, batch)
        print("decode: max |ours - eager| = %.3g (|eager| max %.3g)" % (float((out - oref).abs().max()), float(oref.abs().max())))
        rows = [("embed (Linear + LayerNorm + 3 position gathers)", lambda: glue.embed(fresh()), lambda: eager_embed(fresh()),
                 tokens * (196 + 1024) * 4),
                ("decode (3 position gathers + LayerNorm + proj_out)", dec, lambda: eager_decode(hid, batch),
                 tokens * (196 + 1024) * 4),
                ("quantiser (project_in + sign + indices + project_out)",
                 lambda: glue.vq_model(hid, mask=~batch.key_pad_mask), None, tokens * 2 * 1024 * 4),
                ("forward (encode + decode, identity stacks)", lambda: glue(fresh()), None, None)]
        for name, ours, eager, alg in rows:
            ms = timed(ours)
            line = "%-58s %8.3f ms  %7.1f M tokens/s" % (name, ms, tokens / ms / 1e3)
            if alg:
                line += "  %6.0f GB/s of its algorithmic bytes" % (alg / ms / 1e6)
            if eager is not None:
                line += "   eager PyTorch %8.3f ms" % timed(eager)
            print(line)
            with _lib.profile(dev) as p:
                ours()
            for g, m in p.groups:
                print("      %9.1f us  %s" % (m * 1e3, g))


if __name__ == "__main__":
    main()
