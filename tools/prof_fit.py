"""Per-launch device times of one warm PatchNorm statistic-fit step (pipe.fit_norm: encode + the stat update kernels).

    python tools/prof_fit.py [--images 64] [--size 512]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dct_autoencoder_b200 as D  # noqa: E402
from dct_autoencoder_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--size", type=int, default=512)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    pn = D.PatchNorm(32, 32, 14, 3).cuda()
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).cuda().eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    x = torch.rand(a.images, 3, a.size, a.size, device=dev)
    for _ in range(3):
        pipe.fit_norm(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pipe.fit_norm(x)
    e1.record()
    torch.cuda.synchronize()
    print("fit step ms", e0.elapsed_time(e1) / 10, "images", a.images)
    with _lib.profile(dev) as p:
        pipe.fit_norm(x)
    for name, ms in p.groups:
        print("%9.1f us  %s" % (ms * 1e3, name))


if __name__ == "__main__":
    main()
