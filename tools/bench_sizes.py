"""Fused round trip (encode -> PatchNorm -> LFQ -> decode) at several image sizes: img/s and ns per pixel.

    python tools/bench_sizes.py
"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import dct_autoencoder_b200 as D
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    base = None
    for (h, w, b) in [(512, 512, 256), (300, 451, 480), (300, 452, 480), (304, 464, 480), (255, 255, 512), (256, 256, 512),
                      (1024, 1024, 64), (720, 1280, 64), (721, 1281, 64)]:
        fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
        pn = D.PatchNorm(32, 32, 14, 3).to(dev)
        lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).to(dev).eval()
        pipe = D.TransformPipeline(fe, pn, lfq)
        pipe.fit_norm(torch.rand(8, 3, h, w, device=dev, generator=g))
        x = torch.rand(b, 3, h, w, device=dev, generator=g)
        for _ in range(3):
            pipe.roundtrip(x)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pipe.roundtrip(x)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        ns_px = ms * 1e6 / (b * h * w)
        base = base or ns_px
        print(f"{h:5d} x {w:5d}  batch {b:4d}: {ms:7.3f} ms  {b / ms * 1e3:10.0f} img/s  {ns_px * 1e3:7.2f} ps/pixel  "
              f"({ns_px / base:4.2f}x the 512^2 per-pixel time)  fusable={pipe.fusable()}", flush=True)
        del x, pipe


if __name__ == "__main__":
    main()
