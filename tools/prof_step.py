"""One warm pipeline step for profiling (ncu launch lists / --set full captures).

    python tools/prof_step.py [--batch 64] [--size 512] [--steps 1] [--warmup 2] [--dct-impl tc]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--dct-impl", default="tc")
    ap.add_argument("--staged", action="store_true", help="the drop-in modules one by one instead of the fused step")
    ap.add_argument("--width", type=int, default=0, help="image width (default: --size)")
    ap.add_argument("--max-seq-len", type=int, default=3072)
    ap.add_argument("--launches", action="store_true", help="print the library's per-launch CUDA-event times of one step")
    a = ap.parse_args()
    import torch
    import dct_autoencoder_b200 as D
    dev = torch.device("cuda", 0)
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, a.max_seq_len, dct_impl=a.dct_impl)
    pn = D.PatchNorm(32, 32, 14, 3).to(dev)
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).to(dev).eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    pipe.fit_norm(torch.rand(min(a.batch, 16), 3, a.size, a.width or a.size, device=dev, generator=g))
    x = torch.rand(a.batch, 3, a.size, a.width or a.size, device=dev, generator=g)
    step = (lambda: pipe.roundtrip_staged(x)) if a.staged else (lambda: pipe.roundtrip(x))
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    l0 = D._lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"batch {a.batch} size {a.size}: {e0.elapsed_time(e1) / a.steps:.3f} ms/step, "
          f"{(D._lib.launch_count - l0) // a.steps} launches/step")
    if a.launches:
        with D._lib.profile(dev) as prof:
            step()
        for name, ms in prof.groups:
            print(f"  {ms * 1e3:9.1f} us  {name}")


if __name__ == "__main__":
    main()
