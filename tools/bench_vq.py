"""Config 4 of BASELINE.json: VectorQuantize nearest-codebook search, codebook 8192 x 256, T = 512 * 3072 tokens.

    python tools/bench_vq.py [--tokens 1572864] [--codes 8192] [--dim 256]
"""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=512 * 3072)
    ap.add_argument("--codes", type=int, default=8192)
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=3072 * 64)
    a = ap.parse_args()
    import torch
    from dct_autoencoder_b200.vector_quantize import nearest_code
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    e = torch.randn(a.codes, a.dim, device=dev, generator=g)
    x = torch.randn(a.chunk, a.dim, device=dev, generator=g)
    n_chunks = (a.tokens + a.chunk - 1) // a.chunk
    for impl in ("tc", "fp32"):
        if impl == "fp32" and a.tokens > 200000:
            n = 1
        else:
            n = n_chunks
        nearest_code(x, e, impl=impl)
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                nearest_code(x, e, impl=impl)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        tok = n * a.chunk
        flops = 2.0 * tok * a.codes * a.dim
        print(f"{impl}: {tok} tokens in {ms:.2f} ms -> {tok / ms * 1e3 / 1e6:.2f} M tokens/s, "
              f"{flops / ms / 1e9:.1f} algorithmic TFLOP/s")


if __name__ == "__main__":
    main()
