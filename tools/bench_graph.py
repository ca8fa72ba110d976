"""Eager launches vs CUDA-graph replay of the fused round trip at several batch sizes.

    python tools/bench_graph.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dct_autoencoder_b200 as D
dev = torch.device('cuda', 0)
fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
pn = D.PatchNorm(32, 32, 14, 3).to(dev)
lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).to(dev).eval()
pipe = D.TransformPipeline(fe, pn, lfq)
g = torch.Generator(device=dev).manual_seed(0)
for B in (256, 64, 16):
    x = torch.rand(B, 3, 512, 512, device=dev, generator=g)
    pipe.fit_norm(x[:16])
    gr = pipe.graphed(x)
    def t(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    print(f"B={B}: eager {t(lambda: pipe.roundtrip(x)):.3f} ms, graph {t(gr):.3f} ms, launches {gr.launches}")
