"""Chunk-size sweep of the compact host round trip (uint8 in, uint8 + wire records out).
    python tools/bench_compact.py [--batch 256]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dct_autoencoder_b200 as D

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=256); a = ap.parse_args()
dev = torch.device("cuda", 0)
fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
pn = D.PatchNorm(32, 32, 14, 3).to(dev)
lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).to(dev).eval()
pipe = D.TransformPipeline(fe, pn, lfq)
pipe.fit_norm(torch.rand(32, 3, 512, 512, device=dev))
B = a.batch
D.util.bind_to_gpu_numa(dev)
hx8 = torch.randint(0, 256, (B, 3, 512, 512), dtype=torch.uint8).pin_memory()
o8 = torch.empty((B, 3, 512, 512), dtype=torch.uint8).pin_memory()
ow = torch.empty((B, 3072, 27), dtype=torch.uint8).pin_memory()
oc = torch.empty(B, dtype=torch.int32).pin_memory()
hx = (hx8.float() / 255).pin_memory()
of = torch.empty((B, 3, 512, 512), dtype=torch.float32).pin_memory()
ocod = torch.empty((B, 3072, 14), dtype=torch.int64).pin_memory()
for chunk in (16, 32, 64, 128, 256):
    for name, fn in (("compact", lambda: pipe.roundtrip_host(hx8, o8, ow, chunk=chunk, compact=True, out_counts=oc)),
                     ("fp32", lambda: pipe.roundtrip_host(hx, of, ocod, chunk=chunk))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"chunk {chunk:4d} {name:8s}: {dt * 1e3:7.2f} ms per {B} images -> {B / dt / 1e3:6.1f} k img/s")
