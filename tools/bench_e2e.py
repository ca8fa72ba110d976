"""Host-to-host round trip (bench.py's `e2e`) against the PCIe link it is bound by.

    python tools/bench_e2e.py [--batch 256] [--chunks 8,16,32,64]

Prints the pinned-memory copy bandwidth of this box (H2D alone, D2H alone, both at once) and the
images/s of TransformPipeline.roundtrip_host for several chunk sizes.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--chunks", default="8,16,32,64")
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--bind", action="store_true", help="bind to the CPUs of the GPU's NUMA node before allocating")
    a = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, S = a.batch, 512
    import os
    from dct_autoencoder_b200 import util
    cpus = util.gpu_numa_cpus(dev)
    print(f"affinity {len(os.sched_getaffinity(0))} cpus; GPU NUMA cpus: {None if cpus is None else len(cpus)}; "
          f"nodes: {[d for d in os.listdir('/sys/devices/system/node') if d.startswith('node')]}")
    if a.bind:
        print("bound:", util.bind_to_gpu_numa(dev))

    # ---- the link
    n = 512 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def wall(fn, reps=4):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d()
        d2h()

    print(f"H2D alone      {n / wall(h2d) / 1e9:7.1f} GB/s")
    print(f"D2H alone      {n / wall(d2h) / 1e9:7.1f} GB/s")
    print(f"both at once   {n / wall(both) / 1e9:7.1f} GB/s per direction")
    del h_in, h_out, d_in, d_out

    # ---- the round trip
    import dct_autoencoder_b200 as D
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    pn = D.PatchNorm(32, 32, 14, 3).to(dev)
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).to(dev).eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    g = torch.Generator(device=dev)
    g.manual_seed(1000)
    pipe.fit_norm(torch.rand(64, 3, S, S, device=dev, generator=g))
    g.manual_seed(0)
    x = torch.rand(B, 3, S, S, device=dev, generator=g)
    hx = torch.empty((B, 3, S, S), dtype=torch.float32).pin_memory()
    hx.copy_(x)
    h_rec = torch.empty((B, 3, S, S), dtype=torch.float32).pin_memory()
    h_codes = torch.empty((B, 3072, 14), dtype=torch.int64).pin_memory()
    bytes_in = hx.numel() * 4
    bytes_out = h_rec.numel() * 4 + h_codes.numel() * 8
    for c in [int(v) for v in a.chunks.split(",")]:
        dt = wall(lambda: pipe.roundtrip_host(hx, h_rec, h_codes, chunk=c), a.reps)
        print(f"chunk {c:4d}: {dt * 1e3:7.2f} ms / {B} images = {B / dt:9.0f} img/s   "
              f"({bytes_in / dt / 1e9:5.1f} GB/s in, {bytes_out / dt / 1e9:5.1f} GB/s out)")


if __name__ == "__main__":
    main()
