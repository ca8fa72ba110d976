"""Timings of the two callers next to the path: the code wire format and the shard writer.

    python tools/bench_next.py [--batch 256]
"""
import argparse
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    import torch
    import dct_autoencoder_b200 as D
    from dct_autoencoder_b200 import _lib
    dev = torch.device("cuda", 0)
    B, S = a.batch, 512
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    pn = D.PatchNorm(32, 32, 14, 3).to(dev)
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).to(dev).eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(B, 3, S, S, device=dev, generator=g)
    pipe.fit_norm(x[:64])
    batch, codes = pipe.encode_codes(x)

    # ---- wire format: kernel alone, then the whole host call, against to_dict
    b, s, c = codes.shape
    rec = D.dct_patches.wire_record_bytes(c, 14)
    out = torch.empty((b, s, rec), dtype=torch.uint8, device=dev)
    counts = torch.empty((b, s), dtype=torch.int32, device=dev)

    def pack():
        _lib.call("dcta_wire_pack", _lib.ptr(codes), _lib.ptr(batch.patch_positions), _lib.ptr(batch.patch_channels),
                  _lib.ptr(batch.batched_image_ids), batch.key_pad_mask.data_ptr(), b, s, c, 14, rec, _lib.ptr(out),
                  _lib.ptr(counts), _lib.stream_ptr(dev))
    for _ in range(3):
        pack()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pack()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    traffic = b * s * (8 * c + 16 + 8 + 8 + 1 + rec + 4)
    print(f"wire_pack_kernel: {us:.1f} us per {b} images ({b * s} tokens), {traffic / us / 1e3:.0f} GB/s")
    t0 = time.perf_counter()
    blobs = D.to_bytes(batch, codes, 2 ** 14)
    t_bytes = time.perf_counter() - t0
    n_dict = min(B, 8)
    sub = batch.shallow_copy()
    t0 = time.perf_counter()
    for f in ("key_pad_mask", "batched_image_ids", "patch_channels", "patch_positions"):
        setattr(sub, f, getattr(batch, f)[:n_dict])
    sub.patches = torch.zeros(n_dict, s, 1, device=dev)
    sub._row_num_images = [1] * n_dict
    objs = D.to_dict(sub, codes[:n_dict])
    t_dict = time.perf_counter() - t0
    print(f"to_bytes: {t_bytes * 1e3 / B:.3f} ms/image, {sum(map(len, blobs)) / B / 1e3:.1f} KB/image;  "
          f"to_dict (bulk copy + Python lists): {t_dict * 1e3 / n_dict:.2f} ms/image")
    del objs, blobs

    # ---- shard writer
    for dtype, compress, writers in ((None, False, 1), (torch.float16, False, 1), (torch.float16, False, 4),
                                     (torch.float16, True, 1), (torch.float16, True, 12)):
        d = tempfile.mkdtemp()
        try:
            n_b = 4 if not compress or writers > 1 else 1
            t0 = time.perf_counter()
            info = D.shards.preprocess_to_shards((x[:64] for _ in range(n_b)), fe, d, dtype=dtype, compress=compress, writers=writers)
            dt = time.perf_counter() - t0
            print(f"preprocess_to_shards dtype={dtype} compress={compress} writers={writers}: {info['samples'] / dt:.0f} images/s, "
                  f"{info['bytes'] / info['samples'] / 1e6:.2f} MB/image serialised")
        finally:
            shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
