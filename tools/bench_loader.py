"""Front end of the offline shard writer: JPEG streams -> nvJPEG batch decode on the GPU -> the reference's
down-scaling rule (antialiased resize kernel) -> preprocess.  Prints images/s of each stage and of the whole chain.

    python tools/bench_loader.py [--n 512] [--height 1000] [--width 1500]
"""
import argparse
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--height", type=int, default=1000)
    ap.add_argument("--width", type=int, default=1500)
    ap.add_argument("--decode-batch", type=int, default=64)
    a = ap.parse_args()
    import numpy as np
    import torch
    import torch.nn.functional as F
    from PIL import Image
    import dct_autoencoder_b200 as D
    from dct_autoencoder_b200 import dataset as DS
    from dct_autoencoder_b200 import shards
    dev = torch.device("cuda", 0)
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "config1.npz"))
    pics = torch.from_numpy(g["images"]).float()
    streams = []
    for i in range(13):
        t = F.interpolate(pics[i][None], size=(a.height, a.width), mode="bicubic", align_corners=False)[0]
        buf = io.BytesIO()
        Image.fromarray(t.clamp(0, 255).round().to(torch.uint8).permute(1, 2, 0).numpy()).save(buf, format="JPEG", quality=90)
        streams.append(buf.getvalue())
    meta = json.dumps({"height": a.height, "width": a.width}).encode()
    samples = [{"__key__": f"{i:06d}", "jpg": streams[i % 13], "json": meta} for i in range(a.n)]
    print(f"{a.n} JPEG streams of {a.height} x {a.width}, {sum(len(s['jpg']) for s in samples) / a.n / 1024:.0f} KiB each")
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    max_size = DS.max_image_size(fe)

    def timed(fn, label, n):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"  {label:58s} {n / dt:9.0f} img/s  ({dt * 1e3 / n:.3f} ms/img)")
        return out

    sub = [s["jpg"] for s in samples[:a.decode_batch]]
    timed(lambda: DS.decode_jpegs(sub, dev, decoder="nvjpeg"), "nvJPEG batch decode (torchvision.io.decode_jpeg, cuda)", len(sub))
    timed(lambda: [np.asarray(Image.open(io.BytesIO(s)).convert("RGB")) for s in sub[:16]], "PIL decode on one host core (the reference's decoder)", 16)
    dec = timed(lambda: DS.decode_jpegs(sub, dev), f"libjpeg decode on {min(32, os.cpu_count())} host threads + pinned upload (default)", len(sub))
    crops = timed(lambda: [DS.crop(im, max_size) for im in dec], f"crop: antialiased resize to <= {max_size} (dcta_resize_bilinear_aa_u8)", len(dec))
    print("    resized to", tuple(crops[0].shape))
    timed(lambda: [fe.preprocess(im) for im in crops], "preprocess, image by image (the reference's API)", len(crops))
    batch = torch.stack(crops)
    timed(lambda: fe.preprocess_batch(batch), "preprocess_batch (one size per batch)", len(crops))
    timed(lambda: sum(1 for _ in DS.load_and_transform_dataset(samples, fe, device=dev, decode_batch=a.decode_batch)),
          "load_and_transform_dataset: decode + crop + preprocess", a.n)
    timed(lambda: sum(b.shape[0] for b in DS.image_batches(samples, fe, device=dev, decode_batch=a.decode_batch)),
          "image_batches: decode + crop, grouped by size", a.n)
    import tempfile
    for writers, dtype in ((4, torch.float16),):
        with tempfile.TemporaryDirectory() as tmp:
            t0 = time.perf_counter()
            st = shards.preprocess_to_shards(DS.image_batches(samples, fe, device=dev, decode_batch=a.decode_batch), fe, tmp,
                                             dtype=dtype, compress=False, writers=writers)
            dt = time.perf_counter() - t0
            print(f"  JPEG -> shards (fp16, uncompressed tar, {writers} writer threads): {st['samples'] / dt:9.0f} img/s, "
                  f"{st['bytes'] / 1e6 / dt:.0f} MB/s written")


if __name__ == "__main__":
    main()
