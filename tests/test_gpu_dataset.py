"""The loader front end (dataset.py:35-89 of the reference): antialiased resize against PyTorch's own
``F.interpolate(antialias=True)`` (what torchvision's Resize runs on tensors), the down-scaling rule against
torchvision's Resize used as the reference uses it, GPU JPEG decode against PIL (libjpeg, the reference's decoder),
and the generator over webdataset-style samples / tar shards."""
import io
import json
import os
import tarfile

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import dct_autoencoder_b200 as d
    d._lib.load()
    return d


@pytest.mark.parametrize("ih,iw,oh,ow", [(900, 1300, 531, 767), (1300, 900, 768, 531), (513, 771, 300, 451),
                                         (256, 256, 300, 340), (97, 131, 31, 40), (64, 64, 64, 64)])
def test_resize_matches_torch_antialiased_bilinear(D, ih, iw, oh, ow):
    from dct_autoencoder_b200.dataset import resize_antialias
    torch.manual_seed(ih + ow)
    x = torch.rand(2, 3, ih, iw, device="cuda")
    ref = F.interpolate(x, size=(oh, ow), mode="bilinear", antialias=True, align_corners=False)
    got = resize_antialias(x, (oh, ow))
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 2e-6
    u = (x * 255).round().to(torch.uint8)
    ref8 = F.interpolate(u.float() / 255, size=(oh, ow), mode="bilinear", antialias=True, align_corners=False)
    assert float((resize_antialias(u, (oh, ow)) - ref8).abs().max()) < 2e-6


@pytest.mark.parametrize("h,w", [(900, 1300), (1300, 900), (1000, 1000), (700, 500), (768, 768), (769, 40)])
def test_crop_follows_the_reference_rule(D, h, w):
    """dataset.py:59-73 with torchvision's Resize as the reference calls it."""
    from torchvision import transforms
    from dct_autoencoder_b200.dataset import crop, max_image_size
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    max_size = max_image_size(fe)
    assert max_size == 768
    torch.manual_seed(h)
    x = torch.rand(3, h, w, device="cuda")

    def ref_crop(pixel_values):
        _, hh, ww = pixel_values.shape
        if max(hh, ww) > max_size:
            ar = hh / ww
            if hh > ww:
                hh = max_size
                ww = int(hh / ar)
            else:
                ww = max_size
                hh = int(ar * ww)
            pixel_values = transforms.Resize(min(hh, ww), antialias=True)(pixel_values)
        return pixel_values
    ref = ref_crop(x)
    got = crop(x, max_size)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 2e-6
    u = (x * 255).round().to(torch.uint8)
    got8 = crop(u, max_size)
    assert got8.dtype == torch.float32 and float((got8 - ref_crop(u.float() / 255)).abs().max()) < 2e-6


def _jpeg_bytes(u8_chw: np.ndarray, quality=92, subsampling=None) -> bytes:
    from PIL import Image
    buf = io.BytesIO()
    kw = {} if subsampling is None else {"subsampling": subsampling}
    Image.fromarray(np.ascontiguousarray(u8_chw.transpose(1, 2, 0))).save(buf, format="JPEG", quality=quality, **kw)
    return buf.getvalue()


def _pictures(golden, sizes):
    g = golden("config1")
    ims = torch.from_numpy(g["images"]).float()                   # (13, 3, 256, 256) natural images
    out = []
    for i, (h, w) in enumerate(sizes):
        t = F.interpolate(ims[i % 13][None], size=(h, w), mode="bicubic", align_corners=False)[0]
        out.append(t.clamp(0, 255).round().to(torch.uint8).numpy())
    return out


def test_gpu_jpeg_decode_agrees_with_libjpeg(D, golden):
    from PIL import Image
    from dct_autoencoder_b200.dataset import decode_jpegs
    pics = _pictures(golden, [(380, 500), (512, 512), (301, 451)])
    streams = [_jpeg_bytes(p) for p in pics]                                # 4:2:0 chroma (PIL's default)
    streams.append(_jpeg_bytes(pics[0], subsampling=0))                     # 4:4:4: no chroma up-sampling involved
    grey = io.BytesIO()
    Image.fromarray(pics[0][0]).save(grey, format="JPEG", quality=90)      # 1-channel stream: replicated to RGB
    streams.append(grey.getvalue())
    # the default decoder is the reference's own (libjpeg through PIL, on host threads): identical pixels
    for s_, d_ in zip(streams, decode_jpegs(streams, "cuda")):
        ref = np.asarray(Image.open(io.BytesIO(s_)).convert("RGB")).transpose(2, 0, 1)
        assert d_.is_cuda and d_.dtype == torch.uint8 and np.array_equal(d_.cpu().numpy(), ref)
    dec = decode_jpegs(streams, "cuda", decoder="nvjpeg")
    assert len(dec) == 5
    stats = []
    for i, (s, d) in enumerate(zip(streams, dec)):
        ref = np.asarray(Image.open(io.BytesIO(s)).convert("RGB")).transpose(2, 0, 1).astype(np.int32)
        got = d.cpu().numpy().astype(np.int32)
        assert got.shape == ref.shape and d.dtype == torch.uint8 and d.is_cuda
        diff = np.abs(got - ref)
        psnr = 10 * np.log10(255.0 ** 2 / max(np.mean(diff.astype(np.float64) ** 2), 1e-12))
        print(f"stream {i}: mean |diff| {diff.mean():.3f} LSB, max {diff.max()}, within 1 LSB {(diff <= 1).mean():.4f}, PSNR {psnr:.1f} dB")
        stats.append((i, diff.mean(), (diff <= 2).mean(), (diff <= 3).mean(), psnr))
    for i, mean, w2, w3, psnr in stats:
        if i < 3:
            # two conforming decoders: different rounding in the IDCT, and a different interpolation of the sub-sampled
            # chroma planes (libjpeg's "fancy" triangle filter vs nvJPEG's), which shows at sharp colour edges
            assert w3 > 0.90 and psnr > 38.0, stats
        else:
            # without chroma sub-sampling only the IDCT / colour-conversion rounding differs
            assert w2 > 0.98 and psnr > 44.0, stats


def test_load_and_transform_dataset(D, golden, tmp_path):
    from dct_autoencoder_b200.dataset import crop, decode_jpegs, load_and_transform_dataset, max_image_size
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    sizes = [(380, 500), (900, 1300), (160, 400), (512, 512), (1024, 640), (100, 100)]
    pics = _pictures(golden, sizes)
    samples = [{"__key__": f"{i:04d}", "jpg": _jpeg_bytes(p), "json": json.dumps({"height": s[0], "width": s[1]}).encode()}
               for i, (p, s) in enumerate(zip(pics, sizes))]
    samples.append({"__key__": "nometa", "jpg": samples[0]["jpg"], "json": json.dumps({"height": None, "width": 5}).encode()})
    # min_res = 14 * 12 = 168: (160, 400) and (100, 100) are filtered out, as is the sample without a size
    keep = [0, 1, 3, 4]
    got = list(load_and_transform_dataset(samples, fe, device="cuda", decode_batch=3))
    assert len(got) == len(keep)
    for item, i in zip(got, keep):
        im = crop(decode_jpegs([samples[i]["jpg"]], "cuda")[0], max_image_size(fe))
        want = fe.preprocess(im)
        assert item["original_sizes"] == want["original_sizes"] and item["patch_sizes"] == want["patch_sizes"]
        assert max(item["original_sizes"]) <= 768
        assert torch.equal(item["positions"], want["positions"]) and torch.equal(item["channels"], want["channels"])
        assert torch.equal(item["patches"], want["patches"])
    # the same through a webdataset-style tar shard
    path = os.path.join(tmp_path, "shard-0000.tar")
    with tarfile.open(path, "w") as tf:
        for s in samples[:-1]:
            for ext in ("jpg", "json"):
                ti = tarfile.TarInfo(f"{s['__key__']}.{ext}")
                ti.size = len(s[ext])
                tf.addfile(ti, io.BytesIO(s[ext]))
    got_tar = list(load_and_transform_dataset(os.path.join(tmp_path, "shard-{0000..0000}.tar"), fe, device="cuda"))
    assert len(got_tar) == len(keep)
    for a, b in zip(got_tar, got):
        assert torch.equal(a["patches"], b["patches"]) and a["original_sizes"] == b["original_sizes"]
    # and into the batcher, as main.py does with the loader's output
    from dct_autoencoder_b200.dataset import dict_collate
    batch = next(fe.iter_batches(iter([dict_collate(got)]), None))
    assert int((~batch.key_pad_mask).sum()) == sum(int(it["patches"].shape[0]) for it in got)
