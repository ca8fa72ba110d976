"""The data-loader front end of the encode path: JPEG bytes -> RGB pixels on the GPU ->
the reference's resolution filter and down-scaling rule -> ``preprocess``.  Mirrors dataset.py:7-89 of the reference:

    dict_collate / tuple_collate      dataset.py:7-25
    load_preprocessed_dataset(url)    dataset.py:27-33   (shards.load_preprocessed_dataset)
    load_and_transform_dataset(url, dct_processor, device)   dataset.py:35-89

The reference decodes with webdataset's ``decode("torchrgb")`` (PIL / libjpeg on the host, one image at a time) and
resizes with ``torchvision.transforms.Resize(min(h, w), antialias=True)`` on the host.  Here JPEG streams are decoded in
batches on the GPU (``torchvision.io.decode_jpeg(device="cuda")``: nvJPEG, a library call exactly as the reference's
decode is), the 8-bit pixels stay bytes until the resize / colour kernels read them as ``u / 255``, and the resize is
libdcta's antialiased bilinear kernel (csrc/resize.cu).  ``webdataset`` itself is only a convention on tar files
(members of a sample share the basename up to the first dot), read here with ``tarfile``.
"""
import glob
import io
import json
import os
import re
import tarfile
from typing import Dict, Iterable, Iterator, List, Optional, Tuple

import torch

from . import _lib
from .shards import load_preprocessed_dataset  # noqa: F401  (dataset.py:27-33)


def dict_collate(x: List[Dict]):
    """dataset.py:7-14."""
    assert len(x) > 0
    out = {k: [] for k in x[0].keys()}
    for row in x:
        for k in out:
            out[k].append(row[k])
    return out


def tuple_collate(x: List[Tuple]):
    """dataset.py:17-25."""
    assert len(x) > 0
    lists = [[] for _ in range(len(x[0]))]
    for row in x:
        for i, col in enumerate(row):
            lists[i].append(col)
    return lists


# ------------------------------------------------------------------------------------ resize
def resize_antialias(img: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """(..., h, w) CUDA float32 or uint8 (read as u / 255) -> (..., oh, ow) float32, antialiased bilinear
    (``F.interpolate(mode="bilinear", antialias=True, align_corners=False)``; dataset.py:71-72)."""
    _lib.require_cuda(img)
    assert img.dtype in (torch.float32, torch.uint8)
    img = img.contiguous()
    ih, iw = img.shape[-2:]
    oh, ow = int(size[0]), int(size[1])
    n_planes = img.numel() // (ih * iw)
    out = torch.empty(img.shape[:-2] + (oh, ow), dtype=torch.float32, device=img.device)
    fn = "dcta_resize_bilinear_aa_u8" if img.dtype == torch.uint8 else "dcta_resize_bilinear_aa"
    with torch.cuda.device(img.device):
        _lib.call(fn, _lib.ptr(img), _lib.ptr(out), n_planes, ih, iw, oh, ow, _lib.stream_ptr(img.device))
    return out


def _resize_smaller_edge(h: int, w: int, size: int) -> Tuple[int, int]:
    """Output size of ``torchvision.transforms.Resize(size)`` with an int: the smaller edge becomes ``size``, the other
    ``int(size * long / short)``."""
    if h <= w:
        return size, int(size * w / h)
    return int(size * h / w), size


def max_image_size(dct_processor) -> int:
    """dataset.py:53-57: some room above the largest token grid gives better DCT features."""
    return max(dct_processor.patch_size * max(dct_processor.max_patch_w, dct_processor.max_patch_h), 768)


def crop(pixel_values: torch.Tensor, max_size: int) -> torch.Tensor:
    """dataset.py:59-73: images whose longer edge exceeds ``max_size`` are scaled down (antialiased) so that it does not.
    (c, h, w) CUDA uint8 or float32 -> float32 in [0, 1] (uint8 pixels become u / 255, exactly as torch's division)."""
    _, h, w = pixel_values.shape
    if max(h, w) > max_size:
        ar = h / w
        if h > w:
            h = max_size
            w = int(h / ar)
        else:
            w = max_size
            h = int(ar * w)
        oh, ow = _resize_smaller_edge(pixel_values.shape[1], pixel_values.shape[2], min(h, w))
        return resize_antialias(pixel_values, (oh, ow))
    if pixel_values.dtype == torch.uint8:
        from .util import to_device_pixels
        return to_device_pixels(pixel_values, pixel_values.device)
    return pixel_values


# ------------------------------------------------------------------------------------ decode
_POOL = None


def _pool(threads: int):
    global _POOL
    if _POOL is None or _POOL._max_workers != threads:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=threads)
    return _POOL


def _host_decode(stream: bytes):
    """One JPEG stream -> ((h, w, 3) uint8 array, channel order).  OpenCV's imdecode (libjpeg-turbo, the GIL released for
    the whole call, EXIF orientation ignored as PIL does) when it is importable, else PIL (which holds the GIL while it
    hands the pixels over)."""
    import numpy as np
    try:
        import cv2
    except ImportError:
        cv2 = None
    if cv2 is not None:
        a = cv2.imdecode(np.frombuffer(stream, dtype=np.uint8), cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION)
        if a is None:
            raise ValueError("not a decodable image stream")
        return a, "bgr"
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(stream)).convert("RGB")), "rgb"


class _Staging:
    """Two pinned byte buffers used alternately: the uploads of one decode batch are still in flight while the next batch
    is being decoded into the other buffer."""

    def __init__(self):
        self.bufs = [None, None]
        self.events = [None, None]
        self.turn = 0

    def get(self, nbytes: int, device):
        i = self.turn
        self.turn ^= 1
        if self.events[i] is not None:
            self.events[i].synchronize()
        if self.bufs[i] is None or self.bufs[i].numel() < nbytes:
            self.bufs[i] = torch.empty(max(nbytes, 1 << 24), dtype=torch.uint8).pin_memory()
        return i, self.bufs[i]

    def done(self, i: int, device):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self.events[i] = ev


_STAGING = _Staging()


def decode_jpegs(streams: List[bytes], device, decoder: str = "host", threads: Optional[int] = None) -> List[torch.Tensor]:
    """JPEG byte streams -> (3, h, w) uint8 RGB tensors on ``device`` (grey images replicated, as ``decode("torchrgb")``
    does).

    decoder="host" (default): libjpeg on a pool of ``threads`` host threads -- the decoder family of the reference
        (webdataset ``decode("torchrgb")`` = PIL / libjpeg; pixels identical on the test streams) -- the 8-bit
        interleaved pixels copied into a pinned staging buffer by the worker threads, uploaded asynchronously and
        transposed to planes on the device.
    decoder="nvjpeg": ``torchvision.io.decode_jpeg(device="cuda")`` on the whole batch (pixels within the tolerance two
        conforming decoders have: tests/test_gpu_dataset.py)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.DctaError("decode_jpegs delivers its pixels to the GPU; got device " + str(device))
    if decoder == "nvjpeg":
        from torchvision.io import ImageReadMode, decode_jpeg
        datas = [torch.frombuffer(bytearray(s), dtype=torch.uint8) for s in streams]
        return decode_jpeg(datas, mode=ImageReadMode.RGB, device=device)
    assert decoder == "host", decoder
    threads = threads or min(32, os.cpu_count() or 1)
    decoded = list(_pool(threads).map(_host_decode, streams))
    sizes = [int(a.size) for a, _ in decoded]
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + (n + 255) // 256 * 256)
    slot, stage = _STAGING.get(offs[-1], device)
    stage_np = stage.numpy()

    def put(i):
        a = decoded[i][0]
        stage_np[offs[i]:offs[i] + sizes[i]].reshape(a.shape)[...] = a       # memcpy outside the GIL
    list(_pool(threads).map(put, range(len(decoded))))
    out = []
    with torch.cuda.device(device):
        for i, (a, order) in enumerate(decoded):
            t = stage[offs[i]:offs[i] + sizes[i]].view(a.shape).to(device, non_blocking=True)    # (h, w, 3) bytes cross PCIe
            t = t.permute(2, 0, 1)
            out.append((t.flip(0) if order == "bgr" else t).contiguous())
        _STAGING.done(slot, device)
    return out


# ------------------------------------------------------------------------------------ webdataset-style tar shards
def expand_urls(url: str) -> List[str]:
    """A path, a glob, or a brace range ``a-{0000..0012}.tar`` (the webdataset notation) -> sorted file list."""
    m = re.search(r"\{(\d+)\.\.(\d+)\}", url)
    if m:
        lo, hi, width = int(m.group(1)), int(m.group(2)), len(m.group(1))
        return [url[:m.start()] + str(i).zfill(width) + url[m.end():] for i in range(lo, hi + 1)]
    files = sorted(glob.glob(url))
    return files if files else [url]


def iter_tar_samples(url: str, extensions=("jpg", "jpeg", "json")) -> Iterator[Dict]:
    """Samples of webdataset tar shards: members that share the basename up to the first dot form one sample
    ``{"__key__": ..., "jpg": bytes, "json": bytes}``."""
    for path in expand_urls(url):
        with tarfile.open(path, "r:*") as tf:
            cur_key, cur = None, {}
            for m in tf:
                if not m.isfile():
                    continue
                base = os.path.basename(m.name)
                key, _, ext = base.partition(".")
                key = os.path.join(os.path.dirname(m.name), key)
                if key != cur_key:
                    if cur:
                        yield cur
                    cur_key, cur = key, {"__key__": key}
                if ext.lower() in extensions:
                    cur[ext.lower()] = tf.extractfile(m).read()
            if cur:
                yield cur


def load_and_transform_dataset(dataset_url, dct_processor, device="cuda", decode_batch: int = 32,
                               decoder: str = "host") -> Iterator[Dict]:
    """dataset.py:35-89: an iterable over what ``dct_processor.preprocess`` returns for every image of the shards that
    passes the resolution filter.  ``dataset_url``: tar shard path / glob / brace range, or any iterable of sample dicts
    with ``jpg`` (bytes) and ``json`` (dict or bytes with ``height`` / ``width``).  Samples that fail to decode are
    skipped with a warning (``wds.handlers.warn_and_continue``)."""
    import warnings
    min_res = dct_processor.patch_size * 12                                  # dataset.py:46
    max_size = max_image_size(dct_processor)
    samples = iter_tar_samples(dataset_url) if isinstance(dataset_url, str) else iter(dataset_url)

    def filter_res(meta) -> bool:                                            # dataset.py:48-52
        h, w = meta.get("height"), meta.get("width")
        if h is None or w is None:
            return False
        return not (h < min_res or w < min_res)

    def flush(batch):
        try:
            images = decode_jpegs([s["jpg"] if "jpg" in s else s["jpeg"] for s in batch], device, decoder)
        except Exception as e:                                               # one bad stream: decode the rest one by one
            images = []
            for s in batch:
                try:
                    images.append(decode_jpegs([s["jpg"] if "jpg" in s else s["jpeg"]], device, decoder)[0])
                except Exception as e1:
                    warnings.warn(f"skipping {s.get('__key__')}: {e1!r}")
                    images.append(None)
            del e
        for im in images:
            if im is None:
                continue
            yield dct_processor.preprocess(crop(im, max_size))

    pending = []
    for s in samples:
        try:
            meta = s.get("json")
            if isinstance(meta, (bytes, bytearray, str)):
                meta = json.loads(meta)
            if meta is None or not filter_res(meta) or ("jpg" not in s and "jpeg" not in s):
                continue
        except Exception as e:
            warnings.warn(f"skipping {s.get('__key__')}: {e!r}")
            continue
        pending.append(s)
        if len(pending) == decode_batch:
            yield from flush(pending)
            pending = []
    if pending:
        yield from flush(pending)


def image_batches(dataset_url, dct_processor, device="cuda", decode_batch: int = 64, group_by_size: bool = True,
                  decoder: str = "host") -> Iterator[torch.Tensor]:
    """Decoded, filtered and down-scaled images of the shards as (b, 3, h, w) float32 batches of ONE size each, for
    ``shards.preprocess_to_shards`` (the offline shard writer, preproc_dataset.py:59-84).  ``group_by_size``: images of a
    decode batch that share their size travel together (their order inside the batch is kept, across sizes it is not);
    otherwise every image is its own batch, in the reference's order."""
    import warnings
    min_res = dct_processor.patch_size * 12
    max_size = max_image_size(dct_processor)
    samples = iter_tar_samples(dataset_url) if isinstance(dataset_url, str) else iter(dataset_url)

    def flush(streams):
        try:
            images = decode_jpegs(streams, device, decoder)
        except Exception as e:
            warnings.warn(f"skipping a decode batch of {len(streams)} images: {e!r}")
            return
        images = [crop(im, max_size) for im in images]
        if not group_by_size:
            for im in images:
                yield im[None]
            return
        groups: Dict[Tuple[int, int], List[torch.Tensor]] = {}
        for im in images:
            groups.setdefault(tuple(im.shape[-2:]), []).append(im)
        for ims in groups.values():
            yield torch.stack(ims)

    pending = []
    for s in samples:
        meta = s.get("json")
        if isinstance(meta, (bytes, bytearray, str)):
            meta = json.loads(meta)
        h, w = (meta or {}).get("height"), (meta or {}).get("width")
        if h is None or w is None or h < min_res or w < min_res or ("jpg" not in s and "jpeg" not in s):
            continue
        pending.append(s["jpg"] if "jpg" in s else s["jpeg"])
        if len(pending) == decode_batch:
            yield from flush(pending)
            pending = []
    if pending:
        yield from flush(pending)
