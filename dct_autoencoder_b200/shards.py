"""Offline shard preprocessing: the production caller of ``preprocess``.

The reference runs ``preprocess`` image by image and stores every image's tokens as one sample of a
webdataset shard (preproc_dataset.py:62-84):

    {key}.patches.pth   torch.save of (k, p*p) patches in the chosen dtype
    {key}.positions.pth torch.save of (k, 2) int64 [h, w]
    {key}.channels.pth  torch.save of (k,) int64
    {key}.original_size.pyd / {key}.patch_size.pyd   pickled (h, w) / (ph, pw) tuples

in gzip-compressed tar files ``%06d.tar`` of at most ``maxsize`` bytes, and reads them back with
``load_preprocessed_dataset`` (dataset.py:27-33) as dicts with the keys ``iter_batches`` consumes.
``webdataset`` is a thin convention on top of tar files (members of one sample are adjacent and
share the basename up to the first dot; the extension selects the codec), so the container format is
written and read here with ``tarfile`` and the files are interchangeable with the reference's.

Here the encode runs on the GPU for a whole batch of same-size images at once
(``DCTAutoencoderFeatureExtractor.preprocess_batch``), the tokens come back to the host in one copy
per batch, and serialisation + compression run on a writer thread while the next batch is encoded.
"""
import glob
import gzip
import io
import os
import pickle
import queue
import re
import tarfile
import threading
import time
from typing import Dict, Iterable, Iterator, List, Optional, Sequence

import torch

_SAMPLE_KEYS = ("patches.pth", "positions.pth", "channels.pth", "original_size.pyd", "patch_size.pyd")


def _encode(ext: str, value) -> bytes:
    """webdataset's default codecs for the two extensions the format uses."""
    if ext == "pth":
        buf = io.BytesIO()
        torch.save(value, buf)
        return buf.getvalue()
    if ext == "pyd":
        return pickle.dumps(value)
    raise ValueError(f"no encoder for .{ext}")


def _decode(ext: str, data: bytes):
    if ext == "pth":
        return torch.load(io.BytesIO(data), weights_only=True)
    if ext == "pyd":
        return pickle.loads(data)
    return data           # decode(partial=True): unknown extensions stay bytes


class ShardWriter:
    """``wds.ShardWriter(pattern, maxsize=, maxcount=, compress=)`` as preproc_dataset.py:66 uses
    it: samples are dicts with ``__key__`` plus ``name.ext`` entries; a new shard is started when
    the current one holds ``maxcount`` samples or more than ``maxsize`` (uncompressed) bytes."""

    def __init__(self, pattern: str, maxcount: int = 100000, maxsize: float = 3e9, compress: bool = False,
                 start_shard: int = 0, shard_stride: int = 1):
        self.pattern, self.maxcount, self.maxsize, self.compress = pattern, maxcount, maxsize, compress
        self.shard, self.stride = start_shard, shard_stride
        self.count = self.size = self.total = 0
        self.tar: Optional[tarfile.TarFile] = None
        self.fileobj = None
        self.paths: List[str] = []

    def _next(self):
        self.finish()
        path = self.pattern % self.shard
        self.shard += self.stride
        self.paths.append(path)
        self.fileobj = gzip.open(path, "wb", compresslevel=6) if self.compress else open(path, "wb")
        self.tar = tarfile.open(fileobj=self.fileobj, mode="w|", format=tarfile.USTAR_FORMAT)
        self.count = self.size = 0

    def write(self, sample: Dict) -> int:
        if self.tar is None or self.count >= self.maxcount or self.size >= self.maxsize:
            self._next()
        key = sample["__key__"]
        now = time.time()
        n = 0
        for name, value in sample.items():
            if name.startswith("__"):
                continue
            data = value if isinstance(value, bytes) else _encode(name.rsplit(".", 1)[-1], value)
            info = tarfile.TarInfo(f"{key}.{name}")
            info.size, info.mtime, info.mode = len(data), now, 0o444
            info.uname = info.gname = "bigdata"
            self.tar.addfile(info, io.BytesIO(data))
            n += len(data)
        self.count += 1
        self.total += 1
        self.size += n
        return n

    def finish(self):
        if self.tar is not None:
            self.tar.close()
            self.fileobj.close()
            self.tar = self.fileobj = None

    close = finish

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.finish()


def expand_urls(url) -> List[str]:
    """A shard spec as the reference passes it to ``wds.WebDataset``: a list, a glob, a directory,
    or a brace range such as ``out/{000000..000012}.tar``."""
    if isinstance(url, (list, tuple)):
        return [p for u in url for p in expand_urls(u)]
    m = re.search(r"\{(\d+)\.\.(\d+)\}", url)
    if m:
        lo, hi, width = int(m.group(1)), int(m.group(2)), len(m.group(1))
        return [p for i in range(lo, hi + 1)
                for p in expand_urls(url[:m.start()] + str(i).zfill(width) + url[m.end():])]
    if os.path.isdir(url):
        return sorted(glob.glob(os.path.join(url, "*.tar")))
    if any(ch in url for ch in "*?["):
        return sorted(glob.glob(url))
    return [url]


def iter_samples(url) -> Iterator[Dict]:
    """Decoded samples of the shards, in file order (``wds.WebDataset(url).decode(partial=True)``).
    Gzip-compressed and plain tars are both accepted, as in webdataset."""
    for path in expand_urls(url):
        with tarfile.open(path, mode="r|*") as tar:
            cur_key, cur = None, {}
            for member in tar:
                if not member.isfile():
                    continue
                base = os.path.basename(member.name)
                stem, _, name = base.partition(".")
                key = os.path.join(os.path.dirname(member.name), stem)
                if key != cur_key:
                    if cur:
                        yield cur
                    cur_key, cur = key, {"__key__": key, "__url__": path}
                cur[name] = _decode(name.rsplit(".", 1)[-1], tar.extractfile(member).read())
            if cur:
                yield cur


def load_preprocessed_dataset(dataset_url) -> Iterator[Dict]:
    """dataset.py:27-33: samples renamed to what ``iter_batches`` consumes; samples missing a field
    are skipped (the reference's ``warn_and_continue`` handler)."""
    for row in iter_samples(dataset_url):
        if not all(k in row for k in _SAMPLE_KEYS):
            continue
        yield dict(patches=row["patches.pth"], positions=row["positions.pth"], channels=row["channels.pth"],
                   original_sizes=row["original_size.pyd"], patch_sizes=row["patch_size.pyd"])


def batched(samples: Iterable[Dict], n: int) -> Iterator[Dict]:
    """dict_collate (dataset.py:8-15) over groups of ``n`` samples: what a DataLoader with
    ``collate_fn=dict_collate`` hands to ``iter_batches``."""
    cols: Dict[str, list] = {}
    count = 0
    for s in samples:
        for k, v in s.items():
            cols.setdefault(k, []).append(v)
        count += 1
        if count == n:
            yield cols
            cols, count = {}, 0
    if count:
        yield cols


@torch.no_grad()
def preprocess_to_shards(image_batches: Iterable[torch.Tensor], processor, output_dir: str,
                         dtype: Optional[torch.dtype] = None, maxsize: float = 1e9, compress: bool = True,
                         ks: Optional[Iterable[Sequence[int]]] = None, first_key: int = 0,
                         queue_depth: int = 2, writers: int = 1) -> Dict:
    """preproc_dataset.py:62-84 for batches of same-size images.

    ``image_batches`` yields (b, c, h, w) RGB tensors in [0, 1] (host or device).  Each batch is
    encoded on the GPU in fp32; the stored patches take the dtype of the images, as ``preprocess``
    does (FE:139-141), or ``dtype`` when given (fp16 halves the shard size and the copy); one
    device->host transfer per tensor per batch into pinned memory, then a writer thread serialises
    and compresses while the next batch is encoded.  ``writers`` > 1 runs that many writer threads (zlib
    and file I/O release the GIL), writer i owning shards i, i + writers, ...; sample keys stay global,
    which shard a batch lands in then depends on timing.  Returns {"samples", "shards", "bytes"}."""
    os.makedirs(output_dir, exist_ok=True)
    pattern = os.path.join(output_dir, "%06d.tar")
    shard_writers = [ShardWriter(pattern, maxsize=maxsize, compress=compress, start_shard=i, shard_stride=writers)
                     for i in range(writers)]
    q: "queue.Queue" = queue.Queue(maxsize=queue_depth * writers)
    failure: List[BaseException] = []

    def drain(writer: ShardWriter):
        while True:
            item = q.get()
            if item is None:
                return
            if failure:
                continue
            try:
                ev, key0, recs, pt, pos, ch = item
                ev.synchronize()
                for i, (k, osz, psz) in enumerate(recs):
                    writer.size_total = getattr(writer, "size_total", 0) + writer.write({
                        "__key__": f"{key0 + i:08}",
                        "patches.pth": pt[i, :k].clone(), "positions.pth": pos[i, :k].clone(),
                        "channels.pth": ch[i, :k].clone(),
                        "original_size.pyd": osz, "patch_size.pyd": psz})
            except BaseException as e:      # surfaced to the caller after the loop
                failure.append(e)

    threads = [threading.Thread(target=drain, args=(w,), daemon=True) for w in shard_writers]
    for th in threads:
        th.start()
    key = first_key
    ks_iter = iter(ks) if ks is not None else None
    try:
        for images in image_batches:
            if failure:
                break
            pt, pos, ch, kk, osz, psz = processor._preprocess_batch_raw(
                images, next(ks_iter) if ks_iter is not None else None)
            out_dtype = dtype if dtype is not None else images.dtype
            if out_dtype != pt.dtype:
                pt = pt.to(out_dtype)
            host = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in (pt, pos, ch)]
            for h, t in zip(host, (pt, pos, ch)):
                h.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            meta = [(k, osz, psz) for k in kk]
            q.put((ev, key, meta, *host))
            key += len(kk)
    finally:
        for _ in threads:
            q.put(None)
        for th in threads:
            th.join()
        for w in shard_writers:
            w.finish()
    if failure:
        raise failure[0]
    return dict(samples=sum(w.total for w in shard_writers), shards=sorted(p for w in shard_writers for p in w.paths),
                bytes=sum(getattr(w, "size_total", 0) for w in shard_writers))
