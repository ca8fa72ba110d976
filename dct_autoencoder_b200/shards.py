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


class _StoragePickler(pickle.Pickler):
    """Pickles a tensor the way torch.save does: the storage becomes a persistent id naming record data/0."""

    def persistent_id(self, obj):
        if isinstance(obj, torch.storage.TypedStorage):
            return ("storage", getattr(torch, obj.pickle_storage_type()), "0", "cpu", obj._size())
        return None


_FAST_SAVE_OK: Dict[torch.dtype, bool] = {}


def _fast_tensor_bytes(t: torch.Tensor) -> bytes:
    """A torch.load-able zip archive (the torch.save container: data.pkl, byteorder, data/0, version) holding one
    contiguous CPU tensor.  torch.save spends ~7 ms per MB in its bundled CRC-32; writing the four records with
    ``zipfile`` (zlib's CRC-32) takes a fifth of that, and the tensor need not be cloned out of its batch buffer
    first: the pickled tensor is a view over exactly the bytes that are stored."""
    import warnings
    raw = t.reshape(-1).view(torch.uint8).numpy()
    exact = torch.from_numpy(raw).view(t.dtype).reshape(t.shape)      # same bytes, storage of exactly this extent
    pk = io.BytesIO()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                                # TypedStorage deprecation chatter
        _StoragePickler(pk, protocol=2).dump(exact)
    return _stored_zip([("archive/data.pkl", pk.getvalue()), ("archive/byteorder", b"little"),
                        ("archive/data/0", memoryview(raw)), ("archive/version", b"3\n")])


def _stored_zip(records) -> bytes:
    """A zip archive of uncompressed members, assembled with one copy of the payload (``zipfile`` costs half a
    millisecond of Python per member, which is most of the time for the small position / channel tensors)."""
    import struct
    import zlib
    parts, central, offset = [], [], 0
    for name, data in records:
        nb = name.encode()
        size, crc = len(data), zlib.crc32(data) & 0xffffffff
        # local file header: signature, version 20, flags 0, method 0 (stored), dos time/date (1980-01-01), crc, sizes
        head = struct.pack("<IHHHHHIIIHH", 0x04034b50, 20, 0, 0, 0, 0x21, crc, size, size, len(nb), 0) + nb
        central.append(struct.pack("<IHHHHHHIIIHHHHHII", 0x02014b50, 20, 20, 0, 0, 0, 0x21, crc, size, size, len(nb),
                                   0, 0, 0, 0, 0o600 << 16, offset) + nb)
        parts += [head, data]
        offset += len(head) + size
    cd = b"".join(central)
    parts += [cd, struct.pack("<IHHHHIIH", 0x06054b50, 0, 0, len(records), len(records), len(cd), offset, 0)]
    return b"".join(parts)


def _save_tensor(t: torch.Tensor) -> bytes:
    """``torch.save`` bytes of a tensor; the fast writer is used once it has round-tripped a probe of this dtype
    through ``torch.load(weights_only=True)`` in this process, torch.save itself otherwise."""
    if isinstance(t, torch.Tensor) and t.device.type == "cpu" and t.is_contiguous() and not t.requires_grad \
            and t.layout == torch.strided and t.numel() > 0:
        ok = _FAST_SAVE_OK.get(t.dtype)
        if ok is None:
            try:
                probe = torch.arange(24).reshape(2, 3, 4).to(t.dtype)[1]
                back = torch.load(io.BytesIO(_fast_tensor_bytes(probe)), weights_only=True)
                ok = back.dtype == probe.dtype and back.shape == probe.shape and torch.equal(back, probe)
            except Exception:
                ok = False
            _FAST_SAVE_OK[t.dtype] = ok
        if ok:
            return _fast_tensor_bytes(t)
    buf = io.BytesIO()
    torch.save(t, buf)
    return buf.getvalue()


def _encode(ext: str, value) -> bytes:
    """webdataset's default codecs for the two extensions the format uses."""
    if ext == "pth":
        return _save_tensor(value)
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
        self.tar = tarfile.open(fileobj=self.fileobj, mode="w", format=tarfile.USTAR_FORMAT, copybufsize=4 << 20)
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
    pool: Dict = {}              # (shape, dtype) -> free pinned buffers; list append / pop are atomic under the GIL

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
                        "patches.pth": pt[i, :k], "positions.pth": pos[i, :k], "channels.pth": ch[i, :k],
                        "original_size.pyd": osz, "patch_size.pyd": psz})
                for h in (pt, pos, ch):                 # written out: hand the staging buffers back
                    pool.setdefault((tuple(h.shape), h.dtype), []).append(h)
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
            # pinned staging buffers are recycled (pinning 100+ MB per batch costs more than the copy)
            host = []
            for t in (pt, pos, ch):
                sig = (tuple(t.shape), t.dtype)
                try:
                    h = pool[sig].pop()
                except (KeyError, IndexError):
                    h = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                h.copy_(t, non_blocking=True)
                host.append(h)
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
