"""The exchange type of the transform path.

Same field names and methods as the reference's ``DCTPatches`` dataclass (dct_patches.py:6-51)
so that callers (``PatchNorm``, the model glue, ``to_dict``) work unchanged.  Differences, both
deliberate:

* ``attn_mask`` is built lazily.  It is ``b*s*s`` bytes (9.4 MB per 3072-token row, 4x the patch
  data) and nothing on the encode/decode transform path reads it; the reference materialises it
  eagerly in ``_batch_groups`` (feature_extraction_dct_autoencoder.py:580-584).  The lazily built
  tensor is bit-identical, including the reference's polarity
  ``attn_mask[b,0,i,j] = (id_i == id_j) & key_pad_mask[b,j]``.
* tensors live on the GPU.
"""
from typing import Any, Dict, List, Optional, Tuple

import torch


class DCTPatches:
    def __init__(
        self,
        patches: torch.Tensor,
        key_pad_mask: torch.Tensor,
        attn_mask: Optional[torch.Tensor] = None,
        batched_image_ids: torch.Tensor = None,
        patch_channels: torch.Tensor = None,
        patch_positions: torch.Tensor = None,
        patch_sizes: List[Tuple] = None,
        original_sizes: List[Tuple] = None,
        _data: Optional[Dict[str, List[Any]]] = None,
        _row_num_images: Optional[List[int]] = None,
    ):
        self.patches = patches                        # (b, s, z)
        self.key_pad_mask = key_pad_mask              # (b, s) bool, True = padding
        self._attn_mask = attn_mask                   # (b, 1, s, s) bool, lazy
        self.batched_image_ids = batched_image_ids    # (b, s) int64
        self.patch_channels = patch_channels          # (b, s) int64
        self.patch_positions = patch_positions        # (b, s, 2) int64 [h, w]
        self.patch_sizes = patch_sizes                # per image (ph, pw)
        self.original_sizes = original_sizes          # per image (h, w) in pixels
        self._data = _data
        # host-side bookkeeping (not in the reference): images per row, known at packing time, so
        # that decoding never has to read batched_image_ids back from the device
        self._row_num_images = _row_num_images

    @property
    def attn_mask(self) -> torch.Tensor:
        if self._attn_mask is None:
            ids = self.batched_image_ids
            same = ids[:, None, :, None] == ids[:, None, None, :]
            self._attn_mask = same & self.key_pad_mask[:, None, None, :]
        return self._attn_mask

    @attn_mask.setter
    def attn_mask(self, value):
        self._attn_mask = value

    @property
    def h_indices(self):
        return self.patch_positions[..., 0]

    @property
    def w_indices(self):
        return self.patch_positions[..., 1]

    def shallow_copy(self) -> "DCTPatches":
        return DCTPatches(
            patches=self.patches, key_pad_mask=self.key_pad_mask, attn_mask=self._attn_mask,
            batched_image_ids=self.batched_image_ids, patch_channels=self.patch_channels,
            patch_positions=self.patch_positions, patch_sizes=self.patch_sizes,
            original_sizes=self.original_sizes, _data=self._data,
            _row_num_images=self._row_num_images)

    def to(self, what) -> "DCTPatches":
        """In place, returns self (dct_patches.py:44-51).  A dtype only applies to ``patches``
        semantics-wise in the reference too (bool/int tensors follow torch's .to rules)."""
        if self.patches is not None:            # None on the fused encode path (codes only)
            self.patches = self.patches.to(what)
        self.key_pad_mask = self.key_pad_mask.to(what)
        if self._attn_mask is not None:
            self._attn_mask = self._attn_mask.to(what)
        self.batched_image_ids = self.batched_image_ids.to(what)
        self.patch_channels = self.patch_channels.to(what)
        self.patch_positions = self.patch_positions.to(what)
        return self

    def row_num_images(self) -> List[int]:
        """Images per row: max valid image id + 1 (what ``image_ids.unique()`` yields in
        feature_extraction_dct_autoencoder.py:628 for rows packed by the extractor)."""
        if self._row_num_images is None:
            ids = self.batched_image_ids.masked_fill(self.key_pad_mask, 0)
            self._row_num_images = (ids.amax(dim=1) + 1).tolist()  # one device->host read
        return self._row_num_images

    def __repr__(self):
        shape = tuple(self.patches.shape) if self.patches is not None else None
        return (f"DCTPatches(patches={shape}, rows x slots={tuple(self.key_pad_mask.shape)}, "
                f"images={len(self.patch_sizes or [])}, device={self.key_pad_mask.device})")


def to_dict(dct_patches: DCTPatches, codes: torch.Tensor):
    """dct_patches.py:54-87: per image ``{size, original_size, codes: [{c, h, w, data}]}``.
    One bulk device->host copy instead of ``.item()`` per token."""
    b, s, _ = codes.shape
    # the reference asserts against patches.shape[:2]; key_pad_mask has the same (rows, slots) and is
    # also there for the code-only batches of the fused encode path (patches is None)
    assert (b, s) == tuple(dct_patches.key_pad_mask.shape)
    ids = dct_patches.batched_image_ids.cpu()
    pad = dct_patches.key_pad_mask.cpu()
    ch = dct_patches.patch_channels.cpu()
    pos = dct_patches.patch_positions.cpu()
    codes = codes.cpu()
    objs = []
    for r in range(b):
        for image_i in range(int(ids[r].max()) + 1):
            sel = (ids[r] == image_i) & ~pad[r]
            objs.append({
                "size": dct_patches.patch_sizes[len(objs)],
                "original_size": dct_patches.original_sizes[len(objs)],
                "codes": [
                    {"c": c, "h": h, "w": w, "data": d}
                    for c, (h, w), d in zip(ch[r][sel].tolist(), pos[r][sel].tolist(), codes[r][sel].tolist())
                ],
            })
    return objs


def from_dict(obj: dict, device=None):
    """dct_patches.py:90-122."""
    entries = obj["codes"]
    n = len(entries)
    dp = DCTPatches(
        patches=torch.zeros(1, device=device),
        key_pad_mask=torch.zeros(1, n, dtype=torch.bool, device=device),
        attn_mask=torch.ones(1, n, n, dtype=torch.bool, device=device),
        batched_image_ids=torch.zeros(1, n, dtype=torch.long, device=device),
        patch_channels=torch.tensor([e["c"] for e in entries], dtype=torch.long, device=device).reshape(1, n),
        patch_positions=torch.tensor([[e["h"], e["w"]] for e in entries], dtype=torch.long,
                                     device=device).reshape(1, n, 2),
        patch_sizes=[obj["size"]],
        original_sizes=[obj["original_size"]],
        _row_num_images=[1],
    )
    codes = torch.tensor([e["data"] for e in entries], dtype=torch.long, device=device)
    return dp, codes


# ----------------------------------------------------------------------------------------------
# compact wire format (csrc/wire.cu): the binary counterpart of to_dict / from_dict
# ----------------------------------------------------------------------------------------------
_WIRE_MAGIC = b"DCTW"
_WIRE_HEADER = "<4sBBBBHHIII"      # magic, version, c, d, reserved, ph, pw, H, W, n_tokens
WIRE_HEADER_BYTES = 24


def wire_record_bytes(num_codebooks: int, bits: int) -> int:
    return 2 + (num_codebooks * bits + 7) // 8


def to_bytes(dct_patches: DCTPatches, codes: torch.Tensor, codebook_size: int) -> List[bytes]:
    """One byte string per image: a 24-byte header (magic ``DCTW``, version, codebooks, bits per code,
    patch grid (ph, pw), original size (H, W), token count) followed by one fixed-size record per
    token, ``u16 c<<12|h<<6|w`` + the code words bit-packed MSB first (include/dcta.h
    ``dcta_wire_pack``).  Carries exactly what ``to_dict`` (dct_patches.py:54-87) carries, in
    27 instead of ~400 bytes per token for the 14x14-bit quantiser, and is produced by one kernel and
    one device->host copy instead of one ``.item()`` per field."""
    import struct
    b, s, c = codes.shape
    d = max(1, (int(codebook_size) - 1).bit_length())
    assert 2 ** d == codebook_size, "codebook_size must be a power of two"
    if int(dct_patches.patch_channels.max()) >= 16 or int(dct_patches.patch_positions.max()) >= 64:
        raise ValueError("wire format holds channel < 16 and h, w < 64")
    out, counts = wire_records(dct_patches, codes, d)
    n_img = dct_patches.row_num_images()
    h_out = out.cpu().numpy()
    h_counts = counts[:, :max(n_img)].cpu().numpy()
    blobs = []
    for r in range(b):
        off = 0
        for i in range(n_img[r]):
            k = int(h_counts[r, i])
            ph, pw = dct_patches.patch_sizes[len(blobs)]
            oh, ow = dct_patches.original_sizes[len(blobs)]
            head = struct.pack(_WIRE_HEADER, _WIRE_MAGIC, 1, c, d, 0, ph, pw, oh, ow, k)
            blobs.append(head + h_out[r, off:off + k].tobytes())
            off += k
    return blobs


def wire_records(dct_patches: DCTPatches, codes: torch.Tensor, bits: int):
    """Device half of ``to_bytes``: (rows, s, codebooks) codes + the batch's positions / channels / ids / mask ->
    records (rows, s, 2 + ceil(codebooks*bits/8)) uint8, one per slot (the images of a row follow each other and the
    padding trails, FE:455-513, so image i of row r owns records [sum(counts[r, :i]), +counts[r, i])), and
    counts (rows, s) int32 with counts[r, i] = tokens of image i of row r.  One kernel
    (csrc/wire.cu); the caller must make sure channel < 16 and h, w < 64 (``to_bytes`` checks)."""
    from . import _lib
    _lib.require_cuda(codes)
    b, s, c = codes.shape
    assert (b, s) == tuple(dct_patches.key_pad_mask.shape)
    rec = wire_record_bytes(c, bits)
    dev = codes.device
    codes = codes.to(torch.int64).contiguous()
    pos = dct_patches.patch_positions.to(dev, torch.int64).contiguous()
    ch = dct_patches.patch_channels.to(dev, torch.int64).contiguous()
    ids = dct_patches.batched_image_ids.to(dev, torch.int64).contiguous()
    pad = dct_patches.key_pad_mask.to(dev).contiguous()
    out = torch.empty((b, s, rec), dtype=torch.uint8, device=dev)
    counts = torch.empty((b, s), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dcta_wire_pack", _lib.ptr(codes), _lib.ptr(pos), _lib.ptr(ch), _lib.ptr(ids),
                  pad.data_ptr(), b, s, c, bits, rec, _lib.ptr(out), _lib.ptr(counts), _lib.stream_ptr(dev))
    return out, counts


def from_bytes(blob: bytes, device=None):
    """Inverse of ``to_bytes`` for one image -> (DCTPatches, codes (n, c) int64), the pair
    ``from_dict`` (dct_patches.py:90-122) returns."""
    import struct

    import numpy as np

    from . import _lib
    from .util import default_device
    magic, version, c, d, _, ph, pw, oh, ow, n = struct.unpack_from(_WIRE_HEADER, blob, 0)
    if magic != _WIRE_MAGIC or version != 1:
        raise ValueError("not a DCTW version-1 record")
    rec = wire_record_bytes(c, d)
    if len(blob) != WIRE_HEADER_BYTES + n * rec:
        raise ValueError(f"truncated record: {len(blob)} bytes for {n} tokens of {rec} bytes")
    dev = torch.device(device) if device is not None else default_device()
    body = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8, offset=WIRE_HEADER_BYTES).copy()).to(dev)
    codes = torch.empty((n, c), dtype=torch.int64, device=dev)
    pos = torch.empty((1, n, 2), dtype=torch.int64, device=dev)
    ch = torch.empty((1, n), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dcta_wire_unpack", _lib.ptr(body), n, c, d, rec, _lib.ptr(codes), _lib.ptr(pos), _lib.ptr(ch),
                  _lib.stream_ptr(dev))
    dp = DCTPatches(
        patches=torch.zeros(1, device=dev),
        key_pad_mask=torch.zeros(1, n, dtype=torch.bool, device=dev),
        attn_mask=torch.ones(1, n, n, dtype=torch.bool, device=dev),
        batched_image_ids=torch.zeros(1, n, dtype=torch.long, device=dev),
        patch_channels=ch, patch_positions=pos,
        patch_sizes=[(ph, pw)], original_sizes=[(oh, ow)], _row_num_images=[1],
    )
    return dp, codes
