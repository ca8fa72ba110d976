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
        return (f"DCTPatches(patches={tuple(self.patches.shape)}, images={len(self.patch_sizes or [])}, "
                f"device={self.patches.device})")


def to_dict(dct_patches: DCTPatches, codes: torch.Tensor):
    """dct_patches.py:54-87: per image ``{size, original_size, codes: [{c, h, w, data}]}``.
    One bulk device->host copy instead of ``.item()`` per token."""
    b, s, _ = codes.shape
    assert b == dct_patches.patches.shape[0]
    assert s == dct_patches.patches.shape[1]
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
