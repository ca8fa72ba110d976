"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the reference's encode/decode
transform path.  This is the parity oracle: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product package (``dct_autoencoder_b200``) never does; it fails loudly without its CUDA
library instead.

Every function cites the reference lines it restates (paths relative to
``/root/reference/dct_autoencoder``; FE = feature_extraction_dct_autoencoder.py,
UT = util.py, PN = patchnorm.py, DP = dct_patches.py, VQ = vector_quantize.py).

PINNING
  * Everything except the DCT arithmetic is pinned to the reference itself:
    ``tests/golden/make_golden.py`` imports the unmodified reference in the build
    container, runs it on seeded inputs and commits the outputs under
    ``tests/golden/``; ``tests/test_oracle.py`` checks this file against them.
  * DCT arithmetic: **parity unpinned at the bit level.**  It lives in the third-party
    package ``torch_dct==0.1.6`` (requirements.txt:11; call sites UT:9, UT:333-338,
    FE:140, FE:149), which is absent from /root/reference and not installable here, and
    the reference's tests hold no vector for it.  The oracle uses the mathematical
    definition (orthonormal DCT-II / DCT-III) in float64 via ``scipy.fft``; the golden
    fixtures were produced with ``oracle/torch_dct_standin.py`` (the package's published
    fp32 FFT algorithm).  The two agree to <= 1e-7 * max|Y| (test_oracle.py).

All arithmetic that the reference does in float32 is done in float32 here, in the same
operation order, so that integer outputs (selection order, codes) can be compared
bit-exactly for identical inputs.
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.fft

f32 = np.float32

# --------------------------------------------------------------------------------------
# colour space (UT:21-43 constants, UT:46-47 channel_mult, UT:56-97)
# --------------------------------------------------------------------------------------
# The reference builds Trgb2lms = MHPE @ MsRGB and Tlms2rgb = Trgb2lms.inverse() in
# float32 at import time (UT:40-41) and Mipt.inverse() per call (UT:91).  The literals
# below are those float32 results as observed from the reference in the build container
# (tests/test_oracle.py::test_constants_match_reference re-derives them when the
# reference is present).
_h = float.fromhex
TRGB2LMS = np.array(
    [_h("0x1.418728p-2"), _h("0x1.476fccp-1"), _h("0x1.7dafbcp-5"),
     _h("0x1.36b76ap-3"), _h("0x1.7f130cp-1"), _h("0x1.99aebp-4"),
     _h("0x1.22eb2p-6"), _h("0x1.c05f58p-4"), _h("0x1.beda4ep-1")], dtype=f32).reshape(3, 3)
TLMS2RGB = np.array(
    [_h("0x1.5b9eaep+2"), _h("-0x1.2b6fe6p+2"), _h("0x1.f84438p-3"),
     _h("-0x1.1aebbcp+0"), _h("0x1.27d972p+1"), _h("-0x1.a5a628p-3"),
     _h("0x1.cc8d4cp-6"), _h("-0x1.8ec412p-3"), _h("0x1.2aa69ep+0")], dtype=f32).reshape(3, 3)
MIPT = np.array(
    [_h("0x1.99999ap-2"), _h("0x1.99999ap-2"), _h("0x1.99999ap-3"),
     _h("0x1.1d1eb8p+2"), _h("-0x1.3676c8p+2"), _h("0x1.958106p-2"),
     _h("0x1.9c779ap-1"), _h("0x1.6dc5d6p-2"), _h("-0x1.29ad42p+0")], dtype=f32).reshape(3, 3)
MIPT_INV = np.array(
    [_h("0x1.0p+0"), _h("0x1.8fa46ep-4"), _h("0x1.a44dc2p-3"),
     _h("0x1.0p+0"), _h("-0x1.d27028p-4"), _h("0x1.10d428p-3"),
     _h("0x1.0p+0"), _h("0x1.0b2ed8p-5"), _h("-0x1.5a90f6p-1")], dtype=f32).reshape(3, 3)
IPT_GAMMA = 0.43  # UT:43


def channel_mult(M: np.ndarray, x: np.ndarray) -> np.ndarray:
    """UT:46-47: ``einsum('i j, ... j h w -> ... i h w')`` in the dtype of ``x``."""
    M = M.astype(x.dtype)
    return np.einsum("ij,...jhw->...ihw", M, x).astype(x.dtype)


def rgb_to_ipt(x: np.ndarray) -> np.ndarray:
    """UT:70-82 (rgb_to_lms UT:56-60)."""
    x = channel_mult(TRGB2LMS, x)
    neg = x < 0
    x = np.power(np.abs(x), x.dtype.type(IPT_GAMMA)).astype(x.dtype)
    x[neg] = -x[neg]
    return channel_mult(MIPT, x)


def ipt_to_rgb(x: np.ndarray) -> np.ndarray:
    """UT:85-97 (lms_to_rgb UT:63-67)."""
    x = channel_mult(MIPT_INV, x)
    neg = x < 0
    x = np.power(np.abs(x), x.dtype.type(1 / IPT_GAMMA)).astype(x.dtype)
    x[neg] = -x[neg]
    return channel_mult(TLMS2RGB, x)


# --------------------------------------------------------------------------------------
# DCT (UT:333-338 -> torch_dct 0.1.6, "ortho")
# --------------------------------------------------------------------------------------
def dct2(x: np.ndarray) -> np.ndarray:
    """Orthonormal 2-D DCT-II over the last two axes (FE:140).  float64 inside, result in
    the dtype of ``x`` (the reference forces float32, FE:139)."""
    y = scipy.fft.dctn(x.astype(np.float64), type=2, norm="ortho", axes=(-2, -1))
    return y.astype(x.dtype)


def idct2(x: np.ndarray) -> np.ndarray:
    """Orthonormal 2-D DCT-III (inverse of dct2) over the last two axes (FE:149)."""
    y = scipy.fft.idctn(x.astype(np.float64), type=2, norm="ortho", axes=(-2, -1))
    return y.astype(x.dtype)


def dct_basis(n: int, k: Optional[int] = None, dtype=np.float64) -> np.ndarray:
    """Rows 0..k-1 of the orthonormal DCT-II matrix C_n (definition behind UT:333-334):
    ``C[q, m] = s_q * cos(pi * (2m + 1) * q / (2n))``, ``s_0 = sqrt(1/n)``, else ``sqrt(2/n)``."""
    k = n if k is None else k
    q = np.arange(k, dtype=np.float64)[:, None]
    m = np.arange(n, dtype=np.float64)[None, :]
    c = np.cos(np.pi * (2 * m + 1) * q / (2 * n)) * math.sqrt(2.0 / n)
    c[0, :] = math.sqrt(1.0 / n)
    return c.astype(dtype)


def transform_image_in(x: np.ndarray) -> np.ndarray:
    """FE:130-142."""
    return dct2(rgb_to_ipt(x).astype(f32)).astype(x.dtype)


def transform_image_out(x: np.ndarray) -> np.ndarray:
    """FE:144-152."""
    return ipt_to_rgb(idct2(x.astype(f32))).astype(x.dtype)


# --------------------------------------------------------------------------------------
# feature extractor (FE:107-656)
# --------------------------------------------------------------------------------------
def exp_trunc_dist(a: float) -> float:
    """UT:167-172 -- consumes one draw of the module-global Python RNG."""
    x = random.random()
    return -1 / a * math.log(x)


def power_of_two(target: int) -> int:
    """UT:184-189."""
    if target > 1:
        for i in range(1, int(target)):
            if 2 ** i >= target:
                return 2 ** i
    return 1


def get_max_seq_length(max_patch_h, max_patch_w, image_channels, sample_patches_beta, cdf_p=0.95):
    """factory.py:11-33."""
    full = max_patch_h * max_patch_w * image_channels
    if sample_patches_beta <= 0:
        return full
    n = round(-1 * math.log(1 - cdf_p) / sample_patches_beta)
    return min(full, power_of_two(n))


@dataclass
class Patches:
    """Restatement of the boundary type DCTPatches (DP:6-51) with numpy fields."""
    patches: np.ndarray            # (b, s, z)
    key_pad_mask: np.ndarray       # (b, s) bool, True = padding      (FE:574-576)
    batched_image_ids: np.ndarray  # (b, s) int64, pad = 0            (FE:580)
    patch_channels: np.ndarray     # (b, s) int64
    patch_positions: np.ndarray    # (b, s, 2) int64 [h, w]
    patch_sizes: List[Tuple[int, int]]
    original_sizes: List[Tuple[int, int]]
    _data: Optional[Dict[str, List[Any]]] = None

    @property
    def attn_mask(self) -> np.ndarray:
        """FE:580-584: ``(id_i == id_j) & key_pad_mask[j]`` (True only where key j is padding)."""
        ids = self.batched_image_ids
        m = ids[:, None, :, None] == ids[:, None, None, :]
        return m & self.key_pad_mask[:, None, None, :]

    @property
    def h_indices(self):
        return self.patch_positions[..., 0]

    @property
    def w_indices(self):
        return self.patch_positions[..., 1]


class FeatureExtractor:
    """Restatement of DCTAutoencoderFeatureExtractor (FE:107-656)."""

    def __init__(self, channels, patch_size, sample_patches_beta, max_patch_h, max_patch_w,
                 max_seq_len, channel_importances=(8.0, 1.0, 1.0),
                 patch_sample_magnitude_weight=0.1):
        self.channels = channels
        self.patch_size = patch_size
        self.sample_patches_beta = sample_patches_beta
        self.max_patch_h = max_patch_h
        self.max_patch_w = max_patch_w
        self.max_seq_len = max_seq_len
        self.channel_importances = np.asarray(channel_importances, dtype=f32)  # torch.Tensor -> fp32
        self.patch_sample_magnitude_weight = patch_sample_magnitude_weight
        # hooks so that tests can replace the transform by the identity (testpatching.py:42-43)
        self._transform_image_in = transform_image_in
        self._transform_image_out = transform_image_out

    # FE:312-345
    def _get_crop_dims(self, h: int, w: int):
        assert h >= self.patch_size
        assert w >= self.patch_size
        p_h = max(int(h / self.patch_size), 1)
        p_w = max(int(w / self.patch_size), 1)
        return p_h * self.patch_size, p_w * self.patch_size

    # FE:348-362
    def _crop_image(self, x):
        c, h, w = x.shape
        assert c == self.channels
        c_h, c_w = self._get_crop_dims(h, w)
        return x[:, :c_h, :c_w]

    def importance_scores(self, x: np.ndarray):
        """FE:374-416.  ``x``: cropped coefficient plane (c, h, w).  Returns the tiles
        ``(n_tiles, c, p*p)``, ``h_idx``/``w_idx`` ``(n_tiles,)`` int64 and the float32 scores
        ``(n_tiles, c)`` for the in-bounds tiles, in the reference's pre-sort order."""
        c, h, w = x.shape
        p = self.patch_size
        assert h % p == 0 and w % p == 0
        ph, pw = h // p, w // p
        t = x.reshape(c, ph, p, pw, p).transpose(1, 3, 0, 2, 4).reshape(ph * pw, c, p * p)
        h_idx, w_idx = np.meshgrid(np.arange(ph), np.arange(pw), indexing="ij")
        keep = (h_idx < self.max_patch_h) & (w_idx < self.max_patch_w)
        t = t[keep.reshape(-1)]
        h_idx = h_idx[keep].astype(np.int64)
        w_idx = w_idx[keep].astype(np.int64)
        pos = (-1 * (h_idx + w_idx)).astype(np.int64)[:, None]
        mags = np.abs(t).max(-1)                                   # (n, c) dtype of x
        mags = (mags * mags.dtype.type(self.patch_sample_magnitude_weight))
        # int64 / float32 tensor -> float32 division (torch type promotion)
        dist = pos.astype(f32) / self.channel_importances[None, :c]
        scores = (mags.astype(f32) + dist).astype(f32) if mags.dtype != np.float64 else mags + dist
        return t, h_idx, w_idx, scores

    def _choose_k(self, n: int) -> int:
        """FE:429-435."""
        k = n
        if self.sample_patches_beta > 0.0:
            k = min(round(exp_trunc_dist(self.sample_patches_beta)), k)
            k = max(1, k)
        return min(k, self.max_seq_len)

    # FE:365-452
    def _patch_image(self, x: np.ndarray, k: Optional[int] = None):
        c = x.shape[0]
        t, h_idx, w_idx, scores = self.importance_scores(x)
        flat = scores.reshape(-1)
        # descending; ties in ascending flat index (the reference's order among exact ties
        # is implementation-defined: torch.sort(stable=False), FE:418)
        order = np.argsort(-flat.astype(np.float64), kind="stable")
        if k is None:
            k = self._choose_k(len(order))
        sel = order[:k]
        hs = np.repeat(h_idx[:, None], c, 1).reshape(-1)[sel]
        ws = np.repeat(w_idx[:, None], c, 1).reshape(-1)[sel]
        cs = np.repeat(np.arange(c, dtype=np.int64)[None, :], len(h_idx), 0).reshape(-1)[sel]
        patches = t.reshape(-1, self.patch_size ** 2)[sel]
        assert patches.shape[0] <= self.max_seq_len
        return patches, np.stack([hs, ws], -1), cs

    # FE:155-177
    def preprocess(self, im: np.ndarray, k: Optional[int] = None):
        im = self._transform_image_in(im)
        _, h, w = im.shape
        im = self._crop_image(im)
        _, ch, cw = im.shape
        p = self.patch_size
        patches, pos, channels = self._patch_image(im, k)
        return dict(patches=patches, positions=pos, channels=channels,
                    original_sizes=(h, w), patch_sizes=(ch // p, cw // p))

    # FE:455-513 -- next-fit packing.  Returns the list of CLOSED rows and the open row,
    # each row being a list of image indices.
    def group_by_max_seq_len(self, ks: Sequence[int], state=None):
        if state is None:
            state = dict(groups=[], group=[], seq_len=0, n_seen=0)
        for k in ks:
            assert k <= self.max_patch_h * self.max_patch_w * self.channels and k <= self.max_seq_len
            if state["seq_len"] + k > self.max_seq_len:
                state["groups"].append(state["group"])
                state["group"] = []
                state["seq_len"] = 0
            state["group"].append(state["n_seen"])
            state["n_seen"] += 1
            state["seq_len"] += k
        return state

    # FE:516-605 (+ pad_sequence UT:149-164)
    def batch_groups(self, groups: List[List[int]], items: List[dict], first_image: int = 0) -> Patches:
        """``groups``: rows of GLOBAL image indices; ``items[i - first_image]`` is image i."""
        b, s, z = len(groups), self.max_seq_len, self.patch_size ** 2
        dtype = items[0]["patches"].dtype if items else f32
        patches = np.zeros((b, s, z), dtype)
        pos = np.zeros((b, s, 2), np.int64)
        chan = np.zeros((b, s), np.int64)
        ids = np.zeros((b, s), np.int64)
        lengths = np.zeros((b,), np.int64)
        osz, psz = [], []
        for r, row in enumerate(groups):
            o = 0
            for image_id, gi in enumerate(row):
                it = items[gi - first_image]
                k = it["patches"].shape[0]
                patches[r, o:o + k] = it["patches"]
                pos[r, o:o + k] = it["positions"]
                chan[r, o:o + k] = it["channels"]
                ids[r, o:o + k] = image_id
                o += k
                osz.append(tuple(it["original_sizes"]))
                psz.append(tuple(it["patch_sizes"]))
            lengths[r] = o
        key_pad_mask = lengths[:, None] <= np.arange(s)[None, :]
        return Patches(patches, key_pad_mask, ids, chan, pos, psz, osz)

    # FE:180-287.  ``loader`` yields dicts of lists (dataset.py:8-15 dict_collate).
    def iter_batches(self, loader, batch_size: Optional[int] = None):
        state = None
        items: List[dict] = []   # images not yet emitted, aligned with global index `first`
        first = 0
        for d in loader:
            n = len(d["patches"])
            for i in range(n):
                items.append({k: d[k][i] for k in ("patches", "positions", "channels",
                                                   "original_sizes", "patch_sizes")})
            state = self.group_by_max_seq_len([p.shape[0] for p in d["patches"]], state)
            if batch_size is None and len(state["group"]) > 0:   # FE:220-229
                state["groups"].append(state["group"])
                state["group"] = []
                state["seq_len"] = 0
            if batch_size is None or len(state["groups"]) > batch_size:   # FE:233
                emit = state["groups"][:batch_size]
                state["groups"] = state["groups"][batch_size:] if batch_size is not None else state["groups"]
                n_items = sum(len(g) for g in emit)
                batch = self.batch_groups(emit, items[:n_items], first)
                items = items[n_items:]
                first += n_items
                if batch_size is None:
                    # FE:236 `groups[None:]` keeps every row: the reference is single-shot in
                    # this mode (every caller takes next(iter(...)) once); we stop here.
                    yield batch
                    return
                yield batch
        # FE:190-194: the tail (open row and <= batch_size closed rows) is dropped.

    # FE:607-656
    def revert_patching(self, out: Patches) -> List[np.ndarray]:
        x = out.patches
        z = x.shape[-1]
        p = self.patch_size
        images = []
        for r in range(x.shape[0]):
            ids, mask = out.batched_image_ids[r], out.key_pad_mask[r]
            for image_id in np.unique(ids):
                sel = (ids == image_id) & ~mask
                ph, pw = out.patch_sizes[len(images)]
                img = np.zeros((self.channels, ph, pw, z), x.dtype)
                # sequential "last write wins" (FE:639-643); numpy fancy assignment keeps the
                # last of duplicate indices, as the loop does
                img[out.patch_channels[r][sel], out.patch_positions[r][sel, 0],
                    out.patch_positions[r][sel, 1]] = x[r][sel]
                img = img.reshape(self.channels, ph, pw, p, p).transpose(0, 1, 3, 2, 4)
                images.append(img.reshape(self.channels, ph * p, pw * p))
        return images

    # FE:289-310
    def postprocess(self, x: Patches) -> List[np.ndarray]:
        outs = []
        for image, (h, w) in zip(self.revert_patching(x), x.original_sizes):
            ch, cw = image.shape[-2:]
            pad = np.zeros((self.channels, h, w), image.dtype)
            pad[:, :ch, :cw] = image
            outs.append(self._transform_image_out(pad))
        return outs


# --------------------------------------------------------------------------------------
# PatchNorm (PN:32-177)
# --------------------------------------------------------------------------------------
class PatchNorm:
    def __init__(self, max_patch_h, max_patch_w, patch_size, channels, eps=1e-6,
                 max_val=6.0, min_val=-6.0, dtype=f32):
        self.eps, self.max_val, self.min_val = eps, max_val, min_val
        self.patch_size, self.channels = patch_size, channels
        self.max_patch_h, self.max_patch_w = max_patch_h, max_patch_w
        self.n = np.zeros((channels, max_patch_h, max_patch_w), dtype)                    # PN:52-59
        self.median = np.zeros((channels, max_patch_h, max_patch_w, patch_size ** 2), dtype)  # PN:61-64
        self.b = np.ones((channels, max_patch_h, max_patch_w, patch_size ** 2), dtype)    # PN:66-69
        self.frozen = False
        self.training = True

    def _std(self, c, h, w):
        # PN:158: b * 2**0.5 + eps -- python float scalars, tensor dtype kept
        t = self.b.dtype.type
        return self.b[c, h, w] * t(2 ** 0.5) + t(self.eps)

    # PN:81-165
    def forward(self, dp: Patches) -> np.ndarray:
        x = dp.patches
        c, h, w, pad = dp.patch_channels, dp.h_indices, dp.w_indices, dp.key_pad_mask
        if self.training and not self.frozen:
            self.update_stats(x[~pad], c[~pad], h[~pad], w[~pad])
            out = np.zeros_like(x)
            out[~pad] = x[~pad]
            return out                                                                   # PN:153-155
        y = (x - self.median[c, h, w]) / self._std(c, h, w)                              # PN:157-161
        return np.clip(y, x.dtype.type(self.min_val), x.dtype.type(self.max_val))        # PN:163

    # PN:101-150
    def update_stats(self, x, c, h, w):
        """``x`` (T, z) valid tokens only, in flattened (row, slot) order."""
        dt = x.dtype
        C, H, W, Z = self.median.shape
        flat = (c * H * W + h * W + w).astype(np.int64)
        batch_n = np.bincount(flat, minlength=C * H * W).astype(dt).reshape(C, H, W)     # PN:112-119
        batch_median = np.zeros_like(self.median).reshape(C * H * W, Z)
        order = np.argsort(flat, kind="stable")
        bounds = np.searchsorted(flat[order], np.arange(C * H * W + 1))
        for pid in np.nonzero(batch_n.reshape(-1))[0]:                                   # PN:123-130
            rows = x[order[bounds[pid]:bounds[pid + 1]]]
            n = rows.shape[0]
            # torch.median = LOWER middle element for even counts
            batch_median[pid] = np.partition(rows, (n - 1) // 2, axis=0)[(n - 1) // 2]
        batch_median = batch_median.reshape(C, H, W, Z)
        one = dt.type(1)
        denom = np.maximum(self.n + batch_n, one)[..., None]
        self.median = ((self.median * self.n[..., None] + batch_median * batch_n[..., None]) / denom).astype(dt)  # PN:135-138
        dist = np.abs(x - self.median[c, h, w])                                           # PN:140
        batch_b = np.zeros((C * H * W, Z), dt)
        # scatter_add_ on CPU accumulates in token order (PN:142-143)
        np.add.at(batch_b, flat, dist)
        batch_b = batch_b.reshape(C, H, W, Z) / np.maximum(batch_n, one)[..., None]       # PN:144
        self.b = ((self.b * self.n[..., None] + batch_b * batch_n[..., None]) / denom).astype(dt)  # PN:146-148
        self.n = (self.n + batch_n).astype(dt)                                            # PN:150

    # PN:167-177
    def inverse_norm(self, dp: Patches) -> np.ndarray:
        c, h, w = dp.patch_channels, dp.h_indices, dp.w_indices
        return dp.patches * self._std(c, h, w) + self.median[c, h, w]


# --------------------------------------------------------------------------------------
# LFQ (lfq.py:35-227)
# --------------------------------------------------------------------------------------
class LFQ:
    """Lookup-free quantizer.  ``w_in/b_in/w_out/b_out`` are the optional nn.Linear
    parameters (lfq.py:60-62); None means Identity."""

    def __init__(self, dim=None, codebook_size=None, num_codebooks=1, codebook_scale=1.0,
                 keep_num_codebooks_dim=None, w_in=None, b_in=None, w_out=None, b_out=None):
        assert dim is not None or codebook_size is not None
        assert codebook_size is None or math.log2(codebook_size).is_integer()
        codebook_size = codebook_size if codebook_size is not None else 2 ** dim
        self.codebook_dim = int(math.log2(codebook_size))
        self.codebook_dims = self.codebook_dim * num_codebooks
        self.dim = dim if dim is not None else self.codebook_dims
        self.has_projections = self.dim != self.codebook_dims
        self.num_codebooks = num_codebooks
        self.keep_num_codebooks_dim = (num_codebooks > 1) if keep_num_codebooks_dim is None else keep_num_codebooks_dim
        assert not (num_codebooks > 1 and not self.keep_num_codebooks_dim)
        self.codebook_scale = codebook_scale
        self.mask = (2 ** np.arange(self.codebook_dim - 1, -1, -1)).astype(np.int64)      # lfq.py:87 MSB first
        self.w_in, self.b_in, self.w_out, self.b_out = w_in, b_in, w_out, b_out
        self.training = False

    @property
    def codebook(self):
        """lfq.py:92-96: all sign patterns, MSB first."""
        codes = np.arange(2 ** self.codebook_dim, dtype=np.int64)
        bits = ((codes[:, None] & self.mask) != 0).astype(f32)
        return bits * f32(self.codebook_scale * 2) - f32(self.codebook_scale)

    def _proj(self, x, w, b):
        if w is None:
            return x
        y = x @ w.T
        return y + b if b is not None else y

    # lfq.py:136-227
    def forward(self, x: np.ndarray, mask: Optional[np.ndarray]):
        if mask is None:
            raise NotImplementedError("mask")                                              # lfq.py:153-154
        assert x.shape[-1] == self.dim
        x = self._proj(x, self.w_in, self.b_in) if self.has_projections else x
        b, n, _ = x.shape
        x = x.reshape(b, n, self.num_codebooks, self.codebook_dim)
        s = x.dtype.type(self.codebook_scale)
        q = np.where(x > 0, s, -s).astype(x.dtype)                                          # lfq.py:174-175
        indices = ((x > 0).astype(np.int64) * self.mask).sum(-1)                           # lfq.py:187 (int64 out)
        if self.training:
            distance = -2 * np.einsum("...id,jd->...ij", x, self.codebook.astype(x.dtype))  # lfq.py:191
            m = mask.astype(x.dtype)
            se = (x - q) ** 2                                                               # lfq.py:197
            # masked_mean(dim=0).sum(0).mean() (lfq.py:199, UT:346-353)
            commit = ((se * m[:, :, None, None]) / m.sum()).sum(0).sum(0).mean()
        else:
            distance = f32(0.0)
            commit = f32(0.0)
        out = q.reshape(b, n, self.codebook_dims)
        out = self._proj(out, self.w_out, self.b_out) if self.has_projections else out
        if not self.keep_num_codebooks_dim:
            indices = indices[..., 0]
        return out, indices, commit, distance

    # lfq.py:105-134 (3-D indices, the only form the model passes)
    def indices_to_codes(self, indices: np.ndarray, project_out=True, dtype=f32):
        if not self.keep_num_codebooks_dim:
            indices = indices[..., None]
        bits = ((indices[..., None].astype(np.int32) & self.mask.astype(np.int32)) != 0).astype(dtype)
        codes = bits * dtype(self.codebook_scale * 2) - dtype(self.codebook_scale)
        codes = codes.reshape(*codes.shape[:-2], -1)
        if project_out and self.has_projections:
            codes = self._proj(codes, self.w_out, self.b_out)
        return codes


# --------------------------------------------------------------------------------------
# entropy / perplexity terms (UT:341-410)
# --------------------------------------------------------------------------------------
def masked_mean(x, m, dim=None):
    """UT:346-353 (mult_along_first_dims UT:341-344)."""
    m = m.astype(x.dtype)
    x = x * m.reshape(m.shape + (1,) * (x.ndim - m.ndim))
    x = x / m.sum()
    return x.sum() if dim is None else x.sum(axis=dim)


def compute_entropy_loss(affinity: np.ndarray, mask: np.ndarray, temperature=0.01, eps=1e-9):
    """UT:355-387.  ``affinity`` (b, s, d, z), ``mask`` (b, s) with False at padding."""
    og = affinity.dtype
    a = affinity.astype(f32)
    b, s, d, z = a.shape
    m = mask.reshape(b * s)
    a = a.reshape(b * s, d, z)
    logits = (a / f32(temperature)) + f32(eps)
    logits = logits - logits.max(-1, keepdims=True)
    e = np.exp(logits)
    probs = e / e.sum(-1, keepdims=True)
    log_probs = logits - np.log(e.sum(-1, keepdims=True))
    avg_probs = masked_mean(probs, m, dim=0).mean(axis=0)
    avg_entropy = -1 * (avg_probs * np.log(avg_probs + f32(eps))).sum()
    sample_entropy = -1 * masked_mean((probs * log_probs).sum(-1), m)
    return (sample_entropy - avg_entropy).astype(og)


def calculate_perplexity(codes: np.ndarray, codebook_size: int, null_index=-1):
    """UT:391-410 (counts and probabilities in the integer dtype of ``codes`` are then
    promoted by torch's true division; we use float32 as torch does for int64/int)."""
    codes = codes.reshape(-1)
    codes = codes[codes != null_index]
    counts = np.bincount(codes, minlength=codebook_size).astype(np.int64)
    probs = (counts / f32(codes.size)).astype(f32)
    logits = np.zeros_like(probs)
    nz = probs != 0
    logits[nz] = np.log2(probs[nz])
    entropy = -np.sum(probs * logits)
    return f32(2) ** entropy


# --------------------------------------------------------------------------------------
# VectorQuantize, eval nearest-code path (VQ:29-33, VQ:61-83, VQ:222-226, VQ:436-507, VQ:837-1050)
# --------------------------------------------------------------------------------------
def vq_nearest(x: np.ndarray, embed: np.ndarray, chunk: int = 4096):
    """``x`` (T, d), ``embed`` (C, d).  Returns (indices int64 (T,), d2_best, d2_second) where
    d2 are float32 squared distances in the reference's operation order (VQ:29-33):
    ``x2 + y2 + (-2 * x.y)``; the reference then takes ``argmax(-sqrt(.))`` (VQ:467-469),
    i.e. the FIRST minimum in the sqrt domain."""
    x = x.astype(f32)
    embed = embed.astype(f32)
    y2 = (embed ** 2).sum(-1)
    idx = np.empty((x.shape[0],), np.int64)
    best = np.empty((x.shape[0],), f32)
    second = np.empty((x.shape[0],), f32)
    for s in range(0, x.shape[0], chunk):
        xs = x[s:s + chunk]
        x2 = (xs ** 2).sum(-1)
        d2 = (x2[:, None] + y2[None, :]) + (xs @ embed.T) * f32(-2)
        with np.errstate(invalid="ignore"):
            dist = -np.sqrt(d2)
        i = np.argmax(dist, axis=-1)       # first max, NaN wins (as torch.argmax)
        idx[s:s + chunk] = i
        part = np.partition(d2, 1, axis=-1)
        best[s:s + chunk] = d2[np.arange(len(i)), i]
        second[s:s + chunk] = np.where(part[:, 0] == best[s:s + chunk], part[:, 1], part[:, 0])
    return idx, best, second


class VectorQuantize:
    """Eval forward of VectorQuantize with an EuclideanCodebook (VQ:837-1050, VQ:436-507)."""

    def __init__(self, dim, codebook_size, codebook_dim=None, heads=1, embed=None,
                 w_in=None, b_in=None, w_out=None, b_out=None):
        self.dim, self.codebook_size, self.heads = dim, codebook_size, heads
        self.codebook_dim = dim if codebook_dim is None else codebook_dim
        self.has_projections = self.codebook_dim * heads != dim                             # VQ:729
        self.embed = embed          # (codebook_size, codebook_dim) shared codebook
        self.w_in, self.b_in, self.w_out, self.b_out = w_in, b_in, w_out, b_out

    def forward(self, x: np.ndarray, mask: Optional[np.ndarray] = None):
        orig = x
        b, n, _ = x.shape
        if self.has_projections:
            x = x @ self.w_in.T + self.b_in                                                # VQ:869
        h, d = self.heads, self.codebook_dim
        # '1 (b h) n d' fold (VQ:873-875): token order inside the codebook call is (b, h, n)
        xh = x.reshape(b, n, h, d).transpose(0, 2, 1, 3).reshape(b * h * n, d)
        ind, _, _ = vq_nearest(xh, self.embed)
        q = self.embed[ind]                                                                 # VQ:477
        ind = ind.reshape(b, h, n).transpose(0, 2, 1)                                       # VQ:994
        q = q.reshape(b, h, n, d).transpose(0, 2, 1, 3).reshape(b, n, h * d)               # VQ:1024
        if self.has_projections:
            q = q @ self.w_out.T + self.b_out                                              # VQ:1028
        if h == 1:
            ind = ind[..., 0]
        if mask is not None:
            q = np.where(mask[..., None], q, orig)                                          # VQ:1043-1048
        return q.astype(orig.dtype), ind, np.zeros((1,), f32)


# --------------------------------------------------------------------------------------
# VectorQuantize codebook learning: k-means initialisation and the EMA update
# --------------------------------------------------------------------------------------
def laplace_smoothing(x: np.ndarray, n_categories: int, eps=1e-5) -> np.ndarray:
    """VQ:100-102."""
    return (x + f32(eps)) / (x.sum(-1, keepdims=True) + f32(n_categories * eps))


def vq_kmeans(samples: np.ndarray, means0: np.ndarray, num_iters: int = 10):
    """VQ:180-220 for one codebook from given initial means (the reference draws them with ``sample_fn``):
    ``samples`` (N, d), ``means0`` (C, d) -> (means (C, d), bins (C,)).  Empty clusters keep their mean."""
    samples = samples.astype(f32)
    means = means0.astype(f32).copy()
    C = means.shape[0]
    bins = np.zeros(C, np.int64)
    for _ in range(num_iters):
        buckets, _, _ = vq_nearest(samples, means)                                          # argmax(-cdist), VQ:198-200
        bins = np.bincount(buckets, minlength=C)                                            # VQ:201
        new = np.zeros_like(means)
        np.add.at(new, buckets, samples)                                                    # scatter_add_, VQ:209
        new = new / np.maximum(bins, 1)[:, None].astype(f32)                                # VQ:205, 210
        means = np.where((bins == 0)[:, None], means, new)                                  # VQ:216-220
    return means, bins


def vq_ema_update(x: np.ndarray, ind: np.ndarray, mask: Optional[np.ndarray], cluster_size: np.ndarray,
                  embed_avg: np.ndarray, decay: float, eps: float):
    """VQ:479-500: ``x`` (T, d) tokens, ``ind`` (T,) their codes, ``mask`` (T,) True where the token counts.
    Returns the updated (embed, cluster_size, embed_avg)."""
    C = cluster_size.shape[0]
    keep = np.ones(x.shape[0], bool) if mask is None else mask.astype(bool)
    batch_n = np.bincount(ind[keep], minlength=C).astype(f32)                               # embed_onehot.sum, VQ:486
    batch_sum = np.zeros_like(embed_avg, dtype=f32)
    np.add.at(batch_sum, ind[keep], x[keep].astype(f32))                                   # einsum 'h n d, h n c -> h c d', VQ:491
    w = f32(1 - decay)
    cluster_size = cluster_size + w * (batch_n - cluster_size)                              # lerp_, VQ:38-42, 489
    embed_avg = embed_avg + w * (batch_sum - embed_avg)                                     # VQ:493
    smoothed = laplace_smoothing(cluster_size, C, eps) * cluster_size.sum(-1, keepdims=True)   # VQ:495
    embed = embed_avg / smoothed[:, None]                                                   # VQ:497-498
    return embed.astype(f32), cluster_size.astype(f32), embed_avg.astype(f32)


def vq_train_step(x: np.ndarray, mask: Optional[np.ndarray], embed: np.ndarray, cluster_size: np.ndarray,
                  embed_avg: np.ndarray, decay=0.8, eps=1e-5, commitment_weight=1.0):
    """One training forward of VectorQuantize with an EMA Euclidean codebook, one head, no projections
    (VQ:837-1050 + VQ:436-507): ``x`` (b, n, d), ``mask`` (b, n).  Returns (quantize, indices, loss, new embed,
    new cluster_size, new embed_avg); dead-code expiry (VQ:417-434) is not part of it."""
    b, n, d = x.shape
    flat = x.reshape(b * n, d).astype(f32)
    ind, _, _ = vq_nearest(flat, embed)
    q = embed[ind].astype(f32)                                                              # uses the codebook BEFORE the update
    m = None if mask is None else mask.reshape(-1)
    new_embed, new_cs, new_avg = vq_ema_update(flat, ind, m, cluster_size, embed_avg, decay, eps)
    se = (q - flat) ** 2                                                                    # F.mse_loss(..., 'none'), VQ:986
    loss = (se[m].mean() if m is not None else se.mean()) * f32(commitment_weight)          # VQ:990-994
    quant = q.reshape(b, n, d)
    if mask is not None:
        quant = np.where(mask[..., None], quant, x)                                         # VQ:1043-1048
    return quant.astype(f32), ind.reshape(b, n), np.asarray([loss], f32), new_embed, new_cs, new_avg


# --------------------------------------------------------------------------------------
# whole path, used by bench.py's cpu_baseline / --impl reference legs
# --------------------------------------------------------------------------------------
def run_pipeline(images: np.ndarray, fe: FeatureExtractor, norm: PatchNorm, lfq: LFQ):
    """preprocess -> iter_batches(None) -> PatchNorm (frozen) -> LFQ (eval) -> inverse_norm ->
    postprocess on a (B, C, H, W) float32 array; returns (list of RGB images, codes)."""
    items = [fe.preprocess(im) for im in images]
    loader = iter([{k: [it[k] for it in items] for k in items[0]}])
    batch = next(fe.iter_batches(loader, None))
    batch.patches = norm.forward(batch)
    q, codes, _, _ = lfq.forward(batch.patches, ~batch.key_pad_mask)
    batch.patches = q
    batch.patches = norm.inverse_norm(batch)
    return fe.postprocess(batch), codes
