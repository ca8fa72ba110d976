"""GPU feature extractor for the DCT autoencoder: image <-> packed rows of DCT-coefficient tokens.

Drop-in for the reference's ``DCTAutoencoderFeatureExtractor``
(feature_extraction_dct_autoencoder.py:107-656): same constructor arguments, same public methods
(``preprocess``, ``iter_batches``, ``postprocess``, ``revert_patching``), same replaceable hooks
(``_transform_image_in`` / ``_transform_image_out``), same quirks (next-fit packing, rows padded to
``max_seq_len``, tail dropping in streaming mode, single-shot ``batch_size=None``).  The work is
done by libdcta kernels; host code only decides geometry, draws ``k`` from Python's RNG in the
reference's order and lays out the packing tables.

Additions that the reference does not have (its API is per-image):
``process_batch`` (a whole same-size batch -> ``DCTPatches`` in 5 kernel launches) and
``postprocess_batch``.  They produce exactly what the per-image path followed by
``iter_batches(batch_size=None)`` / ``postprocess`` produces.
"""
import contextlib
import json
import os
from collections import OrderedDict
from typing import Any, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .dct_patches import DCTPatches
from .util import (_round8, dct2, decode_codes_inv_fold, decode_codes_inv_fold_ok, dct2_fwd_fold, dct2_fwd_fold_codes, dct2_fwd_tc, dct2_inv_fold, dct2_inv_tc, dct2_truncated,
                   exp_trunc_dist, fold_ok, idct2, idct2_truncated, ipt_to_rgb, rgb_to_ipt, rgb_to_ipt_fold,
                   rgb_to_ipt_split, tc_forward_ok, to_device_f32, to_device_pixels, unfold_ipt_to_rgb, unit_to_u8)

_SEG_DTYPE = np.dtype([("row", "<i4"), ("offset", "<i4"), ("k", "<i4"), ("image_id", "<i4"), ("img", "<i8")])
assert _SEG_DTYPE.itemsize == 24

_RESERVED_KEYS = ("patches", "positions", "channels", "original_sizes", "patch_sizes")


class _PackState:
    """Next-fit packing state carried across ``iter_batches`` steps
    (the reference's GroupPatchesState, feature_extraction_dct_autoencoder.py:96-104)."""

    def __init__(self):
        self.rows: List[List[int]] = []   # closed rows, each a list of indices into `items`
        self.row: List[int] = []          # the open row
        self.seq_len = 0


class DCTAutoencoderFeatureExtractor:
    def __init__(
        self,
        channels: int,
        patch_size: int,
        sample_patches_beta: float,
        max_patch_h: int,
        max_patch_w: int,
        max_seq_len: int,
        channel_importances: Tuple[float, ...] = (8.0, 1.0, 1.0),
        patch_sample_magnitude_weight: float = 0.1,
        device=None,
        dct_impl: str = "tc",
    ):
        """``dct_impl``: "tc" = tcgen05 split-precision GEMMs (fp32-class accuracy, images in the
        documented [0, 1] range; |IPT - plane mean| must stay below 2^7) -- folded (half the
        multiply-adds, csrc/dct_fold.cu) whenever the sizes allow it; "tc_plain" = the same without the
        fold; "fp32" = exact-fp32 FFMA GEMMs."""
        assert dct_impl in ("tc", "tc_plain", "fp32")
        self.dct_impl = dct_impl
        self.channels = channels
        self.patch_size = patch_size
        self.sample_patches_beta = sample_patches_beta
        self.max_patch_h = max_patch_h
        self.max_patch_w = max_patch_w
        self.max_seq_len = max_seq_len
        self.channel_importances = torch.Tensor(channel_importances)
        self.patch_sample_magnitude_weight = patch_sample_magnitude_weight
        self.device = torch.device(device) if device is not None else None
        # small device tables (segment lists, row bases) keyed by (device, contents), least recently used
        # evicted first; `_keepalive`, when a list, collects every table handed out (GraphedRoundtrip keeps
        # the ones its captured kernels point at alive for as long as the graph exists)
        self._table_cache: "OrderedDict[Tuple[str, bytes], torch.Tensor]" = OrderedDict()
        self._table_cache_size = 64
        self._keepalive: Optional[List[torch.Tensor]] = None
        self._maxabs: Optional[torch.Tensor] = None
        # (side stream, device, tensors it still uses) of a process_batch_to_codes(pack_stream=) call until join_pack()
        self._pack_pending = None
        # decode from codes: generate the operand of inverse pass 1 in shared memory (no coefficient planes in HBM);
        # False = the separate decode kernel followed by the plain inverse (tests compare the two bit for bit)
        self.decode_in_gemm = True

    # ------------------------------------------------------------------ save / load
    # The reference inherits transformers' FeatureExtractionMixin (FE:80); this class keeps that surface for the
    # constructor arguments without importing transformers: the same `preprocessor_config.json` (sorted keys, indent 2,
    # "feature_extractor_type"), readable by either side.
    _CONFIG_NAME = "preprocessor_config.json"
    _CONFIG_KEYS = ("channels", "patch_size", "sample_patches_beta", "max_patch_h", "max_patch_w", "max_seq_len",
                    "channel_importances", "patch_sample_magnitude_weight", "dct_impl")

    def to_dict(self) -> Dict[str, Any]:
        out = {k: getattr(self, k) for k in self._CONFIG_KEYS}
        out["channel_importances"] = [float(v) for v in self.channel_importances.tolist()]
        out["feature_extractor_type"] = "DCTAutoencoderFeatureExtractor"
        return out

    def to_json_string(self) -> str:
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"

    def save_pretrained(self, save_directory, **kwargs):
        os.makedirs(save_directory, exist_ok=True)
        path = os.path.join(save_directory, self._CONFIG_NAME)
        with open(path, "w", encoding="utf-8") as f:
            f.write(self.to_json_string())
        return [path]

    @classmethod
    def from_dict(cls, config: Dict[str, Any], **kwargs):
        cfg = {k: v for k, v in dict(config, **kwargs).items() if k in cls._CONFIG_KEYS or k == "device"}
        missing = [k for k in cls._CONFIG_KEYS[:6] if k not in cfg]
        if missing:
            raise ValueError(f"feature extractor config lacks {missing}")
        if "channel_importances" in cfg:
            cfg["channel_importances"] = tuple(float(v) for v in cfg["channel_importances"])
        return cls(**cfg)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, **kwargs):
        path = str(pretrained_model_name_or_path)
        if os.path.isdir(path):
            path = os.path.join(path, cls._CONFIG_NAME)
        if not os.path.isfile(path):
            raise EnvironmentError(f"no {cls._CONFIG_NAME} at {pretrained_model_name_or_path} (local paths only: no hub access)")
        with open(path, encoding="utf-8") as f:
            return cls.from_dict(json.load(f), **kwargs)

    def __repr__(self):
        return f"{self.__class__.__name__} {self.to_json_string()}"

    # ------------------------------------------------------------------ helpers
    def _dev(self, like: Optional[torch.Tensor] = None) -> torch.device:
        if like is not None and like.is_cuda:
            return like.device
        if self.device is not None:
            return self.device
        if not torch.cuda.is_available():
            raise _lib.DctaError("no CUDA device: the feature extractor has no CPU path")
        return torch.device("cuda", torch.cuda.current_device())

    def _hooks_overridden(self, name: str) -> bool:
        return name in self.__dict__ or getattr(type(self), name) is not getattr(DCTAutoencoderFeatureExtractor, name)

    def _get_crop_dims(self, h: int, w: int):
        """feature_extraction_dct_autoencoder.py:312-345."""
        assert h >= self.patch_size
        assert w >= self.patch_size
        p_h = max(int(h / self.patch_size), 1)
        p_w = max(int(w / self.patch_size), 1)
        return p_h * self.patch_size, p_w * self.patch_size

    def _geometry(self, h: int, w: int):
        """(ph, pw): tiles of the cropped plane (FE:166); (th, tw): tiles that can be selected
        (FE:393); the kept coefficient block is (th*p, tw*p)."""
        ch, cw = self._get_crop_dims(h, w)
        ph, pw = ch // self.patch_size, cw // self.patch_size
        return ph, pw, min(ph, self.max_patch_h), min(pw, self.max_patch_w)

    # ------------------------------------------------------------------ replaceable transforms
    @torch.no_grad()
    def _transform_image_in(self, x: torch.Tensor) -> torch.Tensor:
        """FE:130-142: RGB in [0, 1] -> IPT -> orthonormal 2-D DCT-II (computed in fp32)."""
        og = x.dtype
        y = dct2(rgb_to_ipt(to_device_f32(x, self._dev(x))), "ortho")
        return y.to(og)

    @torch.no_grad()
    def _transform_image_out(self, x: torch.Tensor) -> torch.Tensor:
        """FE:144-152."""
        og = x.dtype
        y = ipt_to_rgb(idct2(to_device_f32(x, self._dev(x)), "ortho"))
        return y.to(og)

    @torch.no_grad()
    def _crop_image(self, x: torch.Tensor) -> torch.Tensor:
        """FE:348-362."""
        c, h, w = x.shape[-3:]
        assert c == self.channels
        c_h, c_w = self._get_crop_dims(h, w)
        return x[..., :c_h, :c_w]

    # ------------------------------------------------------------------ encode: kernels
    def _token_grid(self, x: torch.Tensor, want_maxabs: bool = False):
        """(b, c, h, w) fp32 CUDA images -> token grid (b, th, tw, c, p*p) in the reference's
        pre-sort order (FE:374-399).  With ``want_maxabs`` returns (tiles, maxabs or None): the folded
        tensor-core path reduces amax|tile| (FE:409) in its GEMM epilogue."""
        if want_maxabs:
            self._maxabs = None
            tiles = self._token_grid(x)
            maxabs, self._maxabs = self._maxabs, None
            return tiles, maxabs
        b, c, h, w = x.shape
        assert c == self.channels
        p = self.patch_size
        _, _, th, tw = self._geometry(h, w)
        if not self._hooks_overridden("_transform_image_in"):
            # fused geometry: crop (FE:360) and max_patch clip (FE:393) only remove high
            # frequencies, so only the first th*p x tw*p coefficients are ever computed
            assert c == 3, "the IPT colour transform is defined for 3 channels"
            if self.dct_impl == "tc" and fold_ok(h, w, th * p, tw * p):
                # folded: the colour transform writes the four mirrored sign combinations of each plane
                hi, lo, dc = rgb_to_ipt_fold(x)
                if th > 64:       # the epilogue's per-token max image covers at most 64 tile rows
                    return dct2_fwd_fold(hi, lo, dc, th * p, tw * p, tile_p=p, channels=c, hw=(h, w))
                tiles, self._maxabs = dct2_fwd_fold(hi, lo, dc, th * p, tw * p, tile_p=p, channels=c, with_maxabs=True, hw=(h, w))
                return tiles
            if self.dct_impl in ("tc", "tc_plain") and tc_forward_ok(h, w):
                # tensor cores: colour transform writes the centred fp16 hi/lo operand planes directly
                hi, lo, dc = rgb_to_ipt_split(x)
                return dct2_fwd_tc(hi, lo, dc, th * p, tw * p, tile_p=p, channels=c)
            ipt = rgb_to_ipt(x)
            return dct2_truncated(ipt, th * p, tw * p, tile_p=p, channels=c)
        planes = torch.stack([to_device_f32(self._transform_image_in(im), x.device) for im in x])
        planes = self._crop_image(planes).contiguous()
        tiles = torch.empty((b, th, tw, c, p * p), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("dcta_patchify", _lib.ptr(planes), _lib.ptr(tiles), b, c, planes.shape[-2],
                      planes.shape[-1], th, tw, p, _lib.stream_ptr(x.device))
        return tiles

    def _sorted_order(self, tiles: torch.Tensor, maxabs: Optional[torch.Tensor] = None) -> torch.Tensor:
        """FE:403-418: importance scores and their descending order, (b, th*tw*c) int32.
        ``maxabs`` (b, th, tw, c): amax|tile| already reduced by the DCT kernel (skips the pass over the tiles)."""
        b, th, tw, c, z = tiles.shape
        n_tok = th * tw * c
        order = torch.empty((b, n_tok), dtype=torch.int32, device=tiles.device)
        imp = _lib.host_floats(self.channel_importances.tolist()[:c] + [1.0] * max(0, c - len(self.channel_importances)))
        if maxabs is not None:
            with torch.cuda.device(tiles.device):
                _lib.call("dcta_sort_tokens_maxabs", _lib.ptr(maxabs), None, _lib.ptr(order), b, th, tw, c,
                          float(self.patch_sample_magnitude_weight), imp, _lib.stream_ptr(tiles.device))
            return order
        scores = torch.empty((b, n_tok), dtype=torch.float32, device=tiles.device)
        with torch.cuda.device(tiles.device):
            st = _lib.stream_ptr(tiles.device)
            _lib.call("dcta_tile_scores", _lib.ptr(tiles), _lib.ptr(scores), b, th, tw, c, z,
                      float(self.patch_sample_magnitude_weight), imp, st)
            _lib.call("dcta_sort_tokens", _lib.ptr(scores), _lib.ptr(order), b, n_tok, st)
        return order

    def _choose_k(self, n: int) -> int:
        """FE:429-435; consumes one draw of Python's global RNG when beta > 0, like the reference."""
        k = n
        if self.sample_patches_beta > 0.0:
            k = min(round(exp_trunc_dist(self.sample_patches_beta)), k)
            k = max(1, k)
        return min(k, self.max_seq_len)

    def _next_fit(self, ks: Sequence[int], state: Optional[_PackState] = None, first: int = 0) -> _PackState:
        """FE:455-513."""
        state = state or _PackState()
        for i, k in enumerate(ks):
            assert (k <= self.max_patch_h * self.max_patch_w * self.channels and k <= self.max_seq_len), \
                f"patch with len {k} exceeds maximum sequence length"
            if state.seq_len + k > self.max_seq_len:
                state.rows.append(state.row)
                state.row = []
                state.seq_len = 0
            state.row.append(first + i)
            state.seq_len += k
        return state

    def _tables(self, rows: List[List[int]], ks: Dict[int, int], device, extra: Optional[np.ndarray] = None):
        """Segment table + row starts (+ optional pointer tables) in ONE host->device copy.
        Returns (device byte buffer, offsets dict)."""
        n_seg = sum(len(r) for r in rows)
        segs = np.zeros(n_seg, dtype=_SEG_DTYPE)
        row_start = np.zeros(len(rows) + 1, dtype=np.int32)
        i = 0
        for r, row in enumerate(rows):
            off = 0
            for image_id, img in enumerate(row):
                segs[i] = (r, off, ks[img], image_id, img)
                off += ks[img]
                i += 1
            row_start[r + 1] = i
        parts = [segs.view(np.uint8).reshape(-1), row_start.view(np.uint8).reshape(-1)]
        if extra is not None:
            parts.append(extra.view(np.uint8).reshape(-1))
        offs, blobs, cur = [], [], 0
        for part in parts:
            offs.append(cur)
            pad = (-len(part)) % 16
            blobs.append(part)
            if pad:
                blobs.append(np.zeros(pad, np.uint8))
            cur += len(part) + pad
        blob = np.concatenate(blobs) if cur else np.zeros(16, np.uint8)
        key = (str(torch.device(device)), blob.tobytes()) if extra is None else None
        return self._cached_table(key, blob, device), offs

    def _cached_table(self, key, host: np.ndarray, device) -> torch.Tensor:
        """Device copy of a small host table, cached per (device, contents) with LRU eviction."""
        dev = self._table_cache.get(key) if key is not None else None
        if dev is None:
            dev = torch.from_numpy(host).pin_memory().to(device, non_blocking=True)
            if key is not None:
                self._table_cache[key] = dev
                while len(self._table_cache) > self._table_cache_size:
                    self._table_cache.popitem(last=False)
        else:
            self._table_cache.move_to_end(key)
        if self._keepalive is not None:
            self._keepalive.append(dev)
        return dev

    def _check_ks(self, ks: Sequence[int], b: int, n_tok: int) -> List[int]:
        """Caller-supplied token counts: one per image, 1 <= k <= min(tokens of the image, max_seq_len)
        (FE:429-435 never produces anything else; the pack kernels index `order[img, :k]`)."""
        ks = [int(k) for k in ks]
        assert len(ks) == b, f"{len(ks)} token counts for {b} images"
        lim = min(n_tok, self.max_seq_len)
        assert all(1 <= k <= lim for k in ks), f"token counts must lie in [1, {lim}]"
        return ks

    def _alloc_batch(self, n_rows: int, device):
        s, z = self.max_seq_len, self.patch_size ** 2
        return (torch.empty((n_rows, s, z), dtype=torch.float32, device=device),
                torch.empty((n_rows, s, 2), dtype=torch.int64, device=device),
                torch.empty((n_rows, s), dtype=torch.int64, device=device),
                torch.empty((n_rows, s), dtype=torch.int64, device=device),
                torch.empty((n_rows, s), dtype=torch.bool, device=device))

    # ------------------------------------------------------------------ encode: public
    @torch.no_grad()
    def preprocess(self, im: torch.Tensor):
        """FE:155-177: one (c, h, w) image -> dict(patches (k, p*p), positions (k, 2),
        channels (k,), original_sizes (h, w), patch_sizes (ph, pw))."""
        og = im.dtype if im.dtype != torch.uint8 else torch.float32     # uint8 pixels stand for x / 255 (fp32)
        x = to_device_pixels(im, self._dev(im))[None]
        _, c, h, w = x.shape
        ph, pw, th, tw = self._geometry(h, w)
        tiles, maxabs = self._token_grid(x, want_maxabs=True)
        order = self._sorted_order(tiles, maxabs)
        n_tok = th * tw * c
        k = self._choose_k(n_tok)
        z = self.patch_size ** 2
        patches = torch.empty((1, k, z), dtype=torch.float32, device=x.device)
        pos = torch.empty((1, k, 2), dtype=torch.int64, device=x.device)
        chan = torch.empty((1, k), dtype=torch.int64, device=x.device)
        tab, offs = self._tables([[0]], {0: k}, x.device)
        with torch.cuda.device(x.device):
            _lib.call("dcta_pack_tiles", _lib.ptr(tiles), _lib.ptr(order), tab.data_ptr() + offs[0],
                      tab.data_ptr() + offs[1], 1, k, th, tw, c, z, _lib.ptr(patches), _lib.ptr(pos),
                      _lib.ptr(chan), None, None, _lib.stream_ptr(x.device))
        patches = patches[0] if og == torch.float32 else patches[0].to(og)
        return dict(patches=patches, positions=pos[0], channels=chan[0],
                    original_sizes=(h, w), patch_sizes=(ph, pw))

    @torch.no_grad()
    def _preprocess_batch_raw(self, images: torch.Tensor, ks: Optional[Sequence[int]] = None):
        """Tokens of b same-size images, one image per row of (b, max k, .) buffers.
        -> (patches fp32, positions, channels, ks, (h, w), (ph, pw))."""
        x = to_device_pixels(images, self._dev(images))
        b, c, h, w = x.shape
        ph, pw, th, tw = self._geometry(h, w)
        tiles, maxabs = self._token_grid(x, want_maxabs=True)
        order = self._sorted_order(tiles, maxabs)
        n_tok = th * tw * c
        if ks is None:
            ks = [self._choose_k(n_tok) for _ in range(b)]
        ks = self._check_ks(ks, b, n_tok)
        kmax, z = max(ks), self.patch_size ** 2
        patches = torch.empty((b, kmax, z), dtype=torch.float32, device=x.device)
        pos = torch.empty((b, kmax, 2), dtype=torch.int64, device=x.device)
        chan = torch.empty((b, kmax), dtype=torch.int64, device=x.device)
        tab, offs = self._tables([[i] for i in range(b)], dict(enumerate(ks)), x.device)
        with torch.cuda.device(x.device):
            _lib.call("dcta_pack_tiles", _lib.ptr(tiles), _lib.ptr(order), tab.data_ptr() + offs[0],
                      tab.data_ptr() + offs[1], b, kmax, th, tw, c, z, _lib.ptr(patches), _lib.ptr(pos),
                      _lib.ptr(chan), None, None, _lib.stream_ptr(x.device))
        return patches, pos, chan, ks, (h, w), (ph, pw)

    @torch.no_grad()
    def preprocess_batch(self, images: torch.Tensor, ks: Optional[Sequence[int]] = None) -> List[dict]:
        """``[preprocess(im) for im in images]`` for (b, c, h, w) same-size images in one set of
        launches (the batched multi-image preprocess the offline shard writer needs,
        preproc_dataset.py:62-84).  The per-image tensors are views of one (b, max k, .) buffer."""
        patches, pos, chan, ks, osz, psz = self._preprocess_batch_raw(images, ks)
        if images.dtype not in (torch.float32, torch.uint8):
            patches = patches.to(images.dtype)
        return [dict(patches=patches[i, :k], positions=pos[i, :k], channels=chan[i, :k],
                     original_sizes=osz, patch_sizes=psz) for i, k in enumerate(ks)]

    @torch.no_grad()
    def process_batch(self, images: torch.Tensor, ks: Optional[Sequence[int]] = None) -> DCTPatches:
        """Whole-batch encode of (b, c, h, w) same-size images: equals
        ``next(iter_batches(iter([dict_collate([preprocess(im) for im in images])]), None))``."""
        x = to_device_pixels(images, self._dev(images))
        b, c, h, w = x.shape
        ph, pw, th, tw = self._geometry(h, w)
        tiles, maxabs = self._token_grid(x, want_maxabs=True)
        order = self._sorted_order(tiles, maxabs)
        n_tok = th * tw * c
        if ks is None:
            ks = [self._choose_k(n_tok) for _ in range(b)]
        ks = self._check_ks(ks, b, n_tok)
        state = self._next_fit(ks)
        rows = state.rows + ([state.row] if state.row else [])
        tab, offs = self._tables(rows, dict(enumerate(ks)), x.device)
        patches, pos, chan, ids, pad = self._alloc_batch(len(rows), x.device)
        with torch.cuda.device(x.device):
            _lib.call("dcta_pack_tiles", _lib.ptr(tiles), _lib.ptr(order), tab.data_ptr() + offs[0],
                      tab.data_ptr() + offs[1], len(rows), self.max_seq_len, th, tw, c,
                      self.patch_size ** 2, _lib.ptr(patches), _lib.ptr(pos), _lib.ptr(chan),
                      _lib.ptr(ids), _lib.ptr(pad), _lib.stream_ptr(x.device))
        if images.dtype not in (torch.float32, torch.uint8):
            patches = patches.to(images.dtype)
        return DCTPatches(patches=patches, key_pad_mask=pad, batched_image_ids=ids,
                          patch_channels=chan, patch_positions=pos,
                          patch_sizes=[(ph, pw)] * b, original_sizes=[(h, w)] * b,
                          _data={}, _row_num_images=[len(r) for r in rows])

    @torch.no_grad()
    def _process_batch_index(self, images: torch.Tensor, ks: Optional[Sequence[int]] = None):
        """``process_batch`` without the packed patches: (DCTPatches with ``patches=None``, token grid (tokens, p*p),
        row_src (rows, s) int32 = the token-grid row of every slot, -1 for padding).  For consumers that stream the token
        rows anyway and gather them themselves (the operand split of the LFQ projection)."""
        x = to_device_pixels(images, self._dev(images))
        b, c, h, w = x.shape
        ph, pw, th, tw = self._geometry(h, w)
        tiles, maxabs = self._token_grid(x, want_maxabs=True)
        order = self._sorted_order(tiles, maxabs)
        n_tok = th * tw * c
        if ks is None:
            ks = [self._choose_k(n_tok) for _ in range(b)]
        ks = self._check_ks(ks, b, n_tok)
        state = self._next_fit(ks)
        rows = state.rows + ([state.row] if state.row else [])
        tab, offs = self._tables(rows, dict(enumerate(ks)), x.device)
        n_rows, s = len(rows), self.max_seq_len
        pos = torch.empty((n_rows, s, 2), dtype=torch.int64, device=x.device)
        chan = torch.empty((n_rows, s), dtype=torch.int64, device=x.device)
        ids = torch.empty((n_rows, s), dtype=torch.int64, device=x.device)
        pad = torch.empty((n_rows, s), dtype=torch.bool, device=x.device)
        row_src = torch.empty((n_rows, s), dtype=torch.int32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("dcta_pack_tiles_index", _lib.ptr(order), tab.data_ptr() + offs[0], tab.data_ptr() + offs[1], n_rows, s,
                      th, tw, c, b, _lib.ptr(pos), _lib.ptr(chan), _lib.ptr(ids), _lib.ptr(pad), _lib.ptr(row_src),
                      _lib.stream_ptr(x.device))
        batch = DCTPatches(patches=None, key_pad_mask=pad, batched_image_ids=ids, patch_channels=chan, patch_positions=pos,
                           patch_sizes=[(ph, pw)] * b, original_sizes=[(h, w)] * b, _data={},
                           _row_num_images=[len(r) for r in rows])
        return batch, tiles.reshape(-1, self.patch_size ** 2), row_src

    # ------------------------------------------------------------------ fused PatchNorm + LFQ path
    def _lfq_fusable(self, norm, lfq) -> bool:
        """Projection-free LFQ in eval mode on frozen (or eval) fp32 PatchNorm tables that were built for
        this extractor's token geometry."""
        # keep_num_codebooks_dim False (a single codebook): the staged module returns (rows, s) indices, the fused kernels
        # (rows, s, 1) -- not fused, so that the output shape never depends on which path ran
        return (getattr(lfq, "has_projections", True) is False and not lfq.training
                and getattr(lfq, "keep_num_codebooks_dim", True)
                and (norm.frozen or not norm.training)
                and norm.channels == self.channels and norm.patch_size == self.patch_size
                and lfq.num_codebooks * lfq.codebook_dim == norm.patch_size ** 2 <= 256
                and norm.median.dtype == torch.float32 and norm.median.is_cuda
                and norm.b.dtype == torch.float32 and norm.b.device == norm.median.device)

    @torch.no_grad()
    def process_batch_to_codes(self, images: torch.Tensor, norm, lfq, ks: Optional[Sequence[int]] = None,
                               return_grid: bool = False, pack_stream: Optional["torch.cuda.Stream"] = None):
        """``process_batch`` -> ``norm(batch)`` -> ``lfq(., mask)`` with the three intermediate patch
        tensors kept in registers (csrc/fused_lfq.cu).  Returns (DCTPatches with ``patches=None``,
        codes (rows, s, codebooks) int64), bit-identical to the staged calls.
        ``return_grid``: also return the code grid (b, th, tw, c, p) int32 the forward pass wrote when EVERY token of every
        image was kept (else None): ``postprocess_codes_batch(..., code_grid=)`` then decodes without a slot map.
        ``pack_stream``: when every token was kept, the sort and the gather into the packed API tensors are launched on
        that stream (the decode that follows only needs the code grid, so both run beside it); the caller MUST call
        ``join_pack()`` on the stream it uses before touching ``codes`` or the batch's tensors."""
        assert self._lfq_fusable(norm, lfq)
        self.join_pack()          # a caller that forgot: never drop the references of work still running on a side stream
        x = to_device_pixels(images, self._dev(images), keep_u8=True)     # uint8 pixels stay bytes until the colour kernel
        if norm.median.device != x.device:
            raise _lib.DctaError(f"PatchNorm tables live on {norm.median.device}, the images on {x.device}")
        b, c, h, w = x.shape
        ph, pw, th, tw = self._geometry(h, w)
        p = self.patch_size
        n_tok = th * tw * c
        # one codebook per patch row on the folded tensor-core path: the DCT epilogue emits the code words itself
        in_epilogue = (self.dct_impl == "tc" and c == 3 and not self._hooks_overridden("_transform_image_in")
                       and lfq.num_codebooks == p and lfq.codebook_dim == p
                       and th <= norm.max_patch_h and tw <= norm.max_patch_w
                       and bool(_lib.load().dcta_fold_codes_supported(h, w, th * p, tw * p, p)))
        if in_epilogue:
            hi, lo, dc = rgb_to_ipt_fold(x)
            maxabs, code_grid = dct2_fwd_fold_codes(hi, lo, dc, th * p, tw * p, p, c, norm, hw=(h, w))
            del hi, lo
        else:
            if x.dtype == torch.uint8:
                x = to_device_pixels(x)
            tiles, maxabs = self._token_grid(x, want_maxabs=True)
            order = self._sorted_order(tiles, maxabs)
        if ks is None:
            ks = [self._choose_k(n_tok) for _ in range(b)]
        ks = self._check_ks(ks, b, n_tok)
        all_kept = in_epilogue and all(k == n_tok for k in ks)
        state = self._next_fit(ks)
        rows = state.rows + ([state.row] if state.row else [])
        tab, offs = self._tables(rows, dict(enumerate(ks)), x.device)
        n_rows, s = len(rows), self.max_seq_len
        codes = torch.empty((n_rows, s, lfq.num_codebooks), dtype=torch.int64, device=x.device)
        pos = torch.empty((n_rows, s, 2), dtype=torch.int64, device=x.device)
        chan = torch.empty((n_rows, s), dtype=torch.int64, device=x.device)
        ids = torch.empty((n_rows, s), dtype=torch.int64, device=x.device)
        pad = torch.empty((n_rows, s), dtype=torch.bool, device=x.device)
        if in_epilogue:
            order = torch.empty((b, n_tok), dtype=torch.int32, device=x.device)
            pad_codes = torch.empty(lfq.num_codebooks, dtype=torch.int64, device=x.device)
        # Every tensor above is allocated on the caller's stream; with a pack stream only the LAUNCHES of the
        # sort and the gather move over, after an event on the caller's stream, and everything they touch stays
        # referenced until join_pack() -- so no block can be handed out again while the side stream still uses it.
        side = pack_stream if (pack_stream is not None and all_kept and not _lib.profile_active()) else None
        if side is not None:
            side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.device(x.device), (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            if in_epilogue:
                imp = _lib.host_floats(self.channel_importances.tolist()[:c] + [1.0] * max(0, c - len(self.channel_importances)))
                _lib.call("dcta_sort_tokens_maxabs", _lib.ptr(maxabs), None, _lib.ptr(order), b, th, tw, c,
                          float(self.patch_sample_magnitude_weight), imp, _lib.stream_ptr(x.device))
                _lib.call("dcta_pack_codes_grid", _lib.ptr(code_grid), _lib.ptr(order), tab.data_ptr() + offs[0],
                          tab.data_ptr() + offs[1], n_rows, s, th, tw, c, _lib.ptr(norm.median.data), _lib.ptr(norm.b.data),
                          norm.max_patch_h, norm.max_patch_w, float(norm.eps), float(norm.min_val), float(norm.max_val),
                          lfq.num_codebooks, lfq.codebook_dim, _lib.ptr(pad_codes), _lib.ptr(codes), _lib.ptr(pos),
                          _lib.ptr(chan), _lib.ptr(ids), _lib.ptr(pad), _lib.stream_ptr(x.device))
            else:
                tame = torch.empty(1, dtype=torch.int32, device=x.device)
                _lib.call("dcta_pack_codes_lfq", _lib.ptr(tiles), _lib.ptr(order), tab.data_ptr() + offs[0],
                          tab.data_ptr() + offs[1], n_rows, s, th, tw, c, self.patch_size ** 2,
                          _lib.ptr(norm.median.data), _lib.ptr(norm.b.data), norm.max_patch_h, norm.max_patch_w,
                          float(norm.eps), float(norm.min_val), float(norm.max_val), lfq.num_codebooks,
                          lfq.codebook_dim, float(lfq.codebook_scale), _lib.ptr(tame), _lib.ptr(codes), _lib.ptr(pos),
                          _lib.ptr(chan), _lib.ptr(ids), _lib.ptr(pad), _lib.stream_ptr(x.device))
        if side is not None:
            self._pack_pending = (side, x.device, [maxabs, code_grid, order, pad_codes, tab, codes, pos, chan, ids, pad])
        batch = DCTPatches(patches=None, key_pad_mask=pad, batched_image_ids=ids, patch_channels=chan,
                           patch_positions=pos, patch_sizes=[(ph, pw)] * b, original_sizes=[(h, w)] * b,
                           _data={}, _row_num_images=[len(r) for r in rows])
        if not lfq.keep_num_codebooks_dim:       # lfq.py:224-225: a single codebook loses its axis
            codes = codes[..., 0]
        if return_grid:
            return batch, codes, (code_grid if all_kept else None)
        return batch, codes

    def join_pack(self):
        """Make the current stream wait for the sort + gather a ``process_batch_to_codes(pack_stream=)`` call left on its
        side stream, and release what that work was reading."""
        pending, self._pack_pending = self._pack_pending, None
        if pending is not None:
            side, dev, _keep = pending
            torch.cuda.current_stream(dev).wait_stream(side)

    @torch.no_grad()
    def postprocess_codes_batch(self, x: DCTPatches, codes: torch.Tensor, norm, lfq, out_dtype=torch.float32,
                                code_grid: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``lfq.indices_to_codes`` -> ``norm.inverse_norm`` -> ``postprocess_batch`` with the
        de-quantised patches kept in registers; same-size batches, tensor-core DCT path.
        ``out_dtype=torch.uint8``: 8-bit pixels, quantised like torchvision's save_image (util.unit_to_u8)."""
        assert self._lfq_fusable(norm, lfq) and self.dct_impl in ("tc", "tc_plain")
        assert len(set(map(tuple, x.original_sizes))) == 1 and len(set(map(tuple, x.patch_sizes))) == 1
        if not lfq.keep_num_codebooks_dim:       # lfq.py:106-107
            codes = codes[..., None]
        if norm.median.device != codes.device:
            raise _lib.DctaError(f"PatchNorm tables live on {norm.median.device}, the codes on {codes.device}")
        (idx, rgb), = self._decode_groups(x, codes=codes.contiguous(), norm=norm, lfq=lfq, out_dtype=out_dtype,
                                          code_grid=code_grid)
        return rgb

    def _group_patches_by_max_seq_len(self, batched_patches, batched_positions=None,
                                      batched_channels=None, state: Optional[_PackState] = None,
                                      first: int = 0) -> _PackState:
        """FE:455-513 (only the token counts matter for the grouping)."""
        return self._next_fit([p.shape[0] for p in batched_patches], state, first)

    @torch.no_grad()
    def _batch_groups(self, rows: List[List[int]], items: Dict[int, dict], **dct_patch_kwargs) -> DCTPatches:
        """FE:516-605 + util.py:149-164: concatenate each row's per-image token lists, right-pad
        to ``max_seq_len``, build masks/ids -- one kernel over pointer tables."""
        order = [i for row in rows for i in row]
        device = self._dev(items[order[0]]["patches"]) if order else self._dev()
        src = {}
        for i in order:
            it = items[i]
            pt = to_device_f32(it["patches"], device)
            if pt.data_ptr() % 16:
                pt = pt.clone()
            src[i] = (pt, it["positions"].to(device, torch.int64).contiguous(),
                      it["channels"].to(device, torch.int64).contiguous())
        local = {g: j for j, g in enumerate(order)}
        ptrs = np.array([[src[g][0].data_ptr() for g in order], [src[g][1].data_ptr() for g in order],
                         [src[g][2].data_ptr() for g in order]], dtype=np.int64).reshape(3, -1)
        rows_local = [[local[g] for g in row] for row in rows]
        ks = {local[g]: src[g][0].shape[0] for g in order}
        tab, offs = self._tables(rows_local, ks, device, extra=ptrs)
        patches, pos, chan, ids, pad = self._alloc_batch(len(rows), device)
        n = len(order)
        base = tab.data_ptr() + offs[2]
        with torch.cuda.device(device):
            _lib.call("dcta_pack_lists", base, base + 8 * n, base + 16 * n, tab.data_ptr() + offs[0],
                      tab.data_ptr() + offs[1], len(rows), self.max_seq_len, self.patch_size ** 2,
                      _lib.ptr(patches), _lib.ptr(pos), _lib.ptr(chan), _lib.ptr(ids), _lib.ptr(pad),
                      _lib.stream_ptr(device))
        # `src` tensors are freed in stream order on the launching stream, after the kernel
        dt = items[order[0]]["patches"].dtype if order else torch.float32
        if dt != torch.float32:
            patches = patches.to(dt)
        return DCTPatches(patches=patches, key_pad_mask=pad, batched_image_ids=ids,
                          patch_channels=chan, patch_positions=pos,
                          _row_num_images=[len(r) for r in rows], **dct_patch_kwargs)

    @torch.no_grad()
    def iter_batches(self, dataloader: Iterator[dict], batch_size: Optional[int] = None):
        """FE:180-287.  ``dataloader`` yields dicts of lists (``dict_collate``, dataset.py:8-15)
        with keys patches, positions, channels, original_sizes, patch_sizes (+ passthrough keys).

        Streaming mode (``batch_size`` given): a batch of exactly ``batch_size`` rows is emitted
        whenever MORE than ``batch_size`` rows are closed (FE:233); the tail is never emitted.
        ``batch_size=None``: everything seen so far, including the open row, is emitted at once;
        like the reference this mode is single-shot (every caller takes ``next(iter(...))``)."""
        dataloader = iter(dataloader)
        state = _PackState()
        items: Dict[int, dict] = {}
        n_seen = 0
        n_emitted = 0
        while True:
            try:
                d = next(dataloader)
            except StopIteration:
                return
            n = len(d["patches"])
            for i in range(n):
                items[n_seen + i] = {k: v[i] for k, v in d.items()}
            state = self._group_patches_by_max_seq_len(d["patches"], state=state, first=n_seen)
            n_seen += n
            if batch_size is None and state.row:                      # FE:220-229
                state.rows.append(state.row)
                state.row, state.seq_len = [], 0
            if batch_size is None or len(state.rows) > batch_size:    # FE:233
                emit = state.rows[:batch_size]
                state.rows = state.rows[batch_size:] if batch_size is not None else state.rows
                idx = [i for row in emit for i in row]
                assert idx == list(range(n_emitted, n_emitted + len(idx)))
                misc = {}
                for i in idx:
                    for k, v in items[i].items():
                        if k not in _RESERVED_KEYS:
                            misc.setdefault(k, []).append(v)
                batch = self._batch_groups(
                    emit, items,
                    original_sizes=[tuple(items[i]["original_sizes"]) for i in idx],
                    patch_sizes=[tuple(items[i]["patch_sizes"]) for i in idx],
                    _data=misc)
                for i in idx:
                    del items[i]
                n_emitted += len(idx)
                yield batch
                if batch_size is None:
                    return

    # ------------------------------------------------------------------ decode
    def _slot_map(self, x: DCTPatches, th: int, tw: int):
        b, s = x.key_pad_mask.shape
        counts = x.row_num_images()
        n_img = int(sum(counts))
        base = np.zeros(b, dtype=np.int32)
        if b > 1:
            base[1:] = np.cumsum(counts[:-1])
        dev = x.key_pad_mask.device
        base_dev = self._cached_table((str(dev), b"base" + base.tobytes()), base, dev)
        slot_map = torch.empty((n_img, self.channels, th, tw), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("dcta_build_slot_map", _lib.ptr(x.patch_channels), _lib.ptr(x.patch_positions),
                      _lib.ptr(x.batched_image_ids), _lib.ptr(x.key_pad_mask), _lib.ptr(base_dev), b, s,
                      n_img, self.channels, th, tw, _lib.ptr(slot_map), _lib.stream_ptr(dev))
        return slot_map, n_img

    def _render_planes(self, x: DCTPatches, clip: bool):
        """Token rows -> coefficient planes, grouped by plane size.  Returns
        [(image indices, planes (n, c, rows, cols))]."""
        _lib.require_cuda(x.patches, x.patch_positions, x.patch_channels, x.batched_image_ids, x.key_pad_mask)
        p = self.patch_size
        patches = to_device_f32(x.patches)
        sizes = [tuple(int(v) for v in ps) for ps in x.patch_sizes]
        if clip:
            sizes = [(min(a, self.max_patch_h), min(b, self.max_patch_w)) for a, b in sizes]
        th = max(a for a, _ in sizes)
        tw = max(b for _, b in sizes)
        slot_map, n_img = self._slot_map(x, th, tw)
        assert n_img == len(sizes), f"{n_img} images in the rows but {len(sizes)} patch_sizes"
        groups: Dict[Tuple[int, int], List[int]] = {}
        for i, sz in enumerate(sizes):
            groups.setdefault(sz, []).append(i)
        out = []
        dev = patches.device
        for (gh, gw), idx in groups.items():
            planes = torch.empty((len(idx), self.channels, gh * p, gw * p), dtype=torch.float32, device=dev)
            sel = None
            if len(groups) > 1 or idx != list(range(n_img)):
                sel = torch.tensor(idx, dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
            with torch.cuda.device(dev):
                _lib.call("dcta_unpatchify", _lib.ptr(patches), _lib.ptr(slot_map), _lib.ptr(sel), len(idx),
                          self.channels, th, tw, p, gh * p, gw * p, _lib.ptr(planes), _lib.stream_ptr(dev))
            out.append((idx, planes))
        return out

    @torch.no_grad()
    def revert_patching(self, output: DCTPatches) -> List[torch.Tensor]:
        """FE:607-656: one (c, ph*p, pw*p) coefficient plane per image, zeros where no token."""
        res: List[Optional[torch.Tensor]] = [None] * len(output.patch_sizes)
        for idx, planes in self._render_planes(output, clip=False):
            for j, i in enumerate(idx):
                res[i] = planes[j] if output.patches.dtype == torch.float32 else planes[j].to(output.patches.dtype)
        return res

    @torch.no_grad()
    def postprocess(self, x: DCTPatches) -> List[torch.Tensor]:
        """FE:289-310: un-normalised ``DCTPatches`` -> list of (c, h, w) RGB images."""
        og = x.patches.dtype
        res: List[Optional[torch.Tensor]] = [None] * len(x.patch_sizes)
        if self._hooks_overridden("_transform_image_out"):
            for i, (plane, (h, w)) in enumerate(zip(self.revert_patching(x), x.original_sizes)):
                ch, cw = plane.shape[-2:]
                pad = torch.zeros(self.channels, h, w, device=plane.device, dtype=plane.dtype)
                pad[:, :ch, :cw] = plane
                res[i] = self._transform_image_out(pad)
            return res
        for idx, rgb in self._decode_groups(x):
            for n, i in enumerate(idx):
                res[i] = rgb[n] if og == torch.float32 else rgb[n].to(og)
        return res

    def _decode_groups(self, x: DCTPatches, codes: Optional[torch.Tensor] = None, norm=None, lfq=None,
                       out_dtype=torch.float32, code_grid: Optional[torch.Tensor] = None, denorm=None):
        """Token rows -> RGB, one kernel sequence per group of images that share (tile grid,
        original size).  Yields (image indices, rgb (n, c, h, w)).  The zero padding of FE:300-304
        is implicit in the truncated inverse basis.  With ``codes`` the tokens are de-quantised and
        de-normalised on the fly (fused LFQ + PatchNorm decode) instead of being read from ``patches``."""
        _lib.require_cuda(x.patch_positions, x.patch_channels, x.batched_image_ids, x.key_pad_mask)
        p, C = self.patch_size, self.channels
        if codes is None:
            _lib.require_cuda(x.patches)
            patches = to_device_f32(x.patches)
            dev = patches.device
            if denorm is not None:
                sizes = {(min(int(a), self.max_patch_h), min(int(b), self.max_patch_w), int(o[0]), int(o[1]))
                         for (a, b), o in zip(x.patch_sizes, x.original_sizes)}
                fusable = (self.dct_impl == "tc" and C == 3 and p >= 8 and p % 2 == 0 and denorm.median.device == dev
                           and denorm.median.dtype == torch.float32
                           and denorm.patch_size == p and denorm.channels == C
                           and all(a <= denorm.max_patch_h and b <= denorm.max_patch_w and fold_ok(h, w, a * p, b * p)
                                   for a, b, h, w in sizes))
                if not fusable:              # the plain order of the reference: de-normalise, then un-patchify
                    y = x.shallow_copy()
                    y.patches = patches
                    patches = denorm.inverse_norm(y)
                    denorm = None
        else:
            patches = None
            dev = codes.device
        tiles = [(min(int(a), self.max_patch_h), min(int(b), self.max_patch_w)) for a, b in x.patch_sizes]
        th, tw = max(a for a, _ in tiles), max(b for _, b in tiles)
        if code_grid is not None:
            # a same-size batch that kept every token, straight from the forward pass's code grid (roundtrip): the sign
            # bits are read in grid order, no slot map
            gh, gw = tiles[0]
            h, w = (int(v) for v in x.original_sizes[0])
            n, kh, kw = len(tiles), gh * p, gw * p
            if (self.dct_impl == "tc" and C == 3 and self.decode_in_gemm and len(set(tiles)) == 1
                    and len(set(map(tuple, x.original_sizes))) == 1 and tuple(code_grid.shape) == (n, gh, gw, C, p)
                    and decode_codes_inv_fold_ok(h, w, kh, kw, p, lfq.num_codebooks, lfq.codebook_dim)):
                z, dc = decode_codes_inv_fold(None, None, None, n, C, gh, gw, p, kh, kw, h, w, norm, lfq.num_codebooks,
                                              lfq.codebook_dim, lfq.codebook_scale, code_grid=code_grid)
                yield list(range(n)), unfold_ipt_to_rgb(z, dc, h, w, out_dtype)
                return
        self.join_pack()          # everything below reads the packed tensors (a no-op unless a pack stream is in use)
        slot_map, n_img = self._slot_map(x, th, tw)
        assert n_img == len(tiles), f"{n_img} images in the rows but {len(tiles)} patch_sizes"
        groups: Dict[Tuple[int, int, int, int], List[int]] = {}
        for i, (t, o) in enumerate(zip(tiles, x.original_sizes)):
            groups.setdefault((t[0], t[1], int(o[0]), int(o[1])), []).append(i)
        for (gh, gw, h, w), idx in groups.items():
            sel = None
            if len(groups) > 1:
                sel = torch.tensor(idx, dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
            n, kh, kw = len(idx), gh * p, gw * p
            with torch.cuda.device(dev):
                st = _lib.stream_ptr(dev)
                if (self.dct_impl == "tc" and C == 3 and codes is not None and self.decode_in_gemm and
                        decode_codes_inv_fold_ok(h, w, kh, kw, p, lfq.num_codebooks, lfq.codebook_dim)):
                    z, dc = decode_codes_inv_fold(codes, slot_map, sel, n, C, th, tw, p, kh, kw, h, w, norm,
                                                  lfq.num_codebooks, lfq.codebook_dim, lfq.codebook_scale)
                    yield idx, unfold_ipt_to_rgb(z, dc, h, w, out_dtype)
                    continue
                if self.dct_impl == "tc" and C == 3 and p >= 8 and fold_ok(h, w, kh, kw):
                    ldq = _round8(kw // 2)
                    y_hi = torch.empty((2, 2, n * C, kh // 2, ldq), dtype=torch.float16, device=dev)
                    y_lo = torch.empty_like(y_hi)
                    dc = torch.empty(n * C, dtype=torch.float32, device=dev)
                    if codes is not None:
                        n_tab = C * norm.max_patch_h * p * ((norm.max_patch_w * p + 7) // 8) * 8
                        tab = torch.empty(2 * n_tab, dtype=torch.int32, device=dev)
                        _lib.call("dcta_decode_codes_fold", _lib.ptr(codes), _lib.ptr(slot_map), _lib.ptr(sel), n, C,
                                  th, tw, p, kh, kw, h, w, _lib.ptr(norm.median.data), _lib.ptr(norm.b.data),
                                  norm.max_patch_h, norm.max_patch_w, float(norm.eps), lfq.num_codebooks,
                                  lfq.codebook_dim, float(lfq.codebook_scale), _lib.ptr(y_hi), _lib.ptr(y_lo),
                                  _lib.ptr(dc), _lib.ptr(tab), st)
                    elif denorm is not None:
                        _lib.call("dcta_unpatchify_denorm_fold", _lib.ptr(patches), _lib.ptr(slot_map), _lib.ptr(sel), n, C,
                                  th, tw, p, kh, kw, h, w, _lib.ptr(denorm.median.data), _lib.ptr(denorm.b.data),
                                  denorm.max_patch_h, denorm.max_patch_w, float(denorm.eps), _lib.ptr(y_hi), _lib.ptr(y_lo),
                                  _lib.ptr(dc), st)
                    else:
                        _lib.call("dcta_unpatchify_fold", _lib.ptr(patches), _lib.ptr(slot_map), _lib.ptr(sel), n, C,
                                  th, tw, p, kh, kw, h, w, _lib.ptr(y_hi), _lib.ptr(y_lo), _lib.ptr(dc), st)
                    yield idx, unfold_ipt_to_rgb(dct2_inv_fold(y_hi, y_lo, kh, kw, h, w), dc, h, w, out_dtype)
                    continue
                if self.dct_impl in ("tc", "tc_plain"):
                    ld = _round8(kw)
                    y_hi = torch.empty((n, C, kh, ld), dtype=torch.float16, device=dev)
                    y_lo = torch.empty_like(y_hi)
                    dc = torch.empty(n * C, dtype=torch.float32, device=dev)
                    if codes is not None:
                        _lib.call("dcta_decode_codes_split", _lib.ptr(codes), _lib.ptr(slot_map), _lib.ptr(sel), n, C,
                                  th, tw, p, kh, kw, ld, h, w, _lib.ptr(norm.median.data), _lib.ptr(norm.b.data),
                                  norm.max_patch_h, norm.max_patch_w, float(norm.eps), lfq.num_codebooks,
                                  lfq.codebook_dim, float(lfq.codebook_scale), _lib.ptr(y_hi), _lib.ptr(y_lo),
                                  _lib.ptr(dc), st)
                    else:
                        _lib.call("dcta_unpatchify_split", _lib.ptr(patches), _lib.ptr(slot_map), _lib.ptr(sel), n, C,
                                  th, tw, p, kh, kw, ld, h, w, _lib.ptr(y_hi), _lib.ptr(y_lo), _lib.ptr(dc), st)
                    ipt = dct2_inv_tc(y_hi, y_lo, dc, kw, h, w)
                else:
                    planes = torch.empty((n, C, kh, kw), dtype=torch.float32, device=dev)
                    _lib.call("dcta_unpatchify", _lib.ptr(patches), _lib.ptr(slot_map), _lib.ptr(sel), n, C, th, tw,
                              p, kh, kw, _lib.ptr(planes), st)
                    ipt = idct2_truncated(planes, h, w)
            rgb = ipt_to_rgb(ipt)
            yield idx, (unit_to_u8(rgb) if out_dtype == torch.uint8 else rgb)

    @torch.no_grad()
    def postprocess_batch(self, x: DCTPatches, out_dtype=torch.float32, denorm=None) -> torch.Tensor:
        """Same as ``torch.stack(postprocess(x))`` for batches whose images share one size.
        ``out_dtype=torch.uint8``: 8-bit pixels, quantised like torchvision's save_image (util.unit_to_u8).
        ``denorm``: a PatchNorm whose ``inverse_norm`` (patchnorm.py:167-177) is applied to ``x.patches`` on the way into
        the coefficient planes (the un-patchify kernel de-normalises what it gathers; same values as
        ``x.patches = denorm.inverse_norm(x)`` first, without writing the de-normalised patches)."""
        assert len(set(map(tuple, x.original_sizes))) == 1 and len(set(map(tuple, x.patch_sizes))) == 1
        (idx, rgb), = self._decode_groups(x, out_dtype=out_dtype, denorm=denorm)
        return rgb
