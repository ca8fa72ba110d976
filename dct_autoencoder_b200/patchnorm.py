"""PatchNorm on the GPU.  Drop-in for the reference's ``PatchNorm`` (patchnorm.py:32-177):
same constructor, same parameters ``n``, ``median``, ``b`` (so checkpoints load), same
``forward(dct_patches)`` / ``inverse_norm(dct_patches)`` / ``frozen`` / ``std``.

New: when ``torch.distributed`` is initialised with world_size > 1 the statistic-fitting step sums
its two batch statistics over the ranks (NCCL over NVLink on the GPU box, gloo in CPU tests), so
every rank ends the step with identical tables.  The reference has no distributed PatchNorm; the
rule is its own count-weighted mean of batch medians (patchnorm.py:135-138) applied across ranks.
"""
from typing import Optional

import torch
from torch import nn

from . import _lib
from .dct_patches import DCTPatches
from .util import to_device_f32


def stats_sync_enabled() -> bool:
    dist = torch.distributed
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def all_reduce_sum_(buf: torch.Tensor) -> torch.Tensor:
    """Sums a reduce-ready statistics buffer over the default process group, in place (NCCL for
    CUDA tensors, gloo for the CPU tests).  No-op in a single process."""
    if stats_sync_enabled():
        torch.distributed.all_reduce(buf, op=torch.distributed.ReduceOp.SUM)
    return buf


class PatchNorm(nn.Module):
    def __init__(self, max_patch_h: int, max_patch_w: int, patch_size: int, channels: int,
                 eps: float = 1e-6, max_val: float = 6.0, min_val: float = -6.0):
        super().__init__()
        self.eps = eps
        self.patch_size = patch_size
        self.channels = channels
        self.max_patch_h = max_patch_h
        self.max_patch_w = max_patch_w
        z = patch_size ** 2
        self.n = nn.Parameter(torch.zeros(channels, max_patch_h, max_patch_w), requires_grad=False)
        self.median = nn.Parameter(torch.zeros(channels, max_patch_h, max_patch_w, z), requires_grad=False)
        # mean absolute deviation from the median
        self.b = nn.Parameter(torch.ones(channels, max_patch_h, max_patch_w, z), requires_grad=False)
        self.frozen = False
        self.max_val = max_val
        self.min_val = min_val
        # sum the batch statistics over the default process group while fitting
        self.sync_stats = True
        # quantities derived from (median, b) that the fused kernels reuse across calls (decode tables, the "every b is
        # a tame divisor" flag): keyed by what they depend on, dropped whenever the statistics change
        self._derived = {}
        self._stats_version = 0

    def invalidate_derived(self):
        """Forget everything derived from the statistics.  Called by every path of this module that changes them
        (statistic fitting, load_state_dict, .to()); call it yourself after writing ``median`` / ``b`` through
        ``.data`` or a raw pointer, which leaves no trace the cache could see."""
        self._derived.clear()
        self._stats_version += 1

    def derived(self, key, build):
        """``build()`` once per ``key`` and state of the statistics (tensor identity and version counters)."""
        full = (key, self._stats_version, self.median.data_ptr(), self.b.data_ptr(), self.median._version, self.b._version)
        hit = self._derived.get(full)
        if hit is None:
            if len(self._derived) > 16:
                self._derived.clear()
            hit = self._derived[full] = build()
        return hit

    def _apply(self, fn, *args, **kwargs):
        self._derived.clear()
        self._stats_version += 1
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self._derived.clear()
        self._stats_version += 1
        return super()._load_from_state_dict(*args, **kwargs)

    @property
    def std(self) -> torch.Tensor:
        return self.b * 2 ** 0.5

    # ------------------------------------------------------------------ helpers
    def _tables(self, device):
        for prm in (self.n, self.median, self.b):
            if prm.device != device or prm.dtype != torch.float32 or not prm.is_contiguous():
                raise _lib.DctaError("PatchNorm tables must be contiguous fp32 tensors on the device of the "
                                     f"patches ({device}); call .to(device).float()")
        return self.n.data, self.median.data, self.b.data

    def _run_apply(self, dct_patches: DCTPatches, inverse: bool) -> torch.Tensor:
        og = dct_patches.patches.dtype
        x = to_device_f32(dct_patches.patches)
        _lib.require_cuda(x, dct_patches.patch_channels, dct_patches.patch_positions)
        _, median, b = self._tables(x.device)
        ch = dct_patches.patch_channels.contiguous()
        pos = dct_patches.patch_positions.contiguous()
        out = torch.empty_like(x)
        z = x.shape[-1]
        with torch.cuda.device(x.device):
            _lib.call("dcta_patchnorm_apply", _lib.ptr(x), _lib.ptr(ch), _lib.ptr(pos), _lib.ptr(median),
                      _lib.ptr(b), _lib.ptr(out), x.numel() // z, z, self.channels, self.max_patch_h,
                      self.max_patch_w, float(self.eps), float(self.min_val), float(self.max_val),
                      1 if inverse else 0, _lib.stream_ptr(x.device))
        return out if og == torch.float32 else out.to(og)

    # ------------------------------------------------------------------ statistic fitting
    @torch.no_grad()
    def _update_stats(self, dct_patches: DCTPatches, want_output: bool = True) -> Optional[torch.Tensor]:
        """patchnorm.py:101-155.  Returns the un-normalised patches with padding zeroed (``want_output=False``: the
        caller discards them, as the fitting loop of main.py:115-149 does, and the copy is not made)."""
        og = dct_patches.patches.dtype
        x = to_device_f32(dct_patches.patches)
        dev = x.device
        n, median, b = self._tables(dev)
        ch = dct_patches.patch_channels.contiguous()
        pos = dct_patches.patch_positions.contiguous()
        pad = dct_patches.key_pad_mask.contiguous()
        z = x.shape[-1]
        n_tok = x.numel() // z
        C, H, W = self.channels, self.max_patch_h, self.max_patch_w
        n_pos = C * H * W
        i32 = dict(dtype=torch.int32, device=dev)
        counts = torch.empty(n_pos, **i32)
        offsets = torch.empty(n_pos + 1, **i32)
        cursor = torch.empty(n_pos, **i32)
        lists = torch.empty(2 * max(n_tok, 1), **i32)
        packed = torch.empty(n_pos + n_pos * z, dtype=torch.float32, device=dev)
        abs_dev = torch.empty(n_pos * z, dtype=torch.float32, device=dev)
        out = torch.empty_like(x) if want_output else None
        sync = self.sync_stats and stats_sync_enabled()
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.call("dcta_patchnorm_build_lists", _lib.ptr(ch), _lib.ptr(pos), _lib.ptr(pad), n_tok, C, H, W,
                      _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(cursor), _lib.ptr(lists), st)
            _lib.call("dcta_patchnorm_batch_median", _lib.ptr(x), _lib.ptr(offsets), _lib.ptr(lists), n_pos, z,
                      _lib.ptr(packed), st)
            if sync:  # phase 1: [batch_n | batch_median * batch_n], one collective
                all_reduce_sum_(packed)
            _lib.call("dcta_patchnorm_update_median", _lib.ptr(median), _lib.ptr(n), _lib.ptr(packed), n_pos, z, st)
            _lib.call("dcta_patchnorm_abs_dev", _lib.ptr(x), _lib.ptr(offsets), _lib.ptr(lists), _lib.ptr(median),
                      n_pos, z, _lib.ptr(abs_dev), st)
            if sync:  # phase 2: sum |x - median|
                all_reduce_sum_(abs_dev)
            _lib.call("dcta_patchnorm_update_b", _lib.ptr(b), _lib.ptr(n), _lib.ptr(packed), _lib.ptr(abs_dev),
                      n_pos, z, st)
            if want_output:
                _lib.call("dcta_zero_padding", _lib.ptr(x), _lib.ptr(pad), _lib.ptr(out), n_tok, z, st)
        self.invalidate_derived()
        if not want_output:
            return None
        return out if og == torch.float32 else out.to(og)

    # ------------------------------------------------------------------ public
    @torch.no_grad()
    def fit_step(self, dct_patches: DCTPatches) -> None:
        """One update of the running statistics (what ``forward`` does in training mode, patchnorm.py:101-150) without
        forming the returned patches."""
        self._update_stats(dct_patches, want_output=False)

    def forward(self, dct_patches: DCTPatches) -> torch.Tensor:
        """patchnorm.py:81-165.  Training and not frozen: update the running statistics and
        return the patches un-normalised (padding zeroed).  Otherwise
        ``clamp((x - median) / (b*sqrt(2) + eps), min_val, max_val)`` -- padding rows included,
        with the statistics at (0, 0, 0), as in the reference."""
        if self.training and not self.frozen:
            return self._update_stats(dct_patches)
        return self._run_apply(dct_patches, inverse=False)

    def inverse_norm(self, dct_patches: DCTPatches) -> torch.Tensor:
        """patchnorm.py:167-177."""
        return self._run_apply(dct_patches, inverse=True)
