"""Linear layers of the path on the split-precision tcgen05 GEMM, with the row-wise work around them (LayerNorm,
bias, position-embedding gathers) in csrc/glue.cu.  Inference only (``torch.no_grad``): used by the model glue
(modeling_dct_autoencoder.py) and by the quantisers' ``project_in`` / ``project_out`` in eval mode
(lfq.py:60-62, 164, 212; vector_quantize.py:869, 1028)."""
from typing import Optional

import torch
from torch import nn

from . import _lib
from .util import _round8, to_device_f32

_WEIGHT_CACHE = {}


def _split_weight(weight: torch.Tensor):
    """(N, K) fp32 weight -> cached fp16 hi/lo planes (N, round8(K)) of W * s, s a power of two from max|W|
    (one device->host read per weight version), and 1/s."""
    key = (weight.data_ptr(), weight._version, tuple(weight.shape), str(weight.device))
    hit = _WEIGHT_CACHE.get(key)
    if hit is not None:
        return hit
    w = to_device_f32(weight.detach())
    n, k = w.shape
    amax = float(w.abs().max())
    import math
    e = 10 - math.frexp(amax)[1] if amax > 0 and math.isfinite(amax) else 0
    s = 2.0 ** e
    ld = _round8(k)
    hi = torch.zeros((n, ld), dtype=torch.float16, device=w.device)
    lo = torch.zeros_like(hi)
    if ld == k:
        with torch.cuda.device(w.device):
            _lib.call("dcta_split_f32", _lib.ptr(w), _lib.ptr(hi), _lib.ptr(lo), w.numel(), float(s), _lib.stream_ptr(w.device))
    else:
        wp = torch.zeros((n, ld), dtype=torch.float32, device=w.device)
        wp[:, :k] = w
        with torch.cuda.device(w.device):
            _lib.call("dcta_split_f32", _lib.ptr(wp), _lib.ptr(hi), _lib.ptr(lo), wp.numel(), float(s), _lib.stream_ptr(w.device))
    if len(_WEIGHT_CACHE) > 64:
        _WEIGHT_CACHE.clear()
    out = (hi, lo, 1.0 / s)
    _WEIGHT_CACHE[key] = out
    return out


@torch.no_grad()
def linear_rows(x: torch.Tensor, weight: torch.Tensor, ln: Optional[nn.LayerNorm] = None) -> torch.Tensor:
    """``F.linear(LN(x), weight)`` (no bias) for x (..., K) on the split-precision tensor-core GEMM.
    ``ln``: a LayerNorm applied to the rows first, fused into the operand preparation (proj_out)."""
    _lib.require_cuda(x)
    k = x.shape[-1]
    n = weight.shape[0]
    assert weight.shape[1] == k
    x2 = to_device_f32(x).reshape(-1, k)
    t = x2.shape[0]
    dev = x2.device
    w_hi, w_lo, w_inv = _split_weight(weight)
    ld = _round8(k)
    a_hi = torch.empty((t, ld), dtype=torch.float16, device=dev)
    a_lo = torch.empty_like(a_hi)
    row_scale = torch.empty(t, dtype=torch.float32, device=dev)
    out = torch.empty((t, n), dtype=torch.float32, device=dev)
    gamma = beta = None
    eps = 0.0
    if ln is not None:
        gamma, beta, eps = to_device_f32(ln.weight.detach()), to_device_f32(ln.bias.detach()), float(ln.eps)
    step = 65535 * 128            # rows per launch of the basic tcgen05 GEMM (grid.y limit)
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        _lib.call("dcta_split_rows_rowscale", _lib.ptr(x2), _lib.ptr(gamma), _lib.ptr(beta), eps, _lib.ptr(a_hi), _lib.ptr(a_lo),
                  _lib.ptr(row_scale), float(w_inv), t, k, ld, st)
        for r0 in range(0, t, step):
            rows = min(step, t - r0)
            _lib.call("dcta_gemm_split", a_hi[r0:].data_ptr(), a_lo[r0:].data_ptr(), rows, ld, 0, _lib.ptr(w_hi), _lib.ptr(w_lo),
                      n, ld, 0, k, 1, row_scale[r0:].data_ptr(), 1.0, out[r0:].data_ptr(), n, 0, st)
    return out.reshape(x.shape[:-1] + (n,))


@torch.no_grad()
def ln_pos_rows(x: torch.Tensor, ln: Optional[nn.LayerNorm] = None, bias: Optional[torch.Tensor] = None,
                pos=None, channels: Optional[torch.Tensor] = None, positions: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One pass over token rows x (..., F): optional LayerNorm, optional bias, optional position embeddings
    ``pos = (pos_channel (C, F), pos_height (H, F), pos_width (W, F))`` gathered by ``channels`` / ``positions``
    (modeling_dct_autoencoder.py:101-112: x + h_pos + w_pos + c_pos)."""
    _lib.require_cuda(x)
    f = x.shape[-1]
    x2 = to_device_f32(x).reshape(-1, f)
    out = torch.empty_like(x2)
    g = b = pc = ph = pw = ch = ps = None
    eps = 0.0
    if ln is not None:
        g, b, eps = to_device_f32(ln.weight.detach()), to_device_f32(ln.bias.detach()), float(ln.eps)
    if bias is not None:
        bias = to_device_f32(bias.detach())
    if pos is not None:
        pc, ph, pw = (to_device_f32(p.detach()) for p in pos)
        ch = channels.reshape(-1).to(torch.int64).contiguous()
        ps = positions.reshape(-1, 2).to(torch.int64).contiguous()
        assert ch.numel() == x2.shape[0] and ps.shape[0] == x2.shape[0]
    with torch.cuda.device(x2.device):
        _lib.call("dcta_ln_pos_rows", _lib.ptr(x2), _lib.ptr(g), _lib.ptr(b), eps, _lib.ptr(bias), _lib.ptr(pc), _lib.ptr(ph),
                  _lib.ptr(pw), _lib.ptr(ch), _lib.ptr(ps), _lib.ptr(out), x2.shape[0], f, _lib.stream_ptr(x2.device))
    return out.reshape(x.shape)


@torch.no_grad()
def linear_bias_rows(x: torch.Tensor, linear: nn.Linear) -> torch.Tensor:
    """``linear(x)`` for an nn.Linear with (optional) bias: tensor-core GEMM + one bias pass."""
    y = linear_rows(x, linear.weight)
    return y if linear.bias is None else ln_pos_rows(y, bias=linear.bias)
