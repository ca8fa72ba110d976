"""Linear layers of the path on the split-precision tcgen05 GEMM, with the row-wise work around them (LayerNorm,
bias, position-embedding gathers) in csrc/glue.cu.  Inference only (``torch.no_grad``): used by the model glue
(modeling_dct_autoencoder.py) and by the quantisers' ``project_in`` / ``project_out`` in eval mode
(lfq.py:60-62, 164, 212; vector_quantize.py:869, 1028)."""
from typing import Optional

import contextlib

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


def _split_rows(x2: torch.Tensor, w_inv: float, ln: Optional[nn.LayerNorm] = None, patchnorm=None, row_src=None):
    """fp32 rows (t, k) -> fp16 hi/lo planes (t, round8(k)) scaled per row by a power of two + the factors that undo the
    scaling (times ``w_inv``); ``ln``: a LayerNorm applied to the rows first; ``patchnorm = (PatchNorm, channels (t,),
    positions (t, 2))``: the rows are un-normalised patches and ``PatchNorm.forward`` (frozen statistics) is applied to
    them in the same pass (``dcta_split_rows_patchnorm``); ``row_src`` (t,) int32 (with ``patchnorm``): operand row i is
    row ``row_src[i]`` of ``x2`` (zeros where -1) -- ``x2`` is then the token grid and the packed patches never exist."""
    t, k = x2.shape
    if row_src is not None:
        assert patchnorm is not None and row_src.dtype == torch.int32 and row_src.is_contiguous()
        t = row_src.numel()
    dev = x2.device
    ld = _round8(k)
    a_hi = torch.empty((t, ld), dtype=torch.float16, device=dev)
    a_lo = torch.empty_like(a_hi)
    row_scale = torch.empty(t, dtype=torch.float32, device=dev)
    gamma = beta = None
    eps = 0.0
    if ln is not None:
        gamma, beta, eps = to_device_f32(ln.weight.detach()), to_device_f32(ln.bias.detach()), float(ln.eps)
    if patchnorm is not None:
        assert ln is None
        norm, channels, positions = patchnorm
        ch = channels.reshape(-1).to(torch.int64).contiguous()
        ps = positions.reshape(-1, 2).to(torch.int64).contiguous()
        assert ch.numel() == t and ps.shape[0] == t and k == norm.patch_size ** 2
        with torch.cuda.device(dev):
            _lib.call("dcta_split_rows_patchnorm", _lib.ptr(x2), _lib.ptr(row_src), _lib.ptr(ch), _lib.ptr(ps), _lib.ptr(norm.median.data),
                      _lib.ptr(norm.b.data), norm.channels, norm.max_patch_h, norm.max_patch_w, float(norm.eps),
                      float(norm.min_val), float(norm.max_val), _lib.ptr(a_hi), _lib.ptr(a_lo), _lib.ptr(row_scale),
                      float(w_inv), t, k, ld, _lib.stream_ptr(dev))
        return a_hi, a_lo, row_scale
    with torch.cuda.device(dev):
        _lib.call("dcta_split_rows_rowscale", _lib.ptr(x2), _lib.ptr(gamma), _lib.ptr(beta), eps, _lib.ptr(a_hi), _lib.ptr(a_lo),
                  _lib.ptr(row_scale), float(w_inv), t, k, ld, _lib.stream_ptr(dev))
    return a_hi, a_lo, row_scale


_side_streams = {}


def _side_stream(dev) -> "torch.cuda.Stream":
    """One extra stream per device for output-only work that runs beside the next GEMM."""
    key = str(dev)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(dev)
    return _side_streams[key]


_GEMM_ROWS = 65535 * 128            # rows per launch of the basic tcgen05 GEMM (grid.y limit)


@torch.no_grad()
def linear_rows(x: torch.Tensor, weight: torch.Tensor, ln: Optional[nn.LayerNorm] = None,
                bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``F.linear(LN(x), weight, bias)`` for x (..., K) on the split-precision tensor-core GEMM.
    ``ln``: a LayerNorm applied to the rows first, fused into the operand preparation (proj_out);
    ``bias``: added in the GEMM epilogue."""
    _lib.require_cuda(x)
    k = x.shape[-1]
    n = weight.shape[0]
    assert weight.shape[1] == k
    x2 = to_device_f32(x).reshape(-1, k)
    t = x2.shape[0]
    dev = x2.device
    w_hi, w_lo, w_inv = _split_weight(weight)
    ld = _round8(k)
    a_hi, a_lo, row_scale = _split_rows(x2, w_inv, ln)
    out = torch.empty((t, n), dtype=torch.float32, device=dev)
    if bias is not None:
        bias = to_device_f32(bias.detach())
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        for r0 in range(0, t, _GEMM_ROWS):
            rows = min(_GEMM_ROWS, t - r0)
            _lib.call("dcta_gemm_split", a_hi[r0:].data_ptr(), a_lo[r0:].data_ptr(), rows, ld, 0, _lib.ptr(w_hi), _lib.ptr(w_lo),
                      n, ld, 0, k, 1, row_scale[r0:].data_ptr(), 1.0, _lib.ptr(bias), out[r0:].data_ptr(), n, 0, st)
    return out.reshape(x.shape[:-1] + (n,))


@torch.no_grad()
def lfq_project_quantize(x: torch.Tensor, project_in: nn.Linear, project_out: Optional[nn.Linear], num_codebooks: int,
                         codebook_dim: int, codebook_scale: float, patchnorm=None, row_src=None):
    """LFQ with projections in eval (lfq.py:136-227): ``project_in`` + bias + sign + index packing in ONE GEMM kernel
    (``dcta_lfq_project_sign``; the (t, c*d) activations are never written as fp32), the +-scale codes as an exact fp16
    operand, ``project_out`` + bias as a two-MMA GEMM on it.  Returns (project_out(q) (..., dim) fp32 or the codes
    (..., c*d) when ``project_out`` is None, indices (..., c) int64)."""
    _lib.require_cuda(x)
    k = x.shape[-1]
    n = num_codebooks * codebook_dim
    assert project_in.weight.shape == (n, k)
    x2 = to_device_f32(x).reshape(-1, k)
    t = x2.shape[0] if row_src is None else row_src.numel()
    lead = x.shape[:-1] if row_src is None else tuple(row_src.shape)       # row_src (rows, s): the packed batch's shape
    dev = x2.device
    w_hi, w_lo, w_inv = _split_weight(project_in.weight)
    ld = _round8(k)
    a_hi, a_lo, row_scale = _split_rows(x2, w_inv, patchnorm=patchnorm,      # patchnorm, row_src: see _split_rows
                                        row_src=None if row_src is None else row_src.reshape(-1))
    ldq = _round8(n)
    q_hi = torch.empty((t, ldq), dtype=torch.float16, device=dev) if ldq == n else torch.zeros((t, ldq), dtype=torch.float16, device=dev)
    n_tiles = (n + 127) // 128
    bits = torch.empty((t, n_tiles, 4), dtype=torch.int32, device=dev)
    idx = torch.empty((t, num_codebooks), dtype=torch.int64, device=dev)
    b_in = None if project_in.bias is None else to_device_f32(project_in.bias.detach())
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        for r0 in range(0, t, _GEMM_ROWS):
            rows = min(_GEMM_ROWS, t - r0)
            _lib.call("dcta_lfq_project_sign", a_hi[r0:].data_ptr(), a_lo[r0:].data_ptr(), rows, ld, _lib.ptr(w_hi), _lib.ptr(w_lo),
                      n, ld, k, row_scale[r0:].data_ptr(), _lib.ptr(b_in), float(codebook_scale), q_hi[r0:].data_ptr(), ldq,
                      bits[r0:].data_ptr(), st)
        # the indices are an output only (project_out reads the +-scale operand): pack them on a side stream beside
        # the second GEMM; `bits` and `idx` were allocated on this stream and stay referenced until the join below
        side = None if (project_out is None or _lib.profile_active()) else _side_stream(dev)
        if side is not None:
            side.wait_stream(torch.cuda.current_stream(dev))
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            _lib.call("dcta_lfq_bits_to_codes", _lib.ptr(bits), t, n, num_codebooks, codebook_dim, _lib.ptr(idx),
                      _lib.stream_ptr(dev))
        del a_hi, a_lo
        if project_out is None:
            out = q_hi[:, :n].float()
        else:
            dim = project_out.weight.shape[0]
            assert project_out.weight.shape[1] == n
            o_hi, o_lo, o_inv = _split_weight(project_out.weight)
            b_out = None if project_out.bias is None else to_device_f32(project_out.bias.detach())
            out = torch.empty((t, dim), dtype=torch.float32, device=dev)
            for r0 in range(0, t, _GEMM_ROWS):
                rows = min(_GEMM_ROWS, t - r0)
                # A = q_hi (exact fp16, no lo plane); the weight's scale is undone by alpha
                _lib.call("dcta_gemm_split", q_hi[r0:].data_ptr(), None, rows, ldq, 0, _lib.ptr(o_hi), _lib.ptr(o_lo), dim, ldq, 0,
                          n, 1, None, float(o_inv), _lib.ptr(b_out), out[r0:].data_ptr(), dim, 0, st)
        if side is not None:
            torch.cuda.current_stream(dev).wait_stream(side)
    return out.reshape(tuple(lead) + (out.shape[-1],)), idx.reshape(tuple(lead) + (num_codebooks,))


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
    """``linear(x)`` for an nn.Linear with (optional) bias: tensor-core GEMM, bias added in its epilogue."""
    return linear_rows(x, linear.weight, bias=linear.bias)
