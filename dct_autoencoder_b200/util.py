"""Math utilities of the transform path, backed by libdcta kernels.

Mirrors the functions of the reference's ``dct_autoencoder/util.py`` that the hot path uses
(same names and argument meaning): ``rgb_to_ipt``/``ipt_to_rgb`` (util.py:70-97), ``dct2``/``idct2``
(util.py:333-338), ``exp_trunc_dist`` (util.py:167-172), ``power_of_two`` (util.py:184-189),
``masked_mean`` (util.py:346-353), ``compute_entropy_loss`` (util.py:355-387) and
``calculate_perplexity`` (util.py:391-410).  All tensor work runs on the GPU.
"""
import math
import random
from functools import lru_cache
from typing import Optional

import numpy as np
import torch

from . import _lib

# --- colour-space constants, built exactly as the reference builds them: fp32 torch ops on the
# --- host at import time (util.py:21-43), so the kernels see bit-identical matrices.
MsRGB = torch.tensor([[0.4124564, 0.3575761, 0.1804375],
                      [0.2126729, 0.7151522, 0.0721750],
                      [0.0193339, 0.1191920, 0.9503041]], dtype=torch.float32)
MHPE = torch.tensor([[0.4002, 0.7076, -0.0807],
                     [-0.2280, 1.1500, 0.0612],
                     [0.0, 0.0, 0.9184]], dtype=torch.float32)
Mipt = torch.tensor([[0.4, 0.4, 0.2],
                     [4.455, -4.851, 0.3960],
                     [0.8056, 0.3572, -1.1628]], dtype=torch.float32)
Trgb2lms = MHPE @ MsRGB
Tlms2rgb = Trgb2lms.inverse()
IPT_GAMMA = 0.43

_M_RGB2LMS = _lib.host_floats(Trgb2lms.flatten().tolist())
_M_IPT = _lib.host_floats(Mipt.flatten().tolist())
_M_IPT_INV = _lib.host_floats(Mipt.inverse().flatten().tolist())
_M_LMS2RGB = _lib.host_floats(Tlms2rgb.flatten().tolist())


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.DctaError("no CUDA device: dct_autoencoder_b200 has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def gpu_numa_cpus(device=None):
    """CPUs of the NUMA node the GPU hangs off (from sysfs), or None when the host has one node or does not say.
    Pinned buffers are placed on the node of the thread that allocates them; a round trip through the far
    socket costs a third of the PCIe bandwidth, so the host streaming path binds to these CPUs first."""
    import os
    try:
        dev = torch.device(device) if device is not None else default_device()
        bus = torch.cuda.get_device_properties(dev).pci_bus_id
        dom = torch.cuda.get_device_properties(dev).pci_domain_id
        devn = torch.cuda.get_device_properties(dev).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devn:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_numa(device=None) -> bool:
    """Restrict this process to the CPUs next to the GPU (see gpu_numa_cpus); returns whether anything changed."""
    import os
    cpus = gpu_numa_cpus(device)
    if not cpus or cpus == os.sched_getaffinity(0):
        return False
    os.sched_setaffinity(0, cpus)
    return True


def to_device_f32(x: torch.Tensor, device=None) -> torch.Tensor:
    """fp32, contiguous, on the GPU (host tensors are staged through pinned memory)."""
    if not x.is_cuda:
        device = device or default_device()
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        if not x.is_pinned():
            x = x.pin_memory()
        return x.to(device, non_blocking=True)
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous()


def to_device_pixels(x: torch.Tensor, device=None, keep_u8: bool = False) -> torch.Tensor:
    """Images for the encoder.  Floating tensors: ``to_device_f32``.  uint8 tensors are 8-bit pixels -- what
    ``torchvision.io.read_image`` returns -- and stand for ``x / 255`` exactly as the reference's callers compute it
    (decode_gif.py:22, testpipe.py:17): they cross PCIe as bytes and are converted on the device, either by the
    consuming kernel (``keep_u8``: returned as a contiguous uint8 CUDA tensor) or by ``u8_to_unit``."""
    if x.dtype != torch.uint8:
        return to_device_f32(x, device)
    if not x.is_cuda:
        device = device or default_device()
        x = x.contiguous()
        if not x.is_pinned():
            x = x.pin_memory()
        x = x.to(device, non_blocking=True)
    x = x.contiguous()
    return x if keep_u8 else u8_to_unit(x)


def u8_to_unit(x: torch.Tensor) -> torch.Tensor:
    """uint8 CUDA tensor -> fp32 ``x / 255`` (IEEE division, bit-identical to torch's)."""
    _lib.require_cuda(x)
    assert x.dtype == torch.uint8
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("dcta_u8_to_unit_f32", _lib.ptr(x), _lib.ptr(out), x.numel(), _lib.stream_ptr(x.device))
    return out


def unit_to_u8(x: torch.Tensor) -> torch.Tensor:
    """fp32 CUDA image -> uint8 the way ``torchvision.utils.save_image`` stores it (testpipe.py:74-75):
    ``floor(clamp(x, 0, 1) * 255 + 0.5)``."""
    x = to_device_f32(x)
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("dcta_unit_f32_to_u8", _lib.ptr(x), _lib.ptr(out), x.numel(), _lib.stream_ptr(x.device))
    return out


def _colorspace(x: torch.Tensor, fn: str, a, b) -> torch.Tensor:
    og_dtype = x.dtype
    x = to_device_f32(x)
    if x.ndim < 3 or x.shape[-3] != 3:
        raise ValueError(f"expected (..., 3, h, w), got {tuple(x.shape)}")
    out = torch.empty_like(x)
    plane = x.shape[-1] * x.shape[-2]
    n_img = x.numel() // (3 * plane) if plane else 0
    with torch.cuda.device(x.device):
        _lib.call(fn, _lib.ptr(x), _lib.ptr(out), n_img, plane, a, b, _lib.stream_ptr(x.device))
    return out if og_dtype == torch.float32 else out.to(og_dtype)


def rgb_to_ipt(x: torch.Tensor) -> torch.Tensor:
    """util.py:70-82."""
    return _colorspace(x, "dcta_rgb_to_ipt", _M_RGB2LMS, _M_IPT)


def ipt_to_rgb(x: torch.Tensor) -> torch.Tensor:
    """util.py:85-97."""
    return _colorspace(x, "dcta_ipt_to_rgb", _M_IPT_INV, _M_LMS2RGB)


def _basis_tables(layout: int, n: int, k: int, shape, with_scale: bool):
    """Host tables from ``dcta_basis_init`` (csrc/api.cu: the one place the DCT basis is generated, in double)."""
    lib = _lib.load()
    elems = int(lib.dcta_basis_elems(layout, n, k))
    assert elems == int(np.prod(shape)), (layout, n, k, shape, elems)
    f32 = layout == _lib.BASIS_F32
    hi = np.empty(shape, np.float32 if f32 else np.float16)
    lo = None if f32 else np.empty(shape, np.float16)
    rs = np.empty(int(np.prod(shape[:-1])), np.float32) if with_scale else None
    rc = lib.dcta_basis_init(layout, n, k, hi.ctypes.data, None if lo is None else lo.ctypes.data,
                             None if rs is None else rs.ctypes.data)
    if rc != 0:
        raise _lib.DctaError(f"dcta_basis_init failed ({rc}): {_lib.last_error()}")
    return hi, lo, rs


@lru_cache(maxsize=64)
def _basis_host(n: int, k: int) -> np.ndarray:
    return _basis_tables(_lib.BASIS_F32, n, k, (k, n), False)[0]


_BASIS_CACHE = {}


def dct_basis(n: int, k: int, device) -> torch.Tensor:
    """First ``k`` rows of the orthonormal DCT-II matrix of size ``n`` as an fp32 (k, n) device
    tensor; generated in float64 on the host, cached per (n, k, device)."""
    device = torch.device(device)
    key = (n, k, device.type, device.index)
    t = _BASIS_CACHE.get(key)
    if t is None:
        t = torch.from_numpy(_basis_host(n, k)).to(device)
        _BASIS_CACHE[key] = t
    return t


def dct2_truncated(x: torch.Tensor, kh: int, kw: int, tile_p: int = 0, channels: int = 1) -> torch.Tensor:
    """``CH[:kh] . x . CW[:kw]^T`` over the last two axes of an fp32 CUDA tensor (..., h, w).
    ``tile_p > 0`` writes the token-grid layout (n_img, kh/p, kw/p, channels, p*p) instead."""
    h, w = x.shape[-2:]
    n_planes = x.numel() // (h * w)
    ch, cw = dct_basis(h, kh, x.device), dct_basis(w, kw, x.device)
    work = torch.empty((n_planes, h, kw), dtype=torch.float32, device=x.device)
    if tile_p:
        y = torch.empty((n_planes // channels, kh // tile_p, kw // tile_p, channels, tile_p * tile_p),
                        dtype=torch.float32, device=x.device)
    else:
        y = torch.empty(x.shape[:-2] + (kh, kw), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("dcta_dct2_fwd", _lib.ptr(x), _lib.ptr(ch), _lib.ptr(cw), _lib.ptr(work), _lib.ptr(y),
                  n_planes, h, w, kh, kw, tile_p, channels, _lib.stream_ptr(x.device))
    return y


def idct2_truncated(y: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """``CH[:kh]^T . y . CW[:kw]`` : (..., kh, kw) coefficients -> (..., h, w) samples."""
    kh, kw = y.shape[-2:]
    n_planes = y.numel() // (kh * kw)
    ch, cw = dct_basis(h, kh, y.device), dct_basis(w, kw, y.device)
    work = torch.empty((n_planes, h, kw), dtype=torch.float32, device=y.device)
    x = torch.empty(y.shape[:-2] + (h, w), dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        _lib.call("dcta_dct2_inv", _lib.ptr(y), _lib.ptr(ch), _lib.ptr(cw), _lib.ptr(work), _lib.ptr(x),
                  n_planes, h, w, kh, kw, _lib.stream_ptr(y.device))
    return x


def _check_norm(norm):
    if norm != "ortho":
        raise NotImplementedError("only norm='ortho' is implemented (the only mode the reference uses: "
                                  "feature_extraction_dct_autoencoder.py:140,149)")


def dct2(x: torch.Tensor, norm: Optional[str] = None) -> torch.Tensor:
    """util.py:333-334 (torch_dct.dct_2d): orthonormal 2-D DCT-II over the last two axes."""
    _check_norm(norm)
    og = x.dtype
    x = to_device_f32(x)
    y = dct2_truncated(x, x.shape[-2], x.shape[-1])
    return y if og == torch.float32 else y.to(og)


def idct2(x: torch.Tensor, norm: Optional[str] = None) -> torch.Tensor:
    """util.py:337-338 (torch_dct.idct_2d)."""
    _check_norm(norm)
    og = x.dtype
    x = to_device_f32(x)
    y = idct2_truncated(x, x.shape[-2], x.shape[-1])
    return y if og == torch.float32 else y.to(og)


def exp_trunc_dist(a: float) -> float:
    """util.py:167-172: one draw of Python's module-global RNG (kept on the host so that the
    number of kept patches follows the reference's random stream exactly)."""
    x = random.random()
    return -1 / a * math.log(x)


def power_of_two(target: int) -> int:
    """util.py:184-189."""
    if target > 1:
        for i in range(1, int(target)):
            if 2 ** i >= target:
                return 2 ** i
    return 1


def masked_mean(x: torch.Tensor, m: torch.Tensor, dim=None):
    """util.py:346-353 (host-side composition of torch ops; tiny tensors only)."""
    m = m.to(x.dtype)
    x = x * m.reshape(m.shape + (1,) * (x.ndim - m.ndim))
    x = x / m.sum()
    return x.sum() if dim is None else x.sum(dim=dim)


class FactorizedDistance:
    """What ``LFQ.forward`` returns as ``distance`` at scale: the quantiser's input ``x`` (b, n, c, d) and the codebook
    scale instead of the dense ``-2 x . codebook^T`` tensor (b, n, c, 2^d) of lfq.py:191 (721 GB at 786 k tokens x 14 x
    2^14).  ``compute_entropy_loss`` -- its one consumer (main.py:63-65, modeling_dct_autoencoder.py:195-199) -- evaluates
    the same loss from it through the factorisation of the softmax over a +-s codebook; ``dense()`` materialises the
    reference's tensor for small shapes."""

    def __init__(self, x: torch.Tensor, codebook_scale: float):
        self.x, self.codebook_scale = x, float(codebook_scale)
        self.dtype, self.device = x.dtype, x.device

    @property
    def shape(self):
        return tuple(self.x.shape[:-1]) + (2 ** self.x.shape[-1],)

    def float(self):
        return self

    def to(self, *_args, **_kw):
        return self

    def dense(self) -> torch.Tensor:
        x = to_device_f32(self.x.detach())
        b, n, c, d = x.shape
        out = torch.empty((b, n, c, 2 ** d), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("dcta_lfq_distance", _lib.ptr(x), _lib.ptr(out), b * n, c, d, self.codebook_scale,
                      _lib.stream_ptr(x.device))
        return out


class _FactorizedEntropy(torch.autograd.Function):
    """util.py:355-387 on a FactorizedDistance, forward and backward in csrc/lfq_entropy.cu."""

    @staticmethod
    def forward(ctx, x, mask, scale, temperature, eps):
        xf = to_device_f32(x.detach())
        b, n, c, d = xf.shape
        dev = xf.device
        m = mask.to(dev).reshape(b * n).to(torch.uint8).contiguous()
        ctas = int(_lib.load().dcta_lfq_entropy_ctas())
        partial = torch.empty(ctas * (2 ** d + 2), dtype=torch.float32, device=dev)
        tables = torch.empty(2 * 2 ** d, dtype=torch.float32, device=dev)
        result = torch.empty(4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("dcta_lfq_entropy_factorized", _lib.ptr(xf), _lib.ptr(m), b * n, c, d, float(scale), float(temperature),
                      float(eps), _lib.ptr(partial), _lib.ptr(tables), _lib.ptr(result), _lib.stream_ptr(dev))
        ctx.save_for_backward(xf, m, tables, result)
        ctx.cfg = (float(scale), float(temperature), x.dtype)
        return result[0].clone()

    @staticmethod
    def backward(ctx, g):
        xf, m, tables, result = ctx.saved_tensors
        scale, temperature, og = ctx.cfg
        b, n, c, d = xf.shape
        gx = torch.empty_like(xf)
        gout = g.detach().reshape(1).to(torch.float32).contiguous()
        with torch.cuda.device(xf.device):
            _lib.call("dcta_lfq_entropy_factorized_backward", _lib.ptr(xf), _lib.ptr(m), b * n, c, d, scale, temperature,
                      _lib.ptr(tables), _lib.ptr(result), _lib.ptr(gout), _lib.ptr(gx), _lib.stream_ptr(xf.device))
        return gx.to(og), None, None, None, None


def compute_entropy_loss(affinity, mask: torch.Tensor, temperature=0.01, eps=1e-9):
    """util.py:355-387.  ``affinity``: the dense (b, s, d, z) tensor, or the ``FactorizedDistance`` that ``LFQ.forward``
    returns in training mode at scale (differentiable w.r.t. the quantiser's input); ``mask`` (b, s) False at padding."""
    if isinstance(affinity, FactorizedDistance):
        return _FactorizedEntropy.apply(affinity.x, mask, affinity.codebook_scale, temperature, eps).to(affinity.dtype)
    if torch.is_grad_enabled() and affinity.requires_grad:
        # a DENSE affinity that has to be differentiated (small by construction: LFQ hands out the factorised form beyond
        # dense_distance_limit): the reference's own torch expression, so autograd sees it
        a = affinity.to(torch.float32).reshape(-1, affinity.shape[-2], affinity.shape[-1])
        m = mask.reshape(-1).to(a.device)
        logits = a / temperature + eps
        probs, log_probs = logits.softmax(dim=-1), torch.nn.functional.log_softmax(logits, dim=-1)
        avg_probs = masked_mean(probs, m, dim=0).mean(dim=0)
        avg_entropy = -1 * (avg_probs * (avg_probs + eps).log()).sum()
        sample_entropy = -1 * masked_mean((probs * log_probs).sum(dim=-1), m)
        return (sample_entropy - avg_entropy).to(affinity.dtype)
    og = affinity.dtype
    a = to_device_f32(affinity)
    b, s, d, z = a.shape
    m = mask.to(a.device).reshape(b * s).to(torch.uint8).contiguous()
    scratch = torch.empty(z + 2, dtype=torch.float32, device=a.device)
    result = torch.empty(1, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call("dcta_entropy_loss", _lib.ptr(a), _lib.ptr(m), _lib.ptr(scratch), _lib.ptr(result),
                  b * s, d, z, float(temperature), float(eps), _lib.stream_ptr(a.device))
    return result[0].to(og)


def calculate_perplexity(codes: torch.Tensor, codebook_size: int, null_index=-1):
    """util.py:391-410."""
    _lib.require_cuda(codes)
    c = codes.reshape(-1).to(torch.int64).contiguous()
    counts = torch.empty(codebook_size, dtype=torch.int64, device=c.device)
    result = torch.empty(1, dtype=torch.float32, device=c.device)
    with torch.cuda.device(c.device):
        _lib.call("dcta_perplexity", _lib.ptr(c), c.numel(), codebook_size, int(null_index),
                  _lib.ptr(counts), _lib.ptr(result), _lib.stream_ptr(c.device))
    return result[0]


# ----------------------------------------------------------------------------------------------
# tensor-core (split-precision) DCT path -- see csrc/gemm_tc.cu
# ----------------------------------------------------------------------------------------------
_SCALE_BASIS = 1024.0   # 2^10, must match kScaleBasis in gemm_tc.cu
_SPLIT_CACHE = {}


def _round8(v: int) -> int:
    return (v + 7) // 8 * 8


def _split_host(m: np.ndarray):
    """float64 matrix -> (hi, lo) fp16 with hi + lo == m to ~2^-22 relative."""
    hi = m.astype(np.float16)
    lo = (m - hi.astype(np.float64)).astype(np.float16)
    return hi, lo


def split_basis(n: int, k: int, device, transposed: bool):
    """fp16 hi/lo planes of the scaled orthonormal DCT-II basis C_n[:k].

    forward  (transposed=False): (k, ld=round8(n)) planes of C*2^10, with row 0 (the constant
             1/sqrt(n)) stored as exactly 32 so that the DC sum carries no systematic rounding;
             also returns the fp32 per-row factors that undo the scaling.
    inverse  (transposed=True):  (n, ld=round8(k)) planes of (C*2^10)^T; returns row_scale None."""
    device = torch.device(device)
    key = (n, k, transposed, device.type, device.index)
    hit = _SPLIT_CACHE.get(key)
    if hit is not None:
        return hit
    if not transposed:
        ld = _round8(n)
        hi, lo, rs = _basis_tables(_lib.BASIS_SPLIT_FWD, n, k, (k, ld), True)
        row_scale = torch.from_numpy(rs).to(device)
    else:
        ld = _round8(k)
        hi, lo, _ = _basis_tables(_lib.BASIS_SPLIT_INV, n, k, (n, ld), False)
        row_scale = None
    out = (torch.from_numpy(hi).to(device), torch.from_numpy(lo).to(device), row_scale, ld)
    _SPLIT_CACHE[key] = out
    return out


def split_f32(x: torch.Tensor, scale: float):
    """fp32 CUDA tensor -> (hi, lo) fp16 tensors with x*scale == hi + lo (22 bits)."""
    x = x.contiguous()
    hi = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    with torch.cuda.device(x.device):
        _lib.call("dcta_split_f32", _lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), x.numel(), float(scale),
                  _lib.stream_ptr(x.device))
    return hi, lo


def rgb_to_ipt_split(x: torch.Tensor):
    """util.py:70-82 fused with the operand split: (b, 3, h, w) fp32 RGB -> centred IPT*2^8 as fp16
    hi/lo planes + dc (b*3,) = plane mean * sqrt(h*w) (added back to coefficient [0,0])."""
    b, c, h, w = x.shape
    hi = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    dc = torch.empty(b * 3, dtype=torch.float32, device=x.device)
    scratch = torch.empty(b * 3 * 33, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("dcta_rgb_to_ipt_split", _lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(dc), _lib.ptr(scratch),
                  b, h, w, _M_RGB2LMS, _M_IPT, _lib.stream_ptr(x.device))
    return hi, lo, dc


def split_planes_centered(x: torch.Tensor):
    """fp32 planes (..., h, w) -> centred hi/lo (scale 2^8) + dc (n_planes,)."""
    x = x.contiguous()
    h, w = x.shape[-2:]
    n_planes = x.numel() // (h * w)
    hi = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    dc = torch.empty(n_planes, dtype=torch.float32, device=x.device)
    scratch = torch.empty(n_planes * 33, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("dcta_split_planes_centered", _lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(dc),
                  _lib.ptr(scratch), n_planes, h, w, _lib.stream_ptr(x.device))
    return hi, lo, dc


def tc_forward_ok(h: int, w: int) -> bool:
    """The tensor-core forward path needs TMA-legal image planes (pitch multiple of 8 fp16)."""
    return w % 8 == 0 and (h * w) % 4 == 0


def dct2_fwd_tc(x_hi: torch.Tensor, x_lo: torch.Tensor, dc: Optional[torch.Tensor], kh: int, kw: int,
                tile_p: int = 0, channels: int = 1):
    """Truncated forward DCT on tensor cores from centred split planes (..., h, w) (scale 2^8)."""
    h, w = x_hi.shape[-2:]
    n_planes = x_hi.numel() // (h * w)
    dev = x_hi.device
    bw_hi, bw_lo, rs_w, _ = split_basis(w, kw, dev, False)
    bh_hi, bh_lo, rs_h, ld_h = split_basis(h, kh, dev, False)
    work_hi = torch.empty((n_planes, kw, ld_h), dtype=torch.float16, device=dev)
    work_lo = torch.empty_like(work_hi)
    if tile_p:
        y = torch.empty((n_planes // channels, kh // tile_p, kw // tile_p, channels, tile_p * tile_p),
                        dtype=torch.float32, device=dev)
    else:
        y = torch.empty(x_hi.shape[:-2] + (kh, kw), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dcta_dct2_fwd_tc", _lib.ptr(x_hi), _lib.ptr(x_lo), _lib.ptr(dc), _lib.ptr(bw_hi), _lib.ptr(bw_lo),
                  _lib.ptr(rs_w), _lib.ptr(bh_hi), _lib.ptr(bh_lo), _lib.ptr(rs_h), _lib.ptr(work_hi),
                  _lib.ptr(work_lo), _lib.ptr(y), n_planes, h, w, kh, kw, ld_h, tile_p, channels, _lib.stream_ptr(dev))
    return y


def dct2_inv_tc(y_hi: torch.Tensor, y_lo: torch.Tensor, dc: Optional[torch.Tensor], kw: int, h: int, w: int):
    """Truncated inverse DCT on tensor cores from split coefficient planes (..., kh, ld_kw) (scale 2^4,
    DC coefficient carried separately in ``dc``)."""
    kh, ld_kw = y_hi.shape[-2:]
    n_planes = y_hi.numel() // (kh * ld_kw)
    dev = y_hi.device
    bwt_hi, bwt_lo, _, ld_kw_b = split_basis(w, kw, dev, True)
    bht_hi, bht_lo, _, ld_kh = split_basis(h, kh, dev, True)
    assert ld_kw_b == ld_kw
    work_hi = torch.empty((n_planes, w, ld_kh), dtype=torch.float16, device=dev)
    work_lo = torch.empty_like(work_hi)
    x = torch.empty(y_hi.shape[:-2] + (h, w), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dcta_dct2_inv_tc", _lib.ptr(y_hi), _lib.ptr(y_lo), _lib.ptr(dc), _lib.ptr(bwt_hi), _lib.ptr(bwt_lo),
                  _lib.ptr(bht_hi), _lib.ptr(bht_lo), _lib.ptr(work_hi), _lib.ptr(work_lo), _lib.ptr(x),
                  n_planes, h, w, kh, kw, ld_kh, ld_kw, _lib.stream_ptr(dev))
    return x


def dct2_truncated_tc(x: torch.Tensor, kh: int, kw: int, tile_p: int = 0, channels: int = 1):
    """fp32 planes -> truncated DCT through the tensor-core path (|x - mean| must stay below 2^7)."""
    hi, lo, dc = split_planes_centered(x)
    return dct2_fwd_tc(hi, lo, dc, kh, kw, tile_p, channels)


def idct2_truncated_tc(y: torch.Tensor, h: int, w: int):
    """fp32 coefficient planes (..., kh, kw) -> samples through the tensor-core path (AC |y| < 2^11)."""
    y = y.contiguous()
    kh, kw = y.shape[-2:]
    ld = _round8(kw)
    n_planes = y.numel() // (kh * kw)
    hi = torch.empty(y.shape[:-1] + (ld,), dtype=torch.float16, device=y.device)
    lo = torch.empty_like(hi)
    dc = torch.empty(n_planes, dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        _lib.call("dcta_split_coef_planes", _lib.ptr(y), _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(dc), n_planes, kh, kw,
                  ld, h, w, _lib.stream_ptr(y.device))
    return dct2_inv_tc(hi, lo, dc, kw, h, w)


# ----------------------------------------------------------------------------------------------
# folded tensor-core DCT path (half the multiply-adds) -- see csrc/dct_fold.cu
# ----------------------------------------------------------------------------------------------
_FOLD_CACHE = {}


def fold_ok(h: int, w: int, kh: int, kw: int) -> bool:
    """True if the folded path takes these sizes (any h, w >= 2; kh, kw even; basis fits in shared memory)."""
    return bool(_lib.load().dcta_fold_supported(int(h), int(w), int(kh), int(kw)))


def fold_basis(n: int, k: int, device, transposed: bool):
    """fp16 hi/lo planes of the folded orthonormal DCT-II basis: group g holds the rows of parity g,
    C_n[2j+g, :n/2] * 2^10.

    forward  (transposed=False): (2, k/2, round8(ceil(n/2))); row 0 of group 0 (the constant 1/sqrt(n)) is stored as
             exactly 32; also returns the fp32 (2, k/2) factors that undo the scaling.
    inverse  (transposed=True):  (2, ceil(n/2), round8(k/2)) = the transposes; row_scale None.
    Odd n: the middle sample pairs with itself (it sits in the even rows once, the odd rows hold 0 there)."""
    device = torch.device(device)
    key = (n, k, transposed, device.type, device.index)
    hit = _FOLD_CACHE.get(key)
    if hit is not None:
        return hit
    k2, n2 = k // 2, (n + 1) // 2
    if not transposed:
        hi, lo, rs = _basis_tables(_lib.BASIS_FOLD_FWD, n, k, (2, k2, _round8(n2)), True)
        row_scale = torch.from_numpy(rs.reshape(2, k2)).to(device)
    else:
        hi, lo, _ = _basis_tables(_lib.BASIS_FOLD_INV, n, k, (2, n2, _round8(k2)), False)
        row_scale = None
    out = (torch.from_numpy(hi).to(device), torch.from_numpy(lo).to(device), row_scale)
    _FOLD_CACHE[key] = out
    return out


def fold_half(n: int) -> int:
    """Samples per axis of a folded quadrant: ceil(n / 2) (odd n: the middle sample pairs with itself)."""
    return (n + 1) // 2


def rgb_to_ipt_fold(x: torch.Tensor):
    """util.py:70-82 fused with centring, the 2-D fold and the operand split: (b, 3, h, w) fp32 RGB ->
    quadrants (2, 2, b*3, ceil(h/2), round8(ceil(w/2))) fp16 hi/lo (scale 2^6) + dc (b*3,).  Any h, w >= 2."""
    b, c, h, w = x.shape
    hi = torch.empty((2, 2, b * 3, fold_half(h), _round8(fold_half(w))), dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    dc = torch.empty(b * 3, dtype=torch.float32, device=x.device)
    scratch = torch.empty(b * 3 * 33, dtype=torch.float32, device=x.device)
    fn = "dcta_rgb_u8_to_ipt_fold" if x.dtype == torch.uint8 else "dcta_rgb_to_ipt_fold"   # uint8: read as x / 255
    with torch.cuda.device(x.device):
        _lib.call(fn, _lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(dc), _lib.ptr(scratch),
                  b, h, w, _M_RGB2LMS, _M_IPT, _lib.stream_ptr(x.device))
    return hi, lo, dc


def fold_planes(x: torch.Tensor):
    """fp32 planes (..., h, w) -> centred folded quadrants (2, 2, n_planes, ceil(h/2), round8(ceil(w/2))) hi/lo + dc."""
    x = x.contiguous()
    h, w = x.shape[-2:]
    n_planes = x.numel() // (h * w)
    hi = torch.empty((2, 2, n_planes, fold_half(h), _round8(fold_half(w))), dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    dc = torch.empty(n_planes, dtype=torch.float32, device=x.device)
    scratch = torch.empty(n_planes * 33, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("dcta_fold_planes", _lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(dc), _lib.ptr(scratch),
                  n_planes, h, w, _lib.stream_ptr(x.device))
    return hi, lo, dc


def dct2_fwd_fold(xq_hi: torch.Tensor, xq_lo: torch.Tensor, dc: Optional[torch.Tensor], kh: int, kw: int,
                  tile_p: int = 0, channels: int = 1, out_shape=None, with_maxabs: bool = False, hw=None):
    """Truncated forward DCT from folded quadrants (2, 2, n_planes, ceil(h/2), round8(ceil(w/2))).  ``hw``: the plane
    size (h, w); needed when it is not a multiple of 16 (default: twice the quadrant shape).  ``with_maxabs`` (token
    grid only) also returns amax|tile| per token (n_img, kh/p, kw/p, channels), reduced in the GEMM epilogue."""
    n_planes, h2, w2 = xq_hi.shape[-3:]
    h, w = hw if hw is not None else (2 * h2, 2 * w2)
    assert fold_half(h) == h2 and _round8(fold_half(w)) == w2, "quadrant shape does not belong to this plane size"
    dev = xq_hi.device
    bw_hi, bw_lo, rs_w = fold_basis(w, kw, dev, False)
    bh_hi, bh_lo, rs_h = fold_basis(h, kh, dev, False)
    work_hi = torch.empty((2, n_planes, kw, _round8(h2)), dtype=torch.float16, device=dev)
    work_lo = torch.empty_like(work_hi)
    if tile_p:
        y = torch.empty((n_planes // channels, kh // tile_p, kw // tile_p, channels, tile_p * tile_p),
                        dtype=torch.float32, device=dev)
    else:
        y = torch.empty((out_shape or (n_planes,)) + (kh, kw), dtype=torch.float32, device=dev)
    maxabs = None
    if with_maxabs:
        assert tile_p > 0
        maxabs = torch.empty(y.shape[:4], dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dcta_dct2_fwd_fold", _lib.ptr(xq_hi), _lib.ptr(xq_lo), _lib.ptr(dc), _lib.ptr(bw_hi),
                  _lib.ptr(bw_lo), _lib.ptr(rs_w), _lib.ptr(bh_hi), _lib.ptr(bh_lo), _lib.ptr(rs_h), _lib.ptr(work_hi),
                  _lib.ptr(work_lo), _lib.ptr(y), _lib.ptr(maxabs), n_planes, h, w, kh, kw, tile_p,
                  channels, _lib.stream_ptr(dev))
    return (y, maxabs) if with_maxabs else y


def dct2_fwd_fold_codes(xq_hi: torch.Tensor, xq_lo: torch.Tensor, dc: torch.Tensor, kh: int, kw: int, tile_p: int,
                        channels: int, norm, hw=None):
    """Forward DCT from folded quadrants straight to LFQ code words (one codebook per patch row) of the
    PatchNorm-normalised coefficients: returns (maxabs (n_img, kh/p, kw/p, channels),
    code_grid (n_img, kh/p, kw/p, channels, p) int32) -- the token grid itself is never written."""
    n_planes, h2, w2 = xq_hi.shape[-3:]
    h, w = hw if hw is not None else (2 * h2, 2 * w2)
    assert fold_half(h) == h2 and _round8(fold_half(w)) == w2, "quadrant shape does not belong to this plane size"
    dev = xq_hi.device
    bw_hi, bw_lo, rs_w = fold_basis(w, kw, dev, False)
    bh_hi, bh_lo, rs_h = fold_basis(h, kh, dev, False)
    work_hi = torch.empty((2, n_planes, kw, _round8(h2)), dtype=torch.float16, device=dev)
    work_lo = torch.empty_like(work_hi)
    shape = (n_planes // channels, kh // tile_p, kw // tile_p, channels)
    maxabs = torch.empty(shape, dtype=torch.float32, device=dev)
    code_grid = torch.empty(shape + (tile_p,), dtype=torch.int32, device=dev)
    # the "every b is a tame divisor" flag depends on the statistics only: checked once per state of the tables
    fresh = []
    tame = norm.derived(("b_tame", str(dev)), lambda: fresh.append(1) or torch.empty(1, dtype=torch.int32, device=dev))
    with torch.cuda.device(dev):
        _lib.call("dcta_dct2_fwd_fold_codes", _lib.ptr(xq_hi), _lib.ptr(xq_lo), _lib.ptr(dc), _lib.ptr(bw_hi),
                  _lib.ptr(bw_lo), _lib.ptr(rs_w), _lib.ptr(bh_hi), _lib.ptr(bh_lo), _lib.ptr(rs_h), _lib.ptr(work_hi),
                  _lib.ptr(work_lo), _lib.ptr(maxabs), _lib.ptr(code_grid), _lib.ptr(norm.median.data),
                  _lib.ptr(norm.b.data), norm.max_patch_h, norm.max_patch_w, float(norm.eps), float(norm.min_val),
                  float(norm.max_val), _lib.ptr(tame), 0 if fresh else 1, n_planes, h, w, kh, kw, tile_p, channels,
                  _lib.stream_ptr(dev))
    if not fresh:
        _lib.launch_count -= 1
    return maxabs, code_grid


def dct2_inv_fold(yq_hi: torch.Tensor, yq_lo: torch.Tensor, kh: int, kw: int, h: int, w: int) -> torch.Tensor:
    """Folded coefficient quadrants (2, 2, n_planes, kh/2, round8(kw/2)) -> quadrant transforms
    z (4, n_planes, ceil(h/2), ceil(w/2)) fp32 (to be un-folded by ``unfold_ipt_to_rgb`` / ``unfold_planes``)."""
    n_planes = yq_hi.shape[2]
    dev = yq_hi.device
    bwt_hi, bwt_lo, _ = fold_basis(w, kw, dev, True)
    bht_hi, bht_lo, _ = fold_basis(h, kh, dev, True)
    ldi = _round8(kh // 2)
    work_hi = torch.empty((2, 2, n_planes, fold_half(w), ldi), dtype=torch.float16, device=dev)
    work_lo = torch.empty_like(work_hi)
    z = torch.empty((4, n_planes, fold_half(h), fold_half(w)), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dcta_dct2_inv_fold", _lib.ptr(yq_hi), _lib.ptr(yq_lo), _lib.ptr(bwt_hi), _lib.ptr(bwt_lo),
                  _lib.ptr(bht_hi), _lib.ptr(bht_lo), _lib.ptr(work_hi), _lib.ptr(work_lo), _lib.ptr(z),
                  n_planes, h, w, kh, kw, _lib.stream_ptr(dev))
    return z


DECODE_COEF_LIMIT = 2.0 ** 11   # |AC coefficient| the fp16 hi/lo decode operand represents (numeric range)


def decode_codes_inv_fold_ok(h: int, w: int, kh: int, kw: int, p: int, c: int, d: int) -> bool:
    return bool(_lib.load().dcta_decode_codes_inv_fold_supported(h, w, kh, kw, p, c, d))


def decode_codes_inv_fold(codes, slot_map, sel, n_img: int, channels: int, th: int, tw: int, p: int, kh: int, kw: int,
                          h: int, w: int, norm, c: int, d: int, scale: float, code_grid=None):
    """LFQ codes -> quadrant transforms z (4, n_img*channels, h/2, w/2) + dc (n_img*channels): de-quantisation,
    inverse PatchNorm and un-patchify (lfq.py:105-134, patchnorm.py:167-177, FE:607-656) happen in the operand
    producer of inverse pass 1 (``dcta_decode_codes_inv_fold``); the coefficient planes are never written.
    The two-value table is a function of the PatchNorm statistics: built once per state of ``norm``.
    ``code_grid``: the (n_img, kh/p, kw/p, channels, p) int32 code grid of ``dct2_fwd_fold_codes`` for a batch that kept
    every token -- the sign bits are then read from it directly (no slot map, no gather through ``codes``)."""
    dev = codes.device if codes is not None else code_grid.device
    n_planes = n_img * channels
    lib = _lib.load()
    bwt_hi, bwt_lo, _ = fold_basis(w, kw, dev, True)
    bht_hi, bht_lo, _ = fold_basis(h, kh, dev, True)
    ldi = _round8(kh // 2)
    work_hi = torch.empty((2, 2, n_planes, fold_half(w), ldi), dtype=torch.float16, device=dev)
    work_lo = torch.empty_like(work_hi)
    z = torch.empty((4, n_planes, fold_half(h), fold_half(w)), dtype=torch.float32, device=dev)
    dc = torch.empty(n_planes, dtype=torch.float32, device=dev)
    scratch = torch.empty(lib.dcta_decode_codes_inv_fold_scratch_bytes(n_img, channels, kh, kw), dtype=torch.uint8, device=dev)
    median, b = norm.median.data, norm.b.data
    H, W, eps = norm.max_patch_h, norm.max_patch_w, float(norm.eps)

    def build_table():
        # the fp16 hi/lo operand carries |AC coefficient| < 2^11 (scale 16); the two values a code bit selects are
        # median +- scale * (b * sqrt(2) + eps), so the statistics decide it: checked once per state, where the table is built
        if not torch.cuda.is_current_stream_capturing():
            peak = median.abs() + (b * 2 ** 0.5 + eps) * abs(float(scale))
            peak[:, 0, 0, 0] = 0.0                                        # the DC term of every channel travels in fp32
            if not bool((peak < DECODE_COEF_LIMIT).all()):
                raise ValueError("PatchNorm statistics put de-normalised coefficients at or beyond 2^11 (max %.4g): the "
                                 "tensor-core decode would overflow its fp16 operand planes; use dct_impl=\"fp32\" "
                                 "for such statistics" % float(peak.max()))
        tab = torch.empty(lib.dcta_decode_gen_tables_bytes(channels, kh, kw), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.call("dcta_decode_gen_tables", _lib.ptr(median), _lib.ptr(b), channels, H, W, eps, p, kh, kw, float(scale),
                      _lib.ptr(tab), _lib.stream_ptr(dev))
        return tab

    tab = norm.derived(("decode_tab", str(dev), channels, p, kh, kw, float(scale)), build_table)
    with torch.cuda.device(dev):
        if code_grid is not None:
            assert tuple(code_grid.shape) == (n_img, kh // p, kw // p, channels, p) and code_grid.dtype == torch.int32
            _lib.call("dcta_decode_grid_inv_fold", _lib.ptr(code_grid), n_img, channels, p, kh, kw, h, w, _lib.ptr(median),
                      _lib.ptr(b), H, W, eps, float(scale), _lib.ptr(bwt_hi), _lib.ptr(bwt_lo), _lib.ptr(bht_hi),
                      _lib.ptr(bht_lo), _lib.ptr(work_hi), _lib.ptr(work_lo), _lib.ptr(z), _lib.ptr(dc), _lib.ptr(tab),
                      _lib.ptr(scratch), _lib.stream_ptr(dev))
        else:
            _lib.call("dcta_decode_codes_inv_fold", _lib.ptr(codes), _lib.ptr(slot_map), _lib.ptr(sel), n_img, channels, th, tw,
                      p, kh, kw, h, w, _lib.ptr(median), _lib.ptr(b), H, W, eps, c, d, float(scale), _lib.ptr(bwt_hi),
                      _lib.ptr(bwt_lo), _lib.ptr(bht_hi), _lib.ptr(bht_lo), _lib.ptr(work_hi), _lib.ptr(work_lo), _lib.ptr(z),
                      _lib.ptr(dc), _lib.ptr(tab), _lib.ptr(scratch), _lib.stream_ptr(dev))
    return z, dc


def unfold_ipt_to_rgb(z: torch.Tensor, dc: Optional[torch.Tensor], h: int, w: int, out_dtype=torch.float32) -> torch.Tensor:
    """Final butterfly of the folded inverse fused with util.py:85-97: z (4, n_img*3, h/2, w/2) -> RGB (n_img, 3, h, w).
    ``out_dtype=torch.uint8``: 8-bit pixels, quantised in the same kernel as ``unit_to_u8`` does."""
    n_img = z.shape[1] // 3
    assert out_dtype in (torch.float32, torch.uint8)
    rgb = torch.empty((n_img, 3, h, w), dtype=out_dtype, device=z.device)
    fn = "dcta_unfold_ipt_to_rgb_u8" if out_dtype == torch.uint8 else "dcta_unfold_ipt_to_rgb"
    with torch.cuda.device(z.device):
        _lib.call(fn, _lib.ptr(z), _lib.ptr(dc), _lib.ptr(rgb), n_img, h, w, _M_IPT_INV,
                  _M_LMS2RGB, _lib.stream_ptr(z.device))
    return rgb


def dct2_truncated_fold(x: torch.Tensor, kh: int, kw: int, tile_p: int = 0, channels: int = 1):
    """fp32 planes (..., h, w) -> truncated DCT through the folded tensor-core path."""
    hi, lo, dc = fold_planes(x)
    return dct2_fwd_fold(hi, lo, dc, kh, kw, tile_p, channels, out_shape=tuple(x.shape[:-2]), hw=tuple(x.shape[-2:]))


def idct2_truncated_fold(y: torch.Tensor, h: int, w: int):
    """fp32 coefficient planes (..., kh, kw) -> samples (..., h, w) through the folded tensor-core path."""
    y = y.contiguous()
    kh, kw = y.shape[-2:]
    n_planes = y.numel() // (kh * kw)
    ldq = _round8(kw // 2)
    hi = torch.empty((2, 2, n_planes, kh // 2, ldq), dtype=torch.float16, device=y.device)
    lo = torch.empty_like(hi)
    dc = torch.empty(n_planes, dtype=torch.float32, device=y.device)
    x = torch.empty(y.shape[:-2] + (h, w), dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        st = _lib.stream_ptr(y.device)
        _lib.call("dcta_fold_coef_planes", _lib.ptr(y), _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(dc), n_planes, kh, kw,
                  h, w, st)
        z = dct2_inv_fold(hi, lo, kh, kw, h, w)
        _lib.call("dcta_unfold_planes", _lib.ptr(z), _lib.ptr(dc), _lib.ptr(x), n_planes, h, w, st)
    return x
