"""Lookup-free quantiser on the GPU.  Drop-in for the reference's ``LFQ`` (lfq.py:35-227):
same keyword-only constructor, ``forward(x, mask) -> (x, indices, commit_loss, distance)``,
``indices_to_codes``, buffers ``mask`` (persistent), ``zero`` / ``codebook`` (not persistent).

The sign quantisation, the MSB-first bit packing into int64 indices, ``indices_to_codes``, the
masked commitment loss and the ``distance`` tensor are libdcta kernels.  The optional ``project_in`` /
``project_out`` are ``nn.Linear`` modules (same state-dict keys as the reference); without autograd (eval under
``torch.no_grad``) they run on the split-precision tcgen05 GEMM (linear.py), with autograd PyTorch differentiates
them as in the reference.
"""
from math import ceil, log2

import torch
from torch import nn

from . import _lib
from .linear import lfq_project_quantize, linear_bias_rows
from .util import FactorizedDistance, to_device_f32


def _exists(v):
    return v is not None


class _StraightThrough(torch.autograd.Function):
    """lfq.py:179-181 ``x + (quantized - x).detach()``: the value is ``quantized`` (x - x is an exact zero for finite
    x), the gradient goes to ``x`` unchanged -- without the two elementwise passes over (b, n, c*d) the expression costs
    in eager mode.  (A non-finite x gives +-scale here where the expression would give NaN.)"""

    @staticmethod
    def forward(ctx, x, quantized):
        return quantized.view_as(quantized)

    @staticmethod
    def backward(ctx, g):
        return g, None


class _CommitLoss(torch.autograd.Function):
    """lfq.py:195-200: sum over valid tokens of (x - q)^2 / (n_valid * c * d)."""

    @staticmethod
    def forward(ctx, x, mask, scale):
        xf = to_device_f32(x.detach())
        cd = xf.shape[-1]
        m = mask.reshape(-1).to(torch.uint8).contiguous()
        result = torch.empty(1, dtype=torch.float32, device=xf.device)
        scratch = torch.empty(_lib.REDUCE_SCRATCH, dtype=torch.float32, device=xf.device)
        with torch.cuda.device(xf.device):
            _lib.call("dcta_lfq_commit_loss", _lib.ptr(xf), _lib.ptr(m), _lib.ptr(result), _lib.ptr(scratch),
                      xf.numel() // cd, cd, float(scale), _lib.stream_ptr(xf.device))
        ctx.save_for_backward(x, mask)
        ctx.scale = scale
        return result[0].to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        x, mask = ctx.saved_tensors
        xf = to_device_f32(x.detach())
        cd = xf.shape[-1]
        m = mask.reshape(-1).to(torch.uint8).contiguous()
        gx = torch.empty_like(xf)
        nv = torch.empty(1, dtype=torch.float32, device=xf.device)
        gout = g.detach().reshape(1).to(torch.float32).contiguous()
        with torch.cuda.device(xf.device):
            _lib.call("dcta_lfq_commit_backward", _lib.ptr(xf), _lib.ptr(m), _lib.ptr(gout), _lib.ptr(nv), _lib.ptr(gx),
                      xf.numel() // cd, cd, float(ctx.scale), _lib.stream_ptr(xf.device))
        return gx.to(x.dtype), None, None


class _Distance(torch.autograd.Function):
    """lfq.py:191: -2 * einsum('... i d, j d -> ... i j', x, codebook) for the +-scale codebook."""

    @staticmethod
    def forward(ctx, x, scale, codebook):
        xf = to_device_f32(x.detach())          # (b, n, c, d)
        b, n, c, d = xf.shape
        out = torch.empty((b, n, c, 2 ** d), dtype=torch.float32, device=xf.device)
        with torch.cuda.device(xf.device):
            _lib.call("dcta_lfq_distance", _lib.ptr(xf), _lib.ptr(out), b * n, c, d, float(scale),
                      _lib.stream_ptr(xf.device))
        ctx.save_for_backward(codebook)
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        (codebook,) = ctx.saved_tensors
        return -2 * torch.einsum("...ij,jd->...id", g, codebook.to(g.dtype)), None, None


class LFQ(nn.Module):
    def __init__(
        self,
        *,
        dim=None,
        codebook_size=None,
        diversity_gamma=2.5,
        straight_through_activation=nn.Identity(),
        num_codebooks=1,
        keep_num_codebooks_dim=None,
        codebook_scale=1.0,
        dense_distance_limit=2 ** 28,
    ):
        """``dense_distance_limit`` (not in the reference): in training mode ``forward`` returns the dense ``distance``
        tensor of lfq.py:191 while it has at most this many elements, and a ``util.FactorizedDistance`` -- which
        ``util.compute_entropy_loss`` consumes without ever forming (b, n, c, 2^d) -- beyond it."""
        super().__init__()
        self.dense_distance_limit = dense_distance_limit
        assert _exists(dim) or _exists(codebook_size), "either dim or codebook_size must be specified for LFQ"
        assert not _exists(codebook_size) or log2(codebook_size).is_integer(), (
            f"your codebook size must be a power of 2 for lookup free quantization "
            f"(suggested {2 ** ceil(log2(codebook_size))})")
        codebook_size = codebook_size if _exists(codebook_size) else 2 ** dim
        codebook_dim = int(log2(codebook_size))
        codebook_dims = codebook_dim * num_codebooks
        dim = dim if _exists(dim) else codebook_dims

        has_projections = dim != codebook_dims
        self.project_in = nn.Linear(dim, codebook_dims) if has_projections else nn.Identity()
        self.project_out = nn.Linear(codebook_dims, dim) if has_projections else nn.Identity()
        self.has_projections = has_projections

        self.dim = dim
        self.codebook_dim = codebook_dim
        self.num_codebooks = num_codebooks
        keep_num_codebooks_dim = keep_num_codebooks_dim if _exists(keep_num_codebooks_dim) else num_codebooks > 1
        assert not (num_codebooks > 1 and not keep_num_codebooks_dim)
        self.keep_num_codebooks_dim = keep_num_codebooks_dim
        self.activation = straight_through_activation
        self.diversity_gamma = diversity_gamma
        self.codebook_scale = codebook_scale

        self.register_buffer("mask", 2 ** torch.arange(codebook_dim - 1, -1, -1))
        self.register_buffer("zero", torch.tensor(0.0), persistent=False)
        self._codebook_cache = None

    # lfq.py:92-96: all sign patterns, MSB first.  Built on first use (2^d x d floats).
    @property
    def codebook(self) -> torch.Tensor:
        cb = self._codebook_cache
        if cb is None or cb.device != self.mask.device:
            codes = torch.arange(2 ** self.codebook_dim, device=self.mask.device)
            bits = ((codes[..., None].int() & self.mask) != 0).float()
            cb = self.bits_to_codes(bits)
            self._codebook_cache = cb
        return cb

    def bits_to_codes(self, bits):
        return bits * self.codebook_scale * 2 - self.codebook_scale

    @property
    def dtype(self):
        return self.zero.dtype

    def _project(self, lin: nn.Module, x: torch.Tensor) -> torch.Tensor:
        """project_in / project_out: libdcta GEMM when nothing has to be differentiated."""
        if not self.has_projections:
            return x
        if torch.is_grad_enabled() or not x.is_cuda:
            return lin(x)
        return linear_bias_rows(x, lin).to(x.dtype)

    def _flat_tokens(self, t: torch.Tensor) -> int:
        return t.numel() // t.shape[-1]

    def indices_to_codes(self, indices: torch.Tensor, project_out=True):
        """lfq.py:105-134."""
        is_img_or_video = indices.ndim >= (3 + int(self.keep_num_codebooks_dim))
        if not self.keep_num_codebooks_dim:
            indices = indices[..., None]
        _lib.require_cuda(indices)
        idx = indices.to(torch.int64).contiguous()
        c, d = self.num_codebooks, self.codebook_dim
        assert idx.shape[-1] == c
        codes = torch.empty(idx.shape[:-1] + (c * d,), dtype=torch.float32, device=idx.device)
        with torch.cuda.device(idx.device):
            _lib.call("dcta_lfq_indices_to_codes", _lib.ptr(idx), _lib.ptr(codes), idx.numel() // c, c, d,
                      float(self.codebook_scale), _lib.stream_ptr(idx.device))
        codes = codes.to(self.dtype)
        if project_out:
            codes = self._project(self.project_out, codes)
        if is_img_or_video:
            codes = codes.movedim(-1, 1)     # 'b ... d -> b d ...'
        return codes

    def forward(self, x: torch.Tensor, mask=None, *, patchnorm=None):
        """lfq.py:136-227.  ``mask`` (b, n): False where padding is (required, lfq.py:153-154).
        ``patchnorm`` (not in the reference; eval with projections only): ``(PatchNorm, channels, positions)`` -- ``x`` holds
        UN-normalised patches and the frozen PatchNorm is applied inside the operand split of ``project_in``; the result
        equals ``forward(norm(batch), mask)`` bit for bit."""
        is_img_or_video = x.ndim >= 4
        if mask is None:
            raise NotImplementedError("mask")
        spatial = None
        if is_img_or_video:
            x = x.movedim(1, -1)             # 'b d ... -> b ... d'
            spatial = x.shape[1:-1]
            x = x.reshape(x.shape[0], -1, x.shape[-1])
        assert x.shape[-1] == self.dim, f"expected dimension of {self.dim} but received {x.shape[-1]}"
        _lib.require_cuda(x)

        c, d = self.num_codebooks, self.codebook_dim
        if self.has_projections and not self.training and not torch.is_grad_enabled() and x.is_cuda:
            # eval with projections: project_in + sign + index packing in one GEMM kernel, project_out on the exact
            # +-scale operand (linear.lfq_project_quantize); nothing of size (b, n, c*d) is written as fp32
            out, indices = lfq_project_quantize(x, self.project_in, self.project_out, c, d, self.codebook_scale,
                                                patchnorm=patchnorm)
            out = out.to(x.dtype)
            if is_img_or_video:
                out = out.reshape(out.shape[0], *spatial, out.shape[-1]).movedim(-1, 1)
                indices = indices.reshape(indices.shape[0], *spatial, c)
            if not self.keep_num_codebooks_dim:
                indices = indices[..., 0]
            return out, indices, self.zero, self.zero

        assert patchnorm is None, "patchnorm= is fused only into the eval path with projections (no_grad, CUDA input)"
        x = self._project(self.project_in, x)
        b, n, _ = x.shape
        original_input = x

        xf = to_device_f32(x.detach())
        quantized = torch.empty_like(xf)
        indices = torch.empty((b, n, c), dtype=torch.int64, device=xf.device)
        with torch.cuda.device(xf.device):
            _lib.call("dcta_lfq_quantize", _lib.ptr(xf), _lib.ptr(quantized), _lib.ptr(indices), b * n, c, d,
                      float(self.codebook_scale), _lib.stream_ptr(xf.device))
        quantized = quantized.to(x.dtype)

        if self.training:
            xa = self.activation(x)
            out = _StraightThrough.apply(xa, quantized) if xa.requires_grad else quantized   # straight-through (lfq.py:179-181)
            if b * n * c * 2 ** d <= self.dense_distance_limit:
                distance = _Distance.apply(original_input.reshape(b, n, c, d), self.codebook_scale, self.codebook)
            else:
                distance = FactorizedDistance(original_input.reshape(b, n, c, d), self.codebook_scale)
            commit_loss = _CommitLoss.apply(original_input, mask.to(x.device), self.codebook_scale)
        else:
            out = quantized
            distance = self.zero
            commit_loss = self.zero

        out = self._project(self.project_out, out)

        if is_img_or_video:
            out = out.reshape(out.shape[0], *spatial, out.shape[-1]).movedim(-1, 1)
            indices = indices.reshape(indices.shape[0], *spatial, c)
        if not self.keep_num_codebooks_dim:
            indices = indices[..., 0]
        return out, indices, commit_loss, distance
