"""VectorQuantize, eval (nearest-codebook) path, on the GPU.

Mirrors the inference behaviour of the reference's ``VectorQuantize`` with a Euclidean codebook
(vector_quantize.py:680-1050 -> EuclideanCodebook.forward :436-507 -> cdist :29-33, argmax :469,
gather :222-226): ``forward(x, indices=None, mask=None, ...) -> (quantize, embed_ind, loss)``,
``codebook`` property, ``get_codes_from_indices``, ``get_output_from_indices``; module/buffer
names follow the reference (``_codebook.embed`` ...) so its checkpoints load.

Codebook LEARNING (k-means init, EMA updates, dead-code expiry, affine re-parametrisation, cosine
codebooks, gumbel/reinmax sampling, orthogonal regularisation and their all-reduces) is outside
the transform path and is not implemented: ``forward`` in training mode raises.
"""
import collections
from typing import Optional

import torch
from torch import nn

from . import _lib
from .linear import linear_bias_rows
from .util import to_device_f32


_CODEBOOK_CACHE = collections.OrderedDict()      # key -> (embed, e_hi, e2, s); at most 8 codebooks


def invalidate_codebook_cache():
    """Forget every cached codebook operand (call after writing a codebook through ``.data``, which does not bump
    the version counter the cache keys on)."""
    _CODEBOOK_CACHE.clear()


def _codebook_operand(embed: torch.Tensor):
    """(C, d) fp32 codebook -> cached (fp16 plane (C, ld) of embed * s, |e|^2 (C,), s): s a power of two from max|embed|
    (one device->host read per codebook version; the codebook is a parameter, the tokens never cause a sync).
    The entry keeps a reference to ``embed``: while it is cached its memory cannot be recycled for another tensor, so
    (address, version counter) identifies the contents."""
    key = (embed.data_ptr(), embed._version, tuple(embed.shape), tuple(embed.stride()), str(embed.device))
    hit = _CODEBOOK_CACHE.get(key)
    if hit is not None:
        _CODEBOOK_CACHE.move_to_end(key)
        return hit[1:]
    import math
    C, d = embed.shape
    ld = (d + 7) // 8 * 8
    amax = float(embed.abs().max())
    s = 2.0 ** (10 - math.frexp(amax)[1]) if amax > 0 and math.isfinite(amax) else 1.0
    e_hi = torch.zeros((C, ld), dtype=torch.float16, device=embed.device)
    e_hi[:, :d] = (embed * s).to(torch.float16)            # round-to-nearest fp16 of the scaled codebook (one-off)
    e2 = torch.empty(C, dtype=torch.float32, device=embed.device)
    with torch.cuda.device(embed.device):
        _lib.call("dcta_row_sumsq", _lib.ptr(embed), _lib.ptr(e2), C, d, _lib.stream_ptr(embed.device))
    _CODEBOOK_CACHE[key] = (embed, e_hi, e2, s)
    while len(_CODEBOOK_CACHE) > 8:
        _CODEBOOK_CACHE.popitem(last=False)
    return e_hi, e2, s


def nearest_code(x: torch.Tensor, embed: torch.Tensor, return_quantized: bool = True, impl: str = "tc"):
    """x (T, d), embed (C, d) fp32 CUDA -> (indices int64 (T,), embed[indices] or None).

    impl="tc": approximate x.e on tensor cores (ONE fp16 tcgen05 MMA per product, rows scaled by powers of two on the
    device), the two best codes per token and half of the code slices kept in the epilogue, then an exact fp32 re-rank
    of those four candidates (csrc/vq_tc.cu); impl="fp32": exact-fp32 FFMA kernel in the reference's operation order."""
    T, d = x.shape
    C = embed.shape[0]
    dev = x.device
    idx = torch.empty(T, dtype=torch.int64, device=dev)
    q = torch.empty_like(x) if return_quantized else None
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        ld = (d + 7) // 8 * 8
        # the token operand (16 KB per 64 columns) and |e|^2 of the codebook stay in shared memory next to a ring of >= 3 stages
        fits = (227 * 1024 - 6144 - (d + 63) // 64 * 16384 - (C + 255) // 256 * 1024) // 16384 >= 3
        if impl == "fp32" or not fits:
            e2 = torch.empty(C, dtype=torch.float32, device=dev)
            _lib.call("dcta_vq_nearest", _lib.ptr(x), _lib.ptr(embed), _lib.ptr(e2), _lib.ptr(idx), _lib.ptr(q),
                      T, C, d, st)
            return idx, q
        e_hi, e2, s_e = _codebook_operand(embed)
        x_hi = torch.empty((T, ld), dtype=torch.float16, device=dev)
        row_alpha = torch.empty(T, dtype=torch.float32, device=dev)
        _lib.call("dcta_split_rows_rowscale", _lib.ptr(x), None, None, 0.0, _lib.ptr(x_hi), None, _lib.ptr(row_alpha),
                  -2.0 / s_e, T, d, ld, st)
        cand = torch.empty((T, 4), dtype=torch.int32, device=dev)
        _lib.call("dcta_vq_nearest_tc", _lib.ptr(x), _lib.ptr(x_hi), _lib.ptr(row_alpha), _lib.ptr(embed), _lib.ptr(e_hi),
                  _lib.ptr(e2), _lib.ptr(cand), _lib.ptr(idx), _lib.ptr(q), T, C, d, ld, st)
    return idx, q


class EuclideanCodebook(nn.Module):
    """State holder with the reference's buffer names (vector_quantize.py:287-296)."""

    def __init__(self, dim, codebook_size, num_codebooks=1):
        super().__init__()
        embed = torch.empty(num_codebooks, codebook_size, dim)
        nn.init.kaiming_uniform_(embed)                       # uniform_init, vector_quantize.py:52-55
        self.codebook_size = codebook_size
        self.num_codebooks = num_codebooks
        self.register_buffer("initted", torch.Tensor([True]))
        self.register_buffer("cluster_size", torch.zeros(num_codebooks, codebook_size))
        self.register_buffer("embed_avg", embed.clone())
        self.register_buffer("embed", embed)


class VectorQuantize(nn.Module):
    def __init__(self, dim, codebook_size, codebook_dim=None, heads=1, separate_codebook_per_head=False,
                 channel_last=True, accept_image_fmap=False, use_cosine_sim=False, affine_param=False,
                 vq_impl: str = "tc", **training_kwargs):
        super().__init__()
        assert vq_impl in ("tc", "fp32")
        self.vq_impl = vq_impl      # "tc": tcgen05 split-precision distance GEMM; "fp32": exact FFMA kernel
        if use_cosine_sim or affine_param:
            raise NotImplementedError("cosine-similarity / affine-parametrised codebooks are training "
                                      "machinery outside the transform path")
        self.dim = dim
        self.heads = heads
        self.separate_codebook_per_head = separate_codebook_per_head
        codebook_dim = codebook_dim if codebook_dim is not None else dim
        codebook_input_dim = codebook_dim * heads
        requires_projection = codebook_input_dim != dim
        self.project_in = nn.Linear(dim, codebook_input_dim) if requires_projection else nn.Identity()
        self.project_out = nn.Linear(codebook_input_dim, dim) if requires_projection else nn.Identity()
        self.has_projections = requires_projection
        self._codebook = EuclideanCodebook(codebook_dim, codebook_size,
                                           num_codebooks=heads if separate_codebook_per_head else 1)
        self.codebook_size = codebook_size
        self.accept_image_fmap = accept_image_fmap
        self.channel_last = channel_last
        self.training_kwargs = training_kwargs

    @property
    def codebook(self):
        cb = self._codebook.embed
        return cb if self.separate_codebook_per_head else cb[0]

    @codebook.setter
    def codebook(self, codes):
        if not self.separate_codebook_per_head:
            codes = codes[None]
        self._codebook.embed.copy_(codes)

    def get_codes_from_indices(self, indices):
        """vector_quantize.py:814-831."""
        codebook = self.codebook
        if codebook.ndim <= 2:
            codes = codebook[indices]
            return codes.reshape(*codes.shape[:-2], -1) if self.heads > 1 else codes
        b = indices.shape[0]
        flat = indices.reshape(b, -1, indices.shape[-1])                    # b n h
        h = flat.shape[-1]
        codes = torch.stack([codebook[i][flat[..., i]] for i in range(h)], dim=2)  # b n h d
        return codes.reshape(*indices.shape[:-1], -1)

    def _project(self, lin: nn.Module, x: torch.Tensor) -> torch.Tensor:
        """project_in / project_out (vector_quantize.py:869, 1028) on the libdcta GEMM when no autograd is needed."""
        if not self.has_projections:
            return x
        if torch.is_grad_enabled() or not x.is_cuda:
            return lin(x)
        return linear_bias_rows(x, lin).to(x.dtype)

    def get_output_from_indices(self, indices):
        return self._project(self.project_out, self.get_codes_from_indices(indices))

    def forward(self, x, indices=None, mask=None, sample_codebook_temp=None, freeze_codebook=False):
        """vector_quantize.py:837-1050, eval branch."""
        if self.training:
            raise NotImplementedError("VectorQuantize training (codebook learning) is outside the transform "
                                      "path; call .eval()")
        if indices is not None:
            raise NotImplementedError("cross-entropy loss on given indices is a training feature")
        orig_input = x
        only_one = x.ndim == 2
        if only_one:
            assert mask is None
            x = x[:, None]
        need_transpose = not self.channel_last and not self.accept_image_fmap
        if self.accept_image_fmap:
            height, width = x.shape[-2:]
            x = x.flatten(2).transpose(1, 2)                                # 'b c h w -> b (h w) c'
        if need_transpose:
            x = x.transpose(1, 2)
        _lib.require_cuda(x)
        x = self._project(self.project_in, x)
        b, n, _ = x.shape
        h, embed = self.heads, self._codebook.embed
        d = embed.shape[-1]
        xf = to_device_f32(x)
        ef = to_device_f32(embed)
        if h > 1 and self.separate_codebook_per_head:
            xs = xf.reshape(b, n, h, d).permute(2, 0, 1, 3).contiguous()    # 'h b n d'
            outs = [nearest_code(xs[i].reshape(b * n, d), ef[i].contiguous(), impl=self.vq_impl) for i in range(h)]
            ind = torch.stack([o[0].reshape(b, n) for o in outs], dim=-1)   # b n h
            q = torch.stack([o[1].reshape(b, n, d) for o in outs], dim=2).reshape(b, n, h * d)
        elif h > 1:
            # shared codebook: heads are folded into the batch, '1 (b h) n d' (vector_quantize.py:874-875)
            idx, qf = nearest_code(xf.reshape(b * n * h, d), ef[0].contiguous(), impl=self.vq_impl)
            ind = idx.reshape(b, n, h)
            q = qf.reshape(b, n, h * d)
        else:
            idx, qf = nearest_code(xf.reshape(b * n, d), ef[0].contiguous(), impl=self.vq_impl)
            ind = idx.reshape(b, n)
            q = qf.reshape(b, n, d)
        q = q.to(x.dtype)
        if self.accept_image_fmap:
            ind = ind.reshape(b, height, width, *ind.shape[2:])
        if only_one:
            ind = ind[:, 0]
        loss = torch.tensor([0.0], device=x.device)
        q = self._project(self.project_out, q)
        if need_transpose:
            q = q.transpose(1, 2)
        if self.accept_image_fmap:
            q = q.transpose(1, 2).reshape(b, -1, height, width)
        if only_one:
            q = q[:, 0]
        if mask is not None:
            q = torch.where(mask[..., None], q, orig_input)                 # vector_quantize.py:1043-1048
        return q, ind, loss
