"""VectorQuantize, eval (nearest-codebook) path, on the GPU.

Mirrors the inference behaviour of the reference's ``VectorQuantize`` with a Euclidean codebook
(vector_quantize.py:680-1050 -> EuclideanCodebook.forward :436-507 -> cdist :29-33, argmax :469,
gather :222-226): ``forward(x, indices=None, mask=None, ...) -> (quantize, embed_ind, loss)``,
``codebook`` property, ``get_codes_from_indices``, ``get_output_from_indices``; module/buffer
names follow the reference (``_codebook.embed`` ...) so its checkpoints load.

Codebook LEARNING for the default configuration of the reference's class -- an EMA Euclidean codebook
(``ema_update=True``, ``learnable_codebook=False``) with optional k-means initialisation, dead-code expiry, the
commitment loss and the straight-through estimator (vector_quantize.py:180-220, :417-434, :479-500, :944-1003) -- runs on
libdcta kernels (csrc/vq_train.cu) with the reference's all-reduces of the batch statistics.  Affine
re-parametrisation, cosine codebooks, learnable codebooks, gumbel/reinmax sampling and orthogonal regularisation
are not implemented and raise.
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
    """(C, d) fp32 codebook -> cached (fp16 plane (C, ld) of embed * s, |e|^2 (C,), s, max |e|^2 (1,)): s a power of two from max|embed|
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
    e2_max = e2.max().reshape(1).contiguous()            # for the re-rank shortcut's error bound (stays on the device)
    _CODEBOOK_CACHE[key] = (embed, e_hi, e2, s, e2_max)
    while len(_CODEBOOK_CACHE) > 8:
        _CODEBOOK_CACHE.popitem(last=False)
    return e_hi, e2, s, e2_max


def nearest_code(x: torch.Tensor, embed: torch.Tensor, return_quantized: bool = True, impl: str = "tc", keep=None):
    """x (T, d), embed (C, d) fp32 CUDA -> (indices int64 (T,), embed[indices] or None).
    ``keep`` (T,) bool (tensor-core path only): tokens where it is False get their own row ``x[t]`` back instead of the
    code (vector_quantize.py:1043-1048 folded into the gather).

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
            if keep is not None and q is not None:
                q = torch.where(keep.reshape(-1, 1), q, x)
            return idx, q
        e_hi, e2, s_e, e2_max = _codebook_operand(embed)
        x_hi = torch.empty((T, ld), dtype=torch.float16, device=dev)
        row_alpha = torch.empty(T, dtype=torch.float32, device=dev)
        _lib.call("dcta_split_rows_rowscale", _lib.ptr(x), None, None, 0.0, _lib.ptr(x_hi), None, _lib.ptr(row_alpha),
                  -2.0 / s_e, T, d, ld, st)
        cand = torch.empty((T, 4), dtype=torch.int32, device=dev)
        cand_val = torch.empty((T, 4), dtype=torch.float32, device=dev)
        keep_u8 = None if keep is None else keep.reshape(-1).contiguous().view(torch.uint8)
        _lib.call("dcta_vq_nearest_tc_masked", _lib.ptr(x), _lib.ptr(x_hi), _lib.ptr(row_alpha), _lib.ptr(embed), _lib.ptr(e_hi),
                  _lib.ptr(e2), _lib.ptr(cand), _lib.ptr(cand_val), _lib.ptr(e2_max), _lib.ptr(keep_u8), _lib.ptr(idx),
                  _lib.ptr(q), T, C, d, ld, st)
    return idx, q


def _all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """vector_quantize.py:283-284 ``all_reduce_fn``: sum over the default process group (NCCL / gloo); no-op alone."""
    dist = torch.distributed
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t


def _sample_vectors(samples: torch.Tensor, num: int) -> torch.Tensor:
    """vector_quantize.py:104-112."""
    n = samples.shape[0]
    if n >= num:
        idx = torch.randperm(n, device=samples.device)[:num]
    else:
        idx = torch.randint(0, n, (num,), device=samples.device)
    return samples[idx]


class EuclideanCodebook(nn.Module):
    """State with the reference's buffer names (vector_quantize.py:287-296) + the learning steps on libdcta kernels."""

    def __init__(self, dim, codebook_size, num_codebooks=1, kmeans_init=False, kmeans_iters=10, decay=0.8, eps=1e-5,
                 threshold_ema_dead_code=0, reset_cluster_size=None, sync=None, ema_update=True):
        super().__init__()
        embed = torch.empty(num_codebooks, codebook_size, dim)
        if kmeans_init:
            embed.zero_()                                      # vector_quantize.py:262-263
        else:
            nn.init.kaiming_uniform_(embed)                   # uniform_init, vector_quantize.py:52-55
        self.codebook_size = codebook_size
        self.num_codebooks = num_codebooks
        self.kmeans_iters = kmeans_iters
        self.decay, self.eps = decay, eps
        self.threshold_ema_dead_code = threshold_ema_dead_code
        self.reset_cluster_size = reset_cluster_size if reset_cluster_size is not None else threshold_ema_dead_code
        self.sync = sync                                       # None: whenever torch.distributed has more than one rank
        self.ema_update = ema_update
        self.register_buffer("initted", torch.Tensor([not kmeans_init]))
        # host mirror of `initted` once it is known to be set (it never goes back): the reference reads the device flag on
        # every training forward (vector_quantize.py:334-336), a host synchronisation per step
        self._initted_host = not kmeans_init
        self.register_buffer("cluster_size", torch.zeros(num_codebooks, codebook_size))
        self.register_buffer("embed_avg", embed.clone())
        self.register_buffer("embed", embed)

    def is_initted(self) -> bool:
        if not self._initted_host:
            self._initted_host = bool(self.initted)           # one device read, e.g. after load_state_dict
        return self._initted_host

    def _load_from_state_dict(self, *args, **kwargs):
        self._initted_host = False                           # whatever the checkpoint says: read it once
        return super()._load_from_state_dict(*args, **kwargs)

    def _reduce(self, t):
        return _all_reduce_sum_(t) if self.sync in (None, True) else t

    def cluster_stats(self, head: int, x: torch.Tensor, ind: torch.Tensor, mask: Optional[torch.Tensor]):
        """(counts (C,), sums (C, d)) of the tokens ``x`` (T, d) assigned to ``ind`` (T,), masked tokens left out."""
        C, d = self.embed.shape[1:]
        counts = torch.empty(C, dtype=torch.float32, device=x.device)
        sums = torch.empty((C, d), dtype=torch.float32, device=x.device)
        m = None if mask is None else mask.reshape(-1).to(torch.uint8).contiguous()
        with torch.cuda.device(x.device):
            _lib.call("dcta_vq_cluster_stats", _lib.ptr(x), _lib.ptr(ind), _lib.ptr(m), x.shape[0], d, C, _lib.ptr(counts),
                      _lib.ptr(sums), _lib.stream_ptr(x.device))
        return counts, sums

    @torch.no_grad()
    def init_embed_(self, head: int, x: torch.Tensor, mask: Optional[torch.Tensor], impl: str):
        """k-means initialisation on the first training batch (vector_quantize.py:180-220, :334-355)."""
        data = x if mask is None else x[mask.reshape(-1)]
        means = _sample_vectors(data, self.codebook_size).contiguous()
        d = means.shape[-1]
        counts = None
        for _ in range(self.kmeans_iters):
            buckets, _ = nearest_code(data, means, return_quantized=False, impl=impl)
            counts, sums = self.cluster_stats(head, data, buckets, None)
            self._reduce(counts)
            # the reference all-reduces the MEANS of the ranks (new_means / bins, then all_reduce: vector_quantize.py:210-211)
            with torch.cuda.device(x.device):
                if self.sync in (None, True) and torch.distributed.is_available() and torch.distributed.is_initialized() \
                        and torch.distributed.get_world_size() > 1:
                    local = torch.zeros_like(means)
                    _lib.call("dcta_vq_kmeans_means", _lib.ptr(local), _lib.ptr(counts), _lib.ptr(sums), self.codebook_size, d,
                              _lib.stream_ptr(x.device))
                    _all_reduce_sum_(local)
                    means = torch.where((counts == 0)[:, None], means, local)
                else:
                    _lib.call("dcta_vq_kmeans_means", _lib.ptr(means), _lib.ptr(counts), _lib.ptr(sums), self.codebook_size, d,
                              _lib.stream_ptr(x.device))
            invalidate_codebook_cache()
        self.embed.data[head].copy_(means)
        self.embed_avg.data[head].copy_(means * counts[:, None])
        self.cluster_size.data[head].copy_(counts)

    @torch.no_grad()
    def ema_update_(self, head: int, x: torch.Tensor, ind: torch.Tensor, mask: Optional[torch.Tensor]):
        """vector_quantize.py:479-500 for one codebook: batch statistics (all-reduced), lerp of the running statistics,
        Laplace-smoothed normalisation."""
        counts, sums = self.cluster_stats(head, x, ind, mask)
        self._reduce(counts)
        self._reduce(sums)
        C, d = sums.shape
        total = torch.empty(1, dtype=torch.float32, device=x.device)
        emb, cs, avg = self.embed.data[head], self.cluster_size.data[head], self.embed_avg.data[head]
        assert emb.is_contiguous() and cs.is_contiguous() and avg.is_contiguous()
        with torch.cuda.device(x.device):
            _lib.call("dcta_vq_ema_update", _lib.ptr(emb), _lib.ptr(cs), _lib.ptr(avg), _lib.ptr(counts), _lib.ptr(sums), C, d,
                      float(self.decay), float(self.eps), _lib.ptr(total), _lib.stream_ptr(x.device))
        invalidate_codebook_cache()           # the kernel wrote the codebook behind torch's version counter

    @torch.no_grad()
    def expire_codes_(self, head: int, x: torch.Tensor):
        """vector_quantize.py:403-434: codes whose running cluster size fell below the threshold are replaced by random
        batch vectors (one device->host read of the number of expired codes, as in the reference)."""
        if self.threshold_ema_dead_code == 0:
            return
        expired = self.cluster_size[head] < self.threshold_ema_dead_code
        n = int(expired.sum())
        if n == 0:
            return
        sampled = _sample_vectors(x, n)
        self.embed.data[head][expired] = sampled
        self.cluster_size.data[head][expired] = self.reset_cluster_size
        self.embed_avg.data[head][expired] = sampled * self.reset_cluster_size
        invalidate_codebook_cache()


class _CommitStraightThrough(torch.autograd.Function):
    """(x, q detached, mask) -> (q with the straight-through gradient, mean over the valid tokens of (q - x)^2):
    vector_quantize.py:944-952 and :976-1003.  A non-finite x gives q where ``x + (q - x).detach()`` would give NaN."""

    @staticmethod
    def forward(ctx, x, q, mask):
        xc, qc = x.detach().contiguous(), q.contiguous()
        dim = xc.shape[-1]
        m = None if mask is None else mask.reshape(-1).to(torch.uint8).contiguous()
        result = torch.empty(2, dtype=torch.float32, device=xc.device)
        scratch = torch.empty(int(_lib.load().dcta_masked_mse_scratch_floats()), dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            _lib.call("dcta_masked_mse", _lib.ptr(xc), _lib.ptr(qc), _lib.ptr(m), xc.numel() // dim, dim, _lib.ptr(scratch),
                      _lib.ptr(result), _lib.stream_ptr(xc.device))
        ctx.save_for_backward(xc, qc, m, result)
        return q.view_as(q), result[0]

    @staticmethod
    def backward(ctx, gq, gloss):
        xc, qc, m, result = ctx.saved_tensors
        dim = xc.shape[-1]
        gq = None if gq is None else gq.to(torch.float32).contiguous()
        gl = (torch.zeros(1, device=xc.device) if gloss is None else gloss.detach().reshape(1).to(torch.float32).contiguous())
        gx = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _lib.call("dcta_masked_mse_backward", _lib.ptr(xc), _lib.ptr(qc), _lib.ptr(m), _lib.ptr(result), _lib.ptr(gl),
                      _lib.ptr(gq), _lib.ptr(gx), xc.numel() // dim, dim, _lib.stream_ptr(xc.device))
        return gx, None, None


class VectorQuantize(nn.Module):
    def __init__(self, dim, codebook_size, codebook_dim=None, heads=1, separate_codebook_per_head=False, decay=0.8,
                 eps=1e-5, freeze_codebook=False, kmeans_init=False, kmeans_iters=10, sync_kmeans=True,
                 use_cosine_sim=False, threshold_ema_dead_code=0, channel_last=True, accept_image_fmap=False,
                 commitment_weight=1.0, sync_codebook=None, ema_update=True, learnable_codebook=False,
                 affine_param=False, vq_impl: str = "tc", **unsupported):
        super().__init__()
        assert vq_impl in ("tc", "fp32")
        self.vq_impl = vq_impl      # "tc": tcgen05 nearest-code search; "fp32": exact FFMA kernel
        if use_cosine_sim or affine_param or learnable_codebook or not ema_update:
            raise NotImplementedError("cosine-similarity, affine-parametrised and learnable (non-EMA) codebooks are "
                                      "not implemented: only the EMA Euclidean codebook is")
        bad = {k: v for k, v in unsupported.items()
               if k in ("commitment_use_cross_entropy_loss", "orthogonal_reg_weight", "stochastic_sample_codes",
                        "straight_through", "reinmax", "in_place_codebook_optimizer", "sync_update_v") and v}
        if bad:
            raise NotImplementedError(f"VectorQuantize options outside the implemented codebook learning: {sorted(bad)}")
        self.dim = dim
        self.heads = heads
        self.separate_codebook_per_head = separate_codebook_per_head
        codebook_dim = codebook_dim if codebook_dim is not None else dim
        codebook_input_dim = codebook_dim * heads
        requires_projection = codebook_input_dim != dim
        self.project_in = nn.Linear(dim, codebook_input_dim) if requires_projection else nn.Identity()
        self.project_out = nn.Linear(codebook_input_dim, dim) if requires_projection else nn.Identity()
        self.has_projections = requires_projection
        self.commitment_weight = commitment_weight
        self.freeze_codebook = freeze_codebook
        self._codebook = EuclideanCodebook(codebook_dim, codebook_size,
                                           num_codebooks=heads if separate_codebook_per_head else 1,
                                           kmeans_init=kmeans_init, kmeans_iters=kmeans_iters, decay=decay, eps=eps,
                                           threshold_ema_dead_code=threshold_ema_dead_code, sync=sync_codebook)
        self.codebook_size = codebook_size
        self.accept_image_fmap = accept_image_fmap
        self.channel_last = channel_last

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
        """vector_quantize.py:837-1050.  In training mode the EMA codebook is updated (k-means initialisation on the first
        batch, EMA of the all-reduced batch statistics, dead-code expiry), ``quantize`` carries the straight-through
        gradient and ``loss`` is the commitment loss."""
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
        xf = to_device_f32(x.detach())
        cb = self._codebook
        learn = self.training and not (freeze_codebook or self.freeze_codebook)
        # tokens and mask as the codebook sees them: 'h b n d' per head, or heads folded into the batch '(b h) n d'
        if h > 1 and self.separate_codebook_per_head:
            views = [xf.reshape(b, n, h, d)[:, :, i].reshape(b * n, d).contiguous() for i in range(h)]
            masks = [None if mask is None else mask.reshape(-1)] * h
        elif h > 1:
            views = [xf.reshape(b, n, h, d).permute(0, 2, 1, 3).reshape(b * h * n, d).contiguous()]
            masks = [None if mask is None else mask[:, None, :].expand(b, h, n).reshape(-1)]
        else:
            views = [xf.reshape(b * n, d)]
            masks = [None if mask is None else mask.reshape(-1)]
        if self.training and not cb.is_initted():
            for i, (v, m) in enumerate(zip(views, masks)):
                cb.init_embed_(i, v, m, self.vq_impl)                       # vector_quantize.py:334-355
            cb.initted.fill_(1.0)
            cb._initted_host = True
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
            # eval, no projections, tokens-last input: the final torch.where(mask, quantize, orig_input)
            # (vector_quantize.py:1043-1048) is folded into the gather of the re-rank kernel
            mask_in_gather = (mask is not None and not self.training and x is orig_input and x.dtype == torch.float32
                              and xf.data_ptr() == x.data_ptr() and isinstance(self.project_out, nn.Identity)
                              and mask.dtype == torch.bool and mask.device == x.device)
            idx, qf = nearest_code(xf.reshape(b * n, d), ef[0].contiguous(), impl=self.vq_impl,
                                   keep=mask if mask_in_gather else None)
            ind = idx.reshape(b, n)
            q = qf.reshape(b, n, d)
            if mask_in_gather:
                mask = None                  # already applied
        q = q.to(x.dtype)
        # (torch.zeros: an asynchronous fill -- torch.tensor([0.0], device=...) is a blocking copy from pageable memory
        # that makes the host wait for the stream on every call)
        loss = torch.zeros(1, device=x.device, requires_grad=self.training)
        if self.training:
            if learn:
                # EMA update from this batch's assignments (the quantised vectors above still come from the OLD codebook)
                if h > 1 and self.separate_codebook_per_head:
                    flat_ind = [ind[..., i].reshape(-1).contiguous() for i in range(h)]
                elif h > 1:
                    flat_ind = [ind.permute(0, 2, 1).reshape(-1).contiguous()]
                else:
                    flat_ind = [ind.reshape(-1)]
                for i, (v, m, fi) in enumerate(zip(views, masks, flat_ind)):
                    cb.ema_update_(i, v, fi, m)
                    cb.expire_codes_(i, v)
            # commitment loss against the detached codes, straight-through estimator (vector_quantize.py:944-952, :976-1003)
            if (self.commitment_weight > 0 and x.is_cuda and x.dtype == torch.float32 and q.dtype == torch.float32
                    and (mask is None or mask.shape == x.shape[:-1])):
                # both in one flat pass forward and one backward (csrc/vq_train.cu: masked_mse_*); the eager expressions
                # below cost six passes over (b, n, dim) and a host synchronisation for se[mask]
                q, commit = _CommitStraightThrough.apply(x, q.detach(), mask)
                loss = loss + commit * self.commitment_weight
            else:
                if self.commitment_weight > 0:
                    se = (q.detach() - x) ** 2
                    commit = se[mask].mean() if mask is not None else se.mean()
                    loss = loss + commit * self.commitment_weight
                q = x + (q - x).detach()
        if self.accept_image_fmap:
            ind = ind.reshape(b, height, width, *ind.shape[2:])
        if only_one:
            ind = ind[:, 0]
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
