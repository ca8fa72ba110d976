"""The benchmarked pipeline: image -> DCT tokens -> PatchNorm -> quantiser -> inverse PatchNorm ->
image, assembled from the drop-in classes exactly as a user of the reference would
(main.py:183-190 normalize_, modeling_dct_autoencoder.py:149-153 quantiser call with
``mask=~key_pad_mask``, main.py:96-98 inv_normalize_ + postprocess), minus the transformer."""
import math
from typing import List, Optional, Sequence

import torch

from .dct_patches import DCTPatches, wire_records
from .feature_extraction_dct_autoencoder import DCTAutoencoderFeatureExtractor
from .patchnorm import PatchNorm
from .util import power_of_two, unit_to_u8


def dict_collate(x: List[dict]) -> dict:
    """dataset.py:8-15: list of dicts -> dict of lists."""
    out = {}
    for d in x:
        for k, v in d.items():
            out.setdefault(k, []).append(v)
    return out


def get_max_seq_length(max_patch_h: int, max_patch_w: int, image_channels: int,
                       sample_patches_beta: float, cdf_p: float = 0.95) -> int:
    """factory.py:11-33: 95 % point of the exponential patch-count distribution, rounded up to a
    power of two, capped at the full token count."""
    full = max_patch_h * max_patch_w * image_channels
    if sample_patches_beta <= 0:
        return full
    n = round(-1 * math.log(1 - cdf_p) / sample_patches_beta)
    return min(full, power_of_two(n))


class TransformPipeline:
    def __init__(self, extractor: DCTAutoencoderFeatureExtractor, norm: PatchNorm, quantizer):
        self.extractor = extractor
        self.norm = norm
        self.quantizer = quantizer
        # fused round trip: launch the sort and the gather into the packed API tensors on a side stream beside the decode
        # when every token was kept (False: every launch on the caller's stream)
        self.overlap_pack = True
        self._pack_streams = {}

    @torch.no_grad()
    def fit_norm(self, images: torch.Tensor, ks: Optional[Sequence[int]] = None) -> None:
        """One statistic-fitting step (main.py:115-149 train_patch_norm body), then freeze."""
        was_training, was_frozen = self.norm.training, self.norm.frozen
        self.norm.train()
        self.norm.frozen = False
        self.norm.fit_step(self.extractor.process_batch(images, ks))
        self.norm.frozen = True
        self.norm.train(was_training)

    @torch.no_grad()
    def encode(self, images: torch.Tensor, ks: Optional[Sequence[int]] = None):
        """-> (batch with NORMALISED patches, quantiser output, codes)."""
        batch = self.extractor.process_batch(images, ks)
        batch.patches = self.norm(batch)
        out = self.quantizer(batch.patches, mask=~batch.key_pad_mask)
        return batch, out[0], out[1]

    @torch.no_grad()
    def decode(self, batch: DCTPatches, quantized: torch.Tensor) -> torch.Tensor:
        """quantised (normalised) patches -> (n, c, h, w) RGB for same-size batches."""
        b = batch.shallow_copy()
        b.patches = quantized
        b.patches = self.norm.inverse_norm(b)
        return self.extractor.postprocess_batch(b)

    def fusable(self) -> bool:
        """True when PatchNorm + LFQ can run inside the pack / un-patchify kernels (projection-free
        LFQ in eval mode, frozen fp32 statistics, tensor-core DCT path)."""
        fe = self.extractor
        return (fe.dct_impl in ("tc", "tc_plain") and not fe._hooks_overridden("_transform_image_in")
                and not fe._hooks_overridden("_transform_image_out")
                and fe._lfq_fusable(self.norm, self.quantizer))

    @torch.no_grad()
    def encode_codes(self, images: torch.Tensor, ks: Optional[Sequence[int]] = None):
        """Images -> (batch metadata, codes); uses the fused kernel when ``fusable()``."""
        if self.fusable():
            return self.extractor.process_batch_to_codes(images, self.norm, self.quantizer, ks)
        batch, _, codes = self.encode(images, ks)
        return batch, codes

    @torch.no_grad()
    def decode_codes(self, batch: DCTPatches, codes: torch.Tensor, out_dtype=torch.float32) -> torch.Tensor:
        """(batch metadata, codes) -> (n, c, h, w) RGB (modeling_dct_autoencoder.py:157-163
        decode_from_codes without the transformer); fused when ``fusable()``.
        ``out_dtype=torch.uint8``: 8-bit pixels as torchvision's save_image would store the float result."""
        if self.fusable():
            return self.extractor.postprocess_codes_batch(batch, codes, self.norm, self.quantizer, out_dtype)
        b = batch.shallow_copy()
        b.patches = self.quantizer.indices_to_codes(codes)
        b.patches = self.norm.inverse_norm(b)
        return self.extractor.postprocess_batch(b, out_dtype)

    @torch.no_grad()
    def roundtrip_staged(self, images: torch.Tensor, ks: Optional[Sequence[int]] = None):
        """The round trip through the drop-in modules one by one (every intermediate tensor
        materialised, as a user of the reference's API would write it)."""
        batch, q, codes = self.encode(images, ks)
        return self.decode(batch, q), codes

    @torch.no_grad()
    def roundtrip(self, images: torch.Tensor, ks: Optional[Sequence[int]] = None, fused: Optional[bool] = None,
                  out_dtype=torch.float32):
        """images -> (reconstructed images, codes).  ``fused=None`` picks the fused PatchNorm+LFQ
        kernels when they apply; results are bit-identical to ``roundtrip_staged``.  uint8 ``images`` are 8-bit
        pixels standing for ``x / 255``; ``out_dtype=torch.uint8`` returns 8-bit pixels (util.unit_to_u8)."""
        if fused is None:
            fused = self.fusable()
            if not fused and self.proj_fusable() and self.extractor._dev(images) == self.norm.median.device:
                return self._proj_roundtrip(images, ks, out_dtype)
        if not fused:
            rec, codes = self.roundtrip_staged(images, ks)
            return (rec if out_dtype == torch.float32 else unit_to_u8(rec)), codes
        batch, codes, rec = self._fused_roundtrip(images, ks, out_dtype)
        return rec, codes

    def proj_fusable(self) -> bool:
        """True for an LFQ WITH projections in eval mode behind frozen fp32 PatchNorm statistics on the folded tensor-core
        path (the conf/patch14-l.json quantiser): PatchNorm then runs inside the operand split of ``project_in`` and its
        inverse inside the un-patchify kernel, so neither the normalised nor the de-normalised patches are written."""
        from .lfq import LFQ
        fe, q, n = self.extractor, self.quantizer, self.norm
        return (isinstance(q, LFQ) and q.has_projections and not q.training and q.dim == fe.patch_size ** 2
                and fe.dct_impl == "tc" and fe.channels == 3 and fe.patch_size >= 8 and fe.patch_size % 2 == 0
                and not fe._hooks_overridden("_transform_image_in") and not fe._hooks_overridden("_transform_image_out")
                and (n.frozen or not n.training) and n.median.dtype == torch.float32 and n.median.is_cuda
                and n.patch_size == fe.patch_size and n.channels == fe.channels)

    def _proj_roundtrip(self, images, ks, out_dtype):
        """roundtrip_staged with the two PatchNorm passes folded into their neighbours (bit-identical results)."""
        from .linear import lfq_project_quantize
        q = self.quantizer
        # the token gather of the packing step happens inside the operand split too: the packed patches never exist
        batch, grid, row_src = self.extractor._process_batch_index(images, ks)
        out, codes = lfq_project_quantize(grid, q.project_in, q.project_out, q.num_codebooks, q.codebook_dim, q.codebook_scale,
                                          patchnorm=(self.norm, batch.patch_channels, batch.patch_positions), row_src=row_src)
        if not q.keep_num_codebooks_dim:         # lfq.py:224-225
            codes = codes[..., 0]
        batch.patches = out
        return self.extractor.postprocess_batch(batch, out_dtype, denorm=self.norm), codes

    def _fused_roundtrip(self, images, ks, out_dtype):
        """encode_codes + decode_codes; the decode of a batch that kept every token reads the sign bits straight from the
        forward pass's code grid (no slot map, no gather through the packed codes; same pixels)."""
        dev = self.extractor._dev(images)             # where the extractor will run (host images are uploaded there)
        side = self._pack_stream(dev) if self.overlap_pack else None
        batch, codes, grid = self.extractor.process_batch_to_codes(images, self.norm, self.quantizer, ks, return_grid=True,
                                                                   pack_stream=side)
        try:
            if grid is not None:
                # the decode needs the code grid only: the sort and the gather into the API's packed tensors (90 us of
                # latency-bound work per 256 images) run on the side stream beside it
                rec = self.extractor.postprocess_codes_batch(batch, codes, self.norm, self.quantizer, out_dtype, code_grid=grid)
            else:
                rec = self.decode_codes(batch, codes, out_dtype)
        finally:
            self.extractor.join_pack()
        return batch, codes, rec

    def _pack_stream(self, dev):
        if dev not in self._pack_streams:
            self._pack_streams[dev] = torch.cuda.Stream(dev)
        return self._pack_streams[dev]

    def graphed(self, images: torch.Tensor, ks: Optional[Sequence[int]] = None, fused: Optional[bool] = None):
        """``roundtrip`` of a fixed-shape device batch captured once in a CUDA graph; see GraphedRoundtrip."""
        return GraphedRoundtrip(self, images, ks, fused)

    @torch.no_grad()
    def roundtrip_host(self, images: torch.Tensor, out_images: Optional[torch.Tensor] = None,
                       out_codes: Optional[torch.Tensor] = None, chunk: int = 32, device=None,
                       compact: bool = False, out_counts: Optional[torch.Tensor] = None):
        """Host-to-host round trip of a (n, c, h, w) HOST batch (pinned memory for full speed):
        the batch is streamed through the GPU in chunks so that the host->device copy of chunk i+1,
        the kernels of chunk i and the device->host copy of chunk i-1 overlap (three streams).
        Image-independent work only, so chunking does not change any result.

        Default (the reference's types): fp32 images in, fp32 images + int64 codes out --
        (out_images (n, c, h, w) fp32, out_codes (n, s, codebooks) int64).
        ``compact=True``: the same computation with 8-bit pixels on both sides of the link and the codes in the
        wire format of ``to_bytes`` -- uint8 images in (read as ``x / 255``, what ``read_image(path) / 255`` gives the
        reference), uint8 images out (what ``save_image`` writes) and, instead of int64 codes, one record
        ``u16 c<<12|h<<6|w`` + bit-packed code words per token: returns (out_images (n, c, h, w) uint8,
        records (n, s, record_bytes) uint8, counts (n,) int32 = valid records per image).  PCIe moves 1.65 MB per
        512^2 image instead of 6.6 MB; codes, token order and pixels are those of the default mode (pixels after
        the 8-bit quantisation).  Every image must fill its own row (k == max_seq_len, as in the benchmark
        configuration); use roundtrip() per chunk for packed rows."""
        assert not images.is_cuda, "roundtrip_host takes host tensors; use roundtrip() for device tensors"
        if compact:
            assert images.dtype == torch.uint8, "compact mode takes 8-bit pixels"
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        n = images.shape[0]
        if not hasattr(self, "_copy_streams"):
            self._copy_streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        s_in, s_out = self._copy_streams
        main = torch.cuda.current_stream(dev)
        s_in.wait_stream(main)
        s_out.wait_stream(main)
        out_dtype = torch.uint8 if compact else torch.float32
        for i in range(0, n, chunk):
            sl = slice(i, min(n, i + chunk))
            with torch.cuda.stream(s_in):
                x = images[sl].to(dev, non_blocking=True)
                ev_in = s_in.record_event()
            main.wait_event(ev_in)
            x.record_stream(main)
            batch, codes, rec = self._fused_roundtrip(x, None, out_dtype) if self.fusable() else (None, None, None)
            if batch is None:
                batch, codes = self.encode_codes(x)
                rec = self.decode_codes(batch, codes, out_dtype)
            assert codes.shape[0] == x.shape[0], "roundtrip_host needs one image per row (k == max_seq_len)"
            counts = None
            if compact:
                codes, counts = wire_records(batch, codes, self.quantizer.codebook_dim)
                counts = counts[:, 0].contiguous()
            if out_images is None:
                out_images = torch.empty((n,) + tuple(rec.shape[1:]), dtype=rec.dtype).pin_memory()
            if out_codes is None:
                out_codes = torch.empty((n,) + tuple(codes.shape[1:]), dtype=codes.dtype).pin_memory()
            if compact and out_counts is None:
                out_counts = torch.empty(n, dtype=torch.int32).pin_memory()
            ev_done = main.record_event()
            s_out.wait_event(ev_done)
            with torch.cuda.stream(s_out):
                out_images[sl].copy_(rec, non_blocking=True)
                out_codes[sl].copy_(codes, non_blocking=True)
                if compact:
                    out_counts[sl].copy_(counts, non_blocking=True)
                    counts.record_stream(s_out)
            rec.record_stream(s_out)
            codes.record_stream(s_out)
        main.wait_stream(s_out)
        return (out_images, out_codes, out_counts) if compact else (out_images, out_codes)


class GraphedRoundtrip:
    """The launch sequence of ``TransformPipeline.roundtrip`` for one batch shape, captured in a CUDA graph.

    The step is ~15 short kernels; launched from Python it costs about a millisecond of host time per
    call, which a graph replay removes (one driver call per step, no allocator traffic, launch gaps of a
    microsecond).  The graph reads ``images`` in place (the tensor given at capture time) and writes the
    same two output tensors on every replay; call it with another tensor of the same shape to have it
    copied in first.  The captured kernels read tables derived from the PatchNorm statistics (decode tables, the
    tame-divisor flag) that were built at capture time: after the statistics change (fitting, ``load_state_dict``,
    ``.to()``, ``invalidate_derived()``) a replay raises -- build a new graph.  Only for a fixed token count per image
    (``sample_patches_beta == 0`` or explicit ``ks``): the packing tables are part of the capture."""

    def __init__(self, pipe: "TransformPipeline", images: torch.Tensor, ks: Optional[Sequence[int]] = None,
                 fused: Optional[bool] = None, warmup: int = 2):
        from . import _lib
        assert images.is_cuda, "GraphedRoundtrip captures device work; use roundtrip_host for host tensors"
        assert ks is not None or pipe.extractor.sample_patches_beta <= 0.0, \
            "a captured step cannot redraw k: pass ks or use sample_patches_beta = 0"
        self.pipe, self.images = pipe, images
        cur = torch.cuda.current_stream(images.device)
        side = torch.cuda.Stream(images.device)
        side.wait_stream(cur)
        # The captured kernels hold raw pointers to the extractor's cached packing tables, which are created
        # outside the capture's private pool: keep a strong reference to every table handed out while warming
        # up and capturing, so that a later cache eviction cannot free memory the graph still reads.
        fe = pipe.extractor
        saved, fe._keepalive = fe._keepalive, []
        norm = pipe.norm
        try:
            with torch.cuda.stream(side):          # warm caches (tables, tensor maps, function attributes)
                for _ in range(warmup):
                    pipe.roundtrip(images, ks, fused)
            cur.wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count
            with torch.cuda.graph(self.graph):
                self.rec, self.codes = pipe.roundtrip(images, ks, fused)
            self._tables = list(fe._keepalive)
            # what the kernels read of the statistics' derived tables: kept alive here, valid while the statistics are
            self._derived = list(getattr(norm, "_derived", {}).values())
            self._stats_version = getattr(norm, "_stats_version", None)
        finally:
            fe._keepalive = saved
        self.launches = _lib.launch_count - l0     # kernels per replay
        self._lib = _lib

    def __call__(self, images: Optional[torch.Tensor] = None):
        if getattr(self.pipe.norm, "_stats_version", None) != self._stats_version:
            raise RuntimeError("the PatchNorm statistics changed since this graph was captured: its decode tables are stale; "
                               "capture a new one (TransformPipeline.graphed)")
        if images is not None and images.data_ptr() != self.images.data_ptr():
            assert images.shape == self.images.shape
            self.images.copy_(images)
        self.graph.replay()
        self._lib.launch_count += self.launches
        return self.rec, self.codes
