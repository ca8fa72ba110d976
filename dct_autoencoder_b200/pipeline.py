"""The benchmarked pipeline: image -> DCT tokens -> PatchNorm -> quantiser -> inverse PatchNorm ->
image, assembled from the drop-in classes exactly as a user of the reference would
(main.py:183-190 normalize_, modeling_dct_autoencoder.py:149-153 quantiser call with
``mask=~key_pad_mask``, main.py:96-98 inv_normalize_ + postprocess), minus the transformer."""
import math
from typing import List, Optional, Sequence

import torch

from .dct_patches import DCTPatches
from .feature_extraction_dct_autoencoder import DCTAutoencoderFeatureExtractor
from .patchnorm import PatchNorm
from .util import power_of_two


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

    @torch.no_grad()
    def fit_norm(self, images: torch.Tensor, ks: Optional[Sequence[int]] = None) -> None:
        """One statistic-fitting step (main.py:115-149 train_patch_norm body), then freeze."""
        was_training, was_frozen = self.norm.training, self.norm.frozen
        self.norm.train()
        self.norm.frozen = False
        self.norm(self.extractor.process_batch(images, ks))
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

    @torch.no_grad()
    def roundtrip(self, images: torch.Tensor, ks: Optional[Sequence[int]] = None):
        batch, q, codes = self.encode(images, ks)
        return self.decode(batch, q), codes
