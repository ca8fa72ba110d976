"""The model glue on either side of the quantiser, on the device and on the packed token layout
(reference: modeling_dct_autoencoder.py:15-200).

``DCTAutoencoderGlue`` carries every learned tensor of the reference's ``DCTAutoencoder`` EXCEPT the two
transformer stacks, under the reference's parameter names (``to_patch_embedding.0.weight``,
``to_patch_embedding.1.weight/bias``, ``encoder_pos_embed_{channel,height,width}``, ``decoder_pos_embed_*``,
``proj_out.0.weight/bias``, ``proj_out.1.weight``, ``patchnorm.*``, ``vq_model.*``), so a reference checkpoint
loads with ``load_state_dict(strict=False)``.  Its methods are the reference's (``normalize_``,
``inv_normalize_``, ``add_pos_embedding_encoder_/decoder_``, ``encode``, ``decode``, ``decode_from_codes``,
``forward``, ``entropy_loss``) with the transformer passed in as a callable
``encoder(hidden (b, s, F), attention_mask) -> hidden`` (default: identity) -- the CLIPEncoder stacks are the
model itself and outside this path.

Inference arithmetic (every method runs under ``torch.no_grad``): the Linear layers go through the split-precision
tcgen05 GEMM (fp32-class accuracy), their operands prepared and their outputs finished by csrc/glue.cu (LayerNorm,
per-row power-of-two scaling, position-embedding gathers, bias).  The parameters are ordinary ``nn.Linear`` /
``nn.LayerNorm`` / ``nn.Parameter`` objects, so a trainer can differentiate the same layers with PyTorch autograd,
as the reference does.
"""
from typing import Callable, Optional

import torch
from torch import nn

from .dct_patches import DCTPatches
from .lfq import LFQ
from .linear import linear_bias_rows, linear_rows, ln_pos_rows
from .patchnorm import PatchNorm
from .util import compute_entropy_loss
from .vector_quantize import VectorQuantize

def _identity_stack(hidden: torch.Tensor, attention_mask=None) -> torch.Tensor:
    return hidden


class DCTAutoencoderGlue(nn.Module):
    def __init__(self, image_channels: int = 3, max_patch_h: int = 32, max_patch_w: int = 32, patch_size: int = 14,
                 feature_dim: int = 1024, vq_type: str = "lfq", vq_codebook_size: int = 8192, vq_num_codebooks: int = 16,
                 encoder: Optional[Callable] = None, decoder: Optional[Callable] = None):
        """Arguments follow ``DCTAutoencoderConfig`` (configuration_dct_autoencoder.py) with ``feature_dim`` =
        ``encoder_config.hidden_size``; ``encoder`` / ``decoder``: the transformer stacks as callables."""
        super().__init__()
        self.image_channels, self.max_patch_h, self.max_patch_w = image_channels, max_patch_h, max_patch_w
        self.patch_size, self.feature_dim, self.vq_type = patch_size, feature_dim, vq_type
        self.patchnorm = PatchNorm(max_patch_h=max_patch_h, max_patch_w=max_patch_w, patch_size=patch_size,
                                   channels=image_channels)
        patch_dim = patch_size ** 2
        self.encoder_pos_embed_channel = nn.Parameter(torch.randn(image_channels, feature_dim))
        self.encoder_pos_embed_height = nn.Parameter(torch.randn(max_patch_h, feature_dim))
        self.encoder_pos_embed_width = nn.Parameter(torch.randn(max_patch_w, feature_dim))
        self.decoder_pos_embed_channel = nn.Parameter(torch.randn(image_channels, feature_dim))
        self.decoder_pos_embed_height = nn.Parameter(torch.randn(max_patch_h, feature_dim))
        self.decoder_pos_embed_width = nn.Parameter(torch.randn(max_patch_w, feature_dim))
        self.to_patch_embedding = nn.Sequential(nn.Linear(patch_dim, feature_dim, bias=False),
                                                nn.LayerNorm(feature_dim, eps=1e-4))
        if vq_type == "lfq":
            self.vq_model = LFQ(dim=feature_dim, num_codebooks=vq_num_codebooks, codebook_size=vq_codebook_size)
        elif vq_type == "vq":
            self.vq_model = VectorQuantize(feature_dim, codebook_size=vq_codebook_size, heads=vq_num_codebooks,
                                           codebook_dim=16)
        else:
            raise ValueError(vq_type)
        self.proj_out = nn.Sequential(nn.LayerNorm(feature_dim, eps=1e-4), nn.Linear(feature_dim, patch_dim, bias=False))
        # not registered as sub-modules: the stacks are the caller's
        object.__setattr__(self, "encoder", encoder or _identity_stack)
        object.__setattr__(self, "decoder", decoder or _identity_stack)

    # ------------------------------------------------------------------ position embeddings
    def get_pos_embedding_decoder(self, dct_patches: DCTPatches) -> torch.Tensor:
        """modeling_dct_autoencoder.py:90-94."""
        zero = torch.zeros(dct_patches.patch_channels.shape + (self.feature_dim,), device=dct_patches.patch_channels.device)
        return ln_pos_rows(zero, pos=(self.decoder_pos_embed_channel, self.decoder_pos_embed_height,
                                      self.decoder_pos_embed_width),
                           channels=dct_patches.patch_channels, positions=dct_patches.patch_positions)

    def add_pos_embedding_decoder_(self, dct_patches: DCTPatches) -> DCTPatches:
        """modeling_dct_autoencoder.py:96-101 (in place on the container)."""
        dct_patches.patches = ln_pos_rows(dct_patches.patches, pos=(self.decoder_pos_embed_channel, self.decoder_pos_embed_height,
                                                                    self.decoder_pos_embed_width),
                                          channels=dct_patches.patch_channels, positions=dct_patches.patch_positions)
        return dct_patches

    def add_pos_embedding_encoder_(self, dct_patches: DCTPatches) -> DCTPatches:
        """modeling_dct_autoencoder.py:103-112."""
        dct_patches.patches = ln_pos_rows(dct_patches.patches, pos=(self.encoder_pos_embed_channel, self.encoder_pos_embed_height,
                                                                    self.encoder_pos_embed_width),
                                          channels=dct_patches.patch_channels, positions=dct_patches.patch_positions)
        return dct_patches

    # ------------------------------------------------------------------ normalisation
    @torch.no_grad()
    def normalize_(self, x: DCTPatches) -> DCTPatches:
        x.patches = self.patchnorm(x)
        return x

    def inv_normalize_(self, x: DCTPatches) -> DCTPatches:
        x.patches = self.patchnorm.inverse_norm(x)
        return x

    # ------------------------------------------------------------------ the two halves
    @torch.no_grad()
    def embed(self, dct_patches: DCTPatches) -> DCTPatches:
        """to_patch_embedding + add_pos_embedding_encoder_ (modeling_dct_autoencoder.py:138-141) in two passes:
        the p*p -> F GEMM, then LayerNorm + the three gathers in ONE kernel over the (rows, slots, F) output."""
        lin, ln = self.to_patch_embedding[0], self.to_patch_embedding[1]
        y = linear_rows(dct_patches.patches, lin.weight)
        dct_patches.patches = ln_pos_rows(y, ln=ln, pos=(self.encoder_pos_embed_channel, self.encoder_pos_embed_height,
                                                         self.encoder_pos_embed_width),
                                          channels=dct_patches.patch_channels, positions=dct_patches.patch_positions)
        return dct_patches

    @torch.no_grad()
    def encode(self, dct_patches: DCTPatches, do_normalize: bool = False):
        """modeling_dct_autoencoder.py:129-155 -> (dct_patches, codes, commit_loss, distances)."""
        if do_normalize:
            dct_patches = self.normalize_(dct_patches)
        dct_patches = self.embed(dct_patches)
        dct_patches.patches = self.encoder(dct_patches.patches, attention_mask=dct_patches.attn_mask) \
            if self.encoder is not _identity_stack else dct_patches.patches
        if self.vq_type == "vq":
            dct_patches.patches, codes, commit_loss = self.vq_model(dct_patches.patches, mask=~dct_patches.key_pad_mask)
            distances = None
        else:
            dct_patches.patches, codes, commit_loss, distances = self.vq_model(dct_patches.patches,
                                                                               mask=~dct_patches.key_pad_mask)
        return dct_patches, codes, commit_loss, distances

    @torch.no_grad()
    def decode(self, x: DCTPatches, do_inv_norm: bool = False) -> DCTPatches:
        """modeling_dct_autoencoder.py:165-178: + decoder position embeddings, decoder stack, proj_out (LayerNorm fused
        into the GEMM's operand preparation), optional inverse PatchNorm."""
        x = self.add_pos_embedding_decoder_(x)
        hidden = self.decoder(x.patches, attention_mask=x.attn_mask) if self.decoder is not _identity_stack else x.patches
        x.patches = linear_rows(hidden, self.proj_out[1].weight, ln=self.proj_out[0])
        if do_inv_norm:
            x = self.inv_normalize_(x)
        return x

    @torch.no_grad()
    def decode_from_codes(self, codes: torch.Tensor, do_inv_norm: bool = False, **dct_patches_kwargs) -> DCTPatches:
        """modeling_dct_autoencoder.py:157-163."""
        if self.vq_type == "vq":
            x = self.vq_model.get_output_from_indices(codes)
        else:
            x = self.vq_model.indices_to_codes(codes)
        x = DCTPatches(patches=x, **dct_patches_kwargs)
        return self.decode(x, do_inv_norm=do_inv_norm)

    @torch.no_grad()
    def forward(self, dct_patches: DCTPatches, do_normalize: bool = False):
        """modeling_dct_autoencoder.py:180-193."""
        dct_patches, codes, commit_loss, distances = self.encode(dct_patches, do_normalize=do_normalize)
        dct_patches = self.decode(dct_patches)
        return dict(dct_patches=dct_patches, commit_loss=commit_loss, codes=codes, distances=distances)

    def entropy_loss(self, distances: torch.Tensor, mask: torch.Tensor):
        """modeling_dct_autoencoder.py:195-199."""
        og = distances.dtype
        return compute_entropy_loss(distances.float(), mask).to(og)
