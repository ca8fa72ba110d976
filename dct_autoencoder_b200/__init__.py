"""B200-native encode/decode transform path of the DCT autoencoder.

Public surface (same names as the reference package ``dct_autoencoder``):
``DCTAutoencoderFeatureExtractor``, ``DCTPatches``, ``PatchNorm``, ``LFQ``, ``VectorQuantize`` and the
``util`` functions on the path.  Everything computes in hand-written sm_100a kernels reached
through the C ABI of ``libdcta.so`` (include/dcta.h); there is no CPU or PyTorch fallback.
"""
from . import _lib, shards, util
from .dct_patches import DCTPatches, from_bytes, from_dict, to_bytes, to_dict
from .feature_extraction_dct_autoencoder import DCTAutoencoderFeatureExtractor
from .lfq import LFQ
from .modeling_dct_autoencoder import DCTAutoencoderGlue
from .patchnorm import PatchNorm
from .pipeline import GraphedRoundtrip, TransformPipeline, dict_collate, get_max_seq_length
from .vector_quantize import VectorQuantize

__all__ = [
    "DCTAutoencoderFeatureExtractor", "DCTPatches", "PatchNorm", "LFQ", "VectorQuantize", "DCTAutoencoderGlue",
    "TransformPipeline", "GraphedRoundtrip", "dict_collate", "get_max_seq_length", "to_dict", "from_dict", "to_bytes", "from_bytes", "util", "shards",
]
