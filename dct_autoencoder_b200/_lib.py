"""ctypes binding of libdcta.so (include/dcta.h).  There is NO fallback: if the library is
missing or a call fails, the product path raises."""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdcta.so")

ABI_VERSION = 3


class DctaError(RuntimeError):
    pass


class Segment(ctypes.Structure):
    """dcta_segment"""
    _fields_ = [("row", c_int32), ("offset", c_int32), ("k", c_int32), ("image_id", c_int32),
                ("img", c_int64)]


P = c_void_p
# name -> argtypes (every function returns int unless listed in _RESTYPES)
SIGNATURES = {
    "dcta_last_error": [],
    "dcta_abi_version": [],
    "dcta_compiled_arch": [],
    "dcta_profile_begin": [P],
    "dcta_profile_end": [P, c_int, P, c_int],
    "dcta_basis_elems": [c_int, c_int, c_int],
    "dcta_basis_init": [c_int, c_int, c_int, P, P, P],
    "dcta_rgb_to_ipt": [P, P, c_int64, c_int64, P, P, P],
    "dcta_ipt_to_rgb": [P, P, c_int64, c_int64, P, P, P],
    "dcta_u8_to_unit_f32": [P, P, c_int64, P],
    "dcta_unit_f32_to_u8": [P, P, c_int64, P],
    "dcta_dct2_fwd": [P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "dcta_dct2_inv": [P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P],
    "dcta_gemm_split": [P, P, c_int, c_int64, c_int64, P, P, c_int, c_int64, c_int64, c_int, c_int64, P, c_float,
                        P, P, c_int64, c_int64, P],
    "dcta_lfq_project_sign": [P, P, c_int64, c_int64, P, P, c_int, c_int64, c_int, P, P, c_float, P, c_int64, P, P],
    "dcta_lfq_bits_to_codes": [P, c_int64, c_int, c_int, c_int, P, P],
    "dcta_split_f32": [P, P, P, c_int64, c_float, P],
    "dcta_split_planes_centered": [P, P, P, P, P, c_int64, c_int, c_int, P],
    "dcta_split_coef_planes": [P, P, P, P, c_int64, c_int, c_int, c_int64, c_int, c_int, P],
    "dcta_rgb_to_ipt_split": [P, P, P, P, P, c_int64, c_int, c_int, P, P, P],
    "dcta_unpatchify_split": [P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int64, c_int, c_int,
                              P, P, P, P],
    "dcta_dct2_fwd_tc": [P, P, P, P, P, P, P, P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int64, c_int,
                         c_int, P],
    "dcta_dct2_inv_tc": [P, P, P, P, P, P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int64, c_int64, P],
    "dcta_fold_supported": [c_int, c_int, c_int, c_int],
    "dcta_fold_codes_supported": [c_int, c_int, c_int, c_int, c_int],
    "dcta_rgb_to_ipt_fold": [P, P, P, P, P, c_int64, c_int, c_int, P, P, P],
    "dcta_rgb_u8_to_ipt_fold": [P, P, P, P, P, c_int64, c_int, c_int, P, P, P],
    "dcta_fold_planes": [P, P, P, P, P, c_int64, c_int, c_int, P],
    "dcta_dct2_fwd_fold": [P, P, P, P, P, P, P, P, P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "dcta_dct2_fwd_fold_codes": [P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_float, c_float, c_float, P,
                                 c_int, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "dcta_pack_codes_grid": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, c_int, c_float, c_float,
                             c_float, c_int, c_int, P, P, P, P, P, P, P],
    "dcta_sort_tokens_maxabs": [P, P, P, c_int64, c_int, c_int, c_int, c_float, P, P],
    "dcta_unpatchify_fold": [P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P],
    "dcta_decode_codes_fold": [P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                               P, P, c_int, c_int, c_float, c_int, c_int, c_float, P, P, P, P, P],
    "dcta_decode_codes_inv_fold_scratch_bytes": [c_int64, c_int, c_int, c_int],
    "dcta_decode_codes_inv_fold_supported": [c_int, c_int, c_int, c_int, c_int, c_int, c_int],
    "dcta_decode_codes_inv_fold": [P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   P, P, c_int, c_int, c_float, c_int, c_int, c_float, P, P, P, P, P, P, P, P, P, P, P],
    "dcta_decode_grid_inv_fold": [P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_int, c_int, c_float, c_float,
                                  P, P, P, P, P, P, P, P, P, P, P],
    "dcta_decode_gen_tables_bytes": [c_int, c_int, c_int],
    "dcta_decode_gen_tables": [P, P, c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_float, P, P],
    "dcta_fold_coef_planes": [P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P],
    "dcta_dct2_inv_fold": [P, P, P, P, P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P],
    "dcta_unfold_ipt_to_rgb": [P, P, P, c_int64, c_int, c_int, P, P, P],
    "dcta_unfold_ipt_to_rgb_u8": [P, P, P, c_int64, c_int, c_int, P, P, P],
    "dcta_unfold_planes": [P, P, P, c_int64, c_int, c_int, P],
    "dcta_patchify": [P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "dcta_tile_scores": [P, P, c_int64, c_int, c_int, c_int, c_int, c_float, P, P],
    "dcta_sort_tokens": [P, P, c_int64, c_int, P],
    "dcta_pack_tiles": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P],
    "dcta_pack_lists": [P, P, P, P, P, c_int, c_int, c_int, P, P, P, P, P, P],
    "dcta_patchnorm_apply": [P, P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_int, P],
    "dcta_patchnorm_build_lists": [P, P, P, c_int64, c_int, c_int, c_int, P, P, P, P, P],
    "dcta_patchnorm_batch_median": [P, P, P, c_int, c_int, P, P],
    "dcta_patchnorm_update_median": [P, P, P, c_int, c_int, P],
    "dcta_patchnorm_abs_dev": [P, P, P, P, c_int, c_int, P, P],
    "dcta_patchnorm_update_b": [P, P, P, P, c_int, c_int, P],
    "dcta_zero_padding": [P, P, P, c_int64, c_int, P],
    "dcta_lfq_quantize": [P, P, P, c_int64, c_int, c_int, c_float, P],
    "dcta_lfq_indices_to_codes": [P, P, c_int64, c_int, c_int, c_float, P],
    "dcta_lfq_commit_loss": [P, P, P, P, c_int64, c_int, c_float, P],
    "dcta_lfq_distance": [P, P, c_int64, c_int, c_int, c_float, P],
    "dcta_entropy_loss": [P, P, P, P, c_int64, c_int, c_int, c_float, c_float, P],
    "dcta_perplexity": [P, c_int64, c_int, c_int64, P, P, P],
    "dcta_vq_nearest": [P, P, P, P, P, c_int64, c_int, c_int, P],
    "dcta_pack_codes_lfq": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_int, c_int, c_float,
                            c_float, c_float, c_int, c_int, c_float, P, P, P, P, P, P, P],
    "dcta_decode_codes_split": [P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int64, c_int, c_int,
                                P, P, c_int, c_int, c_float, c_int, c_int, c_float, P, P, P, P],
    "dcta_lfq_entropy_ctas": [],
    "dcta_lfq_entropy_factorized": [P, P, c_int64, c_int, c_int, c_float, c_float, c_float, P, P, P, P],
    "dcta_lfq_entropy_factorized_backward": [P, P, c_int64, c_int, c_int, c_float, c_float, P, P, P, P, P],
    "dcta_lfq_commit_backward": [P, P, P, P, P, c_int64, c_int, c_float, P],
    "dcta_ln_pos_rows": [P, P, P, c_float, P, P, P, P, P, P, P, c_int64, c_int, P],
    "dcta_split_rows_rowscale": [P, P, P, c_float, P, P, P, c_float, c_int64, c_int, c_int64, P],
    "dcta_split_rows_patchnorm": [P, P, P, P, P, P, c_int, c_int, c_int, c_float, c_float, c_float, P, P, P, c_float, c_int64,
                                  c_int, c_int64, P],
    "dcta_pack_tiles_index": [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int64, P, P, P, P, P, P],
    "dcta_unpatchify_denorm_fold": [P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_int,
                                    c_int, c_float, P, P, P, P],
    "dcta_row_sumsq": [P, P, c_int64, c_int, P],
    "dcta_vq_nearest_tc": [P, P, P, P, P, P, P, P, P, c_int64, c_int, c_int, c_int64, P],
    "dcta_vq_nearest_tc_masked": [P, P, P, P, P, P, P, P, P, P, P, P, c_int64, c_int, c_int, c_int64, P],
    "dcta_resize_bilinear_aa": [P, P, c_int64, c_int, c_int, c_int, c_int, P],
    "dcta_resize_bilinear_aa_u8": [P, P, c_int64, c_int, c_int, c_int, c_int, P],
    "dcta_vq_cluster_stats": [P, P, P, c_int64, c_int, c_int, P, P, P],
    "dcta_vq_ema_update": [P, P, P, P, P, c_int, c_int, c_float, c_float, P, P],
    "dcta_masked_mse_scratch_floats": [],
    "dcta_lfq_entropy_use_tensor_cores": [c_int],
    "dcta_masked_mse": [P, P, P, c_int64, c_int, P, P, P],
    "dcta_masked_mse_backward": [P, P, P, P, P, P, P, c_int64, c_int, P],
    "dcta_vq_kmeans_means": [P, P, P, c_int, c_int, P],
    "dcta_build_slot_map": [P, P, P, P, P, c_int, c_int, c_int64, c_int, c_int, c_int, P, P],
    "dcta_unpatchify": [P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, P, P],
    "dcta_wire_pack": [P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P, P, P],
    "dcta_wire_unpack": [P, c_int64, c_int, c_int, c_int, P, P, P, P],
}
_RESTYPES = {"dcta_last_error": c_char_p, "dcta_basis_elems": c_int64,
             "dcta_decode_codes_inv_fold_scratch_bytes": c_int64, "dcta_decode_gen_tables_bytes": c_int64}
BASIS_F32, BASIS_SPLIT_FWD, BASIS_SPLIT_INV, BASIS_FOLD_FWD, BASIS_FOLD_INV = range(5)
REDUCE_SCRATCH = 2048  # DCTA_REDUCE_SCRATCH

# kernels launched by one call of each entry point (memsets not counted)
KERNELS_PER_CALL = {
    "dcta_rgb_to_ipt": 1, "dcta_ipt_to_rgb": 1, "dcta_u8_to_unit_f32": 1, "dcta_unit_f32_to_u8": 1, "dcta_dct2_fwd": 2, "dcta_dct2_inv": 2, "dcta_patchify": 1,
    "dcta_tile_scores": 1, "dcta_sort_tokens": 1, "dcta_pack_tiles": 1, "dcta_pack_lists": 1,
    "dcta_patchnorm_apply": 1, "dcta_patchnorm_build_lists": 4, "dcta_patchnorm_batch_median": 1,
    "dcta_patchnorm_update_median": 1, "dcta_patchnorm_abs_dev": 1, "dcta_patchnorm_update_b": 2,
    "dcta_zero_padding": 1, "dcta_lfq_quantize": 1, "dcta_lfq_indices_to_codes": 1, "dcta_lfq_commit_loss": 2,
    "dcta_lfq_distance": 1, "dcta_entropy_loss": 2, "dcta_perplexity": 2, "dcta_vq_nearest": 2,
    "dcta_build_slot_map": 1, "dcta_unpatchify": 1, "dcta_wire_pack": 1, "dcta_wire_unpack": 1,
    "dcta_gemm_split": 1, "dcta_lfq_project_sign": 1, "dcta_lfq_bits_to_codes": 1, "dcta_split_f32": 1, "dcta_rgb_to_ipt_split": 2, "dcta_unpatchify_split": 1,
    "dcta_split_planes_centered": 2, "dcta_split_coef_planes": 1,
    "dcta_dct2_fwd_tc": 2, "dcta_dct2_inv_tc": 2,
    "dcta_ln_pos_rows": 1, "dcta_split_rows_rowscale": 1, "dcta_split_rows_patchnorm": 1, "dcta_unpatchify_denorm_fold": 1, "dcta_pack_tiles_index": 1, "dcta_lfq_entropy_ctas": 0, "dcta_lfq_entropy_factorized": 2,
    "dcta_lfq_entropy_factorized_backward": 1, "dcta_lfq_commit_backward": 2,
    "dcta_row_sumsq": 1, "dcta_vq_nearest_tc": 2, "dcta_vq_nearest_tc_masked": 2, "dcta_vq_cluster_stats": 1, "dcta_resize_bilinear_aa": 1, "dcta_resize_bilinear_aa_u8": 1, "dcta_vq_ema_update": 2, "dcta_vq_kmeans_means": 1, "dcta_masked_mse": 2, "dcta_masked_mse_backward": 1, "dcta_masked_mse_scratch_floats": 0, "dcta_lfq_entropy_use_tensor_cores": 0,
    "dcta_pack_codes_lfq": 2, "dcta_decode_codes_split": 1,
    "dcta_fold_supported": 0, "dcta_fold_codes_supported": 0, "dcta_rgb_to_ipt_fold": 2, "dcta_rgb_u8_to_ipt_fold": 2, "dcta_unfold_ipt_to_rgb_u8": 1, "dcta_fold_planes": 3, "dcta_dct2_fwd_fold": 2,
    "dcta_unpatchify_fold": 1, "dcta_decode_codes_fold": 2, "dcta_decode_codes_inv_fold": 3, "dcta_decode_grid_inv_fold": 3, "dcta_decode_gen_tables": 1, "dcta_decode_gen_tables_bytes": 0,
    "dcta_decode_codes_inv_fold_scratch_bytes": 0, "dcta_decode_codes_inv_fold_supported": 0, "dcta_fold_coef_planes": 1, "dcta_dct2_inv_fold": 2,
    "dcta_unfold_ipt_to_rgb": 1, "dcta_unfold_planes": 1, "dcta_sort_tokens_maxabs": 1, "dcta_dct2_fwd_fold_codes": 3, "dcta_pack_codes_grid": 2,
}
launch_count = 0

_lib = None


def load():
    """Loads libdcta.so once; raises DctaError if it is absent or its ABI differs."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DctaError(
            f"{LIB_PATH} not found: build it with `python dct_autoencoder_b200/csrc/build.py` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    if lib.dcta_abi_version() != ABI_VERSION:
        raise DctaError(f"libdcta.so ABI {lib.dcta_abi_version()} != binding {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def last_error() -> str:
    return load().dcta_last_error().decode()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DctaError("libdcta kernels need CUDA tensors (no CPU fallback); got a tensor on " + str(t.device))


def call(name: str, *args):
    """Calls an entry point; raises DctaError with dcta_last_error() on a non-zero status."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    launch_count += KERNELS_PER_CALL.get(name, 0)
    if rc != 0:
        raise DctaError(f"{name} failed ({rc}): {lib.dcta_last_error().decode()}")


def host_floats(values):
    """A ctypes float array (host pointer argument)."""
    vals = [float(v) for v in values]
    return (c_float * len(vals))(*vals)


_profile_depth = 0


def profile_active() -> bool:
    """True inside a ``profile`` block: per-launch times are differences of events on ONE stream, so callers that
    would spread their launches over two streams keep them on one meanwhile."""
    return _profile_depth > 0


class profile:
    """``with _lib.profile(device) as p: ...`` -> ``p.groups``: [(launch group name, device ms)] of every libdcta launch
    group inside the block, from CUDA events recorded by the library on the launching stream (dcta_profile_begin / _end)."""

    def __init__(self, device=None):
        self.device = device
        self.groups = []

    def __enter__(self):
        global _profile_depth
        _profile_depth += 1
        load().dcta_profile_begin(stream_ptr(self.device))
        return self

    def __exit__(self, *exc):
        global _profile_depth
        _profile_depth -= 1
        cap, n_max = 1 << 16, 4096
        names = ctypes.create_string_buffer(cap)
        ms = (c_float * n_max)()
        n = load().dcta_profile_end(names, cap, ms, n_max)
        if n < 0:
            raise DctaError("dcta_profile_end failed: " + last_error())
        nm = names.value.decode().split("\n") if n else []
        self.groups = [(nm[i], float(ms[i])) for i in range(min(n, n_max, len(nm)))]
        return False
