"""ctypes binding of libgpnerf_b200.so (include/gpnerf_abi.h).

There is no CPU or PyTorch fallback: if the library has not been built, or a
kernel launch fails, this module raises.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` or
``python gp-nerf_b200/build.py``.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libgpnerf_b200.so")

MAX_VIEWS = 8
N_LEVELS = 4
CNT_PIX, CNT_RAYS, CNT_P1, CNT_P2, N_COUNTERS = 0, 1, 2, 3, 8
PREC_FP32, PREC_BF16 = 0, 1


class Frame(C.Structure):
    """gpnerf_frame_t"""
    _fields_ = [
        ("R", C.c_float * 9), ("Th", C.c_float * 3), ("bounds_min", C.c_float * 3),
        ("voxel_size", C.c_float * 3), ("out_sh", C.c_int32 * 3),
        ("level_dims", (C.c_int32 * 3) * N_LEVELS),
        ("target_pose", C.c_float * 12), ("target_K", C.c_float * 9), ("target_K_inv", C.c_float * 9),
        ("H", C.c_int32), ("W", C.c_int32),
        ("n_views", C.c_int32), ("src_KE", (C.c_float * 16) * MAX_VIEWS),
        ("src_h", C.c_int32), ("src_w", C.c_int32), ("feat_h", C.c_int32), ("feat_w", C.c_int32),
        ("n_samples", C.c_int32), ("neg_ray", C.c_int32), ("mask_threshold", C.c_float),
        ("rank", C.c_int32), ("world", C.c_int32), ("tile_px", C.c_int32),
        ("reserved_", C.c_int32 * 2), ("self_dev", C.c_uint64),
    ]


MAX_PEERS = 8


class Peer(C.Structure):
    """gpnerf_peer_t"""
    _fields_ = [
        ("n_dst", C.c_int32), ("n_flag", C.c_int32),
        ("dst_img", (C.c_uint64 * MAX_PEERS) * 2), ("dst_hit", (C.c_uint64 * MAX_PEERS) * 2),
        ("dst_flag", C.c_uint64 * MAX_PEERS), ("ticket", C.c_uint64), ("seq", C.c_uint64),
    ]


class HeadWeights(C.Structure):
    """gpnerf_head_weights_t"""
    _fields_ = [
        ("geo_w", C.c_void_p), ("geo_b", C.c_void_p),
        ("den_w", C.c_void_p * 4), ("den_b", C.c_void_p * 4),
        ("base_w", C.c_void_p * 2), ("base_b", C.c_void_p * 2),
        ("vis_w", C.c_void_p * 2), ("vis_b", C.c_void_p * 2),
        ("rgb_w", C.c_void_p * 3), ("rgb_b", C.c_void_p * 3),
        ("tc_image", C.c_void_p),
    ]


_P = C.c_void_p
_I = C.c_int
_SIGNATURES = {
    "gpnerf_abi_version": ([], C.c_int),
    "gpnerf_last_error": ([], C.c_char_p),
    "gpnerf_sm_count": ([], C.c_int),
    "gpnerf_struct_bytes": ([_I], C.c_int),
    "gpnerf_workspace_bytes": ([C.c_int64], C.c_int64),
    "gpnerf_k0_level_to_channels_last": ([_P, _I, _I, _I, _I, _I, _P, _P, _P], C.c_int),
    "gpnerf_k0_products_to_f16": ([C.POINTER(_P), C.POINTER(C.c_int32 * 3), _P, _I, _I, _I, C.POINTER(_P),
                                   C.POINTER(_P), _P, _P], C.c_int),
    "gpnerf_k0_sparse_to_f16": ([C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_int32), C.POINTER(_P), _I,
                                 C.POINTER(C.c_int32 * 3), C.POINTER(_P), C.POINTER(_P), _P], C.c_int),
    "gpnerf_k0_sparse_to_f32": ([C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_int32), C.POINTER(_P), _I,
                                 C.POINTER(C.c_int32 * 3), C.POINTER(_P), C.POINTER(_P), _P], C.c_int),
    "gpnerf_sc_index_input": ([_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P], C.c_int),
    "gpnerf_sc_gather_rows": ([_P, _I, _P, _P, _I, _P, _P], C.c_int),
    "gpnerf_sc_strided_sites": ([_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P], C.c_int),
    "gpnerf_sc_neighbours": ([_P, _P, _I, _I, _P, _I, _I, _I, _P, _P, _P], C.c_int),
    "gpnerf_sc_conv": ([_P, _I, _P, _P, _I, _P, _P, _P, _I, _P, _P], C.c_int),
    "gpnerf_sc_conv_tc": ([_P, _I, _P, _P, _I, _P, _P, _P, _I, _P, _P, _P], C.c_int),
    "gpnerf_sc_gather_rows_split": ([_P, _I, _P, _P, _I, _P, _P], C.c_int),
    "gpnerf_attn_smpl_code": ([_P, _P, C.c_longlong, C.c_longlong, _I, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P],
                              C.c_int),
    "gpnerf_k9_instance_norm_act": ([_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, C.c_float, _I, _P, _I, _P, _I, _I, _I, _P],
                                    C.c_int),
    "gpnerf_k9_resample_pad": ([_P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P], C.c_int),
    "gpnerf_k0_build_masks3d": ([C.POINTER(_P), C.POINTER(Frame), _P, _P], C.c_int),
    "gpnerf_k0_featmaps_to_channels_last": ([_P, _I, _I, _I, _I, _I, _P, _P], C.c_int),
    "gpnerf_k0_images_to_rgbx": ([_P, _I, _I, _I, _I, _I, _P, _P], C.c_int),
    "gpnerf_k1_voxel_pixel_mask": ([_P, C.POINTER(Frame), _P, _P, _P], C.c_int),
    "gpnerf_k1_rays_bbox": ([_P, _P, C.POINTER(Frame), _P, _P, _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "gpnerf_k1_dataset_rays": ([_P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "gpnerf_k2_occupancy_compact": ([_P, _P, _P, _P, _P, _P, _P, C.POINTER(Frame), _I, _P, _P, _P, _P, _P, _P],
                                    C.c_int),
    "gpnerf_k2_gather_volume": ([C.POINTER(_P), _I, _P, _P, _P, _P, _P, C.POINTER(Frame), _I, _P, _P, _P], C.c_int),
    "gpnerf_k2_project_gather_meanvar": ([_P, _P, _I, _P, _P, _P, _P, _P, C.POINTER(Frame), _I, _P, _P, _P, _P, _P], C.c_int),
    "gpnerf_k2_mean_variance": ([_P, _I, _I, _P, _P], C.c_int),
    "gpnerf_k3_density_mlp": ([_P, _I, _P, _P, C.POINTER(HeadWeights), _I, _I, _P, _I, _P, _P, _I, _P], C.c_int),
    "gpnerf_k3_color_mlp": ([_P, _P, _P, C.POINTER(HeadWeights), _I, _I, _P, _I, _P, _I, _P], C.c_int),
    "gpnerf_k3_packed_weight_bytes": ([], C.c_int64),
    "gpnerf_k3_pack_weights": ([C.POINTER(HeadWeights), _I, _P, _P], C.c_int),
    "gpnerf_k23_record_bytes": ([_I], C.c_int64),
    "gpnerf_k23_gather_density_tc": ([C.POINTER(_P), _P, _P, _P, _P, _P, _P, C.POINTER(Frame), C.POINTER(HeadWeights),
                                      _I, _P, _P, _P, _P, _P, _P], C.c_int),
    "gpnerf_k23_tile_record_bytes": ([_I], C.c_int64),
    "gpnerf_k23_gather_density_tiles_tc": ([C.POINTER(_P), _P, _P, _P, _P, _P, _P, C.POINTER(Frame),
                                            C.POINTER(HeadWeights), _I, _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "gpnerf_k3_color_tiles_tc": ([_P, _P, C.POINTER(HeadWeights), _I, _I, _P, _I, _P, _P], C.c_int),
    "gpnerf_k3_color_mlp_records": ([_P, _P, C.POINTER(HeadWeights), _I, _I, _P, _I, _P, _P], C.c_int),
    "gpnerf_k3_color_gather_tc": ([_P, _P, _P, _P, _P, _P, _P, C.POINTER(Frame), C.POINTER(HeadWeights), _I, _P, _I, _P, _P,
                                   _P], C.c_int),
    "gpnerf_k2_gather_volume_bwd": ([C.POINTER(_P), _I, _P, _P, _P, _P, _P, C.POINTER(Frame), _I, _P, _P, _P], C.c_int),
    "gpnerf_k2_project_gather_bwd": ([_P, _I, _P, _P, _P, _P, _P, C.POINTER(Frame), _I, _P, _P, _P], C.c_int),
    "gpnerf_k6_linear": ([_P, _I, _I, C.c_float, _P, _I, _P, _I, _I, _I, _P, _I, _P, _I, _P, _I, _I, _I,
                          C.c_longlong, _I, _P], C.c_int),
    "gpnerf_k6_grad_weights": ([_P, _I, _I, C.c_float, _P, _I, _I, _P, _I, _P, _I, _P, C.c_longlong, _I, _P],
                               C.c_int),
    "gpnerf_k6_assemble_raw": ([_P, _P, _P, _I, C.c_longlong, _P, _P], C.c_int),
    "gpnerf_k6_raw_grad_split": ([_P, _P, _P, _P, _I, C.c_longlong, _P, _P, _P], C.c_int),
    "gpnerf_k6_meanvar_bwd": ([_P, _P, _P, _I, C.c_longlong, _P, _P], C.c_int),
    "gpnerf_k6_from_channels_last": ([_P, _I, C.c_longlong, _P, _P], C.c_int),
    "gpnerf_k4_compact_alpha": ([_P, _I, _P, _P, _P, _P, _P], C.c_int),
    "gpnerf_k5_composite": ([_P, _P, _P, _P, _P, C.POINTER(Frame), C.c_float, _P, _P, _P, _P, _P], C.c_int),
    "gpnerf_peer_wait": ([_P, _I, _I, _P, _P], C.c_int),
    "gpnerf_peer_handle_bytes": ([], C.c_int),
    "gpnerf_peer_alloc": ([C.c_int64, C.POINTER(_P), _P], C.c_int),
    "gpnerf_peer_open": ([_P, C.POINTER(_P)], C.c_int),
    "gpnerf_peer_close": ([_P], C.c_int),
    "gpnerf_peer_free": ([_P], C.c_int),
    "gpnerf_k5_raw2outputs": ([_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "gpnerf_k5_raw2outputs_bwd": ([_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P], C.c_int),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class GpnerfError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it is missing – the product
    path has no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpnerfError(
            f"{LIB_PATH} not found: build the CUDA library first "
            "(python gp-nerf_b200/build.py).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = restype
    if lib.gpnerf_abi_version() != 2:
        raise GpnerfError("libgpnerf_b200.so ABI version mismatch")
    for which, st in enumerate((Frame, HeadWeights, Peer)):
        if lib.gpnerf_struct_bytes(which) != C.sizeof(st):
            raise GpnerfError(f"{st.__name__}: ctypes layout ({C.sizeof(st)} B) differs from the library's "
                              f"({lib.gpnerf_struct_bytes(which)} B)")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().gpnerf_last_error().decode(errors="replace")
        raise GpnerfError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None → NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "kernels take contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        assert t.is_cuda and t.is_contiguous()
        arr[i] = t.data_ptr()
    return arr
