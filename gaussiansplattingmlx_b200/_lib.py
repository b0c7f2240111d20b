"""ctypes binding of ``libgsb.so`` — the C ABI declared in ``include/gsb.h``.

This is the same stub a Swift host would get from importing ``gsb.h`` through a module map (see
``INTEGRATION.md``); Python is used here because the image has no Swift toolchain.  There is no
fallback: if the shared library is missing or no B200 is visible, loading / ``gsb_create`` fails
loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libgsb.so"

GSB_OK = 0
GSB_ERR_INVALID, GSB_ERR_CUDA, GSB_ERR_UNSUPPORTED, GSB_ERR_STATE, GSB_ERR_CAPACITY = -1, -2, -3, -4, -5
GSB_FLAG_SORT_CUB = 1
GSB_FLAG_NO_OVERLAP = 2
GSB_FLAG_ASYNC_LOSS = 4
GSB_FLAG_NO_SEGMENTS = 8
GSB_FLAG_NVTX = 16
GSB_PEER_BLOB_BYTES = 256
STAGE_COUNT = 12


class GsbConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("tile_w", C.c_int32), ("tile_h", C.c_int32),
                ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32), ("white_background", C.c_int32),
                ("max_gaussians", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32),
                ("lambda_dssim", C.c_float), ("adam_beta1", C.c_float), ("adam_beta2", C.c_float),
                ("adam_eps", C.c_float)]


class GsbCamera(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("cam_center", C.c_float * 3),
                ("fov_x", C.c_float), ("fov_y", C.c_float), ("focal_x", C.c_float), ("focal_y", C.c_float)]


class GsbStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("pairs_last_view", C.c_uint64), ("pairs_total", C.c_uint64),
                ("views", C.c_uint64), ("pair_capacity", C.c_uint64), ("stage_ms", C.c_double * 16),
                ("stage_calls", C.c_uint64 * 16), ("sb_pairs_last_view", C.c_uint64)]


class GsbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"gsb error {code}: {message}")
        self.code = code


_P = C.c_void_p
_I = C.c_int32
_F = C.c_float

# name -> (restype, argtypes).  Every entry must be declared in include/gsb.h (tests check both ways).
SIGNATURES = {
    "gsb_abi_version": (C.c_int, []),
    "gsb_default_config": (None, [C.POINTER(GsbConfig)]),
    "gsb_create": (C.c_int, [C.POINTER(GsbConfig), C.POINTER(_P)]),
    "gsb_destroy": (None, [_P]),
    "gsb_last_error": (C.c_char_p, [_P]),
    "gsb_set_stream": (C.c_int, [_P, _P]),
    "gsb_synchronize": (C.c_int, [_P]),
    "gsb_set_flags": (C.c_int, [_P, C.c_int32]),
    "gsb_activate_fwd": (C.c_int, [_P, _I] + [_P] * 9),
    "gsb_activate_bwd": (C.c_int, [_P, _I] + [_P] * 12),
    "gsb_project_fwd": (C.c_int, [_P, _I] + [_P] * 4 + [C.POINTER(GsbCamera)] + [_P] * 8),
    "gsb_project_bwd": (C.c_int, [_P, _I] + [_P] * 4 + [C.POINTER(GsbCamera)] + [_P] * 10),
    "gsb_bin": (C.c_int, [_P, _I] + [_P] * 7 + [C.POINTER(C.c_uint32)]),
    "gsb_bin_read": (C.c_int, [_P] + [_P] * 6),
    "gsb_sort_tile_keys": (C.c_int, [_P, C.c_uint32, C.c_uint32] + [_P] * 6 + [_I]),
    "gsb_raster_fwd": (C.c_int, [_P, _I] + [_P] * 5),
    "gsb_raster_bwd": (C.c_int, [_P, _I] + [_P] * 9),
    "gsb_ssim_fwd": (C.c_int, [_P, _I, _I, _I] + [_P] * 8),
    "gsb_ssim_bwd": (C.c_int, [_P, _I, _I, _I] + [_P] * 4),
    "gsb_render_forward": (C.c_int, [_P, _I] + [_P] * 6 + [C.POINTER(GsbCamera)] + [_P] * 5),
    "gsb_render_backward": (C.c_int, [_P] + [_P] * 9 + [_I]),
    "gsb_loss_fwd_bwd": (C.c_int, [_P, _P, _P, _F, _P, _P]),
    "gsb_loss_fwd_bwd_depth": (C.c_int, [_P, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P]),
    "gsb_adam_step": (C.c_int, [_P, _I, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P),
                                C.POINTER(C.c_int64), C.POINTER(_F), _P]),
    "gsb_trainer_init": (C.c_int, [_P, _I] + [_P] * 6),
    "gsb_trainer_param_ptrs": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "gsb_trainer_grad_block": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "gsb_trainer_accumulate": (C.c_int, [_P, _I, C.POINTER(GsbCamera), C.POINTER(_P), _I, _I, _F, _P]),
    "gsb_trainer_accumulate_depth": (C.c_int, [_P, _I, C.POINTER(GsbCamera), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _F, _I, _I,
                                               _F, _P]),
    "gsb_trainer_apply": (C.c_int, [_P, _I, _I, _I]),
    "gsb_trainer_peers_export": (C.c_int, [_P, _P, C.c_int64]),
    "gsb_trainer_peers_import": (C.c_int, [_P, _I, _I, _P, C.c_int64]),
    "gsb_trainer_apply_peers": (C.c_int, [_P, _I, _I, _I]),
    "gsb_trainer_peers_close": (C.c_int, [_P]),
    "gsb_trainer_step_peers": (C.c_int, [_P, _I, C.POINTER(GsbCamera), C.POINTER(_P), _I, _F, _I, _I, _I, _P]),
    "gsb_trainer_peers_check": (C.c_int, [_P]),
    "gsb_trainer_peers_tune": (C.c_int, [_P, _I, _I, _I]),
    "gsb_trainer_attach_symmetric": (C.c_int, [_P, _I, _I, _P, _P, _P, _P, C.c_int64]),
    "gsb_trainer_apply_multicast": (C.c_int, [_P, _I, _I, _I]),
    "gsb_train_step": (C.c_int, [_P, _I, C.POINTER(GsbCamera), C.POINTER(_P), _I, _I, _I, C.POINTER(_F)]),
    "gsb_densify_classify": (C.c_int, [_P, _I, _P, _F, _P, _P, _F, _F, _F, _I, _P, _P]),
    "gsb_densify_map": (C.c_int, [_P, _I, _P, _P, _P, _I, _P, _P, C.POINTER(_I)]),
    "gsb_densify_apply": (C.c_int, [_P, _I, _P, _P, _P, C.c_uint64] + [_P] * 12),
    "gsb_trainer_densify": (C.c_int, [_P, _F, _F, _F, _I, C.c_uint64, _P, C.POINTER(_I)]),
    "gsb_trainer_count": (C.c_int, [_P, C.POINTER(_I), C.POINTER(_I)]),
    "gsb_last_contrib_sum": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "gsb_tile_list_info": (C.c_int, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "gsb_stats_reset": (C.c_int, [_P]),
    "gsb_stats_get": (C.c_int, [_P, C.POINTER(GsbStats)]),
    "gsb_enable_stage_timing": (C.c_int, [_P, _I]),
    "gsb_stage_name": (C.c_char_p, [_I]),
    "gsb_stage_section": (C.c_char_p, [_I]),
    "gsb_bin_generation": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
}

_lib = None


def load() -> C.CDLL:
    """Load ``libgsb.so`` (built in-tree by ``gaussiansplattingmlx_b200.build``).  No fallback."""
    global _lib
    if _lib is None:
        lib_path = Path(os.environ.get("GSB_LIB", str(LIB_PATH)))   # GSB_LIB: tuning builds (tools/)
        if not lib_path.exists():
            raise FileNotFoundError(
                f"{lib_path} is missing: run `python -m gaussiansplattingmlx_b200.build` (there is no CPU fallback)")
        lib = C.CDLL(str(lib_path))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def make_camera(cam) -> GsbCamera:
    """``camera.Camera`` → ``gsb_camera`` (the 7 camera arrays of ``GaussianTrainer.swift:254-272``)."""
    out = GsbCamera()
    blk = cam.pack()
    for i in range(16):
        out.view[i] = float(blk[i])
        out.proj[i] = float(blk[16 + i])
    for i in range(3):
        out.cam_center[i] = float(blk[32 + i])
    out.fov_x, out.fov_y, out.focal_x, out.focal_y = (float(blk[35]), float(blk[36]), float(blk[37]), float(blk[38]))
    return out
