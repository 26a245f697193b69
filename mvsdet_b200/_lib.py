"""ctypes binding of libmvsdet_b200.so -- the C ABI declared in
include/mvsdet_b200.h.  There is no fallback: if the library is missing or a
call fails this raises, loudly."""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MVSDET_B200_LIB") or os.path.join(HERE, "lib", "libmvsdet_b200.so")

ABI_VERSION = 3
OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3
F32, BF16 = 0, 1
CHANNELS_LAST, CHANNELS_FIRST = 0, 1
BP_MEAN, BP_SUM, BP_PER_VIEW = 0, 1, 2

_p, _i, _f, _l = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> argtypes; every symbol include/mvsdet_b200.h declares
SIGNATURES = {
    "mvsd_abi_version": ([], _i),
    "mvsd_build_info": ([], C.c_char_p),
    "mvsd_status_string": ([_i], C.c_char_p),
    "mvsd_last_error": ([], C.c_char_p),
    "mvsd_set_tuning": ([_i, _i], _i),
    "mvsd_launch_count": ([], _l),
    "mvsd_pack_nchw_to_nhwc": ([_p, _p, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_unpack_nhwc_to_nchw": ([_p, _p, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_scene_setup": ([_p, _p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p], _i),
    "mvsd_plane_sweep_fwd": ([_p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_plane_sweep_bwd": ([_p, _i, _i, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_plane_sweep_bwd_det": ([_p, _i, _i, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_fixed_to_float": ([_p, _p, _l, _p], _i),
    "mvsd_backproject_bwd_det": ([_p, _i, _p, _p, _i, _i, _i, _p, _p, _p, _p, _l, _l, _l, _l,
                                  _f, _p, _p, _i, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_plane_sweep_groupcorr_fwd": ([_p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_plane_sweep_groupcorr_bwd": ([_p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_homo_warp_fwd": ([_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_homo_warp_bwd": ([_p, _i, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_depth_topk_fwd": ([_p, _l, _l, _l, _l, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p,
                             _f, _f, _i, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_depth_topk_bwd": ([_p, _l, _l, _l, _l, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p,
                             _f, _f, _i, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_ray_depth_scale": ([_p, _i, _p, _i, _i, _i, _p], _i),
    "mvsd_rgb_downsample4": ([_p, _p, _i, _p, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_backproject_fwd": ([_p, _i, _i, _i, _p, _p, _p, _p, _l, _l, _l, _l, _f, _i, _p, _i,
                              _p, _p, _p, _i, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_backproject_bwd": ([_p, _i, _i, _p, _p, _i, _i, _i, _p, _p, _p, _p, _l, _l, _l, _l,
                              _f, _p, _p, _i, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_prob_norm_bwd": ([_p, _p, _p, _l, _l, _l, _l, _i, _i, _i, _i, _p], _i),
    "mvsd_voxel_normalize": ([_p, _p, _p, _i, _i, _i, _p], _i),
    "mvsd_voxel_reduce_p2p": ([_p, _p, _p, _i, _i, _i, _i, _i, _p], _i),
    "mvsd_halo_reduce_p2p": ([_p, _p, _p, _i, _i, _i, _l, _p], _i),
}

_lock = threading.Lock()
_lib = None


class MvsdError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the library and type every entry point (no compute happens)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise MvsdError(
                    f"{LIB_PATH} is missing. Build it with `python -m mvsdet_b200.build` "
                    "(nvcc, sm_100a). mvsdet_b200 has no CPU or PyTorch fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (argtypes, restype) in SIGNATURES.items():
                fn = getattr(lib, name)          # AttributeError if a symbol is missing
                fn.argtypes = argtypes
                fn.restype = restype
            if lib.mvsd_abi_version() != ABI_VERSION:
                raise MvsdError("libmvsdet_b200.so ABI version mismatch; rebuild")
            _lib = lib
    return _lib


def call(name: str, *args) -> None:
    """Invoke a status-returning entry point; non-zero status -> exception
    (ValueError for caller mistakes, RuntimeError otherwise)."""
    lib = load()
    status = getattr(lib, name)(*args)
    if status != OK:
        msg = lib.mvsd_last_error().decode() or lib.mvsd_status_string(status).decode()
        if status in (ERR_INVALID_ARG, ERR_UNSUPPORTED):
            raise ValueError(f"{name}: {msg}")
        raise MvsdError(f"{name}: {msg}")


def launch_count() -> int:
    return int(load().mvsd_launch_count())


def set_tuning(key: int, value: int) -> int:
    return int(load().mvsd_set_tuning(key, value))
