"""CPU-side checks of the drop-in boundary (no compute calls, no GPU needed):
the C-ABI library loads, exports every symbol include/mvsdet_b200.h declares
and nothing is declared in the ctypes layer that the header does not have;
the host-side mirrors validate their arguments and refuse CPU tensors."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mvsdet_b200.h")


def _header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)            # strip comments
    # every prototype of the form  <type> mvsd_xxx(
    return sorted(set(re.findall(r"\b(mvsd_[a-z0-9_]+)\s*\(", text)))


def _build_if_needed():
    from mvsdet_b200 import _lib, build
    if not os.path.isfile(_lib.LIB_PATH):
        build.build()
    return _lib


def test_header_declares_entry_points():
    names = _header_functions()
    for must in ("mvsd_plane_sweep_fwd", "mvsd_plane_sweep_bwd", "mvsd_depth_topk_fwd",
                 "mvsd_depth_topk_bwd", "mvsd_backproject_fwd", "mvsd_backproject_bwd",
                 "mvsd_homo_warp_fwd", "mvsd_homo_warp_bwd", "mvsd_voxel_normalize"):
        assert must in names


def test_library_exports_every_declared_symbol():
    _lib = _build_if_needed()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_functions():
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"


def test_ctypes_table_matches_header():
    _lib = _build_if_needed()
    assert sorted(_lib.SIGNATURES) == _header_functions()


def test_prototype_arity_matches_ctypes_table():
    """Count the parameters of every prototype in the header and compare with
    the ctypes argtypes: a drifted signature would corrupt the stack silently."""
    _lib = _build_if_needed()
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (argtypes, _) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^)]*)\)" % name, text)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(argtypes), f"{name}: header has {n} parameters, ctypes {len(argtypes)}"


def test_load_info_and_status_strings():
    _lib = _build_if_needed()
    lib = _lib.load()
    assert lib.mvsd_abi_version() == _lib.ABI_VERSION == 3
    assert b"sm_100a" in lib.mvsd_build_info()
    assert lib.mvsd_status_string(0) == b"ok"
    assert lib.mvsd_status_string(1) == b"invalid argument"
    assert lib.mvsd_launch_count() >= 0


def test_invalid_arguments_fail_before_any_launch():
    """Argument validation happens on the host side of the ABI: a null pointer
    or a non-positive dimension is rejected with MVSD_ERR_INVALID_ARG and a
    message, without touching the (absent) device."""
    _lib = _build_if_needed()
    lib = _lib.load()
    before = lib.mvsd_launch_count()
    st = lib.mvsd_plane_sweep_fwd(None, 0, None, None, None, None, 0, 0, 2, 8, 4, 4, 4, 0, 0, 2, None)
    assert st == _lib.ERR_INVALID_ARG and b"null" in lib.mvsd_last_error()
    st = lib.mvsd_plane_sweep_fwd(None, 0, None, None, None, None, 0, 0, 0, 8, 4, 4, 4, 0, 0, 2, None)
    assert st == _lib.ERR_INVALID_ARG
    st = lib.mvsd_plane_sweep_fwd(None, 0, None, None, None, None, 0, 0, 2, 6, 4, 4, 4, 0, 0, 2, None)
    assert st == _lib.ERR_UNSUPPORTED                    # C not a multiple of 4
    one = ctypes.c_void_p(16)                            # non-null dummy: rejected before any dereference
    st = lib.mvsd_plane_sweep_fwd(one, 0, one, one, one, one, 0, 0, 2, 8, 4, 4, 4, 2, 1, 2, None)
    assert st == _lib.ERR_INVALID_ARG and b"exceed" in lib.mvsd_last_error()   # ref views [1,3) of 2
    st = lib.mvsd_depth_topk_fwd(None, 0, 0, 0, 0, None, None, None, None, None, None,
                                 None, 0, None, None, None, None, 0.2, 0.4, 0, 1, 100, 4, 4, 3, None)
    assert st == _lib.ERR_UNSUPPORTED                    # D > 64
    st = lib.mvsd_scene_setup(one, one, 0, one, one, one, one, one, 4, 4, 0, 4, None)
    assert st == _lib.ERR_UNSUPPORTED                    # k must be <= V-1
    st = lib.mvsd_scene_setup(one, one, 0, one, one, one, one, one, 4, 2, 3, 2, None)
    assert st == _lib.ERR_INVALID_ARG                    # reference views [3,5) of 4
    st = lib.mvsd_plane_sweep_groupcorr_fwd(one, 0, one, one, one, one, 2, 24, 4, 4, 4, 1, 2, 0, 2, None)
    assert st == _lib.ERR_UNSUPPORTED and b"per group" in lib.mvsd_last_error()    # 12 channels per group
    st = lib.mvsd_plane_sweep_groupcorr_bwd(one, one, 0, one, one, one, one, 2, 24, 4, 4, 4, 1, 5, 0, 2, None)
    assert st == _lib.ERR_INVALID_ARG                    # 24 channels, 5 groups
    st = lib.mvsd_backproject_fwd(None, 0, 4, 4, None, None, None, None, 0, 0, 0, 0, 0.2, 7, None,
                                  0, None, None, None, 1, 8, 4, 4, 3, 10, None)
    assert st == _lib.ERR_INVALID_ARG                    # bad mode
    with pytest.raises(ValueError):
        _lib.call("mvsd_voxel_normalize", None, None, None, 0, 4, 4, None)
    assert lib.mvsd_launch_count() == before


def test_ops_refuse_cpu_tensors():
    from mvsdet_b200 import functional as F_, ops
    x = torch.randn(2, 8, 4, 4)
    with pytest.raises(ValueError, match="no CPU path"):
        ops.pack_features(x)
    with pytest.raises(ValueError):
        ops.depth_topk(torch.randn(2, 2, 4, 4, 4), 0.2, 0.4, 3)
    with pytest.raises(ValueError):
        F_.backproject_Weigh(x, torch.zeros(3, 2, 2, 2), torch.zeros(2, 3, 4),
                             torch.zeros(2, 16, 1, 3), [0.1, 0.1, 0.1], torch.zeros(2, 16, 1, 3),
                             gt_depth=torch.zeros(1))


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under mvsdet_b200/ may import it."""
    pkg = os.path.join(ROOT, "mvsdet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
