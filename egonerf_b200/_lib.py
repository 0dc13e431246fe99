"""ctypes binding of libegn_b200.so (C ABI declared in include/egn.h).

There is no CPU or PyTorch fallback: if the shared library is missing or does not export the ABI,
import of the compute path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EGN_B200_LIB") or os.path.join(_HERE, "libegn_b200.so")      # override: kernel experiments only
ABI_VERSION = 8

c_float_p = C.POINTER(C.c_float)


class EgnConfig(C.Structure):
    _fields_ = [
        ("grid", C.c_int32 * 3), ("c_sigma", C.c_int32), ("c_app", C.c_int32), ("app_dim", C.c_int32),
        ("shading", C.c_int32), ("view_pe", C.c_int32), ("fea_pe", C.c_int32), ("feature_c", C.c_int32),
        ("fea2dense", C.c_int32), ("n_coarse", C.c_int32), ("n_fine", C.c_int32),
        ("use_coarse_sample", C.c_int32), ("resampling", C.c_int32), ("env_h", C.c_int32),
        ("mlp_mode", C.c_int32),
        ("center", C.c_float * 3), ("near_plane", C.c_float), ("density_shift", C.c_float),
        ("distance_scale", C.c_float), ("ang_near", C.c_float * 2), ("ang_inv", C.c_float * 2),
        ("exp_sampling", C.c_int32), ("far_plane", C.c_float), ("step_size", C.c_float), ("aabb", C.c_float * 6), ("bwd_tc", C.c_int32),
        ("plain_ladders", C.c_int32), ("jitter_ratio", C.c_float), ("jitter_r0", C.c_float),
        ("r_knots", C.c_void_p), ("r_knots_coarse", C.c_void_p), ("z_coarse", C.c_void_p), ("tables_bf16", C.c_void_p),
        ("tables_h", C.c_void_p),
    ]


class EgnParams(C.Structure):
    _fields_ = [
        ("density_plane", (C.c_void_p * 3) * 2), ("density_line", (C.c_void_p * 3) * 2),
        ("app_plane", (C.c_void_p * 3) * 2), ("app_line", (C.c_void_p * 3) * 2),
        ("basis", C.c_void_p * 2), ("mlp_w", C.c_void_p * 3), ("mlp_b", C.c_void_p * 3),
        ("emission", C.c_void_p),
    ]


EgnGrads = EgnParams   # identical layout (include/egn.h)


class EgnOutputs(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("depth", C.c_void_p), ("bg", C.c_void_p), ("env", C.c_void_p),
                ("alpha", C.c_void_p)]


SHADING = {"MLP_Fea": 0, "MLP": 1, "RGB": 2, "SH": 3}
ACT = {"softplus": 0, "relu": 1}
MLP_MODE = {"fp32": 0, "tc_split": 1, "tc_f16": 2, "tc_bf16": 2}   # EGN_MLP_* (include/egn.h); tc_bf16 = pre-ABI-7 name of tc_f16

# name -> (restype, argtypes); every symbol include/egn.h declares
PROTOTYPES = {
    "egn_last_error": (C.c_char_p, []),
    "egn_abi_version": (C.c_int32, []),
    "egn_samples_per_ray": (C.c_int32, [C.POINTER(EgnConfig)]),
    "egn_table_floats": (C.c_int64, [C.POINTER(EgnConfig)]),
    "egn_pack_tables": (C.c_int32, [C.POINTER(EgnConfig), C.POINTER(EgnParams), C.c_void_p, C.c_void_p]),
    "egn_table_bf16_elems": (C.c_int64, [C.POINTER(EgnConfig)]),
    "egn_pack_tables_bf16": (C.c_int32, [C.POINTER(EgnConfig), C.c_void_p, C.c_void_p, C.c_void_p]),
    "egn_table_h_bytes": (C.c_int64, [C.POINTER(EgnConfig)]),
    "egn_pack_tables_h": (C.c_int32, [C.POINTER(EgnConfig), C.c_void_p, C.c_void_p, C.c_void_p]),
    "egn_adam_tables": (C.c_int32, [C.POINTER(EgnConfig), C.POINTER(EgnGrads), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_void_p]),
    "egn_unpack_table_grads": (C.c_int32, [C.POINTER(EgnConfig), C.c_void_p, C.POINTER(EgnGrads), C.c_void_p]),
    "egn_pack_table_grads": (C.c_int32, [C.POINTER(EgnConfig), C.POINTER(EgnGrads), C.c_void_p, C.c_void_p]),
    "egn_regularize_tables": (C.c_int32, [C.POINTER(EgnConfig), C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                          C.c_void_p, C.c_void_p]),
    "egn_workspace_bytes": (C.c_int64, [C.POINTER(EgnConfig), C.c_int64]),
    "egn_workspace_bytes_eval": (C.c_int64, [C.POINTER(EgnConfig), C.c_int64]),
    "egn_render_forward": (C.c_int32, [C.POINTER(EgnConfig), C.POINTER(EgnParams), C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64,
                                       C.POINTER(EgnOutputs), C.c_void_p, C.c_int32, C.c_void_p]),
    "egn_render_forward_timed": (C.c_int32, [C.POINTER(EgnConfig), C.POINTER(EgnParams), C.c_void_p, C.c_void_p,
                                             C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64,
                                             C.POINTER(EgnOutputs), C.c_void_p, C.c_void_p, c_float_p]),
    "egn_sample_rays": (C.c_int32, [C.POINTER(EgnConfig), C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p]),
    "egn_render_samples": (C.c_int32, [C.POINTER(EgnConfig), C.POINTER(EgnParams), C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_void_p, C.POINTER(EgnOutputs), C.c_void_p, C.c_int32, C.c_void_p]),
    "egn_render_backward": (C.c_int32, [C.POINTER(EgnConfig), C.POINTER(EgnParams), C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.POINTER(EgnGrads), C.c_void_p]),
    "egn_render_backward_sparse_env": (C.c_int32, [C.POINTER(EgnConfig), C.POINTER(EgnParams), C.c_void_p, C.c_void_p, C.c_int64,
                                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.POINTER(EgnGrads), C.c_void_p, C.c_void_p]),
    "egn_density_feature": (C.c_int32, [C.POINTER(EgnConfig), C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                        C.c_void_p, C.c_void_p]),
    "egn_app_feature": (C.c_int32, [C.POINTER(EgnConfig), C.POINTER(EgnParams), C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "egn_yinyang_coords": (C.c_int32, [C.POINTER(EgnConfig), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "egn_envmap_radiance": (C.c_int32, [C.POINTER(EgnConfig), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "egn_envmap_backward": (C.c_int32, [C.POINTER(EgnConfig), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "egn_erp_rays": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_float_p, C.c_void_p, C.c_void_p]),
    "egn_resample_factor": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                        C.c_int32, C.c_void_p, C.c_void_p]),
    "egn_host_sample_schedule": (C.c_int32, [C.c_float, C.c_float, C.c_float, C.c_int32, c_float_p]),
    "egn_host_r_knots": (C.c_int32, [C.c_float, C.c_float, C.c_int32, c_float_p]),
    "egn_host_plain_sample_schedule": (C.c_int32, [C.c_float, C.c_float, C.c_int32, c_float_p, c_float_p, c_float_p]),
    "egn_host_plain_r_knots": (C.c_int32, [C.c_float, C.c_float, C.c_int32, c_float_p]),
    "egn_peer_flag_bytes": (C.c_int64, []),
    "egn_peer_alloc": (C.c_int32, [C.c_int64, C.POINTER(C.c_void_p)]),
    "egn_peer_free": (C.c_int32, [C.c_void_p]),
    "egn_peer_export": (C.c_int32, [C.c_void_p, C.c_char_p]),
    "egn_peer_open": (C.c_int32, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "egn_peer_close": (C.c_int32, [C.c_void_p]),
    "egn_peer_allreduce": (C.c_int32, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int64, C.c_float,
                                       C.c_uint32, C.c_int32, C.c_void_p]),
}

_lib = None


def load():
    """Loads the library once; raises if it is absent or stale (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `make -C egonerf_b200/csrc`). egonerf_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)            # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.egn_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libegn_b200.so ABI {lib.egn_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        raise RuntimeError("libegn_b200: " + load().egn_last_error().decode())


def ptr(t):
    """Raw device (or host) pointer of a contiguous fp32 tensor, 0 for None."""
    if t is None:
        return None
    assert t.is_contiguous(), "egonerf_b200 passes raw pointers: tensor must be contiguous"
    return t.data_ptr()
