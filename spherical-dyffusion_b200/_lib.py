"""ctypes binding of ``libsfno_b200.so`` (C ABI declared in ``include/sfno_b200.h``).

The product path has no CPU or PyTorch fallback: if the shared library is missing or a call fails,
a ``SfnoLibraryError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# SFNO_B200_LIB: load another build of the same library (A/B runs of engine experiments, scripts/build_variant.sh);
# unset in normal use
LIB_PATH = os.environ.get("SFNO_B200_LIB") or os.path.join(PKG_DIR, "libsfno_b200.so")

SFNO_GRID = {"legendre-gauss": 0, "equiangular": 1}
SFNO_PREC = {"fp32": 0, "float32": 0, "bf16": 1, "bfloat16": 1, "tf32": 2}
SFNO_OP = {"dhconv": 0, "diagonal": 1}
SFNO_ACT = {"none": 0, "gelu": 1, "relu": 2, "silu": 3}


class SfnoLibraryError(RuntimeError):
    pass


class NetConfig(ctypes.Structure):
    """Mirror of ``sfno_net_config`` (include/sfno_b200.h)."""

    _fields_ = [
        ("struct_size", c_int32), ("precision", c_int32), ("nlat", c_int32), ("nlon", c_int32),
        ("in_chans", c_int32), ("out_chans", c_int32), ("embed_dim", c_int32), ("num_layers", c_int32),
        ("mlp_hidden", c_int32), ("operator_type", c_int32), ("activation", c_int32), ("data_grid", c_int32),
        ("lmax", c_int32), ("mmax", c_int32), ("pos_embed", c_int32), ("big_skip", c_int32),
        ("instance_norm", c_int32), ("with_time_emb", c_int32), ("time_dim", c_int32),
        ("time_scale_shift_before_filter", c_int32), ("time_scaler", c_float), ("time_shift", c_float),
        ("norm_eps", c_float), ("dropout_mlp", c_float), ("drop_path_rate", c_float), ("max_batch", c_int32),
    ]


_SIGNATURES = {
    "sfno_b200_abi_version": (c_int, []),
    "sfno_b200_status_string": (c_char_p, [c_int]),
    "sfno_b200_last_error": (c_char_p, []),
    "sfno_b200_launch_count": (c_int64, []),
    "sfno_b200_set_option": (c_int, [c_char_p, c_int64]),
    "sfno_b200_tc_counters": (c_int, [c_void_p]),
    "sfno_b200_profile_begin": (c_int, [c_void_p]),
    "sfno_b200_profile_end": (c_int, [c_void_p, c_size_t, c_void_p, c_int]),
    "sfno_b200_selftest_gemm": (c_int, [c_int, POINTER(c_int), c_int, POINTER(c_double)]),
    "sfno_sht_tables_host": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_double), POINTER(c_double),
                                     POINTER(c_double), POINTER(c_double)]),
    "sfno_sht_plan_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int]),
    "sfno_sht_plan_destroy": (c_int, [c_void_p]),
    "sfno_sht_workspace_bytes": (c_size_t, [c_void_p, c_int64]),
    "sfno_sht_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "sfno_sht_inverse": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "sfno_spectral_contract": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sfno_instance_norm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64,
                                   c_float, c_void_p, c_size_t, c_void_p]),
    "sfno_instance_norm_workspace_bytes": (c_size_t, [c_int, c_int]),
    "sfno_conv1x1": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int, c_void_p]),
    "sfno_net_create": (c_int, [POINTER(NetConfig), POINTER(c_void_p)]),
    "sfno_net_destroy": (c_int, [c_void_p]),
    "sfno_net_set_param": (c_int, [c_void_p, c_char_p, c_void_p, c_int64, c_void_p]),
    "sfno_net_param_names": (c_char_p, [c_void_p]),
    "sfno_net_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "sfno_net_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_uint64, c_uint64, c_void_p,
                                 c_size_t, c_void_p]),
    "sfno_net_forward_parts": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int), c_int, c_void_p, c_void_p, c_int, c_int,
                                       c_uint64, c_uint64, c_void_p, c_size_t, c_void_p]),
    "sfno_net_set_option": (c_int, [c_void_p, c_char_p, c_int64]),
    "sfno_net_debug_tap": (c_int64, [c_void_p, c_char_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "sfno_net_forward_parts_rng": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int), c_int, c_void_p, c_void_p, c_int, c_int,
                                           c_void_p, c_void_p, c_size_t, c_void_p]),
    "sfno_param_fingerprint": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "sfno_ensemble_local_sum": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "sfno_ensemble_shifted_moments": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "sfno_ensemble_finalize": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "sfno_ensemble_stats": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfno_ensemble_stats_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfno_spectral_weight_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int]),
    "sfno_spectral_weight_set": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfno_spectral_weight_destroy": (c_int, [c_void_p]),
    "sfno_spectral_conv_workspace_bytes": (c_size_t, [c_void_p, c_void_p, c_void_p, c_int]),
    "sfno_spectral_conv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "sfno_conv1x1_backward_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int64, c_int]),
    "sfno_conv1x1_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int,
                                      c_void_p, c_size_t, c_void_p]),
    "sfno_spectral_conv_backward_workspace_bytes": (c_size_t, [c_void_p, c_void_p, c_void_p, c_int]),
    "sfno_spectral_conv_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_int, c_void_p, c_size_t, c_void_p]),
    "sfno_conv1x1_ex_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int64, c_int]),
    "sfno_conv1x1_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int, c_float,
                                c_uint64, c_uint64, c_int, c_void_p, c_size_t, c_void_p]),
    "sfno_ensemble_crps": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "sfno_normalize_pack": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sfno_prescribe_denormalize": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                           c_void_p, c_int, c_int, c_int64, c_void_p]),
    "sfno_sht_forward_adjoint": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "sfno_sht_inverse_adjoint": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "sfno_spectral_contract_backward": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                                c_void_p]),
    "sfno_conv1x1_weight_grad_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int64]),
    "sfno_conv1x1_weight_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_void_p, c_size_t, c_void_p]),
    "sfno_instance_norm_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_float,
                                            c_void_p]),
    "sfno_cold_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load_library(path: str | None = None) -> ctypes.CDLL:
    """Load (once) and type the shared library.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.isfile(path):
        raise SfnoLibraryError(
            f"{path} not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C spherical-dyffusion_b200/csrc`; there is no fallback path")
    cdll = ctypes.CDLL(path)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(cdll, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = cdll
    dbg = os.environ.get("SFNO_TC_DEBUG")   # experiment switches of the tensor-core engine (include/sfno_b200.h), tests / A-B runs only
    if dbg:
        cdll.sfno_b200_set_option(b"tc_debug", int(dbg))
    return cdll


def lib() -> ctypes.CDLL:
    return load_library()


def check(status: int, what: str = "") -> int:
    """Translate a negative status into an exception (the reference raises Python exceptions)."""
    if status >= 0:
        return status
    L = lib()
    msg = L.sfno_b200_last_error().decode(errors="replace")
    kind = L.sfno_b200_status_string(int(status)).decode()
    raise SfnoLibraryError(f"{what or 'sfno_b200'} failed: {kind} ({status}): {msg}")


def launch_count() -> int:
    return int(lib().sfno_b200_launch_count())


def set_option(key: str, value: int) -> None:
    check(lib().sfno_b200_set_option(key.encode(), int(value)), "sfno_b200_set_option")
