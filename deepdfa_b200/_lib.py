"""ctypes binding of libddfa_b200.so (C ABI declared in include/ddfa_b200.h).

There is NO fallback: if the shared library is missing or a symbol is absent, importing the
binding raises.  Every call checks the status code and raises ``DdfaError`` carrying
``ddfa_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

from . import build as _build

HEADER = Path(__file__).resolve().parent.parent / "include" / "ddfa_b200.h"

ENGINE_SIMT = 0
ENGINE_TCGEN05 = 1


class DdfaError(RuntimeError):
    pass


def declared_symbols():
    """Function names declared in include/ddfa_b200.h (used by the CPU export test)."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ddfa_[a-z0-9_]+)\s*\(", text)))


_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
_int = C.c_int

# name -> (restype, argtypes); pointers are passed as integers (tensor.data_ptr()) or ctypes arrays
_SIGNATURES = {
    "ddfa_abi_version": (_int, []),
    "ddfa_last_error": (C.c_char_p, []),
    "ddfa_device_supported": (_int, []),
    "ddfa_launch_count": (C.c_longlong, []),
    "ddfa_engine_available": (_int, [_int]),
    "ddfa_tuning_set": (_int, [_int, _int]),
    "ddfa_tuning_get": (_int, [_int]),
    "ddfa_debug_set": (_int, [_int, _int]),
    "ddfa_debug_read": (_int, [_int, _vp, _sz]),
    "ddfa_build_csr_workspace_bytes": (_sz, [_i64, _i32]),
    "ddfa_build_csr": (_int, [_vp, _vp, _int, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ddfa_graph_ptr": (_int, [_vp, _i32, _vp, _vp]),
    "ddfa_arena_batch_workspace_bytes": (_sz, [_i32]),
    "ddfa_arena_batch": (_int, [_vp, _i32, _i32] + [_vp] * 6 + [_i32, _vp, _i32, _i32] + [_vp] * 7 + [_vp, _sz, _vp]),
    "ddfa_embed_concat_fwd": (_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "ddfa_embed_concat_fwd_image": (_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "ddfa_embed_concat_bwd": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "ddfa_gather_sum": (_int, [_vp, _vp, _vp, _i32, _i32, _vp, _int, _vp]),
    "ddfa_gather_sum_variant": (_int, [_int, _vp, _vp, _vp, _i32, _i32, _vp, _int, _vp]),
    "ddfa_fold_weights_fwd": (_int, [_vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "ddfa_fold_weights_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "ddfa_gru_step_workspace_bytes": (_sz, [_i32, _i32, _int]),
    "ddfa_gru_step_prepare": (_int, [_vp] * 5 + [_i32, _int, _vp, _sz, _vp]),
    "ddfa_gru_step_fwd": (_int, [_vp] * 8 + [_i32, _i32, _vp, _vp, _vp, _sz, _int, _vp]),
    "ddfa_act_image_bytes": (_sz, [_i64]),
    "ddfa_act_to_image": (_int, [_vp, _i32, _i32, _vp, _vp]),
    "ddfa_gather_sum_image": (_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "ddfa_gru_step_fwd_image": (_int, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ddfa_gather_sum_image_src": (_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "ddfa_gru_gates_packed_bytes": (_sz, [_i32, _i32]),
    "ddfa_gru_step_fwd_image_v2": (_int, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ddfa_gru_step_bwd_image_v2": (_int, [_vp] * 9 + [_i32, _i32] + [_vp] * 7 + [_vp, _sz, _int, _vp]),
    "ddfa_gru_step_bwd_image": (_int, [_vp] * 9 + [_i32, _i32] + [_vp] * 7 + [_vp, _sz, _int, _vp]),
    "ddfa_gru_step_bwd_finish": (_int, [_i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "ddfa_gru_step_bwd_workspace_bytes": (_sz, [_i32, _i32, _int]),
    "ddfa_gru_step_bwd_workspace_bytes_steps": (_sz, [_i32, _i32, _int, _i32]),
    "ddfa_gru_bwd_wgrad_batched": (_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "ddfa_gru_step_prepare_bwd": (_int, [_vp, _vp, _i32, _int, _vp, _sz, _vp]),
    "ddfa_gru_step_bwd": (_int, [_vp] * 7 + [_i32, _i32] + [_vp] * 7 + [_vp, _sz, _int, _vp]),
    "ddfa_ggnn_workspace_bytes": (_sz, [_i32, _i32, _i32, _int, _int]),
    "ddfa_ggnn_fwd": (_int, [_vp, _vp, _vp, _i32, _i32, _i32] + [_vp] * 7 + [_vp, _sz, _int, _int, _vp]),
    "ddfa_ggnn_bwd": (_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32] + [_vp] * 12 + [_vp, _sz, _int, _vp]),
    "ddfa_readout_mlp_fwd": (_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _i32] + [_vp] * 6 + [_vp]),
    "ddfa_mlp_bwd": (_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "ddfa_readout_bwd": (_int, [_vp] * 5 + [_i32, _i32] + [_vp] * 8 + [_vp]),
    "ddfa_graph_label_bce": (_int, [_vp, _vp, _vp, _i32, _f32, _f32, _f32, _vp, _vp, _vp, _vp]),
    "ddfa_graph_label_bce_valid": (_int, [_vp, _vp, _vp, _i32, _i32, _f32, _f32, _f32, _vp, _vp, _vp, _vp]),
    "ddfa_adam_flat": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _vp]),
    "ddfa_allreduce_adam_p2p": (_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _f32, _f32, _f32, _f32, _f32, _vp]),
    "ddfa_sgemm": (_int, [_int, _int, _i32, _i32, _i32, _f32, _vp, _i32, _vp, _i32, _f32, _vp, _i32, _i32, _vp]),
}

TUNE_L2_HINTS, TUNE_PDL_MASK, TUNE_GATHER_VARIANT, TUNE_FWD_PAIR, TUNE_GATE_BWD_TMA, TUNE_GATHER_SRC_GROUPS = 0, 1, 2, 3, 4, 5

_NO_STATUS = {"ddfa_gru_gates_packed_bytes", "ddfa_tuning_get", "ddfa_abi_version", "ddfa_last_error", "ddfa_device_supported", "ddfa_launch_count", "ddfa_engine_available",
              "ddfa_build_csr_workspace_bytes", "ddfa_arena_batch_workspace_bytes", "ddfa_gru_step_workspace_bytes", "ddfa_gru_step_bwd_workspace_bytes", "ddfa_gru_step_bwd_workspace_bytes_steps",
              "ddfa_act_image_bytes", "ddfa_ggnn_workspace_bytes"}


class _Lib:
    def __init__(self):
        path = Path(os.environ["DDFA_LIB_PATH"]) if os.environ.get("DDFA_LIB_PATH") else _build.LIB   # override: A/B of two builds
        if not path.exists():
            raise DdfaError(
                f"{path} is missing: build it with `python -m deepdfa_b200.build` (or __graft_entry__.build()). "
                "deepdfa_b200 has no CPU / PyTorch fallback.")
        self.path = path
        self._dll = C.CDLL(str(path))
        missing = []
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(self._dll, name)
            except AttributeError:
                missing.append(name)
                continue
            fn.restype = res
            fn.argtypes = args
        if missing:
            raise DdfaError(f"{path} does not export: {missing}")
        abi = self._dll.ddfa_abi_version()
        if abi != 1:
            raise DdfaError(f"ABI version mismatch: library {abi}, binding 1")
        # A/B scripts select launch configurations through the environment of the PYTHON layer; the library itself reads none
        for env, key in (("DDFA_L2_HINTS", TUNE_L2_HINTS), ("DDFA_PDL", TUNE_PDL_MASK), ("DDFA_GATHER_VARIANT", TUNE_GATHER_VARIANT),
                         ("DDFA_FWD_PAIR", TUNE_FWD_PAIR), ("DDFA_GATE_BWD_TMA", TUNE_GATE_BWD_TMA),
                         ("DDFA_GATHER_SRC_GROUPS", TUNE_GATHER_SRC_GROUPS)):
            if os.environ.get(env) is not None:
                self._dll.ddfa_tuning_set(key, int(os.environ[env]))

    def last_error(self) -> str:
        msg = self._dll.ddfa_last_error()
        return msg.decode() if msg else ""

    def raw(self, name):
        return getattr(self._dll, name)

    def call(self, name, *args):
        rc = getattr(self._dll, name)(*args)
        if name not in _NO_STATUS and rc != 0:
            raise DdfaError(f"{name} failed (status {rc}): {self.last_error()}")
        return rc


_LIB = None


def lib() -> _Lib:
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB


def ptr_array(ptrs):
    """Host array of device pointers (for the `const T* const*` parameters)."""
    arr = (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])
    return arr
