"""ctypes binding of the C ABI declared in include/gecco_b200.h.

The shared library is built in-tree by gecco_b200.build; there is no CPU fallback: if the
library is missing or the device is not sm_100 every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libgecco_b200.so"
_lib = None
_inited_devices: set[int] = set()


class GeccoError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_int64),
        ("w", C.c_void_p), ("ldw", C.c_int64),
        ("m", C.c_int32), ("n_out", C.c_int32), ("k", C.c_int32),
        ("rows_per_cloud", C.c_int32), ("valid_rows", C.c_int32), ("w_rows_per_cloud", C.c_int32),
        ("bias", C.c_void_p), ("bias_stride", C.c_int32),
        ("act", C.c_int32), ("act_alpha", C.c_float),
        ("res", C.c_void_p), ("ldr", C.c_int64),
        ("out_f32", C.c_void_p), ("ldo32", C.c_int64),
        ("out_bf16", C.c_void_p), ("ldo16", C.c_int64),
        ("stats", C.c_void_p),
        ("geom", C.c_void_p),
        ("sigma", C.c_void_p), ("sigma_stride", C.c_int32),
        ("wx", C.c_void_p),
    ]


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Loads libgecco_b200.so (building it first if GECCO_B200_AUTOBUILD=1 and it is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        if os.environ.get("GECCO_B200_AUTOBUILD", "0") == "1":
            from . import build as _build

            _build.build()
        else:
            raise GeccoError(
                f"{_LIB_PATH} is missing: build it with `python -m gecco_b200.build` "
                "(gecco_b200 has no CPU or PyTorch fallback)"
            )
    lib = C.CDLL(str(_LIB_PATH))
    lib.gecco_last_error.restype = C.c_char_p
    lib.gecco_abi_version.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().gecco_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"gecco_b200: {msg}")
        raise GeccoError(f"gecco_b200 (code {rc}): {msg}")


def init(device_index: int) -> C.CDLL:
    lib = load()
    if device_index not in _inited_devices:
        check(lib.gecco_init(C.c_int(device_index)))
        _inited_devices.add(device_index)
    return lib
