"""ctypes binding of libladine (include/ladine.h).  The only place the shared library is loaded.

There is no CPU or PyTorch fallback: if the library is missing or the device is not sm_100 the
call raises (BASELINE.json north_star: "no Triton, no multi-backend dispatch and no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
import os
import threading

LADINE_MAX_CLASSES = 16
LADINE_MAX_GROUP = 8
PREC = {"auto": 0, "fp32": 1, "fp16": 2, "bf16": 3, "fp32x": 4}
PREC_NAME = {v: k for k, v in PREC.items()}

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libladine.so")


class MemberDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("feature_dim", C.c_int32), ("num_classes", C.c_int32),
        ("n_steps", C.c_int32), ("emb_rows", C.c_int32), ("guidance", C.c_int32), ("precision", C.c_int32),
        ("bn_eps", C.c_float),
        ("lin1_w", C.c_void_p), ("lin1_b", C.c_void_p), ("lin2_w", C.c_void_p), ("lin2_b", C.c_void_p),
        ("lin3_w", C.c_void_p), ("lin3_b", C.c_void_p), ("lin4_w", C.c_void_p), ("lin4_b", C.c_void_p),
        ("emb", C.c_void_p * 3), ("bn_w", C.c_void_p * 3), ("bn_b", C.c_void_p * 3),
        ("bn_mean", C.c_void_p * 3), ("bn_var", C.c_void_p * 3),
    ]


class SampleArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("K", C.c_int32), ("N", C.c_int32), ("D", C.c_int32), ("T", C.c_int32),
        ("t_first", C.c_int32), ("t_last", C.c_int32),
        ("xf", C.c_void_p), ("y0hat", C.c_void_p), ("ytmean", C.c_void_p), ("y_init", C.c_void_p),
        ("coef", C.c_void_p), ("noise", C.c_void_p), ("seed", C.c_uint64),
        ("member_ids", C.POINTER(C.c_int32)), ("image_offset", C.c_int32), ("images_total", C.c_int32),
        ("draw_offset", C.c_int32), ("draws_total", C.c_int32),
        ("y_out", C.c_void_p), ("traj_out", C.c_void_p), ("prob_out", C.c_void_p), ("temperature", C.c_float),
        ("stream", C.c_void_p),
    ]


class EncoderDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("data_dim", C.c_int32), ("hidden_dim", C.c_int32), ("feature_dim", C.c_int32),
        ("bn_eps", C.c_float),
        ("lin_w", C.c_void_p * 3), ("lin_b", C.c_void_p * 3), ("bn_w", C.c_void_p * 3), ("bn_b", C.c_void_p * 3),
        ("bn_mean", C.c_void_p * 3), ("bn_var", C.c_void_p * 3),
    ]


# every symbol include/ladine.h declares: (restype, argtypes)
SYMBOLS = {
    "ladine_version": (C.c_int, []),
    "ladine_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "ladine_destroy": (C.c_int, [C.c_void_p]),
    "ladine_last_error": (C.c_char_p, [C.c_void_p]),
    "ladine_pack_member": (C.c_int, [C.c_void_p, C.POINTER(MemberDesc), C.c_void_p, C.POINTER(C.c_void_p)]),
    "ladine_free_member": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ladine_member_bytes": (C.c_uint64, [C.c_void_p]),
    "ladine_member_precision": (C.c_int, [C.c_void_p]),
    "ladine_member_fpad": (C.c_int, [C.c_void_p]),
    "ladine_member_cpad": (C.c_int, [C.c_void_p]),
    "ladine_sample": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(SampleArgs)]),
    "ladine_fill_noise": (C.c_int, [C.c_void_p, C.POINTER(SampleArgs), C.c_int32, C.c_void_p]),
    "ladine_last_launches": (C.c_int64, [C.c_void_p]),
    "ladine_workspace_bytes": (C.c_uint64, [C.c_void_p]),
    "ladine_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "ladine_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "ladine_get_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "ladine_pack_encoder": (C.c_int, [C.c_void_p, C.POINTER(EncoderDesc), C.c_void_p, C.POINTER(C.c_void_p)]),
    "ladine_free_encoder": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ladine_encoder_bytes": (C.c_uint64, [C.c_void_p]),
    "ladine_encode": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                C.c_void_p]),
    "ladine_last_encoder_launches": (C.c_int64, [C.c_void_p]),
    "ladine_member_image_bytes": (C.c_uint64, [C.c_void_p]),
    "ladine_member_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "ladine_member_import": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ladine_encoder_image_bytes": (C.c_uint64, [C.c_void_p]),
    "ladine_encoder_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "ladine_encoder_import": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ladine_image_info": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_char_p)]),
    "ladine_image_checksum": (C.c_uint64, [C.c_void_p, C.c_uint64]),
    "ladine_member_dims": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "ladine_encoder_dims": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "ladine_debug_geometry": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "ladine_debug_plan": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.POINTER(C.c_int32), C.c_int64, C.POINTER(C.c_int32)]),
    "ladine_debug_layer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()
_handles = {}
_call_locks = {}


def call_lock(device_index: int) -> "threading.Lock":
    """Serialises the C-ABI calls that use one handle's workspace (ctypes releases the GIL, so two Python threads could
    otherwise interleave ladine_sample / ladine_encode on the same handle; the header requires serialisation)."""
    with _lock:
        lk = _call_locks.get(device_index)
        if lk is None:
            lk = _call_locks[device_index] = threading.Lock()
    return lk


class LadineError(RuntimeError):
    """Carries the C-ABI status code and ladine_last_error() text."""

    def __init__(self, code, msg):
        super().__init__(f"libladine error {code}: {msg}")
        self.code = code


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load libladine.so (built in-tree by ``__graft_entry__.build()``); fail loudly when absent."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(_LIB_PATH):
                raise RuntimeError(
                    f"{_LIB_PATH} is missing: build the CUDA extension first "
                    "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
            lib = C.CDLL(_LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name)  # AttributeError if the ABI drifted from the header
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def handle(device_index: int) -> int:
    """One ladine_handle per CUDA device per process."""
    lib = load()
    with _lock:
        h = _handles.get(device_index)
        if h is None:
            out = C.c_void_p()
            rc = lib.ladine_create(device_index, C.byref(out))
            if rc != 0:
                why = {-3: "device is not compute capability 10.x (sm_100a kernels only, no fallback)",
                       -2: "CUDA device query failed"}.get(rc, "ladine_create failed")
                raise LadineError(rc, why)
            h = out.value
            _handles[device_index] = h
    return h


def check(h, rc):
    if rc != 0:
        raise LadineError(rc, load().ladine_last_error(h).decode())
