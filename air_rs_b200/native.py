"""ctypes binding of include/airgpu.h and include/airgpu_synth.h.

Loading is strict: a missing libairgpu.so, a missing symbol or a missing GPU is
an error, never a silent fallback (there is no CPU path in this package).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import build as _build

FMT_CS16 = 0
FMT_U8 = 1

# airgpu_frame, 24 bytes
FRAME_DTYPE = np.dtype(
    [("bytes", np.uint8, (14,)), ("fixed_bit", np.uint8), ("reserved", np.uint8), ("offset", np.uint64)],
    align=True,
)
assert FRAME_DTYPE.itemsize == 24

# airgpu_fields, 32 bytes
FIELDS_DTYPE = np.dtype(
    [("icao", np.uint32), ("downlink_format", np.uint8), ("capability", np.uint8), ("msg_type", np.uint8),
     ("kind", np.uint8), ("altitude", np.int32), ("cpr_latitude", np.uint32), ("cpr_longitude", np.uint32),
     ("surveillance_status", np.uint8), ("nic_supplement", np.uint8), ("cpr_time", np.uint8), ("cpr_odd", np.uint8),
     ("callsign", "S8")],
    align=True,
)
assert FIELDS_DTYPE.itemsize == 32

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_NOMEM, ERR_OVERFLOW, ERR_BUSY, ERR_TICKET = 0, -1, -2, -3, -4, -5, -6, -7


class AirgpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"airgpu error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("device", C.c_int32),
        ("format", C.c_uint32),
        ("ring_slots", C.c_uint32),
        ("max_buffer_samples", C.c_uint64),
        ("max_frames", C.c_uint64),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_samples", C.c_uint64),
        ("n_frames", C.c_uint64),
        ("gate_passes", C.c_uint64),
        ("n_tiles", C.c_uint64),
        ("kernel_ms", C.c_float),
        ("h2d_ms", C.c_float),
        ("decode_ms", C.c_float),
        ("decode_launches", C.c_float),
    ]


_vp, _sz, _u64 = C.c_void_p, C.c_size_t, C.c_uint64

MAX_PEERS = 8


class Peers(C.Structure):
    """airgpu_peers: destinations of a fused frame exchange (device-accessible addresses)."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("n_outs", C.c_uint32),
        ("multicast", C.c_uint32),
        ("reserved", C.c_uint32),
        ("outs", _vp * MAX_PEERS),
        ("counts", _vp * MAX_PEERS),
    ]

# every symbol include/airgpu.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "airgpu_version": (C.c_char_p, []),
    "airgpu_last_error": (C.c_char_p, []),
    "airgpu_device_count": (C.c_int, []),
    "airgpu_create": (C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    "airgpu_destroy": (None, [_vp]),
    "airgpu_submit": (C.c_int, [_vp, _vp, _sz, _u64, C.POINTER(_u64)]),
    "airgpu_collect": (C.c_int, [_vp, _u64, _vp, _sz, C.POINTER(_sz)]),
    "airgpu_decode": (C.c_int, [_vp, _vp, _sz, _sz, _u64, _vp, _sz, C.POINTER(_sz)]),
    "airgpu_decode_device": (C.c_int, [_vp, _vp, _sz, _sz, _u64, _vp, _sz, _vp, _vp]),
    "airgpu_sync_count": (C.c_int, [_vp, C.POINTER(_u64)]),
    "airgpu_get_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "airgpu_playback_samples": (_sz, [_sz, _sz]),
    "airgpu_reserve": (C.c_int, [_vp, _sz, _sz, _sz]),
    "airgpu_set_timing": (C.c_int, [_vp, C.c_int]),
    "airgpu_graph_begin": (C.c_int, [_vp, _vp]),
    "airgpu_graph_end": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "airgpu_graph_launch": (C.c_int, [_vp, _vp]),
    "airgpu_set_capturing": (C.c_int, [_vp, C.c_int]),
    "airgpu_graph_destroy": (None, [_vp]),
    "airgpu_decode_device_peers": (C.c_int, [_vp, _vp, _sz, _sz, _u64, C.POINTER(Peers), _sz, _vp]),
    "airgpu_peer_barrier": (C.c_int, [_vp, C.POINTER(_vp), C.c_uint32, C.c_uint32, _u64, _vp]),
    "airgpu_group_create": (C.c_int, [C.POINTER(C.c_int), C.c_uint32, C.c_uint32, C.POINTER(_vp)]),
    "airgpu_group_decode": (C.c_int, [_vp, _vp, _sz, _u64, _vp, _sz, C.POINTER(_sz)]),
    "airgpu_group_stats": (C.c_int, [_vp, C.POINTER(Stats), C.c_uint32]),
    "airgpu_group_destroy": (None, [_vp]),
    "airgpu_decode_sharded": (C.c_int, [C.POINTER(C.c_int), C.c_uint32, C.c_uint32, _vp, _sz, _u64, _vp, _sz, C.POINTER(_sz)]),
    "airgpu_decode_fields": (C.c_int, [_vp, _vp, _sz, _vp, _vp]),
    "airgpu_decode_fields_host": (C.c_int, [_vp, _vp, _sz, _vp]),
    "airgpu_host_alloc": (C.c_int, [_sz, C.POINTER(_vp)]),
    "airgpu_host_free": (C.c_int, [_vp]),
    "airgpu_dbg_levels_u8": (C.c_int, [_vp, _vp]),
    "airgpu_dbg_levels_cs16": (C.c_int, [_vp, _vp, _sz, _vp]),
}

# include/airgpu_synth.h
SYNTH_SYMBOLS = {
    "airgpu_synth_last_error": (C.c_char_p, []),
    "airgpu_synth_table_create": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _sz, C.POINTER(_vp)]),
    "airgpu_synth_table_destroy": (None, [_vp]),
    "airgpu_synth_render": (C.c_int, [_vp, _u64, _u64, _u64, C.c_uint32, C.c_int32, _u64, _vp, _vp]),
}

_lib = None


def library_path() -> Path:
    return _build.LIB


def lib() -> C.CDLL:
    """Load libairgpu.so (building it first if the sources are newer)."""
    global _lib
    if _lib is None:
        path = _build.build()
        L = C.CDLL(str(path))
        for table in (SYMBOLS, SYNTH_SYMBOLS):
            for name, (res, args) in table.items():
                fn = getattr(L, name)  # AttributeError if the ABI drifted: loud by design
                fn.restype = res
                fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise AirgpuError(rc, lib().airgpu_last_error().decode("utf-8", "replace"))


def check_synth(rc: int) -> None:
    if rc != OK:
        raise AirgpuError(rc, lib().airgpu_synth_last_error().decode("utf-8", "replace"))
