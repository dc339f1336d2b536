"""ctypes loader for oracle/_build/libadsb_oracle.so -- TEST INFRASTRUCTURE ONLY.

The C file restates the reference's decode path (see adsb_oracle.h for the
file:line map).  This module only builds and binds it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libadsb_oracle.so"

FMT_CS16 = 0
FMT_U8 = 1

FRAME_DTYPE = np.dtype(
    [("bytes", np.uint8, (14,)), ("fixed_bit", np.uint8), ("reserved", np.uint8), ("offset", np.uint64)],
    align=True,
)
assert FRAME_DTYPE.itemsize == 24

FIELDS_DTYPE = np.dtype(
    [("icao", np.uint32), ("downlink_format", np.uint8), ("capability", np.uint8), ("msg_type", np.uint8),
     ("kind", np.uint8), ("altitude", np.int32), ("cpr_latitude", np.uint32), ("cpr_longitude", np.uint32),
     ("surveillance_status", np.uint8), ("nic_supplement", np.uint8), ("cpr_time", np.uint8), ("cpr_odd", np.uint8),
     ("callsign", "S8")],
    align=True,
)
assert FIELDS_DTYPE.itemsize == 32


def build(force: bool = False) -> Path:
    """Compile the oracle with gcc (a few hundred ms). Safe to call repeatedly."""
    src = _HERE / "adsb_oracle.c"
    hdr = _HERE / "adsb_oracle.h"
    stale = (not _SO.exists()) or _SO.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime)
    if force or stale:
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B"], check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        u8p, u16p, u32p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64))
        L.oracle_get_magnitude.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_get_magnitude.restype = None
        L.oracle_widen_u8.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_widen_u8.restype = None
        L.oracle_check_for_adsb_packet.argtypes = [C.c_void_p, u32p]
        L.oracle_check_for_adsb_packet.restype = C.c_int
        L.oracle_get_adsb_crc.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_get_adsb_crc.restype = C.c_uint32
        L.oracle_try_crc_recovery.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]
        L.oracle_try_crc_recovery.restype = C.c_int
        L.oracle_extract_packet.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.POINTER(C.c_int)]
        L.oracle_extract_packet.restype = C.c_int
        L.oracle_extract_manchester_relative.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
        L.oracle_extract_manchester_relative.restype = C.c_int
        L.oracle_decode_packet.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_decode_packet.restype = C.c_size_t
        L.oracle_process_mags.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p, C.c_size_t, u64p]
        L.oracle_process_mags.restype = C.c_size_t
        common = [C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, C.c_uint64, C.c_void_p, C.c_size_t, u64p]
        L.oracle_decode_literal.argtypes = common
        L.oracle_decode_literal.restype = C.c_size_t
        L.oracle_decode_literal_mt.argtypes = common + [C.c_int]
        L.oracle_decode_literal_mt.restype = C.c_size_t
        L.oracle_decode_fast.argtypes = common + [C.c_int]
        L.oracle_decode_fast.restype = C.c_size_t
        L.oracle_frames_fields.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_frames_fields.restype = None
        L.oracle_syndrome_table.argtypes = [u32p]
        L.oracle_syndrome_table.restype = None
        del u8p, u16p
        _lib = L
    return _lib


def _fmt_of(iq: np.ndarray) -> int:
    if iq.dtype == np.uint8:
        return FMT_U8
    if iq.dtype == np.int16:
        return FMT_CS16
    raise TypeError(f"IQ must be uint8 or int16 interleaved, got {iq.dtype}")


def _decode(which: str, iq: np.ndarray, segment_samples: int, base: int, threads: int | None):
    iq = np.ascontiguousarray(iq).reshape(-1)
    n = iq.size // 2
    L = lib()
    gp = C.c_uint64(0)
    cap = 4096
    while True:
        out = np.zeros(cap, dtype=FRAME_DTYPE)
        args = [iq.ctypes.data, n, _fmt_of(iq), segment_samples, base, out.ctypes.data, cap, C.byref(gp)]
        if which == "literal":
            total = L.oracle_decode_literal(*args)
        elif which == "literal_mt":
            total = L.oracle_decode_literal_mt(*args, threads or os.cpu_count() or 1)
        else:
            total = L.oracle_decode_fast(*args, threads or 1)
        if total <= cap:
            return out[:total].copy(), int(gp.value)
        cap = int(total)


def decode_literal(iq, segment_samples: int = 0, base: int = 0):
    """Reference-literal decode (single thread). Returns (frames, gate_passes)."""
    return _decode("literal", iq, segment_samples, base, 1)


def decode_literal_mt(iq, segment_samples: int = 0, base: int = 0, threads: int | None = None):
    return _decode("literal_mt", iq, segment_samples, base, threads)


def decode_fast(iq, segment_samples: int = 0, base: int = 0, threads: int = 1):
    return _decode("fast", iq, segment_samples, base, threads)


def get_magnitude(iq_cs16: np.ndarray) -> np.ndarray:
    iq = np.ascontiguousarray(iq_cs16, dtype=np.int16).reshape(-1)
    out = np.zeros(iq.size // 2, dtype=np.uint32)
    lib().oracle_get_magnitude(iq.ctypes.data, out.size, out.ctypes.data)
    return out


def widen_u8(iq_u8: np.ndarray) -> np.ndarray:
    iq = np.ascontiguousarray(iq_u8, dtype=np.uint8).reshape(-1)
    out = np.zeros(iq.size, dtype=np.int16)
    lib().oracle_widen_u8(iq.ctypes.data, iq.size // 2, out.ctypes.data)
    return out


def check_for_adsb_packet(buf32) -> int | None:
    b = np.ascontiguousarray(buf32, dtype=np.uint32)
    assert b.size == 32
    high = C.c_uint32(0)
    ok = lib().oracle_check_for_adsb_packet(b.ctypes.data, C.byref(high))
    return int(high.value) if ok else None


def get_adsb_crc(data: bytes) -> int:
    b = np.frombuffer(bytes(data), dtype=np.uint8)
    return int(lib().oracle_get_adsb_crc(b.ctypes.data, b.size))


def extract_packet(mags224, high: int = 0):
    b = np.ascontiguousarray(mags224, dtype=np.uint32)
    out = np.zeros(14, dtype=np.uint8)
    fixed = C.c_int(0)
    ok = lib().oracle_extract_packet(b.ctypes.data, b.size, high, out.ctypes.data, C.byref(fixed))
    return (bytes(out), int(fixed.value)) if ok else None


def try_crc_recovery(packet: bytes, calc_crc: int, packet_crc: int):
    b = np.frombuffer(bytes(packet), dtype=np.uint8).copy()
    flipped = C.c_int(-1)
    ok = lib().oracle_try_crc_recovery(b.ctypes.data, b.size, calc_crc, packet_crc, C.byref(flipped))
    return (bytes(b), int(flipped.value)) if ok else None


def process_mags(mags, base: int = 0):
    m = np.ascontiguousarray(mags, dtype=np.uint32)
    gp = C.c_uint64(0)
    cap = max(16, m.size)
    out = np.zeros(cap, dtype=FRAME_DTYPE)
    n = lib().oracle_process_mags(m.ctypes.data, m.size, base, out.ctypes.data, cap, C.byref(gp))
    return out[:n].copy(), int(gp.value)


def syndrome_table() -> np.ndarray:
    t = np.zeros(88, dtype=np.uint32)
    lib().oracle_syndrome_table(t.ctypes.data_as(C.POINTER(C.c_uint32)))
    return t


def frames_fields(frames: np.ndarray) -> np.ndarray:
    """AdsbPacket::new field derivation for every frame record (N1 oracle)."""
    fr = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
    out = np.zeros(fr.size, dtype=FIELDS_DTYPE)
    lib().oracle_frames_fields(fr.ctypes.data, fr.size, out.ctypes.data)
    return out


def packet_fields(packet: bytes):
    fr = np.zeros(1, dtype=FRAME_DTYPE)
    fr["bytes"][0] = np.frombuffer(bytes(packet), dtype=np.uint8)
    return frames_fields(fr)[0]
