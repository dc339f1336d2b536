"""numpy restatement of air_rs's ADS-B decode path -- TEST INFRASTRUCTURE ONLY.

A second, independently written restatement of the same reference functions as
oracle/adsb_oracle.c, used to cross-check the C oracle (the reference ships no
IQ fixture and cannot be compiled here, so two restatements that must agree is
the strongest pin available; see adsb_oracle.h).

Reference map (file:line in the upstream tree):
  magnitude()            src/utils.rs:46-52
  gate_mask()            src/adsb/demod.rs:17-57, driven by src/adsb.rs:98-104
  slice_bytes()          src/adsb/demod.rs:92-131 + 180-201 (net bit rule)
  crc24()                src/adsb/crc.rs:10-40
  repair()               src/adsb/crc.rs:49-65
  decode()               src/adsb.rs:92-122
"""
from __future__ import annotations

import numpy as np

FRAME_SAMPLES = 240  # 16 + 112 * 2, src/adsb.rs:98
GENERATOR = 0x1FFF409  # src/adsb/crc.rs:11

PRE_HIGHS = (0, 2, 7, 9)  # demod.rs:24
PRE_LOWS = (1, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14, 15)  # demod.rs:23
DF_HIGHS = (0, 3, 5, 7, 8)  # demod.rs:46 (+16)
DF_LOWS = (1, 2, 4, 6, 9)  # demod.rs:45 (+16)

FRAME_DTYPE = np.dtype(
    [("bytes", np.uint8, (14,)), ("fixed_bit", np.uint8), ("reserved", np.uint8), ("offset", np.uint64)],
    align=True,
)


def widen_u8(iq_u8: np.ndarray) -> np.ndarray:
    """re = (2u - 255) * 128, exact in int16."""
    return ((2 * iq_u8.astype(np.int32) - 255) * 128).astype(np.int16)


def magnitude(iq_cs16: np.ndarray) -> np.ndarray:
    """utils.rs:46-52: f64 sqrt of re^2 + im^2, truncated to u32."""
    z = iq_cs16.reshape(-1, 2).astype(np.float64)
    return np.sqrt(z[:, 0] ** 2 + z[:, 1] ** 2).astype(np.uint32)


def gate_mask(m: np.ndarray) -> np.ndarray:
    """Boolean mask over candidate offsets [0, len-240): demod.rs:17-57.

    `high < low -> reject` for every (high, low) pair is the same as
    min(highs) >= max(lows); ties pass.
    """
    n = m.size - FRAME_SAMPLES
    if n <= 0:
        return np.zeros(0, dtype=bool)

    def win(k):
        return m[k : k + n]

    hi = win(PRE_HIGHS[0]).copy()
    for k in PRE_HIGHS[1:]:
        np.minimum(hi, win(k), out=hi)
    lo = win(PRE_LOWS[0]).copy()
    for k in PRE_LOWS[1:]:
        np.maximum(lo, win(k), out=lo)
    ok = hi >= lo
    hi = win(16 + DF_HIGHS[0]).copy()
    for k in DF_HIGHS[1:]:
        np.minimum(hi, win(16 + k), out=hi)
    lo = win(16 + DF_LOWS[0]).copy()
    for k in DF_LOWS[1:]:
        np.maximum(lo, win(16 + k), out=lo)
    ok &= hi >= lo
    return ok


def slice_bytes(d224: np.ndarray) -> bytes:
    """bit k = 1 iff d[2k] > d[2k+1], MSB first (demod.rs:104-118 with 190-197)."""
    bits = (d224[0::2] > d224[1::2]).astype(np.uint8)
    return bytes(np.packbits(bits))


def crc24(data: bytes) -> int:
    """crc.rs:10-40 as integer polynomial division: (msg << 24) mod GENERATOR."""
    v = int.from_bytes(data, "big") << 24
    for shift in range(len(data) * 8 - 1, -1, -1):
        if v >> (shift + 24) & 1:
            v ^= GENERATOR << shift
    return v & 0xFFFFFF


def repair(packet: bytes, packet_crc: int):
    """crc.rs:49-65: first single-bit flip (byte 0..13, bit 7..0) whose CRC over
    bytes 0..11 equals the RECEIVED crc."""
    for num in range(len(packet)):
        for i in range(8):
            aug = bytearray(packet)
            aug[num] ^= 1 << (7 - i)
            if crc24(bytes(aug[: len(aug) - 3])) == packet_crc:
                return bytes(aug), num * 8 + i
    return None


def extract_packet(d224: np.ndarray):
    """demod.rs:65-82 -> (bytes, fixed_bit) or None."""
    pkt = slice_bytes(d224)
    calc = crc24(pkt[:11])
    rx = (pkt[11] << 16) | (pkt[12] << 8) | pkt[13]
    if calc == rx:
        return pkt, 0xFF
    return repair(pkt, rx)


def decode_mags(m: np.ndarray, base: int = 0):
    """adsb.rs:96-116 for one buffer. Returns (list of (bytes, fixed_bit, offset), gate_passes)."""
    frames = []
    hits = np.flatnonzero(gate_mask(m))
    for i in hits:
        r = extract_packet(m[i + 16 : i + 240])
        if r is not None:
            frames.append((r[0], r[1], base + int(i)))
    return frames, int(hits.size)


def decode(iq: np.ndarray, segment_samples: int = 0, base: int = 0):
    """Whole path on interleaved IQ (uint8 = U8 mode, int16 = CS16 mode)."""
    iq = np.ascontiguousarray(iq).reshape(-1)
    if iq.dtype == np.uint8:
        iq = widen_u8(iq)
    elif iq.dtype != np.int16:
        raise TypeError(iq.dtype)
    n = iq.size // 2
    seg = n if segment_samples in (0, None) or segment_samples > n else segment_samples
    frames, passes = [], 0
    for s0 in range(0, n, seg):
        m = magnitude(iq[2 * s0 : 2 * min(n, s0 + seg)])
        f, p = decode_mags(m, base + s0)
        frames += f
        passes += p
    return frames, passes


def to_records(frames) -> np.ndarray:
    out = np.zeros(len(frames), dtype=FRAME_DTYPE)
    for k, (b, fx, off) in enumerate(frames):
        out["bytes"][k] = np.frombuffer(b, dtype=np.uint8)
        out["fixed_bit"][k] = fx
        out["offset"][k] = off
    return out
