"""Shared helpers for the test-suite (inputs + comparison)."""
from __future__ import annotations

import numpy as np

from air_rs_b200 import synth
from oracle import oracle_c


def frames_equal(a: np.ndarray, b: np.ndarray) -> bool:
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def describe_diff(got: np.ndarray, want: np.ndarray) -> str:
    if len(got) != len(want):
        go = set(got["offset"].tolist())
        wo = set(want["offset"].tolist())
        return (f"count {len(got)} != {len(want)}; missing offsets {sorted(wo - go)[:8]}, "
                f"extra offsets {sorted(go - wo)[:8]}")
    for k in range(len(got)):
        if got[k].tobytes() != want[k].tobytes():
            return f"first difference at record {k}: got {got[k]}, want {want[k]}"
    return "equal"


def capture_u8(seed=1090, n=240_000, df17=2000.0, decoy=1000.0, snr=(6.0, 30.0), sigma=2.0, smear=0.0):
    tab = synth.make_traffic(seed, n, df17_per_s=df17, decoy_per_s=decoy, snr_db=snr, sigma=sigma,
                             smear_fraction=smear)
    return tab, synth.render(tab, seed, 0, n, synth.FMT_U8, sigma)


def capture_cs16(seed=2024, n=240_000, df17=2000.0, decoy=1000.0, snr=(6.0, 30.0), sigma=300.0):
    tab = synth.make_traffic(seed, n, df17_per_s=df17, decoy_per_s=decoy, snr_db=snr, sigma=sigma)
    return tab, synth.render(tab, seed, 0, n, synth.FMT_CS16, sigma)


def flip_bit(frame_hex: str, bit: int) -> bytes:
    b = bytearray(bytes.fromhex(frame_hex))
    b[bit >> 3] ^= 0x80 >> (bit & 7)
    return bytes(b)


def oracle_frames(iq, segment_samples=0, base=0):
    return oracle_c.decode_fast(iq, segment_samples, base, threads=4)
