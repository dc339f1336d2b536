"""CPU tests: the oracle against every known-answer test the reference carries
for this path, and the three restatements (C literal, C fast, numpy) against
each other.  No GPU, no product code under test here."""
import numpy as np
import pytest

from air_rs_b200 import synth
from oracle import oracle_c, oracle_np

from common import capture_cs16, capture_u8, flip_bit, frames_equal

GOLDEN = synth.GOLDEN_FRAMES


# ---- reference KATs (file:line of the upstream test in each docstring) --------

def test_crc_kat():
    """demod.rs:337-355 test_get_adsb_crc / test_get_adsb_crc_real"""
    data = bytes.fromhex("8D406B902015A678D4D220")
    assert oracle_c.get_adsb_crc(data) == 0xAA4BDA
    assert oracle_np.crc24(data) == 0xAA4BDA
    assert synth.crc24(data) == 0xAA4BDA


def test_crc_kat_invalid():
    """demod.rs:357-367 test_get_adsb_crc_real_invalid"""
    data = bytes.fromhex("8d406a902015a678d4d220")
    assert oracle_c.get_adsb_crc(data) != 0xAA4BDA
    assert oracle_np.crc24(data) != 0xAA4BDA


def test_crc_of_zeros_is_zero():
    assert oracle_c.get_adsb_crc(bytes(11)) == 0


def test_gate_valid_ties_pass():
    """demod.rs:250-264: highs 1000, lows 500, all-zero DF window; returns 900"""
    buf = np.zeros(32, dtype=np.uint32)
    buf[[0, 2, 7, 9]] = 1000
    buf[[1, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14, 15]] = 500
    assert oracle_c.check_for_adsb_packet(buf) == 900


def test_gate_invalid():
    """demod.rs:266-278"""
    buf = np.zeros(32, dtype=np.uint32)
    buf[[0, 2, 7, 9]] = 500
    buf[[1, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14, 15]] = 1000
    assert oracle_c.check_for_adsb_packet(buf) is None


def test_extract_packet_bad_crc():
    """demod.rs:369-380: alternating 120/50 -> all ones -> CRC 0xD1D94C != 0xFFFFFF, unrepairable"""
    buf = np.tile(np.array([120, 50], dtype=np.uint32), 112)
    assert oracle_c.extract_packet(buf, 100) is None
    assert oracle_np.extract_packet(buf) is None
    assert oracle_c.get_adsb_crc(b"\xff" * 11) == 0xD1D94C


@pytest.mark.parametrize("hexframe", GOLDEN)
def test_golden_frames_are_crc_valid(hexframe):
    """aircraft.rs:188-261, demod.rs:339-344: the reference's own DF17 frames"""
    f = bytes.fromhex(hexframe)
    assert f[0] >> 3 == 17
    assert oracle_c.get_adsb_crc(f[:11]) == int.from_bytes(f[11:], "big")
    assert oracle_np.crc24(f[:11]) == int.from_bytes(f[11:], "big")


def test_stale_reference_tests_documented():
    """demod.rs:322-335 are stale upstream: the CODE yields 0x55 for
    0b1001100110011001 and never returns None.  The oracle follows the code."""
    import ctypes as C
    sym = np.array([0b1001100110011001], dtype=np.uint16)
    out = np.zeros(1, dtype=np.uint8)
    oracle_c.lib().oracle_decode_packet(sym.ctypes.data, 1, out.ctypes.data)
    assert out[0] == 0x55
    sym = np.array([0b1111000011110000], dtype=np.uint16)
    assert oracle_c.lib().oracle_decode_packet(sym.ctypes.data, 1, out.ctypes.data) == 1


# ---- properties of the algorithm the GPU design relies on ----------------------

def test_syndromes_distinct_nonzero_not_single_bit():
    t = oracle_c.syndrome_table()
    assert len(set(t.tolist())) == 88 and 0 not in t
    assert all(bin(int(v)).count("1") > 1 for v in t)
    assert t[0] == 0x3935EA and t[87] == 0xFFF409
    for p in (0, 5, 40, 87):
        e = bytearray(11)
        e[p >> 3] = 0x80 >> (p & 7)
        assert oracle_np.crc24(bytes(e)) == t[p]


def test_repair_data_bits_only():
    """crc.rs:49-65: flips in bits 0..87 are repaired, parity-bit flips are not."""
    good = GOLDEN[6]
    rx = int(good[22:], 16)
    for bit in (5, 17, 63, 87):
        bad = flip_bit(good, bit)
        r = oracle_c.try_crc_recovery(bad, oracle_c.get_adsb_crc(bad[:11]), rx)
        assert r is not None and r[0] == bytes.fromhex(good) and r[1] == bit
        assert oracle_np.repair(bad, rx) == (bytes.fromhex(good), bit)
    for bit in (88, 100, 111):
        bad = flip_bit(good, bit)
        rxb = int.from_bytes(bad[11:], "big")
        assert oracle_c.try_crc_recovery(bad, oracle_c.get_adsb_crc(bad[:11]), rxb) is None


def test_u8_level_is_order_isomorphic_to_reference_magnitude():
    """The GPU compares level = I(255-I) + Q(255-Q) instead of the magnitude.
    Exhaustive over all 65536 byte pairs: level order == reverse magnitude order,
    ties included."""
    i, q = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    iq = np.stack([i.ravel(), q.ravel()], axis=1).astype(np.uint8).reshape(-1)
    mags = oracle_c.get_magnitude(oracle_c.widen_u8(iq)).astype(np.int64)
    level = (i * (255 - i) + q * (255 - q)).ravel().astype(np.int64)
    assert level.max() == 32512 and level.min() == 0
    order = np.argsort(level, kind="stable")
    ls, ms = level[order], mags[order]
    same = ls[1:] == ls[:-1]
    assert np.all(ms[1:][same] == ms[:-1][same])          # equal level -> equal magnitude
    assert np.all(ms[1:][~same] < ms[:-1][~same])         # larger level -> strictly smaller magnitude


def test_magnitude_is_exact_isqrt():
    rng = np.random.default_rng(7)
    iq = rng.integers(-32768, 32768, size=200_000, dtype=np.int16)
    iq[:8] = [-32768, -32768, 32767, 32767, -32768, 0, 0, 0]
    m = oracle_c.get_magnitude(iq).astype(np.int64)
    n = iq[0::2].astype(np.int64) ** 2 + iq[1::2].astype(np.int64) ** 2
    assert np.all(m * m <= n) and np.all((m + 1) * (m + 1) > n)
    assert np.array_equal(m, oracle_np.magnitude(iq))
    assert m.max() <= 46340


# ---- the three restatements agree ---------------------------------------------

def _np_records(iq, seg=0, base=0):
    f, gp = oracle_np.decode(iq, seg, base)
    return oracle_np.to_records(f), gp


@pytest.mark.parametrize("maker,kw", [
    (capture_u8, dict(n=120_000)),
    (capture_u8, dict(n=120_000, seed=5, snr=(0.0, 12.0), smear=0.5)),
    (capture_cs16, dict(n=120_000)),
    (capture_cs16, dict(n=60_000, sigma=3.0, snr=(10.0, 30.0))),   # coarse magnitudes: many ties
])
def test_literal_fast_numpy_agree(maker, kw):
    _, iq = maker(**kw)
    lit, gp1 = oracle_c.decode_literal(iq)
    fast, gp2 = oracle_c.decode_fast(iq, threads=3)
    npr, gp3 = _np_records(iq)
    assert gp1 == gp2 == gp3
    assert frames_equal(lit, fast) and frames_equal(lit, npr)
    assert len(lit) > 0


def test_segmented_and_threaded_equal_single_pass():
    _, iq = capture_u8(n=100_000, seed=11)
    seg = 20_000                                            # the reference's playback chunk, adsb.rs:78
    lit, gp = oracle_c.decode_literal(iq, seg, base=1000)
    fast, gp2 = oracle_c.decode_fast(iq, seg, base=1000, threads=4)
    mt, gp3 = oracle_c.decode_literal_mt(iq, seg, base=1000, threads=4)
    npr, gp4 = _np_records(iq, seg, 1000)
    assert gp == gp2 == gp3 == gp4
    assert frames_equal(lit, fast) and frames_equal(lit, mt) and frames_equal(lit, npr)
    whole, _ = oracle_c.decode_literal(iq)
    # chunking loses candidates in the last 240 samples of every chunk, never adds any
    assert set(lit["offset"] - 1000) <= set(whole["offset"])


def test_edge_cases():
    # shorter than a frame: upstream panics (adsb.rs:98 underflow); defined as zero frames
    for n in (0, 1, 239, 240):
        iq = np.zeros(2 * n, dtype=np.int16)
        assert len(oracle_c.decode_literal(iq)[0]) == 0
        assert len(oracle_c.decode_fast(iq)[0]) == 0
    # constant input: ties pass the gate, bits slice to 0, crc(0)=0 -> a frame at EVERY offset
    iq = np.zeros(2 * 300, dtype=np.int16)
    lit, gp = oracle_c.decode_literal(iq)
    assert len(lit) == 60 and gp == 60 and not lit["bytes"].any()
    assert np.array_equal(lit["offset"], np.arange(60))
    assert frames_equal(lit, oracle_c.decode_fast(iq)[0])
    iq8 = np.full(2 * 300, 200, dtype=np.uint8)
    assert len(oracle_c.decode_literal(iq8)[0]) == 60


def test_frame_touching_last_sample_is_missed():
    """adsb.rs:98: candidates stop at len-241, so a frame whose 240th sample is the
    buffer's last sample is not decoded; one sample of slack and it is."""
    tab = synth.single_frames([GOLDEN[0]], [1000], amp_i=40)
    iq = synth.render(tab, 3, 0, 1240, synth.FMT_U8, sigma=0.5)
    assert 1000 not in oracle_c.decode_literal(iq)[0]["offset"]
    iq = synth.render(tab, 3, 0, 1241, synth.FMT_U8, sigma=0.5)
    f = oracle_c.decode_literal(iq)[0]
    assert 1000 in f["offset"]
    assert bytes(f[f["offset"] == 1000][0]["bytes"]).hex() == GOLDEN[0]


def test_injected_bit_errors():
    """config 3: one flipped bit in 5..87 is repaired with fixed_bit = that bit; in 88..111
    rejected; in 0..4 fails the DF gate."""
    frames, starts = [], []
    bits = [5, 20, 87, 88, 111, 0, 2, 4]
    for k, b in enumerate(bits):
        frames.append(flip_bit(GOLDEN[k % 7], b))
        starts.append(2000 + 600 * k)
    tab = synth.single_frames(frames, starts, amp_i=50, amp_q=20)
    iq = synth.render(tab, 9, 0, 8000, synth.FMT_U8, sigma=1.0)
    f = oracle_c.decode_literal(iq)[0]
    got = {int(r["offset"]): int(r["fixed_bit"]) for r in f}
    assert got == {2000: 5, 2600: 20, 3200: 87}
    for off, k in ((2000, 0), (2600, 1), (3200, 2)):
        assert bytes(f[f["offset"] == off][0]["bytes"]).hex() == GOLDEN[k].lower()
