"""GPU: BASELINE.json configs 1-3 at their STATED sizes, not scaled-down shapes.

The captures are rendered on the device by the integer generator (byte-identical to the numpy one:
tests/test_synth.py), decoded through the C ABI (device-resident call), and compared record for record -- and
gate-pass counter for gate-pass counter -- with the CPU oracle run on a host copy of the same bytes: the fast oracle
on the whole capture, the LITERAL oracle (statement-by-statement restatement of the reference) on a slice.
Config 5's whole-capture check (8.64 G samples) lives in bench.py (`full_capture_check`), config 4 in its `config4`
block and in test_independent_segments[131072]."""
import os

import numpy as np
import pytest
import torch

from air_rs_b200 import synth
from air_rs_b200.decoder import AdsbDecoder
from air_rs_b200.native import FMT_U8, FRAME_DTYPE
from oracle import oracle_c

from common import describe_diff, flip_bit, frames_equal

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 4


def _decode_resident(dec, d_iq, n, cap):
    out, count = dec.decode_tensor(d_iq, cap=cap)
    return AdsbDecoder.frames_from_tensor(out, count), dec.stats()["gate_passes"]


def _check_against_oracle(table, seed, n, sigma=2.0, literal_slice=2_400_000, min_frames=1):
    gen = synth.DeviceSynth(table)
    d_iq = gen.render(seed, 0, n, FMT_U8, sigma)
    torch.cuda.synchronize()
    with AdsbDecoder(fmt=FMT_U8) as dec:
        got, gp = _decode_resident(dec, d_iq, n, max(1 << 16, n // 200))
    host = d_iq.cpu().numpy()
    gen.close()
    del d_iq
    want, wgp = oracle_c.decode_fast(host, threads=THREADS)
    assert frames_equal(got, want), describe_diff(got, want)
    assert gp == wgp and len(want) >= min_frames
    lit, _ = oracle_c.decode_literal(host[: 2 * literal_slice])
    sub = got[got["offset"] < literal_slice - 240]
    assert frames_equal(sub, lit), describe_diff(sub, lit)
    return got


def test_config1_10s_capture_at_stated_size():
    """config 1: 10 s at 2.4 MS/s = 24 000 000 samples u8, ~200 DF17/s at 20 dB, as ONE continuous capture and in the
    reference's 20 000-sample playback chunks (src/adsb.rs:75-89)."""
    n = 24_000_000
    tab = synth.make_traffic(1090, n, df17_per_s=200.0, decoy_per_s=0.0, snr_db=(20.0, 20.0))
    got = _check_against_oracle(tab, 1090, n, min_frames=1500)
    assert 1500 < len(got) < 2600
    # reference-chunked: independent 20 000-sample buffers, the tail chunk never sent
    gen = synth.DeviceSynth(tab)
    d_iq = gen.render(1090, 0, n, FMT_U8, 2.0)
    kept = ((n - 1) // 20_000) * 20_000
    with AdsbDecoder(fmt=FMT_U8) as dec:
        out, count = dec.decode_tensor(d_iq[: 2 * kept], segment_samples=20_000, cap=1 << 16)
        got_c = AdsbDecoder.frames_from_tensor(out, count)
    want_c, _ = oracle_c.decode_fast(d_iq[: 2 * kept].cpu().numpy(), 20_000, 0, threads=THREADS)
    assert frames_equal(got_c, want_c), describe_diff(got_c, want_c)
    assert len(got_c) <= len(got)          # frames straddling a chunk boundary are lost, as in the reference
    gen.close()


def test_config2_60s_dense_capture_at_stated_size():
    """config 2: 60 s = 144 000 000 samples (288 MB), 500-ICAO pool, ~3000 DF17/s + ~3000 decoys/s (DF4/5/11/20/21),
    SNR 8-30 dB, overlaps allowed."""
    n = 144_000_000
    tab = synth.make_traffic(2, n, df17_per_s=3000.0, decoy_per_s=3000.0, snr_db=(8.0, 30.0), n_icao=500)
    got = _check_against_oracle(tab, 2, n, min_frames=60_000)
    # the decoys never come out: every emitted frame starts with DF17 (demod.rs:45-54)
    assert set((got["bytes"][:, 0] >> 3).tolist()) == {17}


def test_config3_snr_sweep_and_bit_errors_at_stated_size():
    """config 3: SNR steps 0..20 dB x 10 000 frames each, then the single-bit-error sets (bits 5..87 must be repaired
    with fixed_bit = that bit, bits 88..111 rejected, bits 0..4 fail the gate), then a pure-noise stretch for
    false-positive parity.  210 000 + 112 frames, 300 samples apart."""
    rng = np.random.default_rng(3)
    sigma = 2.0
    frames, amps = [], []
    for snr in range(0, 21):
        a = int(round(sigma * 10.0 ** (snr / 20.0)))
        for _ in range(10_000):
            me = bytearray(rng.integers(0, 256, size=7, dtype=np.uint8).tobytes())
            me[0] = (int(rng.integers(1, 20)) << 3) | (me[0] & 7)
            frames.append(synth.df17_frame(int(rng.integers(0x100000, 0xFFFFFF)), bytes(me)))
            amps.append(a)
    n_sweep = len(frames)
    base = synth.GOLDEN_FRAMES[0]
    for bit in range(112):
        frames.append(flip_bit(base, bit))
        amps.append(40)
    starts = [1_000 + 300 * k for k in range(len(frames))]
    noise_tail = 4_000_000
    n = starts[-1] + 1_000 + noise_tail
    tab = synth.single_frames(frames, starts, amp_i=np.asarray(amps, dtype=np.int32))
    got = _check_against_oracle(tab, 3, n, sigma=sigma, min_frames=60_000)       # ~40 % of the sweep decodes (none below 10 dB)
    # sensitivity is the oracle's business (parity is what is asserted above); sanity: the 20 dB step decodes fully
    top = got[(got["offset"] >= starts[20 * 10_000]) & (got["offset"] < starts[n_sweep])]
    assert len(np.unique(top["offset"])) >= 9_990
    # the injected single-bit errors at 26 dB
    flip0 = starts[n_sweep]
    by_off = {int(r["offset"]): r for r in got[got["offset"] >= flip0]}
    for bit in range(112):
        r = by_off.get(flip0 + 300 * bit)
        if 5 <= bit < 88:
            assert r is not None and r["fixed_bit"] == bit and bytes(r["bytes"]).hex() == base.lower()
        else:
            assert r is None, bit
