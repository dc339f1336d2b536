"""Synthetic generator: determinism (CPU) and host == device byte equality (GPU)."""
import hashlib

import numpy as np
import pytest

from air_rs_b200 import synth


def test_noise_statistics():
    n = synth.noise(1090, 0, 400_000, synth.noise_gain(2.0)).astype(np.float64)
    assert abs(n.mean() + 0.5) < 0.02          # floor shift centres the noise on -0.5 (127.5 after +128)
    assert abs(n.std() - 2.0) < 0.06
    a = synth.noise(1090, 1000, 64, synth.noise_gain(2.0))
    b = synth.noise(1090, 0, 2000, synth.noise_gain(2.0))[1000:1064]
    assert np.array_equal(a, b)                # counter based: any window, same values


def test_render_is_windowable_and_deterministic():
    tab = synth.make_traffic(7, 50_000, df17_per_s=3000, decoy_per_s=3000, snr_db=(8, 30), smear_fraction=0.3)
    whole = synth.render(tab, 7, 0, 50_000)
    part = synth.render(tab, 7, 12_345, 10_000)
    assert np.array_equal(whole[2 * 12_345 : 2 * 22_345], part)
    assert hashlib.sha256(whole.tobytes()).hexdigest() == hashlib.sha256(synth.render(tab, 7, 0, 50_000).tobytes()).hexdigest()


def test_periodic_schedule():
    tab = synth.make_traffic(3, 20_000, df17_per_s=4000, snr_db=(20, 20))
    a = synth.render(tab, 3, 0, 60_000, period=20_000)
    sig0 = synth.signal(tab, 0, 20_000)
    # same pulses in every period, different noise
    for rep in range(3):
        blk = a[2 * 20_000 * rep : 2 * 20_000 * (rep + 1)].astype(np.int64).reshape(-1, 2)
        nz = synth.noise(3, 20_000 * rep, 20_000, synth.noise_gain(2.0))
        assert np.array_equal(np.clip(nz + sig0 + 128, 0, 255), blk)
    assert not np.array_equal(a[:40_000], a[40_000:80_000])


def test_frame_table_golden_frames_present():
    tab = synth.make_traffic(1090, 2_400_000, df17_per_s=400)
    hexes = {bytes(p).hex() for p in tab.payload[tab.kind == 17]}
    assert sum(h.lower() in hexes for h in (g.lower() for g in synth.GOLDEN_FRAMES)) >= 5


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,sigma,period", [
    (synth.FMT_U8, 2.0, 0), (synth.FMT_CS16, 300.0, 0), (synth.FMT_U8, 2.0, 30_000), (synth.FMT_CS16, 3.0, 17_001)])
def test_device_render_equals_numpy(fmt, sigma, period):
    import torch

    n = 100_000
    tab = synth.make_traffic(5, period or n, df17_per_s=3000, decoy_per_s=3000, snr_db=(8, 30), sigma=sigma,
                             smear_fraction=0.25)
    dev = synth.DeviceSynth(tab)
    for j0 in (0, 77_777):
        got = dev.render(5, j0, n, fmt, sigma, period)
        torch.cuda.synchronize()
        want = synth.render(tab, 5, j0, n, fmt, sigma, period)
        assert np.array_equal(got.cpu().numpy(), want)
    dev.close()
