#!/usr/bin/env python
"""Supplementary measurements (not the bench headline):
  - CS16 (reference-native, 4 B/sample) decode throughput and roofline fraction;
  - N1 fields_kernel throughput (24 B in + 32 B out per frame) against the HBM roofline.
Prints one JSON object."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from air_rs_b200 import synth  # noqa: E402
from air_rs_b200.decoder import AdsbDecoder  # noqa: E402
from air_rs_b200.native import FMT_CS16, FMT_U8  # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    res = {"hbm_peak_gbs": peak}
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    # ---- CS16 ----
    n = 1_200_000_000
    tab = synth.make_traffic(1090, 24_000_000, df17_per_s=3000, decoy_per_s=3000, snr_db=(8, 30), sigma=300.0)
    gen = synth.DeviceSynth(tab)
    iq = gen.render(1090, 0, n, FMT_CS16, 300.0, period=24_000_000)
    dec = AdsbDecoder(fmt=FMT_CS16)
    cap = n // 240
    out = torch.empty((cap, 24), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")

    def run():
        dec.decode_device(iq.data_ptr(), n, out.data_ptr(), cap, 0, 0, cnt.data_ptr(), s.cuda_stream)

    ms = timed(run)
    kms = float(np.mean([(run(), dec.stats()["decode_ms"])[1] for _ in range(5)]))
    frames = int(cnt.item())
    res["cs16"] = {"samples": n, "frames": frames, "ms_per_pass": ms, "Msamples_per_s": n / ms / 1e3,
                   "decode_kernel_ms": kms, "achieved_gbs": (4 * n + 24 * frames) / kms / 1e6,
                   "roofline_frac": (4 * n + 24 * frames) / kms / 1e6 / peak}
    del iq
    # ---- fields kernel ----
    nf = min(frames, cap)
    fo = torch.empty((nf, 32), dtype=torch.uint8, device="cuda")
    big_in = out[:nf].repeat(max(1, 20_000_000 // max(nf, 1)), 1).contiguous()
    big_out = torch.empty((big_in.shape[0], 32), dtype=torch.uint8, device="cuda")
    m = big_in.shape[0]

    def runf():
        dec.decode_fields_device(big_in.data_ptr(), m, big_out.data_ptr(), s.cuda_stream)

    msf = timed(runf, n=20)
    res["fields_kernel"] = {"frames": m, "ms": msf, "Mframes_per_s": m / msf / 1e3, "achieved_gbs": 56 * m / msf / 1e6,
                            "roofline_frac": 56 * m / msf / 1e6 / peak, "bytes_per_frame": 56}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
