#!/usr/bin/env python
"""Smallest run that touches every device code path (for compute-sanitizer):
U8 and CS16, aligned and guarded loaders, multi-segment, the >16-frames overflow path
(constant buffer), streaming ring, fields kernel.  Checks results against the oracle."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from air_rs_b200 import synth  # noqa: E402
from air_rs_b200.decoder import AdsbDecoder  # noqa: E402
from air_rs_b200.native import FMT_CS16, FMT_U8  # noqa: E402
from oracle import oracle_c  # noqa: E402


def main():
    n = 60_000
    tab = synth.make_traffic(5, n, df17_per_s=4000, decoy_per_s=2000, snr_db=(8, 30))
    u8 = synth.render(tab, 5, 0, n, FMT_U8, 2.0)
    tab16 = synth.make_traffic(5, n, df17_per_s=4000, decoy_per_s=2000, snr_db=(8, 30), sigma=300.0)
    cs16 = synth.render(tab16, 5, 0, n, FMT_CS16, 300.0)
    ok = True
    for fmt, iq in ((FMT_U8, u8), (FMT_CS16, cs16)):
        with AdsbDecoder(fmt=fmt, max_buffer_samples=1 << 16, max_frames=1 << 15) as dec:
            for seg in (0, 20_000, 8_433):
                got = dec.decode(iq, segment_samples=seg)
                want, _ = oracle_c.decode_fast(iq, seg)
                ok &= got.tobytes() == want.tobytes()
            const = np.full(2 * 5000, 9, dtype=iq.dtype)
            got = dec.decode(const)
            ok &= len(got) == 5000 - 240
            t = dec.submit(iq[: 2 * 30_000])
            ok &= len(dec.collect(t)) == len(oracle_c.decode_fast(iq[: 2 * 30_000])[0])
            f = dec.decode_fields(got[:100])
            ok &= f.tobytes() == oracle_c.frames_fields(got[:100]).tobytes()
    print("sanity_small:", "OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
