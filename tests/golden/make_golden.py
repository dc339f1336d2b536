"""Regenerates tests/golden/capture_small.* -- run from the repo root:
    python tests/golden/make_golden.py

The reference (Rust) cannot be built or imported in this environment, so these
vectors are the LITERAL C oracle's output (oracle/adsb_oracle.c, which restates
the reference line by line and is pinned on the reference's own KATs in
tests/test_oracle.py) for two small synthetic captures that include the seven
CRC-valid frames the reference's tests carry, three single-bit-error variants
and DF4/5/11/20/21 decoys.  Both restatements (C and numpy) must agree before
anything is written.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from air_rs_b200 import synth  # noqa: E402
from oracle import oracle_c, oracle_np  # noqa: E402
from common import flip_bit  # noqa: E402

OUT = Path(__file__).resolve().parent
N = 40_000


def table(amp, sigma_scale):
    bg = synth.make_traffic(17, N, df17_per_s=2000, decoy_per_s=3000, snr_db=(10, 28), sigma=2.0 * sigma_scale,
                            include_golden=False)
    gold = list(synth.GOLDEN_FRAMES) + [flip_bit(synth.GOLDEN_FRAMES[k], b) for k, b in ((0, 9), (3, 50), (6, 87))]
    fg = synth.single_frames(gold, [1500 + 3500 * k for k in range(len(gold))], amp_i=int(amp * 0.8),
                             amp_q=int(amp * 0.6))
    return synth.FrameTable.concat([bg, fg])


def records(frames):
    return [{"hex": bytes(r["bytes"]).hex(), "offset": int(r["offset"]), "fixed_bit": int(r["fixed_bit"])} for r in frames]


def main():
    meta = {"generator": "tests/golden/make_golden.py", "oracle": "oracle_decode_literal", "n_samples": N}
    for key, fmt, sigma, amp, scale, seg in (("u8", synth.FMT_U8, 2.0, 50, 1.0, 20_000),
                                            ("cs16", synth.FMT_CS16, 250.0, 6000, 125.0, 0)):
        iq = synth.render(table(amp, scale), 17, 0, N, fmt, sigma)
        lit, gp = oracle_c.decode_literal(iq, seg)
        npf, gp2 = oracle_np.decode(iq, seg)
        assert gp == gp2 and lit.tobytes() == oracle_np.to_records(npf).tobytes()
        name = f"capture_small_{key}.bin"
        iq.tofile(OUT / name)
        meta[key] = {"file": name, "segment_samples": seg, "gate_passes": gp, "frames": records(lit)}
        print(key, len(lit), "frames,", gp, "gate passes,", (lit["fixed_bit"] != 255).sum(), "repaired")
    (OUT / "capture_small.json").write_text(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
