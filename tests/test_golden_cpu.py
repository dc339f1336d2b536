"""CPU: the committed golden fixture still equals what both oracle restatements emit."""
import json
from pathlib import Path

import numpy as np

from oracle import oracle_c, oracle_np

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


def test_golden_fixture_matches_oracles():
    meta = json.loads((GOLDEN_DIR / "capture_small.json").read_text())
    for key, dt in (("u8", np.uint8), ("cs16", np.int16)):
        iq = np.fromfile(GOLDEN_DIR / meta[key]["file"], dtype=dt)
        assert iq.size == 2 * meta["n_samples"]
        seg = meta[key]["segment_samples"]
        want = [(f["hex"], f["offset"], f["fixed_bit"]) for f in meta[key]["frames"]]
        for frames, gp in (oracle_c.decode_literal(iq, seg), oracle_c.decode_fast(iq, seg, threads=2)):
            assert [(bytes(r["bytes"]).hex(), int(r["offset"]), int(r["fixed_bit"])) for r in frames] == want
            assert gp == meta[key]["gate_passes"]
        npf, _ = oracle_np.decode(iq, seg)
        assert [(b.hex(), off, fx) for b, fx, off in npf] == want
        # the reference's own frames are in there, unrepaired
        hexes = {w[0] for w in want if w[2] == 255}
        assert "8d406b902015a678d4d220aa4bda" in hexes
