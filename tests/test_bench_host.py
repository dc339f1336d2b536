"""CPU: host-side logic of bench.py that the GPU run depends on (no device needed)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from oracle import oracle_c  # noqa: E402

from common import capture_u8  # noqa: E402


def test_full_capture_check_slices_and_compares():
    """bench.full_capture_check with the oracle's own output standing in for the GPU list: equal -> true, one
    record changed / dropped -> false; slices that do not divide the capture, a slice larger than the capture."""
    n = 1_000_123
    _, iq = capture_u8(seed=8, n=n, df17=4000.0)
    want, _ = oracle_c.decode_fast(iq)
    assert len(want) > 500
    t_iq = torch.from_numpy(iq)
    as_tensor = lambda fr: torch.from_numpy(fr.view(np.uint8).reshape(-1, 24).copy())
    for slice_c in (240_000, 99_999, 5_000_000):
        r = bench.full_capture_check(t_iq, as_tensor(want), len(want), n, slice_c)
        assert r["whole_capture_frames_equal_oracle"] is True and r["frames"] == len(want)
        assert r["single_bit_repairs"] == int((want["fixed_bit"] != 0xFF).sum())
    bad = want.copy()
    bad["bytes"][len(bad) // 2, 3] ^= 1
    assert bench.full_capture_check(t_iq, as_tensor(bad), len(bad), n, 240_000)["whole_capture_frames_equal_oracle"] is False
    assert bench.full_capture_check(t_iq, as_tensor(want[:-1]), len(want) - 1, n, 240_000)["whole_capture_frames_equal_oracle"] is False


def test_shard_plan_covers_every_candidate_once():
    from air_rs_b200 import sharding

    for n in (0, 100, 240, 241, 20_000, 8_640_000_000, 40_012_345):
        for world in (1, 2, 3, 8):
            b = sharding.shard_bounds(n, world)
            assert b[0] == 0 and b[-1] == max(0, n - 240) and all(x <= y for x, y in zip(b, b[1:]))
            assert all(x % sharding.ALIGN == 0 for x in b[:-1])
            for r in range(world):
                first, cnt = sharding.shard_samples(n, world, r)
                assert (cnt == 0) == (b[r + 1] <= b[r]) and (cnt == 0 or cnt == b[r + 1] - b[r] + 240)
    assert sharding.sub_ranges(1_000_000, 3)[0][0] == 0 and sharding.sub_ranges(1_000_000, 3)[-1][1] == 1_000_000
