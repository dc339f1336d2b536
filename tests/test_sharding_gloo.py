"""CPU, world_size 2 and 3 over gloo: the shard plan and the frame all-gather
(air_rs_b200/sharding.py) reproduce the single-pass result.  Each rank decodes its
shard with the CPU oracle here (no GPU in this test); on the GPU the same plan is
exercised by tests/test_parity_gpu.py::test_sharding_with_halo_equals_single_pass
and by bench.py --gpus N."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from air_rs_b200 import sharding, synth  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    from oracle import oracle_c

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tab = synth.make_traffic(77, n, df17_per_s=3000, decoy_per_s=3000, snr_db=(8, 30))
        first, count = sharding.shard_samples(n, world, rank, align=8)
        iq = synth.render(tab, 77, first, count, synth.FMT_U8, 2.0)      # each rank renders only its shard
        frames, _ = oracle_c.decode_fast(iq, 0, base=first)
        cap = 4096
        buf = torch.zeros((cap, 24), dtype=torch.uint8)
        buf[: len(frames)] = torch.from_numpy(frames.view(np.uint8).reshape(-1, 24))
        slab, counts = sharding.allgather_frames(buf, torch.tensor([len(frames)], dtype=torch.int64))
        allf = sharding.concat_gathered(slab, counts).numpy().tobytes()
        if rank == 0:
            q.put((allf, counts.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_decode_equals_single_pass(world):
    from oracle import oracle_c

    n = 300_000
    tab = synth.make_traffic(77, n, df17_per_s=3000, decoy_per_s=3000, snr_db=(8, 30))
    whole, _ = oracle_c.decode_fast(synth.render(tab, 77, 0, n, synth.FMT_U8, 2.0))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, counts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(counts) == len(whole) and all(c > 0 for c in counts)
    assert got == whole.tobytes()


def test_shard_plan_properties():
    for n in (0, 100, 240, 241, 1_000_000, 8_640_000_000):
        for world in (1, 2, 4, 8):
            b = sharding.shard_bounds(n, world)
            assert b[0] == 0 and b[-1] == max(0, n - 240) and all(x <= y for x, y in zip(b, b[1:]))
            assert all(x % sharding.ALIGN == 0 for x in b[:-1])
            cover = 0
            for r in range(world):
                first, count = sharding.shard_samples(n, world, r)
                if count:
                    assert first == b[r] and first + count <= n
                    cover += count - 240
            assert cover == max(0, n - 240)
