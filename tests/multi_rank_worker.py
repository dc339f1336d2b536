"""Worker of tests/test_parity_gpu.py::test_multi_rank_exchange_equals_single_gpu_decode (run under torchrun).

Every rank renders its shard of one capture, the ranks exchange their ordered frame lists with each back end, and
every rank compares the list it ends up with, byte for byte, with rank 0's single-GPU decode of the WHOLE capture
(broadcast over NCCL)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from air_rs_b200 import sharding, synth  # noqa: E402
from air_rs_b200.decoder import AdsbDecoder  # noqa: E402
from air_rs_b200.native import FMT_U8  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    total = 40_000_000 + 12_345
    period = 4_800_000
    tab = synth.make_traffic(77, period, df17_per_s=3000.0, decoy_per_s=3000.0, snr_db=(8.0, 30.0))
    gen = synth.DeviceSynth(tab, device=local)
    a, n_local = sharding.shard_samples(total, world, rank)
    iq = gen.render(77, a, n_local, FMT_U8, 2.0, period=period)
    dec = AdsbDecoder(fmt=FMT_U8, device=local)
    # the truth: rank 0 decodes the whole capture on its one GPU
    cap_all = total // 200
    ref = torch.zeros((cap_all, 24), dtype=torch.uint8, device=dev)
    ref_n = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == 0:
        whole = gen.render(77, 0, total, FMT_U8, 2.0, period=period)
        out, n = dec.decode_tensor(whole, cap=cap_all)
        ref.copy_(out)
        ref_n.fill_(n)
        del whole
    dist.broadcast(ref_n, 0)
    dist.broadcast(ref, 0)
    n_ref = int(ref_n.item())
    assert n_ref > 10_000
    stream = torch.cuda.Stream(device=dev)
    ok_all = True
    for exchange in ("multicast", "peers", "nccl"):
        for use_graph in ((False, True) if exchange != "nccl" else (False,)):
            for pieces in (1, 3):
                with torch.cuda.stream(stream):
                    try:
                        sd = sharding.ShardedDecoder(dec, n_local, a, pieces=pieces, exchange=exchange, use_graph=use_graph)
                    except RuntimeError as e:
                        if exchange == "multicast":          # a box without an NVSwitch multicast mapping
                            if rank == 0:
                                print(f"multicast unavailable: {e}", flush=True)
                            continue
                        raise
                    for _ in range(5):                       # eager steps, then recorded and replayed ones, both parities
                        sd.step(iq)
                    sd.wait()
                    frames, n = sd.finish()
                    good = n == n_ref and bool(torch.equal(frames, ref[:n_ref]))
                    sd.step(iq)                              # and once more after a finish()
                    frames, n = sd.finish()
                    good = good and n == n_ref and bool(torch.equal(frames, ref[:n_ref]))
                    sd.close()
                flag = torch.tensor([1 if good else 0], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if rank == 0:
                    print(f"world={world} exchange={sd.exchange} graph={use_graph} pieces={pieces}: "
                          f"{'equal to the single-GPU decode' if int(flag.item()) else 'MISMATCH'} ({n_ref} frames)", flush=True)
                ok_all = ok_all and bool(int(flag.item()))
                del sd
                torch.cuda.synchronize()
                dist.barrier()
    if rank == 0 and ok_all:
        print(f"ALL OK world={world}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
