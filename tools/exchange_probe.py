"""torchrun probe: where does a multi-GPU step spend its time?  Times, with CUDA events on one stream, increasingly
complete versions of the per-rank step on this rank's shard of the bench capture."""
import os, sys
from pathlib import Path
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from air_rs_b200 import sharding, synth
from air_rs_b200.decoder import AdsbDecoder
from air_rs_b200.native import FMT_U8
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
total = int(os.environ.get("AIRGPU_BENCH_SAMPLES", 8_640_000_000)); period = 24_000_000
tab = synth.make_traffic(1090, period, df17_per_s=3000.0, decoy_per_s=3000.0, snr_db=(8.0, 30.0), sigma=2.0)
a, n_local = sharding.shard_samples(total, world, rank)
iq = synth.DeviceSynth(tab, device=local).render(1090, a, n_local, FMT_U8, 2.0, period=period)
dec = AdsbDecoder(fmt=FMT_U8, device=local)
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)
cap = n_local // 200
out = torch.empty((cap, 24), dtype=torch.uint8, device=dev); cnt = torch.zeros(1, dtype=torch.int64, device=dev)

WAIT = [None]


def timeit(name, fn, reps=20):
    for _ in range(4):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    if WAIT[0] is not None:
        WAIT[0].wait()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    mx = ms.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    mn = ms.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{name:<58} max {mx.item():.4f} ms   min {mn.item():.4f} ms", flush=True)

timeit("plain decode_device, whole shard, local out", lambda: dec.decode_device(iq.data_ptr(), n_local, out.data_ptr(), cap, 0, a, cnt.data_ptr(), s.cuda_stream))
dec.set_timing(False)
timeit("  same, timing events off", lambda: dec.decode_device(iq.data_ptr(), n_local, out.data_ptr(), cap, 0, a, cnt.data_ptr(), s.cuda_stream))
for exchange in os.environ.get("PROBE_EXCHANGE", "multicast,peers").split(","):
    for pieces in [int(x) for x in os.environ.get("PROBE_PIECES", "1,2").split(",")]:
        for graph in (False, True):
            sd = sharding.ShardedDecoder(dec, n_local, a, pieces=pieces, exchange=exchange, use_graph=graph)
            WAIT[0] = sd
            timeit(f"ShardedDecoder {exchange} pieces={pieces} graph={graph}", lambda: sd.step(iq))
            frames, n = sd.finish()
            WAIT[0] = None
            sd.close(); del sd
            torch.cuda.synchronize(); dist.barrier()
dist.destroy_process_group()
