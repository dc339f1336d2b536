"""Probe (torchrun, >= 2 GPUs): symmetric memory with an NVSwitch multicast mapping, and the library's fused exchange on it."""
import os, sys, time
from pathlib import Path
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm
for backend in ("default",):
    try:
        t = symm.empty(16 << 20, dtype=torch.uint8, device=f"cuda:{local}")
        t.zero_()
        hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
        mc = getattr(hdl, "multicast_ptr", None)
        print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast_ptr", hex(mc) if mc else mc,
              "signal pads", [hex(p) for p in getattr(hdl, "signal_pad_ptrs", [])][:2], flush=True)
    except Exception as e:
        import traceback; traceback.print_exc()
        print(rank, "SYMM FAILED", type(e).__name__, e, flush=True)
        dist.destroy_process_group(); sys.exit(0)
from air_rs_b200 import synth
from air_rs_b200.decoder import AdsbDecoder
from air_rs_b200.native import FMT_U8, FRAME_DTYPE
n = 4_000_000
tab = synth.make_traffic(5 + rank, n, df17_per_s=3000, decoy_per_s=1000, snr_db=(10, 30))
iq = torch.from_numpy(synth.render(tab, 5 + rank, 0, n, FMT_U8, 2.0)).cuda()
dec = AdsbDecoder(fmt=FMT_U8, device=local)
cap = 1 << 15
slot = (cap + 1) * 24
flags_off = world * slot
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
ref_out, ref_n = dec.decode_tensor(iq, cap=cap)
ref = ref_out[:ref_n].cpu().numpy().tobytes()
for mode in (["multicast"] if mc else []) + ["peers"]:
    t.zero_(); torch.cuda.synchronize(); dist.barrier()
    if mode == "multicast":
        outs, counts = [mc + rank * slot + 24], [mc + rank * slot]
    else:
        outs = [p + rank * slot + 24 for p in hdl.buffer_ptrs]; counts = [p + rank * slot for p in hdl.buffer_ptrs]
    flag_ptrs = [p + flags_off for p in hdl.buffer_ptrs]
    for epoch in (1, 2, 3):
        dec.decode_device_peers(iq.data_ptr(), n, outs, counts, cap, stream=s.cuda_stream, multicast=(mode == "multicast"))
        dec.peer_barrier(flag_ptrs, rank, epoch, stream=s.cuda_stream)
    torch.cuda.synchronize()
    # every rank now holds every rank's list; gather the reference lists over NCCL to compare
    mine = torch.zeros(cap * 24 + 8, dtype=torch.uint8, device="cuda")
    mine[:8] = torch.tensor([ref_n], dtype=torch.int64).view(torch.uint8).cuda()
    mine[8:8 + ref_n * 24] = ref_out[:ref_n].reshape(-1)
    allr = torch.empty(world * (cap * 24 + 8), dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(allr, mine)
    ok = True
    for q in range(world):
        want_n = int(allr[q * (cap * 24 + 8): q * (cap * 24 + 8) + 8].view(torch.int64).item())
        got_n = int(t[q * slot: q * slot + 8].view(torch.int64).item())
        want = allr[q * (cap * 24 + 8) + 8: q * (cap * 24 + 8) + 8 + want_n * 24]
        got = t[q * slot + 24: q * slot + 24 + want_n * 24]
        ok = ok and want_n == got_n and bool(torch.equal(want, got))
    print(rank, mode, "exchange equals every rank's own decode:", ok, "frames", ref_n, flush=True)
    # time it
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for epoch in range(4, 24):
        dec.decode_device_peers(iq.data_ptr(), n, outs, counts, cap, stream=s.cuda_stream, multicast=(mode == "multicast"))
        dec.peer_barrier(flag_ptrs, rank, epoch, stream=s.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    print(rank, mode, f"{e0.elapsed_time(e1) / 20 * 1e3:.1f} us per decode+exchange+barrier of {n} samples", flush=True)
dist.destroy_process_group()
