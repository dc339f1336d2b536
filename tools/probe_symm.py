"""Probe: does torch symmetric memory work on this box (peer pointers across ranks)?"""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(64 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok", type(hdl).__name__, [a for a in dir(hdl) if not a.startswith("_")][:30], flush=True)
    ptrs = list(hdl.buffer_ptrs)
    print(rank, "ptrs", [hex(p) for p in ptrs], flush=True)
    # write my rank id pattern into every peer's buffer at my slot using get_buffer views
    n = 16 << 20
    src = torch.full((n,), rank + 1, dtype=torch.uint8, device=f"cuda:{local}")
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for it in range(10):
        for peer in range(world):
            dst = hdl.get_buffer(peer, (n,), torch.uint8, storage_offset=rank * n)
            dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    hdl.barrier()
    dist.barrier()
    ok = all(int(t[r * n].item()) == r + 1 and int(t[r * n + n - 1].item()) == r + 1 for r in range(world))
    print(rank, "peer copies ok:", ok, f"{world * n / dt / 1e9:.1f} GB/s out per rank", flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "SYMM FAILED", type(e).__name__, e, flush=True)
dist.destroy_process_group()
