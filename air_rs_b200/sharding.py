"""Long-capture sharding across the GPUs of one node.

A candidate offset i reads samples [i, i+240) and nothing else (reference
src/adsb.rs:98-106: no state is carried between offsets), so a capture of N
samples shards naturally: the candidates [0, N-240) are cut into `world`
contiguous ranges and rank r decodes samples [a_r, b_r + 240) as one segment with
base_offset = a_r.  Every candidate is owned by exactly one rank; concatenating
the per-rank frame lists in rank order is the reference's emission order.  The
only exchange is the all-gather of those lists.

One process per GPU; `torch.distributed` supplies the plumbing (NCCL on GPUs,
gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

HALO = 240          # samples a candidate reads beyond its own offset, plus one (16 + 112*2)
ALIGN = 16384       # shard starts on a CTA-tile boundary (keeps 16-byte aligned loads)

RECORD_BYTES = 24   # sizeof(airgpu_frame)


def shard_bounds(n_samples: int, world: int, align: int = ALIGN) -> List[int]:
    """Candidate-range boundaries b[0..world]: rank r owns candidates [b[r], b[r+1])."""
    cands = max(0, n_samples - HALO)
    b = [min(cands, (cands * r // world) // align * align) for r in range(world)]
    b.append(cands)
    return b


def shard_samples(n_samples: int, world: int, rank: int, align: int = ALIGN) -> Tuple[int, int]:
    """(first sample, number of samples) rank `rank` must hold: its candidates + the halo."""
    b = shard_bounds(n_samples, world, align)
    first, last = b[rank], b[rank + 1]
    if last <= first:
        return first, 0
    return first, last - first + HALO


def allgather_frames(frames, count, group=None):
    """All-gather per-rank ordered frame lists.

    frames: uint8 tensor [cap, 24] on this rank's device (first `count` rows valid)
    count:  int64 tensor [1] on the same device (device-resident count is fine)
    Returns (slab [world, m, 24], counts [world] on the host) where m = max count;
    rank r's frames are slab[r, :counts[r]].  Two collectives: counts, then records
    padded to the largest count.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    counts_dev = torch.empty(world, dtype=torch.int64, device=frames.device)
    dist.all_gather_into_tensor(counts_dev, count.reshape(1), group=group)
    counts = counts_dev.cpu()
    m = int(counts.max())
    slab = torch.empty((world, max(m, 1), RECORD_BYTES), dtype=torch.uint8, device=frames.device)
    if m > frames.shape[0]:
        raise ValueError(f"a rank produced {m} frames but the local buffer holds {frames.shape[0]}")
    dist.all_gather_into_tensor(slab.view(-1), frames[: max(m, 1)].reshape(-1), group=group)
    return slab, counts


def concat_gathered(slab, counts):
    """Global ordered frame list [sum(counts), 24] from an all-gathered slab."""
    import torch

    parts = [slab[r, : int(c)] for r, c in enumerate(counts.tolist())]
    return torch.cat(parts, dim=0) if parts else slab.new_zeros((0, RECORD_BYTES))


class ShardedDecoder:
    """Decode one rank's shard in `pieces` sub-shards and all-gather the frame lists while
    the next sub-shard is being decoded.

    The decode kernels of all pieces are queued back to back on the compute stream; the
    two collectives of piece p (counts, then records padded to the largest count) run on
    a second stream as soon as piece p is done, so on NVLink the exchange hides behind
    the decode of piece p+1 and only the last piece's gather is exposed.
    Result order: rank-major, piece-minor == ascending offset == the reference's order.
    """

    def __init__(self, decoder, n_local: int, first_sample: int, pieces: int = 4, cap_per_piece: int = 0,
                 group=None):
        import torch
        import torch.distributed as dist

        self.dec = decoder
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.dev = torch.device("cuda", decoder.device)
        cands = max(0, n_local - HALO)
        pieces = max(1, min(pieces, max(1, cands // ALIGN)))
        b = [min(cands, (cands * k // pieces) // ALIGN * ALIGN) for k in range(pieces)] + [cands]
        self.ranges = [(b[k], b[k + 1]) for k in range(pieces)]   # same number of collectives on every rank
        self.first = first_sample
        cap = cap_per_piece or max(1 << 14, (max(e - s for s, e in self.ranges) + HALO) // 200)   # ~4x dense traffic
        self.cap = cap
        self.out = [torch.empty((cap, RECORD_BYTES), dtype=torch.uint8, device=self.dev) for _ in self.ranges]
        self.cnt = [torch.zeros(1, dtype=torch.int64, device=self.dev) for _ in self.ranges]
        # high priority: NCCL's few CTAs must get SM slots while the decode kernel still has CTAs queued
        self.comm = torch.cuda.Stream(device=self.dev, priority=-1)
        self.events = [torch.cuda.Event() for _ in self.ranges]

    def step(self, iq, bytes_per_sample: int = 2, concat: bool = True):
        """iq: this rank's shard (CUDA tensor, interleaved IQ).  Returns (frames [n, 24], n)."""
        import torch

        compute = torch.cuda.current_stream(self.dev)
        base_ptr = iq.data_ptr()
        for k, (s, e) in enumerate(self.ranges):
            self.dec.decode_device(base_ptr + s * bytes_per_sample, e - s + HALO, self.out[k].data_ptr(), self.cap,
                                   0, self.first + s, self.cnt[k].data_ptr(), compute.cuda_stream)
            self.events[k].record(compute)
        if self.world == 1:
            counts = torch.cat(self.cnt).cpu().tolist()
            parts = [self.out[k][:c] for k, c in enumerate(counts)]
            total = sum(counts)
            return (torch.cat(parts) if concat else parts), total
        gathered = []
        with torch.cuda.stream(self.comm):
            for k in range(len(self.ranges)):
                self.comm.wait_event(self.events[k])
                gathered.append(allgather_frames(self.out[k], self.cnt[k], self.group))
            # rank-major, piece-minor
            parts, total = [], 0
            for r in range(self.world):
                for slab, counts in gathered:
                    c = int(counts[r])
                    parts.append(slab[r, :c])
                    total += c
            result = torch.cat(parts) if concat else parts
        compute.wait_stream(self.comm)
        return result, total
