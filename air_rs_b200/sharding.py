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

import os
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
    """Decode one rank's shard in `pieces` sub-shards and exchange the frame lists with every
    other rank without ever stalling the GPU on the host.

    Each rank's buffer for a piece starts with a 24-byte header whose first 8 bytes are the
    frame count (the library writes it there itself: `d_count` of airgpu_decode_device points
    at it), followed by the ordered records.  Only `slab` rows travel: not the worst-case
    capacity but what the traffic needs (a little above the largest count seen so far, the
    same on every rank because every rank sees all counts); `finish()` checks after the fact
    that no rank produced more and repeats the exchange with more rows if one did.

    Two exchange back ends:
      "p2p"  (default on NVLink boxes) -- every rank owns a symmetric-memory slab set
             (torch.distributed._symmetric_memory, rendezvous over the NCCL group); a rank
             copies its rows straight into its slot of every peer's slabs with plain device
             copies on a high-priority stream, i.e. over NVLink by the copy engines, so the
             exchange takes no SMs away from the decode kernel of the next piece.  One
             symmetric-memory barrier per step makes the peers' writes visible.
      "nccl" -- one ncclAllGather per piece (used by the CPU/gloo tests and as the fallback).
    Result order: rank-major, piece-minor == ascending offset == the reference's order.
    """

    def __init__(self, decoder, n_local: int, first_sample: int, pieces: int = 2, cap_per_piece: int = 0,
                 group=None, exchange: str = "auto"):
        import torch
        import torch.distributed as dist

        self.dec = decoder
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dev = torch.device("cuda", decoder.device)
        cands = max(0, n_local - HALO)
        pieces = max(1, min(pieces, max(1, cands // ALIGN)))
        # later sub-shards are smaller: only the LAST exchange is exposed (nothing left to overlap it)
        w = [float(pieces - k + 1) for k in range(pieces)] if pieces > 1 else [1.0]
        if os.environ.get("AIRGPU_PIECE_WEIGHTS"):
            w = [float(x) for x in os.environ["AIRGPU_PIECE_WEIGHTS"].split(",")][:pieces]
        acc = [sum(w[:k]) / sum(w) for k in range(pieces)]
        b = [min(cands, int(cands * a) // ALIGN * ALIGN) for a in acc] + [cands]
        self.ranges = [(b[k], b[k + 1]) for k in range(pieces)]   # same number of exchanges on every rank
        self.first = first_sample
        cap = cap_per_piece or max(1 << 14, (max(e - s for s, e in self.ranges) + HALO) // 200)   # ~4x dense traffic
        if self.world > 1:   # every rank must use the same row count
            c = torch.tensor([cap], dtype=torch.int64, device=self.dev)
            dist.all_reduce(c, op=dist.ReduceOp.MAX, group=group)
            cap = int(c.item())
        self.cap = cap
        self.slab = cap                      # rows exchanged per rank and piece; shrinks after the first step
        P, W = len(self.ranges), self.world
        # row 0 = header (frame count in its first 8 bytes), rows 1.. = records
        self.out = [torch.zeros((cap + 1, RECORD_BYTES), dtype=torch.uint8, device=self.dev) for _ in range(P)]
        self.hdr_host = torch.zeros((P, W), dtype=torch.int64).pin_memory()
        # high priority: the exchange must get going while the decode kernel still has CTAs queued
        self.comm = torch.cuda.Stream(device=self.dev, priority=-1)
        self.decoded = [torch.cuda.Event() for _ in range(P)]
        self.gathered = [torch.cuda.Event() for _ in range(P)]
        self._first_step = True
        self.exchange = "nccl"
        self.symm = None
        self.gath = None
        if exchange in ("auto", "p2p") and self.world > 1:
            try:
                import torch.distributed._symmetric_memory as symm

                rows = P * W * (cap + 1)
                self._symm_t = symm.empty(rows * RECORD_BYTES, dtype=torch.uint8, device=self.dev)
                g = group if group is not None else dist.group.WORLD
                self.symm = symm.rendezvous(self._symm_t, g.group_name)
                self.gath_full = self._symm_t.view(P, W, cap + 1, RECORD_BYTES)
                self.peer_slots = [
                    [self.symm.get_buffer(q, (P, W, cap + 1, RECORD_BYTES), torch.uint8)[k][self.rank]
                     for k in range(P)] for q in range(W)]
                self.exchange = "p2p"
            except Exception:
                if exchange == "p2p":
                    raise
                self.symm = None

    def _alloc(self):
        import torch

        P, W = len(self.ranges), self.world
        if self.exchange == "p2p":
            self.gath = [self.gath_full[k] for k in range(P)]
        else:
            self.gath = [torch.empty((W, self.slab + 1, RECORD_BYTES), dtype=torch.uint8, device=self.dev)
                         for _ in range(P)]

    def _exchange(self, k):
        import torch.distributed as dist

        rows = self.slab + 1
        if self.exchange == "p2p":
            for q in range(self.world):      # my rows -> my slot in every rank's slab set (NVLink copy engines)
                self.peer_slots[(self.rank + q) % self.world][k][:rows].copy_(self.out[k][:rows], non_blocking=True)
        elif self.world > 1:
            dist.all_gather_into_tensor(self.gath[k].view(-1), self.out[k][:rows].reshape(-1), group=self.group)
        else:
            self.gath[k][0].copy_(self.out[k][:rows])
        self.gathered[k].record(self.comm)

    def _publish_counts(self):
        """After every piece of a step has been exchanged: barrier (p2p), then the counts of all ranks
        go to pinned host memory for finish()."""
        if self.exchange == "p2p":
            self.symm.barrier(channel=0)
        for k in range(len(self.ranges)):
            self.hdr_host[k].copy_(self.gath[k][:, 0, :8].contiguous().view(-1).view(dtype=self.hdr_host.dtype),
                                   non_blocking=True)

    def step(self, iq, bytes_per_sample: int = 2):
        """Queue one pass over this rank's shard (CUDA tensor of interleaved IQ).  Asynchronous."""
        import torch

        if self.gath is None:
            self._alloc()
        compute = torch.cuda.current_stream(self.dev)
        base_ptr = iq.data_ptr()
        for k, (s, e) in enumerate(self.ranges):
            if not self._first_step:
                compute.wait_event(self.gathered[k])     # out[k] of the previous step has been sent
            o = self.out[k]
            self.dec.decode_device(base_ptr + s * bytes_per_sample, e - s + HALO, o.data_ptr() + RECORD_BYTES,
                                   self.cap, 0, self.first + s, o.data_ptr(), compute.cuda_stream)
            self.decoded[k].record(compute)
        self._first_step = False
        with torch.cuda.stream(self.comm):
            for k in range(len(self.ranges)):
                self.comm.wait_event(self.decoded[k])
                self._exchange(k)
            self._publish_counts()

    def finish(self, concat: bool = True):
        """Wait for the last queued step; returns (frames [n, 24] in global order, n)."""
        import torch

        self.comm.synchronize()
        counts = self.hdr_host.clone()
        m = int(counts.max())
        if m > self.cap:
            raise ValueError(f"a sub-shard produced {m} frames but the buffers hold {self.cap}")
        if m > self.slab:
            # optimistic row count was too small (traffic got denser): exchange again with room to spare
            self.slab = min(self.cap, (int(m * 1.25) + 1023) // 1024 * 1024)
            if self.exchange != "p2p":
                self._alloc()
            with torch.cuda.stream(self.comm):
                for k in range(len(self.ranges)):
                    self._exchange(k)
                self._publish_counts()
            self.comm.synchronize()
            counts = self.hdr_host.clone()
        elif self.slab == self.cap and m < self.cap:
            self.slab = min(self.cap, (int(m * 1.15) + 1023) // 1024 * 1024)   # first step done: size for the traffic
            if self.exchange != "p2p":
                keep = self.gath
                self.gath = [g[:, : self.slab + 1].contiguous() for g in keep]
        parts, total = [], 0
        for r in range(self.world):
            for k in range(len(self.ranges)):
                c = int(counts[k, r])
                parts.append(self.gath[k][r, 1 : 1 + c])
                total += c
        return (torch.cat(parts) if concat else parts), total
