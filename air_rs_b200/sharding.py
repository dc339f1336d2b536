"""Long-capture sharding across the GPUs of one node.

A candidate offset i reads samples [i, i+240) and nothing else (reference
src/adsb.rs:98-106: no state is carried between offsets), so a capture of N
samples shards naturally: the candidates [0, N-240) are cut into `world`
contiguous ranges and rank r decodes samples [a_r, b_r + 240) as one segment with
base_offset = a_r.  Every candidate is owned by exactly one rank; concatenating
the per-rank frame lists in rank order is the reference's emission order.  The
only exchange is that of the per-rank ordered lists.

One process per GPU; `torch.distributed` supplies the rendezvous (NCCL on GPUs,
gloo in the CPU tests).  The exchange itself is done by the library's own ordering
kernels (airgpu_decode_device_peers): they store every ordered record -- and the
frame count -- straight into every rank's slab, through one NVSwitch multicast
address when the box offers one, else through the peer-mapped addresses.  NCCL's
all-gather is kept as the third back end (and is what gloo runs in the CPU tests).
"""
from __future__ import annotations

import os
import sys
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


def sub_ranges(cands: int, pieces: int, weights=None) -> List[Tuple[int, int]]:
    """Cut a shard's candidates [0, cands) into `pieces` contiguous sub-shards on ALIGN boundaries.  Later ones are
    smaller: only the LAST exchange of a step has nothing left to overlap with."""
    pieces = max(1, min(pieces, max(1, cands // ALIGN)))
    w = list(weights)[:pieces] if weights else ([float(pieces - k + 1) for k in range(pieces)] if pieces > 1 else [1.0])
    acc = [sum(w[:k]) / sum(w) for k in range(pieces)]
    b = [min(cands, int(cands * a) // ALIGN * ALIGN) for a in acc] + [cands]
    return [(b[k], b[k + 1]) for k in range(pieces)]


class ShardedDecoder:
    """Decode one rank's shard in `pieces` sub-shards and exchange the ordered frame lists with every other rank,
    without ever stalling the GPU on the host.

    Slab layout (every rank holds the same): [parity 2][piece P][rank W][1 + cap rows of 24 bytes]; row 0 of a slot is
    a header whose first 8 bytes are the frame count, rows 1.. are the records.  Rank r's ordering kernels write
    slot [parity][k][r] of EVERY rank's slab.

    Back ends (`exchange`):
      "multicast" -- the slab set lives in symmetric memory (torch.distributed._symmetric_memory: a CUDA VMM
                  allocation mapped into every rank, plus one NVSwitch multicast mapping).  The library's gather
                  kernel stores each record once, to the multicast address; the switch replicates it to all ranks.
                  Per-rank NVLink egress: 1x its list.
      "peers"  -- same slab, the gather kernel stores each record to every rank's peer-mapped address ((W-1)x egress).
      "nccl"   -- records stay local, one ncclAllGather per sub-shard on a side stream (the portable fallback, and what
                  the CPU tests run over gloo).
      "auto"   -- peers (measured faster than multicast once two steps are in flight: its stores hide behind the next
                  decode), else nccl when the box has no symmetric memory; the reason for a downgrade is kept in
                  `exchange_note` and printed to stderr -- never silent.
    multicast / peers: a step is `pieces` x (memset + decode + scan + gather) and ONE barrier kernel
    (airgpu_peer_barrier: release/acquire flags in the slab), recorded once per parity into a CUDA graph and replayed
    with one launch per step (`use_graph`).  TWO STEPS ARE IN FLIGHT: even and odd steps run on two streams (forked
    from the caller's) with two library contexts (two workspaces) and the two slab parities, so the ordering kernels,
    the NVLink traffic and the barrier of step t -- an all-gather lands (W-1)/W of the whole list in every GPU,
    ~0.15 ms of link time per step at 8 GPUs -- run while the decode kernel of step t+1 has the SMs, and the first
    CTAs of step t+1 fill the tail of step t's decode kernel.  `wait()` joins both streams into the caller's.

    Hazards: slabs are double-buffered by step parity and every step ends in a barrier.  Step t+2 runs on the same
    stream as step t, i.e. after step t's barrier: by then every peer has finished writing step t's slab AND has
    passed its own barrier of step t, hence has executed everything it queued before its step t+2 on that stream --
    including the reads `finish()` queued there.  Consumers of `finish(concat=False)` views must queue their reads
    before calling `step()` twice.

    Result order: rank-major, piece-minor == ascending offset == the reference's order.
    """

    def __init__(self, decoder, n_local: int, first_sample: int, pieces: int = 0, cap_per_piece: int = 0,
                 group=None, exchange: str = "auto", use_graph: bool = True, bytes_per_sample: int = 2):
        import torch
        import torch.distributed as dist

        self.dec = decoder
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dev = torch.device("cuda", decoder.device)
        self.bps = bytes_per_sample
        cands = max(0, n_local - HALO)
        weights = None
        if os.environ.get("AIRGPU_PIECE_WEIGHTS"):
            weights = [float(x) for x in os.environ["AIRGPU_PIECE_WEIGHTS"].split(",")]
        self.first = first_sample
        self.exchange_note = ""
        self.symm = None
        self._slab_t = None
        # the slab must be sized before the back end is probed (the symmetric allocation is the probe), so the
        # number of sub-shards is fixed first: 0 = automatic -- one for the fused back ends (two steps in flight hide
        # the exchange), two for the NCCL fallback (its all-gather of sub-shard 0 overlaps the decode of sub-shard 1)
        want_nccl = exchange == "nccl" or (self.world == 1 and exchange == "auto")
        if pieces <= 0:
            pieces = 2 if want_nccl else 1
        self.ranges = sub_ranges(cands, pieces, weights)          # same number of exchanges on every rank
        cap = cap_per_piece or max(1 << 14, (max(e - s for s, e in self.ranges) + HALO) // 200)   # ~4x dense traffic
        if self.world > 1:   # every rank must use the same row count
            c = torch.tensor([cap], dtype=torch.int64, device=self.dev)
            dist.all_reduce(c, op=dist.ReduceOp.MAX, group=group)
            cap = int(c.item())
        self.cap = cap
        P, W = len(self.ranges), self.world
        self.rows = cap + 1
        self.slot_bytes = self.rows * RECORD_BYTES
        self.parity_bytes = P * W * self.slot_bytes
        self.flags_off = 2 * self.parity_bytes                    # per lane [W + 2] u64: epochs seen from each rank, own counter, time-out mark
        self.flags_bytes = 8 * (W + 2)
        self.hdr_host = torch.zeros((2, P, W), dtype=torch.int64).pin_memory()
        self.step_no = 0
        self.use_graph = use_graph
        self._graphs = [None, None]
        self._graph_key = None
        self._own_stream = None
        self.exchange = self._pick_exchange(exchange)
        total_bytes = self.flags_off + 2 * self.flags_bytes
        if self.exchange in ("multicast", "peers"):
            self.decode_reserved = False
            # two lanes (even / odd steps): a context and a stream each
            self._decs = [decoder, type(decoder)(fmt=decoder.fmt, device=decoder.device, ring_slots=1,
                                                 max_buffer_samples=1024, max_frames=64)]
            self._lanes = [torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev)]
        else:
            # nccl: local records, all-gathered into `gath` on a high-priority side stream
            self._slab_t = torch.zeros(total_bytes, dtype=torch.uint8, device=self.dev)
            self.base_ptrs = [self._slab_t.data_ptr()] * W
            self.comm = torch.cuda.Stream(device=self.dev, priority=-1)
            self.out = [torch.zeros((self.rows, RECORD_BYTES), dtype=torch.uint8, device=self.dev) for _ in range(P)]
            self.decoded = [torch.cuda.Event() for _ in range(P)]
            self.gathered = [torch.cuda.Event() for _ in range(P)]
            self.slab_rows = self.rows                            # rows exchanged per rank and piece; shrinks after the first step
        self._slab = self._slab_t.view(torch.uint8)[: 2 * self.parity_bytes].view(2, P, W, self.rows, RECORD_BYTES)

    # ------------------------------------------------------------------ set-up
    def _pick_exchange(self, want: str) -> str:
        import torch
        import torch.distributed as dist

        if want not in ("auto", "multicast", "peers", "nccl"):
            raise ValueError(f"unknown exchange back end {want!r}")
        if want == "nccl" or self.world == 1 and want == "auto":
            return "nccl"
        total_bytes = self.flags_off + 2 * self.flags_bytes
        try:
            import torch.distributed._symmetric_memory as symm

            t = symm.empty(total_bytes, dtype=torch.uint8, device=self.dev)
            t.zero_()
            g = self.group if self.group is not None else dist.group.WORLD
            hdl = symm.rendezvous(t, g.group_name)
            self._slab_t, self.symm = t, hdl
            self.base_ptrs = [int(p) for p in hdl.buffer_ptrs]
            self.mc_ptr = int(getattr(hdl, "multicast_ptr", 0) or 0)
            torch.cuda.synchronize(self.dev)
            dist.barrier(group=self.group)                        # every rank's slab is zeroed before anyone writes to it
        except Exception as e:                                    # no symmetric memory on this box / build
            if want != "auto":
                raise
            self.exchange_note = f"symmetric memory unavailable ({type(e).__name__}: {e}); NCCL all-gather instead"
            print(f"[airgpu sharding] {self.exchange_note}", file=sys.stderr)
            return "nccl"
        if want == "multicast" and not self.mc_ptr:
            raise RuntimeError("no NVSwitch multicast mapping for the symmetric slab on this box")
        if want == "multicast" and self.mc_ptr:
            return "multicast"
        # "auto" takes plain peer stores even when the box offers a multicast mapping.  Measured on this pool
        # (profiles/r2_exchange_*.txt): with two steps in flight the plain stores to the peers' mapped slabs hide
        # completely behind the next step's decode kernel (2 GPUs: 2.161 ms per step vs 2.164 ms for the decode
        # alone), while multimem.st stores do not (2.245 ms at 2 GPUs, 0.710 ms vs 0.523 ms of decode at 8 GPUs).
        return "peers"

    def _slot_off(self, parity: int, k: int, r: int) -> int:
        return parity * self.parity_bytes + (k * self.world + r) * self.slot_bytes

    # ------------------------------------------------------------------ one step
    def _enqueue_fused(self, base_ptr: int, parity: int, stream: int) -> None:
        """pieces x (decode + ordering kernels storing to every rank) + one barrier, on the lane's stream."""
        dec = self._decs[parity]
        for k, (s, e) in enumerate(self.ranges):
            off = self._slot_off(parity, k, self.rank)
            if self.exchange == "multicast":
                outs, counts = [self.mc_ptr + off + RECORD_BYTES], [self.mc_ptr + off]
            else:
                # own slab first, then the peers starting with the next rank (spreads the NVLink traffic)
                order = [(self.rank + q) % self.world for q in range(self.world)]
                outs = [self.base_ptrs[q] + off + RECORD_BYTES for q in order]
                counts = [self.base_ptrs[q] + off for q in order]
            dec.decode_device_peers(base_ptr + s * self.bps, e - s + HALO, outs, counts, self.cap, 0, self.first + s,
                                    stream, multicast=self.exchange == "multicast")
        # each lane has its own flag arrays: the barriers of two consecutive steps may run at the same time
        dec.peer_barrier([p + self.flags_off + parity * self.flags_bytes for p in self.base_ptrs], self.rank, 0, stream)

    def step(self, iq, bytes_per_sample: int = None):
        """Queue one pass over this rank's shard (CUDA tensor of interleaved IQ).  Asynchronous; ordered after the
        work already queued on the current stream.  Call wait() (or finish()) before the current stream uses the result."""
        import torch

        if bytes_per_sample is not None:
            self.bps = bytes_per_sample
        cur = torch.cuda.current_stream(self.dev)
        parity = self.step_no & 1
        if self.exchange == "nccl":
            stream = cur
            if cur.cuda_stream == 0:
                # the legacy default stream: the library would fall back to its own (non-blocking) compute stream, which
                # is not ordered against it -- run the step on a private stream bracketed by the default stream instead
                if self._own_stream is None:
                    self._own_stream = torch.cuda.Stream(device=self.dev)
                stream = self._own_stream
                stream.wait_stream(cur)
            self._step_nccl(iq, parity, stream)
            if stream is not cur:
                cur.wait_stream(stream)
        else:
            stream = self._lanes[parity]
            stream.wait_stream(cur)                  # the producer of `iq`
            if not self.decode_reserved:
                # size the workspaces once: nothing may allocate inside a graph capture
                for d in self._decs:
                    d.reserve(max(e - s for s, e in self.ranges) + HALO, 0, self.cap)
                    d.set_timing(False)
                self.decode_reserved = True
            key = iq.data_ptr()
            if self.use_graph and self._graph_key != key:
                for g in self._graphs:
                    if g is not None:
                        g.close()
                self._graphs, self._graph_key = [None, None], key
            dec = self._decs[parity]
            if self.use_graph and self._graphs[parity] is None and self.step_no >= 2:
                # record this lane's sequence once (steps 0 and 1 run eagerly: they warm everything up)
                dec.graph_begin(stream.cuda_stream)
                try:
                    self._enqueue_fused(iq.data_ptr(), parity, stream.cuda_stream)
                finally:
                    self._graphs[parity] = dec.graph_end(stream.cuda_stream)
            if self.use_graph and self._graphs[parity] is not None:
                self._graphs[parity].launch(stream.cuda_stream)
            else:
                self._enqueue_fused(iq.data_ptr(), parity, stream.cuda_stream)
            with torch.cuda.stream(stream):
                self.hdr_host[parity].copy_(self._slab[parity, :, :, 0, :8].contiguous().view(-1).view(torch.int64)
                                            .view(len(self.ranges), self.world), non_blocking=True)
        self.last_stream = stream
        self.step_no += 1

    def _step_nccl(self, iq, parity: int, compute) -> None:
        import torch
        import torch.distributed as dist

        P = len(self.ranges)
        base_ptr = iq.data_ptr()
        for k, (s, e) in enumerate(self.ranges):
            if self.step_no > 0:
                compute.wait_event(self.gathered[k])     # out[k] of the previous step has been sent
            o = self.out[k]
            self.dec.decode_device(base_ptr + s * self.bps, e - s + HALO, o.data_ptr() + RECORD_BYTES,
                                   self.cap, 0, self.first + s, o.data_ptr(), compute.cuda_stream)
            self.decoded[k].record(compute)
        rows = self.slab_rows
        with torch.cuda.stream(self.comm):
            for k in range(P):
                self.comm.wait_event(self.decoded[k])
                dst = self._slab[parity, k, :, :rows]
                if self.world > 1:
                    tmp = torch.empty((self.world, rows, RECORD_BYTES), dtype=torch.uint8, device=self.dev)
                    dist.all_gather_into_tensor(tmp.view(-1), self.out[k][:rows].reshape(-1), group=self.group)
                    dst.copy_(tmp)
                else:
                    dst[0].copy_(self.out[k][:rows])
                self.gathered[k].record(self.comm)
            self.hdr_host[parity].copy_(self._slab[parity, :, :, 0, :8].contiguous().view(-1).view(torch.int64)
                                        .view(P, self.world), non_blocking=True)

    def wait(self) -> None:
        """Make the current stream wait for the exchange of the last queued step (a timed region ends here)."""
        import torch

        cur = torch.cuda.current_stream(self.dev)
        if self.exchange == "nccl":
            cur.wait_stream(self.comm)
        else:
            for st in self._lanes:
                cur.wait_stream(st)

    def finish(self, concat: bool = True):
        """Wait for the last queued step; returns (frames [n, 24] in global order, n)."""
        import torch

        if self.step_no == 0:
            raise RuntimeError("finish() before any step()")
        if self.exchange == "nccl":
            self.comm.synchronize()
        else:
            for st in self._lanes:
                st.synchronize()
        self.last_stream.synchronize()
        parity = (self.step_no - 1) & 1
        if self.exchange != "nccl":
            marks = self._slab_t[self.flags_off: self.flags_off + 2 * self.flags_bytes].view(torch.int64).view(2, self.world + 2)
            if int(marks[:, self.world + 1].max().item()) != 0:
                raise RuntimeError("a frame-exchange barrier timed out (a peer rank did not arrive within 4 s)")
        counts = self.hdr_host[parity].clone()
        m = int(counts.max())
        if m > self.cap:
            raise ValueError(f"a sub-shard produced {m} frames but the buffers hold {self.cap}")
        if self.exchange == "nccl":
            if m + 1 > self.slab_rows:
                # optimistic row count was too small (traffic got denser): repeat the exchange with room to spare
                self.slab_rows = min(self.rows, (int(m * 1.25) + 1024) // 1024 * 1024)
                self.step_no -= 1
                with torch.cuda.stream(self.last_stream):
                    self._redo_nccl(parity)
                self.step_no += 1
                self.comm.synchronize()
                counts = self.hdr_host[parity].clone()
            elif self.slab_rows == self.rows and m + 1 < self.rows:
                self.slab_rows = min(self.rows, (int(m * 1.15) + 1024) // 1024 * 1024)   # first step done: size for the traffic
        parts, total = [], 0
        for r in range(self.world):
            for k in range(len(self.ranges)):
                c = int(counts[k, r])
                parts.append(self._slab[parity, k, r, 1: 1 + c])
                total += c
        return (torch.cat(parts) if concat else parts), total

    def _redo_nccl(self, parity: int) -> None:
        import torch
        import torch.distributed as dist

        rows = self.slab_rows
        with torch.cuda.stream(self.comm):
            for k in range(len(self.ranges)):
                dst = self._slab[parity, k, :, :rows]
                if self.world > 1:
                    tmp = torch.empty((self.world, rows, RECORD_BYTES), dtype=torch.uint8, device=self.dev)
                    dist.all_gather_into_tensor(tmp.view(-1), self.out[k][:rows].reshape(-1), group=self.group)
                    dst.copy_(tmp)
                else:
                    dst[0].copy_(self.out[k][:rows])
            self.hdr_host[parity].copy_(self._slab[parity, :, :, 0, :8].contiguous().view(-1).view(torch.int64)
                                        .view(len(self.ranges), self.world), non_blocking=True)

    def launches_per_step(self) -> int:
        """Kernels of this library launched per step on this rank (decode, scan, gather per sub-shard + the barrier)."""
        return 3 * len(self.ranges) + (1 if self.exchange != "nccl" else 0)

    def close(self) -> None:
        for g in self._graphs:
            if g is not None:
                g.close()
        self._graphs = [None, None]
        if self.exchange != "nccl":
            for d in self._decs[1:]:
                d.close()
            self._decs = self._decs[:1]
            self.dec.set_timing(True)
