"""Host-side mirror of the reference's decode-thread interface, on top of the C ABI.

Reference interface being mirrored (jaxsonpd/air_rs):
    fn process_sdr_data_thread(rx: Receiver<Vec<Complex<i16>>>, tx: Sender<AdsbPacket>)
        src/adsb.rs:92-122 -- for every received buffer, emit every CRC-valid
        DF17 frame in ascending offset order; return when the input channel
        closes or the output channel's receiver has gone.

`AdsbDecoder` is the thin object wrapper over include/airgpu.h; the free function
`process_sdr_data_thread` has the reference's name, argument meaning and
termination behaviour, with queue.Queue standing in for std::sync::mpsc.
Nothing here computes on the CPU: without libairgpu.so and a B200 every call
raises.
"""
from __future__ import annotations

import ctypes as C
import queue
from typing import Optional

import numpy as np

from . import native
from .native import FMT_CS16, FMT_U8, FRAME_DTYPE, AirgpuError
from .packet import AdsbPacket

_CLOSED = object()  # what a producer puts on a queue to "drop the Sender"


def _as_iq(buf, fmt: int) -> np.ndarray:
    want = np.uint8 if fmt == FMT_U8 else np.int16
    a = np.asarray(buf)
    if a.dtype == np.complex64 or a.dtype == np.complex128:
        raise TypeError("pass interleaved integer IQ, not complex floats")
    if a.dtype != want:
        raise TypeError(f"decoder format needs {np.dtype(want).name} IQ, got {a.dtype}")
    a = np.ascontiguousarray(a).reshape(-1)
    if a.size % 2:
        raise ValueError("interleaved IQ needs an even number of values")
    return a


class AdsbDecoder:
    """One decode stage on one GPU (airgpu_ctx)."""

    def __init__(self, fmt: int = FMT_CS16, device: int = 0, ring_slots: int = 4,
                 max_buffer_samples: int = 262144, max_frames: int = 8192):
        self._lib = native.lib()
        self.fmt = fmt
        self.device = device
        self.max_frames = max_frames
        cfg = native.Config(C.sizeof(native.Config), device, fmt, ring_slots, max_buffer_samples, max_frames)
        h = C.c_void_p()
        native.check(self._lib.airgpu_create(C.byref(cfg), C.byref(h)))
        self._h = h

    # -- lifetime ---------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.airgpu_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- streaming (the loop body of process_sdr_data_thread) -------------
    def submit(self, buf, base_offset: int = 0) -> int:
        a = _as_iq(buf, self.fmt)
        t = C.c_uint64(0)
        native.check(self._lib.airgpu_submit(self._h, a.ctypes.data, a.size // 2, base_offset, C.byref(t)))
        return int(t.value)

    def collect(self, ticket: int) -> np.ndarray:
        """Frames of one submitted buffer.  Nothing is ever dropped: if the buffer yields more frames than
        max_frames (a constant buffer yields one at every offset) the ticket stays collectable and is
        collected again with room for all of them -- the reference sends every packet (adsb.rs:98-111)."""
        cap = self.max_frames
        while True:
            out = np.zeros(cap, dtype=FRAME_DTYPE)
            n = C.c_size_t(0)
            rc = self._lib.airgpu_collect(self._h, ticket, out.ctypes.data, out.size, C.byref(n))
            if rc == native.ERR_OVERFLOW and n.value > cap:
                cap = int(n.value)
                continue
            native.check(rc)
            return out[: n.value].copy()

    # -- one-shot, host memory ---------------------------------------------
    def decode(self, iq, segment_samples: int = 0, base_offset: int = 0, max_frames: Optional[int] = None,
               n_samples: Optional[int] = None) -> np.ndarray:
        """Decode a whole capture; grows the output buffer and retries on overflow."""
        if isinstance(iq, int):  # raw host pointer (e.g. from host_alloc)
            ptr, n = iq, int(n_samples)
        else:
            a = _as_iq(iq, self.fmt)
            ptr, n = a.ctypes.data, a.size // 2
        cap = max_frames or max(4096, n // 256)
        while True:
            out = np.zeros(cap, dtype=FRAME_DTYPE)
            got = C.c_size_t(0)
            rc = self._lib.airgpu_decode(self._h, ptr, n, segment_samples, base_offset, out.ctypes.data, cap,
                                         C.byref(got))
            if rc == native.ERR_OVERFLOW and max_frames is None:
                cap = int(got.value)
                continue
            native.check(rc)
            return out[: got.value].copy()

    # -- device memory -------------------------------------------------------
    def decode_device(self, d_iq: int, n_samples: int, d_out: int, cap: int, segment_samples: int = 0,
                      base_offset: int = 0, d_count: int = 0, stream: int = 0) -> None:
        """Asynchronous decode of device-resident IQ (raw pointers, see airgpu_decode_device)."""
        native.check(self._lib.airgpu_decode_device(self._h, d_iq, n_samples, segment_samples, base_offset, d_out,
                                                    cap, d_count or None, stream or None))

    def decode_device_peers(self, d_iq: int, n_samples: int, outs, counts, cap: int, segment_samples: int = 0,
                            base_offset: int = 0, stream: int = 0, multicast: bool = False) -> None:
        """Asynchronous decode whose ordering kernels store every record and the frame count straight to each of
        `outs` / `counts` (device-accessible addresses: local, peer-mapped, or ONE multicast address)."""
        pe = native.Peers()
        pe.struct_size = C.sizeof(native.Peers)
        pe.n_outs = len(outs)
        pe.multicast = 1 if multicast else 0
        for j, (o, k) in enumerate(zip(outs, counts)):
            pe.outs[j] = o
            pe.counts[j] = k or None
        native.check(self._lib.airgpu_decode_device_peers(self._h, d_iq, n_samples, segment_samples, base_offset,
                                                          C.byref(pe), cap, stream or None))

    def peer_barrier(self, flag_ptrs, rank: int, epoch: int, stream: int = 0) -> None:
        arr = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        native.check(self._lib.airgpu_peer_barrier(self._h, arr, len(flag_ptrs), rank, epoch, stream or None))

    def reserve(self, n_samples: int, segment_samples: int = 0, cap: int = 0) -> None:
        native.check(self._lib.airgpu_reserve(self._h, n_samples, segment_samples, cap))

    def set_timing(self, enabled: bool) -> None:
        native.check(self._lib.airgpu_set_timing(self._h, 1 if enabled else 0))

    # -- CUDA graphs -----------------------------------------------------------
    def graph_begin(self, stream: int) -> None:
        native.check(self._lib.airgpu_graph_begin(self._h, stream))

    def set_capturing(self, on: bool) -> None:
        native.check(self._lib.airgpu_set_capturing(self._h, 1 if on else 0))

    def graph_end(self, stream: int) -> "Graph":
        g = C.c_void_p()
        native.check(self._lib.airgpu_graph_end(self._h, stream, C.byref(g)))
        return Graph(g)

    def sync_count(self) -> int:
        n = C.c_uint64(0)
        native.check(self._lib.airgpu_sync_count(self._h, C.byref(n)))
        return int(n.value)

    def decode_tensor(self, iq, segment_samples: int = 0, base_offset: int = 0, cap: Optional[int] = None,
                      out=None):
        """Decode a CUDA torch tensor of interleaved IQ; returns (frames[cap, 24] uint8 tensor, count)."""
        import torch

        want = torch.uint8 if self.fmt == FMT_U8 else torch.int16
        if not iq.is_cuda or iq.dtype != want or not iq.is_contiguous():
            raise TypeError(f"need a contiguous CUDA tensor of {want}")
        n = iq.numel() // 2
        cap = cap or max(4096, n // 256)
        if out is None:
            out = torch.empty((cap, FRAME_DTYPE.itemsize), dtype=torch.uint8, device=iq.device)
        cur = torch.cuda.current_stream(iq.device)
        stream = cur.cuda_stream
        if stream == 0:
            # the legacy default stream is not ordered against the context's own (non-blocking)
            # compute stream, which NULL selects: make the producer of `iq` finish first
            cur.synchronize()
        self.decode_device(iq.data_ptr(), n, out.data_ptr(), cap, segment_samples, base_offset, 0, stream)
        return out, self.sync_count()

    @staticmethod
    def frames_from_tensor(out, count: int) -> np.ndarray:
        n = min(count, out.shape[0])
        return out[:n].cpu().numpy().view(FRAME_DTYPE).reshape(-1).copy()

    # -- N1: frame fields ------------------------------------------------------
    def decode_fields(self, frames: np.ndarray) -> np.ndarray:
        """Per-frame derived fields (AdsbPacket::new on the device); host arrays in and out."""
        fr = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        out = np.zeros(fr.size, dtype=native.FIELDS_DTYPE)
        native.check(self._lib.airgpu_decode_fields_host(self._h, fr.ctypes.data, fr.size, out.ctypes.data))
        return out

    def decode_fields_device(self, d_frames: int, n_frames: int, d_out: int, stream: int = 0) -> None:
        native.check(self._lib.airgpu_decode_fields(self._h, d_frames, n_frames, d_out, stream or None))

    def stats(self) -> dict:
        st = native.Stats()
        native.check(self._lib.airgpu_get_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in native.Stats._fields_}

    # -- diagnostics -----------------------------------------------------------
    def levels_u8(self) -> np.ndarray:
        out = np.zeros(65536, dtype=np.uint16)
        native.check(self._lib.airgpu_dbg_levels_u8(self._h, out.ctypes.data))
        return out

    def levels_cs16(self, iq) -> np.ndarray:
        a = _as_iq(iq, FMT_CS16)
        out = np.zeros(a.size // 2, dtype=np.uint16)
        native.check(self._lib.airgpu_dbg_levels_cs16(self._h, a.ctypes.data, out.size, out.ctypes.data))
        return out


class Graph:
    """airgpu_graph: a recorded sequence of device-side calls, replayed with one launch."""

    def __init__(self, handle):
        self._h = handle
        self._lib = native.lib()

    def launch(self, stream: int) -> None:
        native.check(self._lib.airgpu_graph_launch(self._h, stream))

    def close(self) -> None:
        if self._h:
            self._lib.airgpu_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DecoderGroup:
    """airgpu_group: one host thread, several GPUs (SURVEY 8(b) `airgpu_decode_sharded`)."""

    def __init__(self, devices, fmt: int = FMT_CS16):
        self._lib = native.lib()
        self.fmt = fmt
        self.devices = list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        native.check(self._lib.airgpu_group_create(arr, len(self.devices), fmt, C.byref(h)))
        self._h = h

    def decode(self, iq, base_offset: int = 0, max_frames: Optional[int] = None, n_samples: Optional[int] = None) -> np.ndarray:
        if isinstance(iq, int):
            ptr, n = iq, int(n_samples)
        else:
            a = _as_iq(iq, self.fmt)
            ptr, n = a.ctypes.data, a.size // 2
        cap = max_frames or max(4096, n // 256)
        while True:
            out = np.zeros(cap, dtype=FRAME_DTYPE)
            got = C.c_size_t(0)
            rc = self._lib.airgpu_group_decode(self._h, ptr, n, base_offset, out.ctypes.data, cap, C.byref(got))
            if rc == native.ERR_OVERFLOW and max_frames is None:
                cap = int(got.value)
                continue
            native.check(rc)
            return out[: got.value].copy()

    def stats(self):
        arr = (native.Stats * len(self.devices))()
        native.check(self._lib.airgpu_group_stats(self._h, arr, len(self.devices)))
        return [{k: getattr(st, k) for k, _ in native.Stats._fields_} for st in arr]

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.airgpu_group_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def host_alloc(nbytes: int) -> int:
    """Page-locked host buffer (airgpu_host_alloc); returns the address."""
    p = C.c_void_p()
    native.check(native.lib().airgpu_host_alloc(nbytes, C.byref(p)))
    return int(p.value)


def host_free(ptr: int) -> None:
    native.check(native.lib().airgpu_host_free(ptr))


def close_channel(q: "queue.Queue") -> None:
    """drop(tx): tell the consumer no more items will come."""
    q.put(_CLOSED)


def process_sdr_data_thread(rx: "queue.Queue", tx: "queue.Queue", decoder: Optional[AdsbDecoder] = None,
                            fmt: int = FMT_CS16, depth: int = 2) -> int:
    """Drop-in for the reference's decode thread (src/adsb.rs:92-122).

    rx yields interleaved-IQ numpy buffers (one per `Vec<Complex<i16>>` message)
    until `close_channel(rx)`; every decoded frame is pushed to tx as an
    `AdsbPacket` in the reference's order (ascending offset within a buffer,
    buffers in arrival order); tx is closed on return (src/adsb.rs:121).  Up to
    `depth` buffers are in flight on the GPU while the next one is received.
    Returns the number of packets sent.
    """
    own = decoder is None
    dec = decoder or AdsbDecoder(fmt=fmt)
    sent = 0
    pending = []
    try:
        def drain(limit):
            nonlocal sent
            while len(pending) > limit:
                for rec in dec.collect(pending.pop(0)):
                    tx.put(AdsbPacket(bytes(rec["bytes"])))   # src/adsb.rs:107-108
                    sent += 1

        while True:
            buf = rx.get()                                     # src/adsb.rs:95
            if buf is _CLOSED:
                break
            pending.append(dec.submit(buf))
            drain(depth - 1)
        drain(0)
    finally:
        close_channel(tx)                                      # src/adsb.rs:121
        if own:
            dec.close()
    return sent
