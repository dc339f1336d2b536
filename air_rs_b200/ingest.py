"""Host-side mirror of what sits in FRONT of the decode thread in jaxsonpd/air_rs:

  save_data / load_data   SatDump-compatible `.c16` files, little-endian i16 I then Q   src/utils.rs:7-43
  playback_thread         20 000-sample chunks into the decode thread's channel        src/adsb.rs:75-89
  launch_adsb_playback    load -> playback thread -> decode thread (GPU) -> packets     src/adsb.rs:126-173 (playback arm)

Buffers are interleaved-IQ numpy int16 arrays: exactly the bytes `Vec<Complex<i16>>::as_ptr()` points at,
which is what `airgpu_submit` takes.  Pure host code: no CUDA here, and nothing from oracle/.
"""
from __future__ import annotations

import queue
import threading
import time
from typing import Iterator, List, Optional

import numpy as np

from .decoder import AdsbDecoder, close_channel, process_sdr_data_thread
from .native import FMT_CS16

PLAYBACK_CHUNK = 20_000            # samples per message, src/adsb.rs:78
PLAYBACK_SLEEP_S = 1e4 / 2e6       # src/adsb.rs:84: 5 ms per chunk, i.e. a 4 MS/s device


def save_data(data, name: str) -> None:
    """src/utils.rs:7-20: I then Q, little-endian i16, no header."""
    a = np.ascontiguousarray(data)
    if a.dtype != np.int16:
        raise TypeError(f"save_data wants interleaved int16 I,Q, got {a.dtype}")
    a = a.reshape(-1)
    if a.size % 2:
        raise ValueError("odd number of int16 values: not interleaved I,Q")
    a.astype("<i2", copy=False).tofile(name)


def load_data(filename: str) -> np.ndarray:
    """src/utils.rs:23-43.  Returns interleaved int16 (2 values per complex sample)."""
    raw = np.fromfile(filename, dtype=np.uint8)
    if raw.size % 4 != 0:
        raise ValueError("Invalid file length (not divisible by 4)")      # utils.rs:28-30
    return raw.view("<i2").astype(np.int16, copy=False)


def playback_chunks(data: np.ndarray, chunk_samples: int = PLAYBACK_CHUNK) -> Iterator[np.ndarray]:
    """The messages playback_thread sends (src/adsb.rs:76-79): `while i < data.len() - 20000`, so the final
    chunk is dropped even when it is complete.  A capture shorter than one chunk underflows `usize` upstream
    (a panic); here, as in csrc/host/adsb_host.hpp, it simply yields nothing."""
    a = np.ascontiguousarray(data, dtype=np.int16).reshape(-1)
    n = a.size // 2
    i = 0
    while n >= chunk_samples and i < n - chunk_samples:
        yield a[2 * i: 2 * (i + chunk_samples)].copy()                   # `.to_vec()`: an owned buffer per message
        i += chunk_samples


def playback_thread(tx: "queue.Queue", data: np.ndarray, chunk_samples: int = PLAYBACK_CHUNK,
                    realtime: bool = False) -> int:
    """src/adsb.rs:75-89.  `realtime=True` keeps the reference's 5 ms sleep per chunk; the default replays as
    fast as the decode stage accepts.  Closes the channel on return (`drop(tx)`, adsb.rs:88)."""
    sent = 0
    try:
        for buf in playback_chunks(data, chunk_samples):
            tx.put(buf)
            sent += 1
            if realtime:
                time.sleep(PLAYBACK_SLEEP_S)
    finally:
        close_channel(tx)
    return sent


def launch_adsb_playback(filename: str, decoder: Optional[AdsbDecoder] = None, realtime: bool = False) -> List:
    """The playback arm of launch_adsb (src/adsb.rs:126-147) without the display threads: returns the
    AdsbPackets the decode thread sent, in order."""
    data = load_data(filename)                                           # adsb.rs:137 ("Couldn't load playback data file")
    rx_raw: "queue.Queue" = queue.Queue()
    rx_pkt: "queue.Queue" = queue.Queue()
    prod = threading.Thread(target=playback_thread, args=(rx_raw, data, PLAYBACK_CHUNK, realtime))
    cons = threading.Thread(target=process_sdr_data_thread, args=(rx_raw, rx_pkt, decoder, FMT_CS16))
    prod.start()
    cons.start()
    prod.join()
    cons.join()
    out = []
    while True:
        p = rx_pkt.get()
        if not hasattr(p, "packet"):
            break
        out.append(p)
    return out
