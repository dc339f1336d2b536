"""air_rs_b200 -- B200-native ADS-B / Mode S decode stage for jaxsonpd/air_rs.

One hot path, rebuilt for sm_100a behind a C ABI (include/airgpu.h): the decode
thread of the reference (src/adsb.rs:92-122).  This package holds
  csrc/      the CUDA kernels and the C ABI (libairgpu.so)
  native.py  ctypes binding of that ABI (fails loudly without the library / a GPU)
  decoder.py host-side mirror of the reference's decode-thread interface
  packet.py  AdsbPacket mirror (reference src/adsb/packet.rs, src/adsb/msgs.rs)
  ingest.py  .c16 loader / writer and the playback thread (reference src/utils.rs, src/adsb.rs:75-89)
  sharding.py  contiguous candidate shards + frame-list exchange for one process per GPU
  synth.py   integer-only synthetic capture generator (host) + device twin

There is no CPU fallback in this package and it never imports oracle/.
"""
from .native import FRAME_DTYPE, FMT_CS16, FMT_U8, AirgpuError  # noqa: F401

__version__ = "0.1.0"
