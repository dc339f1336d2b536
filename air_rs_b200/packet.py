"""AdsbPacket: the boundary type the decode thread hands downstream.

Mirrors the reference's host-side container so code written against it reads the
same (jaxsonpd/air_rs src/adsb/packet.rs:10-49 and src/adsb/msgs.rs).  The GPU
stage only produces the 14 bytes; these derived fields are computed on the host
exactly as `AdsbPacket::new` does, quirks included (`capability = b0 & 5`).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Union

# msgs.rs:164-169
CHAR_CONVERT = (
    "#ABCDEFGHIJKLMNOPQRSTUVWXYZ#####_###############0123456789######"
)
assert len(CHAR_CONVERT) == 64


@dataclass
class UknownMsg:  # (sic) msgs.rs:32-35
    raw_msg: bytes


@dataclass
class AircraftID:  # msgs.rs:171-201
    msg_type: int
    callsign: str

    @staticmethod
    def msg_id_match(tc: int) -> bool:  # msgs.rs:208-213
        return 1 <= tc <= 4

    @classmethod
    def new(cls, me: bytes) -> "AircraftID":
        acc = int.from_bytes(me[1:7], "big")  # 48 bits -> eight 6-bit characters (msgs.rs:141-162)
        chars = [(acc >> (42 - 6 * k)) & 0x3F for k in range(8)]
        return cls(msg_type=me[0] >> 3, callsign="".join(CHAR_CONVERT[c] for c in chars))


@dataclass
class AircraftPosition:  # msgs.rs:54-102
    msg_type: int
    surveillance_status: int
    nic_supplement: int
    altitude: int
    cpr_time: int
    cpr_odd: bool
    cpr_latitude: int
    cpr_longitude: int

    @staticmethod
    def msg_id_match(tc: int) -> bool:  # msgs.rs:121-125
        return 9 <= tc <= 18

    @classmethod
    def new(cls, me: bytes) -> "AircraftPosition":
        alt_mode_25 = (me[1] & 1) == 1
        altitude = (((me[1] & 0xFE) >> 1) << 4) | ((me[2] & 0xF0) >> 4)
        altitude = altitude * (25 if alt_mode_25 else 100) - 1000
        lat = ((me[2] & 0b11) << 15) | (me[3] << 7) | ((me[4] & 0xFE) >> 1)
        lon = ((me[4] & 1) << 16) | (me[5] << 8) | me[6]
        return cls(
            msg_type=me[0] >> 3,
            surveillance_status=(me[0] & 0b110) >> 1,
            nic_supplement=me[0] & 1,
            altitude=altitude,
            cpr_time=(me[2] & 0b1000) >> 3,
            cpr_odd=bool((me[2] & 0b100) >> 2),
            cpr_latitude=lat,
            cpr_longitude=lon,
        )


@dataclass
class AdsbPacket:  # packet.rs:10-18
    packet: bytes
    downlink_format: int = field(init=False)
    capability: int = field(init=False)
    icao: int = field(init=False)
    msg_type: int = field(init=False)
    msg: Union[AircraftID, AircraftPosition, UknownMsg] = field(init=False)
    time_processed: float = field(init=False)

    def __post_init__(self):  # packet.rs:25-49
        p = bytes(self.packet)
        self.packet = p
        self.downlink_format = p[0] >> 3
        self.capability = p[0] & 5  # sic: the reference masks with 5, not 7
        self.icao = (p[1] << 16) | (p[2] << 8) | p[3]
        self.msg_type = p[4] >> 3
        me = p[4:11]
        if AircraftID.msg_id_match(self.msg_type):
            self.msg = AircraftID.new(me)
        elif AircraftPosition.msg_id_match(self.msg_type):
            self.msg = AircraftPosition.new(me)
        else:
            self.msg = UknownMsg(raw_msg=p[4:])
        self.time_processed = time.time()

    @classmethod
    def from_hex(cls, s: str) -> "AdsbPacket":  # packet.rs:57-70 (_new_from_string)
        return cls(bytes.fromhex(s))

    def get_icao(self) -> int:
        return self.icao
