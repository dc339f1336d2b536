"""CPU: the host-side AdsbPacket mirror (air_rs_b200/packet.py) against the reference's own
unit tests for the boundary type (src/adsb/packet.rs:25-49, src/adsb/msgs.rs:225-321,
src/adsb/aircraft.rs:184-199)."""
from air_rs_b200.packet import AdsbPacket, AircraftID, AircraftPosition, UknownMsg


def test_aircraft_id_callsign_and_type():
    """msgs.rs:229-243 test_aircraft_id / test_aircraft_type"""
    me = bytes([0x20, 0x2C, 0xC3, 0x71, 0xC3, 0x2C, 0xE0])
    m = AircraftID.new(me)
    assert m.callsign == "KLM1023_" and m.msg_type == 4


def test_altitudes():
    """msgs.rs:245-275"""
    cases = [([0x58, 0xC3, 0x82, 0xD6, 0x90, 0xC8, 0xAC], 38000), ([0x58, 0xC2, 0x82, 0xD6, 0x90, 0xC8, 0xAC], 155000),
             ([0x58, 0x01, 0x02, 0xD6, 0x90, 0xC8, 0xAC], -1000), ([0x58, 0x01, 0x12, 0xD6, 0x90, 0xC8, 0xAC], -975)]
    for me, alt in cases:
        assert AircraftPosition.new(bytes(me)).altitude == alt


def test_position_flags_and_cpr():
    """msgs.rs:277-321"""
    even = AircraftPosition.new(bytes([0x58, 0xC3, 0x82, 0xD6, 0x90, 0xC8, 0xAC]))
    odd = AircraftPosition.new(bytes([0x58, 0xC3, 0x86, 0x43, 0x5C, 0xC4, 0x12]))
    for p in (even, odd):
        assert (p.msg_type, p.surveillance_status, p.nic_supplement, p.cpr_time) == (11, 0, 0, 0)
    assert not even.cpr_odd and odd.cpr_odd
    assert (even.cpr_latitude, even.cpr_longitude) == (93000, 51372)
    assert (odd.cpr_latitude, odd.cpr_longitude) == (74158, 50194)


def test_packet_new():
    """packet.rs:25-49 incl. the `& 5` capability mask; aircraft.rs:184-199"""
    p = AdsbPacket.from_hex("8d7c6b3020293532d70820fc8090")
    assert (p.downlink_format, p.capability, p.icao, p.msg_type) == (17, 0x8D & 5, 0x7C6B30, 4)
    assert isinstance(p.msg, AircraftID) and p.msg.callsign == "JST250__"
    p = AdsbPacket.from_hex("8d7c6b30581304f388bb4455896f")
    assert isinstance(p.msg, AircraftPosition) and p.msg.altitude == 2600
    p = AdsbPacket.from_hex("8D40621D58C382D690C8AC2863A7")
    assert p.icao == 0x40621D and p.msg.altitude == 38000 and not p.msg.cpr_odd
    p = AdsbPacket.from_hex("8D406B909915A678D4D220AA4BDA")        # type code 19: neither ID nor position
    assert isinstance(p.msg, UknownMsg) and p.msg.raw_msg == bytes.fromhex("9915A678D4D220AA4BDA")
    assert p.get_icao() == 0x406B90
