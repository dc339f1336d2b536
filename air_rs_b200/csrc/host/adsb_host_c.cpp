// adsb_host_c.cpp -- extern "C" test shims over adsb_host.hpp / adsb_track.hpp so that the host-side
// C++ mirror (rows N2, N4 of SURVEY 8(f)) can be checked from pytest against the reference's unit
// tests (cpr.rs:152-188, aircraft.rs:170-262).  No CUDA in this file.
#include <cstring>

#include "adsb_track.hpp"

using namespace adsb_host;

extern "C" {

unsigned adsb_host_calc_num_zones(double lat) { return calc_num_zones(lat); }

void adsb_host_calculate_latitude(unsigned even_lat, unsigned odd_lat, int first_is_odd, double out[3])
{
    const Latitudes l = calculate_latitude(even_lat, odd_lat, first_is_odd ? CprFormat::Odd : CprFormat::Even);
    out[0] = l.latitude;
    out[1] = l.even_latitude;
    out[2] = l.odd_latitude;
}

double adsb_host_calculate_longitude(unsigned even_lon, unsigned odd_lon, double latitude, int first_is_odd)
{
    return calculate_longitude(even_lon, odd_lon, latitude, first_is_odd ? CprFormat::Odd : CprFormat::Even);
}

void *adsb_host_tracker_new(void) { return new AircraftMap(); }
void adsb_host_tracker_free(void *t) { delete static_cast<AircraftMap *>(t); }

// handle_aircraft_update + get_summary + serde_json::to_string for one 14-byte frame.
// Returns the JSON length (0 if it does not fit).
size_t adsb_host_tracker_update(void *t, const unsigned char packet[14], double now, char *json, size_t cap)
{
    AdsbPacket p(std::vector<uint8_t>(packet, packet + 14));
    const Aircraft a = handle_aircraft_update(p, *static_cast<AircraftMap *>(t), now);
    const std::string s = summary_json(a);
    if (s.size() + 1 > cap) return 0;
    std::memcpy(json, s.c_str(), s.size() + 1);
    return s.size();
}

}  // extern "C"
