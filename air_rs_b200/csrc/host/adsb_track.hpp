// adsb_track.hpp -- C++ mirror of the reference's post-decode host logic:
//   CPR global position          src/adsb/cpr.rs:21-147
//   Aircraft / handle_packet     src/adsb/aircraft.rs:27-165
//   AircraftSummary JSON         src/adsb/aircraft.rs:14-23,141-149 (serde camelCase), web.rs:117-126
// Host-side, sequential per aircraft, f64 -- deliberately NOT on the GPU (a few thousand frames/s).
// Time is passed in explicitly (seconds) instead of chrono::Local::now() so the logic is testable.
#pragma once

#include <charconv>
#include <cmath>
#include <cstdint>
#include <optional>
#include <string>
#include <unordered_map>

#include "adsb_host.hpp"

namespace adsb_host {

struct GeographicPosition {   // cpr.rs:10-16
    double latitude, longitude;
};

enum class CprFormat { Even, Odd };   // msgs.rs:46-51

constexpr double kNumZones = 15.0;    // cpr.rs:19

inline double convert_cpr_to_float(uint32_t cpr) { return double(cpr) / 131072.0; }   // cpr.rs:22-25

inline double normalize_longitude(double lon)   // cpr.rs:27-31
{
    while (lon < -180.0) lon += 360.0;
    while (lon > 180.0) lon -= 360.0;
    return lon;
}

inline uint32_t calc_num_zones(double lat)   // cpr.rs:39-55
{
    if (lat == 0.0) return 59;
    if (lat == 87.0 || lat == -87.0) return 2;
    if (lat < -87.0 || lat > 87.0) return 1;
    const double pi = 3.14159265358979323846;
    const double int1 = 1.0 - std::cos(pi / (2.0 * kNumZones));
    const double int2 = std::cos(pi / 180.0 * lat);
    const double int3 = (2.0 * pi) / std::acos(1.0 - (int1 / (int2 * int2)));
    return (uint32_t)std::floor(int3);
}

struct Latitudes {
    double latitude, even_latitude, odd_latitude;
};

inline Latitudes calculate_latitude(uint32_t even_cpr_lat, uint32_t odd_cpr_lat, CprFormat first)   // cpr.rs:64-90
{
    const double even_div = 360.0 / (4.0 * kNumZones);
    const double odd_div = 360.0 / (4.0 * kNumZones - 1.0);
    const double e = convert_cpr_to_float(even_cpr_lat), o = convert_cpr_to_float(odd_cpr_lat);
    const double j = std::floor(59.0 * e - 60.0 * o + 0.5);
    const double even_latitude = even_div * (std::fmod(j, 60.0) + e);   // Rust `%` on f64 is fmod
    const double odd_latitude = odd_div * (std::fmod(j, 59.0) + o);
    double latitude = first == CprFormat::Even ? odd_latitude : even_latitude;   // newest format wins
    if (latitude > 270.0) latitude -= 360.0;
    return {latitude, even_latitude, odd_latitude};
}

inline double calculate_longitude(uint32_t even_cpr_long, uint32_t odd_cpr_long, double latitude, CprFormat first)   // cpr.rs:92-126
{
    const double e = convert_cpr_to_float(even_cpr_long), o = convert_cpr_to_float(odd_cpr_long);
    const uint32_t nl = calc_num_zones(latitude);
    const double num_zones = first == CprFormat::Even ? double(std::max<uint32_t>(calc_num_zones(latitude - 1.0), 1))
                                                      : double(std::max<uint32_t>(calc_num_zones(latitude), 1));
    const double divisions = 360.0 / num_zones;
    // (nl - 1) is u32 arithmetic upstream; nl >= 1 always
    const double m = std::floor(e * double(nl - 1) - o * double(nl) + 0.5);
    const double lon = first == CprFormat::Even ? divisions * (std::fmod(m, num_zones) + o)
                                                : divisions * (std::fmod(m, num_zones) + e);
    return normalize_longitude(lon);
}

inline std::optional<GeographicPosition> calculate_geographic_position(uint32_t even_lat, uint32_t even_lon,
                                                                       uint32_t odd_lat, uint32_t odd_lon,
                                                                       CprFormat first)   // cpr.rs:135-147
{
    const Latitudes l = calculate_latitude(even_lat, odd_lat, first);
    if (calc_num_zones(l.even_latitude) != calc_num_zones(l.odd_latitude)) return std::nullopt;
    return GeographicPosition{l.latitude, calculate_longitude(even_lon, odd_lon, l.latitude, first)};
}

struct Aircraft {   // aircraft.rs:27-38
    uint32_t icao = 0;
    std::optional<std::string> callsign;
    int32_t altitude = 0;
    std::optional<GeographicPosition> geo_position;
    double last_contact = 0;
    bool has_odd = false, has_even = false;
    uint32_t odd_lat = 0, odd_lon = 0, even_lat = 0, even_lon = 0;
    double last_odd_processed = 0, last_even_processed = 0;

    Aircraft() = default;
    Aircraft(uint32_t i, double now) : icao(i), last_contact(now), last_odd_processed(now), last_even_processed(now) {}

    void handle_packet(const AdsbPacket &msg, double time_processed)   // aircraft.rs:48-111
    {
        if (msg.icao != icao) return;
        if (msg.kind == AdsbPacket::Kind::AircraftPosition) {
            altitude = msg.altitude;
            last_contact = time_processed;
            uint32_t o_lat, o_lon, e_lat, e_lon;
            CprFormat first;
            if (!msg.cpr_odd) {   // CprFormat::Even
                has_even = true;
                even_lat = msg.cpr_latitude;
                even_lon = msg.cpr_longitude;
                last_even_processed = time_processed;
                if (!has_odd) return;
                if (std::fabs(time_processed - last_odd_processed) > 10.0) return;
                o_lat = odd_lat; o_lon = odd_lon; e_lat = even_lat; e_lon = even_lon;
                first = CprFormat::Odd;
            } else {
                has_odd = true;
                odd_lat = msg.cpr_latitude;
                odd_lon = msg.cpr_longitude;
                last_odd_processed = time_processed;
                if (!has_even) return;
                if (std::fabs(time_processed - last_even_processed) > 10.0) return;
                o_lat = odd_lat; o_lon = odd_lon; e_lat = even_lat; e_lon = even_lon;
                first = CprFormat::Even;
            }
            if (auto g = calculate_geographic_position(e_lat, e_lon, o_lat, o_lon, first)) geo_position = g;
        } else if (msg.kind == AdsbPacket::Kind::AircraftID) {
            callsign = msg.callsign;
        }
    }

    std::string get_callsign() const { return callsign.value_or(""); }   // aircraft.rs:117-124
};

// serde_json prints f64 with the shortest round-trip representation and always keeps a ".0"
inline std::string json_f64(double v)
{
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    std::string s(buf, r.ptr);
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
    return s;
}

// AircraftSummary as the web thread broadcasts it (aircraft.rs:14-23 with rename_all = "camelCase";
// field order = declaration order; bindings/AircraftSummary.ts is the consumer's view).
inline std::string summary_json(const Aircraft &a)
{
    std::string s = "{\"icao\":" + std::to_string(a.icao) + ",\"callsign\":\"";
    for (char ch : a.get_callsign()) {
        if (ch == '"' || ch == '\\') s += '\\';
        s += ch;
    }
    s += "\",\"altitude\":" + std::to_string(a.altitude) + ",\"geoPosition\":";
    if (a.geo_position)
        s += "{\"latitude\":" + json_f64(a.geo_position->latitude) + ",\"longitude\":" + json_f64(a.geo_position->longitude) + "}";
    else
        s += "null";
    s += ",\"lastContact\":" + std::to_string((long long)std::floor(a.last_contact)) + "}";
    return s;
}

using AircraftMap = std::unordered_map<uint32_t, Aircraft>;

// aircraft.rs:158-165
inline Aircraft handle_aircraft_update(const AdsbPacket &packet, AircraftMap &aircrafts, double now)
{
    auto it = aircrafts.find(packet.icao);
    if (it == aircrafts.end()) it = aircrafts.emplace(packet.icao, Aircraft(packet.icao, now)).first;
    it->second.handle_packet(packet, now);
    return it->second;
}

}  // namespace adsb_host
