// dump1090_rs_b200/host/sdrconfig.hpp -- the reference binary's SDR configuration layer
// (dump1090_rs/src/sdrconfig.rs:5-41, dump1090_rs/src/main.rs:70-120): a list of per-driver settings read
// from TOML, an embedded default list, and an optional user file whose entries take precedence.
//
//   [[sdrs]]            driver = "name"   channel = 0 (default 0)
//   [[sdrs.gain]]       key = "..."  value = <float>          (at least one table or `gain = []`)
//   [[sdrs.setting]]    key = "..."  value = "..."            (optional)
//   [sdrs.antenna]      name = "..."                          (optional)
//
// Only the TOML subset this schema needs is read (tables, arrays of tables, string / number values,
// comments); anything else is an error, as a serde failure is in the reference.  SoapySDR itself is out of
// scope: the selected configuration is reported, the way main.rs:107-137 narrates what it applies.
#pragma once
#include <cstdlib>
#include <fstream>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace dump1090_rs {

struct Arg {
    std::string key, value;
};
struct Gain {
    std::string key;
    double value = 0;
};
struct Antenna {
    std::string name;
};
struct Sdr {
    std::size_t channel = 0;                    // sdrconfig.rs:11-12,20-24
    std::string driver;
    std::optional<std::vector<Arg>> setting;
    std::vector<Gain> gain;
    std::optional<Antenna> antenna;
};
struct SdrConfig {
    std::vector<Sdr> sdrs;
};

// the default device list compiled into the binary (the reference embeds its config.toml, sdrconfig.rs:3):
// tuner gains for the four front ends the reference ships defaults for
inline const char *default_config()
{
    return "[[sdrs]]\ndriver = \"rtlsdr\"\n[[sdrs.gain]]\nkey = \"TUNER\"\nvalue = 49.6\n"
           "[[sdrs]]\ndriver = \"hackrf\"\n[[sdrs.gain]]\nkey = \"LNA\"\nvalue = 40.0\n[[sdrs.gain]]\nkey = \"VGA\"\nvalue = 52.0\n"
           "[[sdrs]]\ndriver = \"bladerf\"\nchannel = 0\n[[sdrs.gain]]\nkey = \"full\"\nvalue = 35.0\n"
           "[[sdrs]]\ndriver = \"uhd\"\nchannel = 0\n[[sdrs.gain]]\nkey = \"PGA\"\nvalue = 70.0\n[sdrs.antenna]\nname = \"RX2\"\n";
}

namespace detail {
inline std::string trim(const std::string &s)
{
    const auto a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
inline std::string strip_comment(const std::string &line)
{
    bool in_str = false;
    for (std::size_t i = 0; i < line.size(); i++) {
        if (line[i] == '"' && (i == 0 || line[i - 1] != '\\'))
            in_str = !in_str;
        if (line[i] == '#' && !in_str)
            return line.substr(0, i);
    }
    return line;
}
inline std::string unquote(const std::string &v, int ln)
{
    if (v.size() < 2 || v.front() != '"' || v.back() != '"')
        throw std::runtime_error("config line " + std::to_string(ln) + ": expected a string");
    std::string out;
    for (std::size_t i = 1; i + 1 < v.size(); i++) {
        if (v[i] == '\\' && i + 2 < v.size()) {
            const char c = v[++i];
            out += c == 'n' ? '\n' : c == 't' ? '\t' : c;
        } else {
            out += v[i];
        }
    }
    return out;
}
inline double number(const std::string &v, int ln)
{
    char *end = nullptr;
    std::string t;
    for (char c : v)
        if (c != '_')
            t += c;
    const double d = std::strtod(t.c_str(), &end);
    if (t.empty() || *end)
        throw std::runtime_error("config line " + std::to_string(ln) + ": expected a number");
    return d;
}
}  // namespace detail

inline SdrConfig parse_sdr_config(const std::string &text)
{
    using namespace detail;
    SdrConfig cfg;
    enum { NONE, SDR, GAIN, SETTING, ANTENNA } where = NONE;
    std::vector<bool> has_gain_key;      // `gain` is required by the schema (sdrconfig.rs:15)
    std::istringstream in(text);
    std::string raw;
    int ln = 0;
    auto cur = [&]() -> Sdr & {
        if (cfg.sdrs.empty())
            throw std::runtime_error("config line " + std::to_string(ln) + ": no [[sdrs]] table yet");
        return cfg.sdrs.back();
    };
    while (std::getline(in, raw)) {
        ln++;
        const std::string line = trim(strip_comment(raw));
        if (line.empty())
            continue;
        if (line.front() == '[') {
            if (line == "[[sdrs]]") {
                cfg.sdrs.emplace_back();
                has_gain_key.push_back(false);
                where = SDR;
            } else if (line == "[[sdrs.gain]]") {
                cur().gain.emplace_back();
                has_gain_key.back() = true;
                where = GAIN;
            } else if (line == "[[sdrs.setting]]") {
                if (!cur().setting)
                    cur().setting.emplace();
                cur().setting->emplace_back();
                where = SETTING;
            } else if (line == "[sdrs.antenna]") {
                cur().antenna.emplace();
                where = ANTENNA;
            } else {
                throw std::runtime_error("config line " + std::to_string(ln) + ": unknown table " + line);
            }
            continue;
        }
        const auto eq = line.find('=');
        if (eq == std::string::npos)
            throw std::runtime_error("config line " + std::to_string(ln) + ": expected key = value");
        const std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
        switch (where) {
        case SDR:
            if (key == "driver") cur().driver = unquote(val, ln);
            else if (key == "channel") cur().channel = (std::size_t)number(val, ln);
            else if (key == "gain" && val == "[]") has_gain_key.back() = true;
            else throw std::runtime_error("config line " + std::to_string(ln) + ": unknown field " + key);
            break;
        case GAIN:
            if (key == "key") cur().gain.back().key = unquote(val, ln);
            else if (key == "value") cur().gain.back().value = number(val, ln);
            else throw std::runtime_error("config line " + std::to_string(ln) + ": unknown field " + key);
            break;
        case SETTING:
            if (key == "key") cur().setting->back().key = unquote(val, ln);
            else if (key == "value") cur().setting->back().value = unquote(val, ln);
            else throw std::runtime_error("config line " + std::to_string(ln) + ": unknown field " + key);
            break;
        case ANTENNA:
            if (key == "name") cur().antenna->name = unquote(val, ln);
            else throw std::runtime_error("config line " + std::to_string(ln) + ": unknown field " + key);
            break;
        default:
            throw std::runtime_error("config line " + std::to_string(ln) + ": value outside a table");
        }
    }
    for (std::size_t i = 0; i < cfg.sdrs.size(); i++) {
        if (cfg.sdrs[i].driver.empty())
            throw std::runtime_error("config: an [[sdrs]] entry has no driver");
        if (!has_gain_key[i])
            throw std::runtime_error("config: sdr \"" + cfg.sdrs[i].driver + "\" has no gain");
    }
    return cfg;
}

// main.rs:72-86: the embedded list, with the entries of the user's file pushed to the front one by one
// (so the last entry of the file ends up first) -- `find` then meets the user's entries first
inline SdrConfig layered_config(const std::optional<std::string> &custom_path, std::string *note = nullptr)
{
    SdrConfig cfg = parse_sdr_config(default_config());
    if (custom_path) {
        std::ifstream f(*custom_path);
        if (!f)
            throw std::runtime_error("cannot read " + *custom_path);
        std::stringstream ss;
        ss << f.rdbuf();
        SdrConfig custom = parse_sdr_config(ss.str());
        if (note)
            *note = "[-] read in custom config: " + *custom_path;
        for (auto &sdr : custom.sdrs)
            cfg.sdrs.insert(cfg.sdrs.begin(), std::move(sdr));
    }
    return cfg;
}

// main.rs:106: first entry whose driver equals --driver exactly
inline const Sdr *find_sdr(const SdrConfig &cfg, const std::string &driver)
{
    for (const auto &s : cfg.sdrs)
        if (s.driver == driver)
            return &s;
    return nullptr;
}

// main.rs:89-95: "driver=<name>" followed by the --driver-extra values, comma separated
inline std::string driver_args(const std::string &driver, const std::vector<std::string> &extra)
{
    std::string d = "driver=" + driver;
    for (const auto &e : extra)
        d += "," + e;
    return d;
}

// what main.rs:107-137 prints while applying an entry (the SoapySDR calls themselves are out of scope)
inline std::string describe(const Sdr &sdr)
{
    std::ostringstream o;
    o << "[-] using config: driver \"" << sdr.driver << "\" channel " << sdr.channel << "\n";
    for (const auto &g : sdr.gain)
        o << "[-] Writing gain: " << g.key << " = " << g.value << "\n";
    if (sdr.setting)
        for (const auto &s : *sdr.setting)
            o << "[-] Writing setting: " << s.key << " = " << s.value << "\n";
    if (sdr.antenna)
        o << "setting antenna: " << sdr.antenna->name << "\n";
    o << "[-] frequency: 1090000000\n[-] sample rate: 2400000\n";   // main.rs:133,136
    return o.str();
}

}  // namespace dump1090_rs
