// dump1090_rs_b200/host/dump1090_rs.hpp -- C++ host-side mirror of the libdump1090_rs crate
// surface for the demodulation hot path, over the C ABI of libb200adsb.so.
//
// The reference's host language is Rust; this image has no rustc/cargo, so the host side
// above the C ABI is C++ (the reference is compiled code).  Names, argument meaning and
// error behaviour follow the crate so that code written against it reads the same:
//
//   libdump1090_rs::utils::to_mag            (src/utils.rs:43)      -> utils::to_mag
//   libdump1090_rs::utils::read_test_data    (src/utils.rs:23)      -> utils::read_test_data
//   libdump1090_rs::demod_2400::demodulate2400 (src/demod_2400.rs:115) -> demod_2400::demodulate2400
//   libdump1090_rs::demod_2400::ModeSMessage::buffer (:106)          -> ModeSMessage::buffer
//   libdump1090_rs::icao_filter::{icao_flush, icao_hash, icao_filter_add, icao_filter_test}
//   MagnitudeBuffer, MODES_MAG_BUF_SAMPLES, MODES_{LONG,SHORT}_MSG_BYTES (src/lib.rs:22-51)
//
// Where the reference panics (index out of bounds for > 131072 samples, lib.rs:48) these
// throw std::out_of_range; CUDA failures throw std::runtime_error.  There is no CPU path.
#pragma once
#include <array>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200adsb.h"

namespace dump1090_rs {

constexpr std::size_t MODES_MAG_BUF_SAMPLES = B200ADSB_MODES_MAG_BUF_SAMPLES;   // lib.rs:22
constexpr std::size_t TRAILING_SAMPLES = B200ADSB_TRAILING_SAMPLES;             // lib.rs:24
constexpr std::size_t MODES_LONG_MSG_BYTES = B200ADSB_MODES_LONG_MSG_BYTES;     // lib.rs:25
constexpr std::size_t MODES_SHORT_MSG_BYTES = B200ADSB_MODES_SHORT_MSG_BYTES;   // lib.rs:26

// num_complex::Complex<i16> is #[repr(C)] {re, im}
using Complex16 = std::complex<std::int16_t>;
static_assert(sizeof(Complex16) == 4, "Complex<i16> must be two packed int16");

// The reference keeps the ICAO filter in process-wide statics (icao_filter.rs:8-9).  Here
// the state lives on the GPU inside a context; `Context::global()` plays that role.
class Context {
public:
    explicit Context(int device = 0, void *stream = nullptr)
    {
        const int rc = b200adsb_ctx_create(&ctx_, device, stream);
        if (rc != B200ADSB_OK)
            throw std::runtime_error(std::string("b200adsb_ctx_create: ") + b200adsb_strerror(rc) +
                                     " (no CUDA device: there is no CPU fallback)");
    }
    ~Context() { b200adsb_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    b200adsb_ctx *raw() const { return ctx_; }
    static Context &global()
    {
        static Context c(0);
        return c;
    }
    void check(int rc, const char *what) const
    {
        if (rc != B200ADSB_OK)
            throw std::runtime_error(std::string(what) + ": " + b200adsb_strerror(rc) + "; " +
                                     b200adsb_last_error(ctx_));
    }

private:
    b200adsb_ctx *ctx_ = nullptr;
};

// src/lib.rs:29-51
struct MagnitudeBuffer {
    std::array<std::uint16_t, TRAILING_SAMPLES + MODES_MAG_BUF_SAMPLES> data{};
    std::size_t length = 0;
    std::size_t first_sample_timestamp_12mhz = 0;
    void push(std::uint16_t x)   // lib.rs:47-50
    {
        data.at(TRAILING_SAMPLES + length) = x;
        length += 1;
    }
};

namespace utils {

// utils::to_mag (src/utils.rs:43-58)
inline std::unique_ptr<MagnitudeBuffer> to_mag(const Complex16 *data, std::size_t n, Context &ctx = Context::global())
{
    if (n > MODES_MAG_BUF_SAMPLES)
        throw std::out_of_range("index out of bounds: more than 131072 samples (src/lib.rs:48)");
    auto out = std::make_unique<MagnitudeBuffer>();
    ctx.check(b200adsb_to_mag(ctx.raw(), reinterpret_cast<const std::int16_t *>(data), n, out->data.data(),
                              &out->length),
              "to_mag");
    return out;
}
inline std::unique_ptr<MagnitudeBuffer> to_mag(const std::vector<Complex16> &data, Context &ctx = Context::global())
{
    return to_mag(data.data(), data.size(), ctx);
}

// utils::read_test_data (src/utils.rs:23-40): im first, then re, little endian, 0x20000 samples
inline std::vector<Complex16> read_test_data(const std::string &filepath)
{
    std::vector<Complex16> buf(0x20000);
    std::FILE *f = std::fopen(filepath.c_str(), "rb");
    if (!f)
        throw std::runtime_error("read_test_data: cannot open " + filepath);
    std::vector<std::int16_t> raw(2 * 0x20000);
    const std::size_t got = std::fread(raw.data(), sizeof(std::int16_t), raw.size(), f);
    std::fclose(f);
    if (got != raw.size())
        throw std::runtime_error("read_test_data: short file (the reference unwraps the read error)");
    for (std::size_t i = 0; i < buf.size(); i++)
        buf[i] = Complex16(raw[2 * i + 1], raw[2 * i]);
    return buf;
}

// utils::save_test_data (src/utils.rs:8-21)
inline void save_test_data(const std::vector<Complex16> &data, const std::string &name)
{
    std::FILE *f = std::fopen(name.c_str(), "wb");
    if (!f)
        throw std::runtime_error("save_test_data: cannot create " + name);
    for (const auto &d : data) {
        const std::int16_t v[2] = {d.imag(), d.real()};
        std::fwrite(v, sizeof(std::int16_t), 2, f);
    }
    std::fclose(f);
}

}  // namespace utils

namespace demod_2400 {

enum class MsgLen { Short, Long };   // src/demod_2400.rs:87-91

// src/demod_2400.rs:93-112 (signal_level is not computed: it is private and unobservable)
class ModeSMessage {
public:
    explicit ModeSMessage(const b200adsb_frame &f) : f_(f) {}
    MsgLen msglen() const { return f_.len == MODES_LONG_MSG_BYTES ? MsgLen::Long : MsgLen::Short; }
    // ModeSMessage::buffer(): 7 or 14 bytes
    std::pair<const std::uint8_t *, std::size_t> buffer() const { return {f_.msg, f_.len}; }
    std::string hex() const
    {
        static const char *d = "0123456789abcdef";
        std::string s;
        for (unsigned i = 0; i < f_.len; i++) {
            s.push_back(d[f_.msg[i] >> 4]);
            s.push_back(d[f_.msg[i] & 15]);
        }
        return s;
    }
    int score() const { return f_.score; }
    unsigned phase() const { return f_.phase; }
    std::uint32_t j() const { return f_.j; }
    std::uint32_t buffer_index() const { return f_.buffer; }

private:
    b200adsb_frame f_;
};

inline std::vector<ModeSMessage> wrap(const std::vector<b200adsb_frame> &fr, std::size_t n)
{
    std::vector<ModeSMessage> out;
    out.reserve(n);
    for (std::size_t i = 0; i < n; i++)
        out.emplace_back(fr[i]);
    return out;
}

// demod_2400::demodulate2400 (src/demod_2400.rs:115-212).  The reference returns
// Result<Vec<_>, &'static str> that is always Ok (:211); failures here throw.
inline std::vector<ModeSMessage> demodulate2400(const MagnitudeBuffer &mag, Context &ctx = Context::global())
{
    std::vector<b200adsb_frame> fr(4096);
    std::size_t n = 0;
    int rc = b200adsb_demodulate2400(ctx.raw(), mag.data.data(), mag.length, fr.data(), fr.size(), &n);
    if (rc == B200ADSB_ERR_CAPACITY)
        throw std::runtime_error("demodulate2400: more than 4096 frames in one buffer");
    ctx.check(rc, "demodulate2400");
    return wrap(fr, n);
}

// to_mag + demodulate2400 fused, the pair main.rs:166-167 issues per SDR read
inline std::vector<ModeSMessage> demod_iq(const Complex16 *data, std::size_t n, Context &ctx = Context::global())
{
    if (n > MODES_MAG_BUF_SAMPLES)
        throw std::out_of_range("index out of bounds: more than 131072 samples (src/lib.rs:48)");
    std::vector<b200adsb_frame> fr(4096);
    std::size_t k = 0;
    ctx.check(b200adsb_demod_iq(ctx.raw(), reinterpret_cast<const std::int16_t *>(data), n, fr.data(), fr.size(), &k),
              "demod_iq");
    return wrap(fr, k);
}

// a run of n_buffers consecutive buffers of one stream in one call
inline std::vector<ModeSMessage> demod_iq_batch(const Complex16 *data, std::size_t n_buffers,
                                                std::size_t samples_per_buffer, Context &ctx = Context::global(),
                                                std::size_t cap = 1 << 16)
{
    std::vector<b200adsb_frame> fr(cap);
    std::size_t k = 0;
    ctx.check(b200adsb_demod_iq_batch(ctx.raw(), reinterpret_cast<const std::int16_t *>(data), n_buffers,
                                      samples_per_buffer, samples_per_buffer, nullptr, fr.data(), cap, &k, nullptr),
              "demod_iq_batch");
    return wrap(fr, k);
}

}  // namespace demod_2400

namespace icao_filter {

constexpr std::uint32_t ICAO_FILTER_ADSB_NT = B200ADSB_ICAO_FILTER_ADSB_NT;   // icao_filter.rs:6
inline void icao_flush(Context &ctx = Context::global()) { ctx.check(b200adsb_icao_flush(ctx.raw()), "icao_flush"); }
inline std::uint32_t icao_hash(std::uint32_t a32) { return b200adsb_icao_hash(a32); }
inline void icao_filter_add(std::uint32_t addr, Context &ctx = Context::global())
{
    ctx.check(b200adsb_icao_filter_add(ctx.raw(), addr), "icao_filter_add");
}
inline bool icao_filter_test(std::uint32_t addr, Context &ctx = Context::global())
{
    const int rc = b200adsb_icao_filter_test(ctx.raw(), addr);
    if (rc < 0)
        ctx.check(rc, "icao_filter_test");
    return rc != 0;
}

}  // namespace icao_filter

namespace crc {

// crc::modes_checksum (src/crc.rs:263-282)
inline std::uint32_t modes_checksum(const std::uint8_t *message, std::size_t bits, Context &ctx = Context::global())
{
    if (bits / 8 < 3)
        throw std::invalid_argument("assertion failed: n >= 3 (src/crc.rs:267)");
    std::uint8_t m[14] = {0};
    for (std::size_t i = 0; i < bits / 8 && i < 14; i++)
        m[i] = message[i];
    std::uint32_t out = 0;
    ctx.check(b200adsb_modes_checksum(ctx.raw(), m, 1, bits, &out), "modes_checksum");
    return out;
}

}  // namespace crc

}  // namespace dump1090_rs
