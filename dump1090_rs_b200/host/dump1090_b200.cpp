// dump1090_rs_b200/host/dump1090_b200.cpp -- the reference binary's receive loop
// (dump1090_rs/src/main.rs:149-213) over libb200adsb, with the SoapySDR front end (out of scope)
// replaced by a sample source that delivers the same thing stream.read does: up to `mtu`
// Complex<i16> samples per read, memory order (re, im) = SoapySDR CS16 [I, Q] (main.rs:143,161).
//
//   loop { accept one client;  read <= mtu samples;  to_mag + demodulate2400 (fused on the GPU);
//          "*{hex};\n" per frame -> stdout unless --quiet, and to every TCP client;
//          clients that reset their connection are dropped }
//
// The reference's demodulation blocks its next `stream.read` (main.rs:161-167).  Here the loop is double
// buffered: read k+1 fills one pinned buffer while the GPU demodulates read k out of the other
// (b200adsb_demod_iq_batch_submit / _wait); frames leave in the same order as from the serial loop
// (--serial keeps that one for comparison).
//
// Command line (main.rs:33-67): --host --port --driver --driver-extra (repeatable) --custom-config --quiet
// with the reference's defaults and its configuration layering (sdrconfig.hpp; main.rs:72-86,106): the
// entry selected by --driver is reported the way the reference narrates it, and a driver without an
// entry is an error as there.  --print-config stops after that (no GPU needed).
// Sources:  --file capture.iq   the reference's capture format (utils.rs:8-40: im, re pairs)
//           --raw path | -      raw interleaved CS16 (re, im), e.g. a pipe from an SDR tool
// --batch K collects K reads into one b200adsb_demod_iq_batch call (identical frames, in order).
// Exits 0 at end of input (the reference exits 1 on an SDR time-out, main.rs:203-209).
//   g++ -std=c++17 -O2 -o dump1090_b200 dump1090_b200.cpp -L.. -lb200adsb -Wl,-rpath,'$ORIGIN/..'
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "avr_server.hpp"
#include "dump1090_rs.hpp"
#include "sdrconfig.hpp"

using namespace dump1090_rs;

struct Options {
    std::string host = "127.0.0.1";     // main.rs:33-40 defaults
    int port = 30002;
    bool quiet = false;
    std::string driver = "rtlsdr";      // main.rs:49-55
    std::vector<std::string> driver_extra;
    std::optional<std::string> custom_config;
    bool print_config = false;
    std::string file, raw;
    std::size_t mtu = MODES_MAG_BUF_SAMPLES, batch = 1;
    int wait_clients = 0;               // test aid: do not start before this many clients are connected
    bool serial = false;                // one synchronous call per read, as the reference's loop
    std::size_t frame_cap = 1 << 14;    // frame slots per batch
};

static bool parse(int argc, char **argv, Options &o)
{
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&](const char *name) -> const char * {
            if (i + 1 >= argc) {
                std::fprintf(stderr, "%s needs a value\n", name);
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "--host") o.host = val("--host");
        else if (a == "--port") o.port = std::atoi(val("--port"));
        else if (a == "--quiet") o.quiet = true;
        else if (a == "--driver") o.driver = val("--driver");
        else if (a == "--driver-extra") o.driver_extra.push_back(val("--driver-extra"));
        else if (a == "--custom-config") o.custom_config = std::string(val("--custom-config"));
        else if (a == "--print-config") o.print_config = true;
        else if (a == "--file") o.file = val("--file");
        else if (a == "--raw") o.raw = val("--raw");
        else if (a == "--mtu") o.mtu = (std::size_t)std::atoll(val("--mtu"));
        else if (a == "--batch") o.batch = (std::size_t)std::atoll(val("--batch"));
        else if (a == "--wait-clients") o.wait_clients = std::atoi(val("--wait-clients"));
        else if (a == "--serial") o.serial = true;
        else if (a == "--frame-cap") o.frame_cap = (std::size_t)std::atoll(val("--frame-cap"));
        else return false;
    }
    return (o.print_config || !o.file.empty() || !o.raw.empty()) && o.mtu >= 1 && o.mtu <= MODES_MAG_BUF_SAMPLES &&
           o.batch >= 1;
}

int main(int argc, char **argv)
{
    Options opt;
    if (!parse(argc, argv, opt)) {
        std::fprintf(stderr,
                     "usage: %s (--file capture.iq | --raw path|-) [--host 127.0.0.1] [--port 30002] [--quiet]\n"
                     "          [--driver rtlsdr] [--driver-extra k=v]... [--custom-config file.toml] [--print-config]\n"
                     "          [--mtu samples<=131072] [--batch reads] [--wait-clients n] [--serial] [--frame-cap n]\n", argv[0]);
        return 2;
    }
    try {
        // configuration layering and driver selection (main.rs:72-120)
        std::string note;
        const SdrConfig config = layered_config(opt.custom_config, &note);
        if (!note.empty())
            std::printf("%s\n", note.c_str());
        std::printf("[-] using soapysdr driver_args: %s\n", driver_args(opt.driver, opt.driver_extra).c_str());
        const Sdr *sdr = find_sdr(config, opt.driver);
        if (!sdr) {
            std::fprintf(stderr, "[-] selected --driver gain values not found in custom or default config\n");
            return 1;
        }
        std::printf("%s", describe(*sdr).c_str());
        std::fflush(stdout);
        if (opt.print_config)
            return 0;
        // the sample source
        std::vector<Complex16> capture;
        std::size_t cap_pos = 0;
        std::FILE *rawf = nullptr;
        if (!opt.file.empty())
            capture = utils::read_test_data(opt.file);
        else
            rawf = opt.raw == "-" ? stdin : std::fopen(opt.raw.c_str(), "rb");
        if (opt.file.empty() && !rawf)
            throw std::runtime_error("cannot open " + opt.raw);
        auto read = [&](Complex16 *dst) -> std::size_t {      // stream.read(&mut [&mut buf], ..) -> len
            if (rawf)
                return std::fread(dst, sizeof(Complex16), opt.mtu, rawf);
            const std::size_t n = std::min(opt.mtu, capture.size() - cap_pos);
            std::memcpy(dst, capture.data() + cap_pos, n * sizeof(Complex16));
            cap_pos += n;
            return n;
        };

        AvrServer server;
        server.bind(opt.host, opt.port);
        std::fprintf(stderr, "[-] listening on %s:%d\n", opt.host.c_str(), server.port());
        while ((int)server.clients() < opt.wait_clients) {
            server.accept_one();
            ::usleep(1000);
        }

        Context &ctx = Context::global();
        // two pinned staging sets of `batch` reads each (the host entry points accept pageable memory too,
        // only slower and without overlap)
        struct Set {
            Complex16 *buf = nullptr;
            std::vector<std::uint32_t> lengths;
            b200adsb_frame *frames = nullptr;
            std::uint32_t *result = nullptr;
            std::size_t nreads = 0;
        } set[2];
        for (auto &st : set) {
            st.buf = static_cast<Complex16 *>(b200adsb_host_alloc(opt.batch * opt.mtu * sizeof(Complex16)));
            st.frames = static_cast<b200adsb_frame *>(b200adsb_host_alloc(opt.frame_cap * sizeof(b200adsb_frame)));
            st.result = static_cast<std::uint32_t *>(b200adsb_host_alloc(16));
            st.lengths.resize(opt.batch);
            if (!st.buf || !st.frames || !st.result)
                throw std::runtime_error("pinned allocation failed");
        }
        std::size_t total = 0;
        auto emit = [&](const b200adsb_frame *fr, std::size_t n) {
            if (n == 0)
                return;                                        // main.rs:170
            std::vector<std::string> lines;
            lines.reserve(n);
            for (std::size_t i = 0; i < n; i++) {              // main.rs:171-181
                char line[40];
                std::size_t len = 0;
                b200adsb_format_avr(&fr[i], 1, line, sizeof line, &len);
                lines.emplace_back(line, len);
                if (!opt.quiet)
                    std::fwrite(line, 1, len, stdout);
            }
            total += n;
            server.broadcast(lines);                           // main.rs:183-199
        };
        auto fill = [&](Set &st) -> bool {                     // `batch` reads; false at end of input
            st.nreads = 0;
            while (st.nreads < opt.batch) {
                const std::size_t len = read(st.buf + st.nreads * opt.mtu);
                if (len == 0)
                    return false;
                st.lengths[st.nreads++] = (std::uint32_t)len;
            }
            return true;
        };
        auto demod_sync = [&](Set &st) {                       // main.rs:166-167, one blocking call
            std::size_t n = 0;
            if (st.nreads == 1)
                ctx.check(b200adsb_demod_iq(ctx.raw(), reinterpret_cast<const std::int16_t *>(st.buf), st.lengths[0],
                                            st.frames, opt.frame_cap, &n), "demod_iq");
            else
                ctx.check(b200adsb_demod_iq_batch(ctx.raw(), reinterpret_cast<const std::int16_t *>(st.buf), st.nreads,
                                                  opt.mtu, opt.mtu, st.lengths.data(), st.frames, opt.frame_cap, &n,
                                                  nullptr), "demod_iq_batch");
            emit(st.frames, n);
        };
        if (opt.serial) {
            for (bool more = true; more;) {
                server.accept_one();                           // main.rs:154-157
                more = fill(set[0]);
                if (set[0].nreads)
                    demod_sync(set[0]);
            }
        } else {
            // the first call is synchronous (it sizes the candidate pool for this kind of traffic)
            int cur = 0;
            bool more = fill(set[cur]);
            if (set[cur].nreads)
                demod_sync(set[cur]);
            bool pending = false;                              // set[cur ^ 1] is on the GPU
            while (more || pending) {
                server.accept_one();
                Set &st = set[cur];
                st.nreads = 0;
                if (more)
                    more = fill(st);                           // overlaps the demodulation of the other set
                const bool submitted = st.nreads > 0;
                if (submitted)
                    ctx.check(b200adsb_demod_iq_batch_submit(ctx.raw(), cur, reinterpret_cast<const std::int16_t *>(st.buf),
                                                             st.nreads, opt.mtu, opt.mtu, st.lengths.data(), st.frames,
                                                             opt.frame_cap, st.result), "demod_iq_batch_submit");
                if (pending) {
                    Set &pv = set[cur ^ 1];
                    ctx.check(b200adsb_demod_iq_batch_wait(ctx.raw(), cur ^ 1), "demod_iq_batch_wait");
                    if (pv.result[1] != 0 || pv.result[3] != 0) {
                        // the queued batch failed (candidate pool too small for it, or more frames than slots):
                        // redo it with the blocking call, and the one queued behind it, in order
                        if (submitted)
                            ctx.check(b200adsb_demod_iq_batch_wait(ctx.raw(), cur), "demod_iq_batch_wait");
                        if (pv.result[3] != 0 && pv.result[1] == 0)
                            throw std::runtime_error("more frames in one batch than --frame-cap");
                        demod_sync(pv);
                        if (submitted)
                            demod_sync(st);
                        pending = false;
                        cur ^= 1;
                        continue;
                    }
                    emit(pv.frames, pv.result[0]);
                }
                pending = submitted;
                cur ^= 1;
            }
        }
        std::fflush(stdout);
        std::fprintf(stderr, "[-] end of input: %zu frames\n", total);
        for (auto &st : set) {
            b200adsb_host_free(st.buf);
            b200adsb_host_free(st.frames);
            b200adsb_host_free(st.result);
        }
        if (rawf && rawf != stdin)
            std::fclose(rawf);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "[!] %s\n", e.what());
        return 1;
    }
    return 0;
}
