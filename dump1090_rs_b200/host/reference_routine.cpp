// The reference's test / bench routine (tests/test.rs:7-17, benches/demod_benchmark.rs:7-12)
// written against the C++ mirror: icao_flush, read_test_data, to_mag, demodulate2400.
//   g++ -std=c++17 -O2 -o reference_routine reference_routine.cpp -L.. -lb200adsb -Wl,-rpath,'$ORIGIN/..'
//   ./reference_routine test_1641427457780.iq
#include <chrono>
#include <cstdio>

#include "dump1090_rs.hpp"

using namespace dump1090_rs;

int main(int argc, char **argv)
{
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s capture.iq [iterations]\n", argv[0]);
        return 2;
    }
    const int iters = argc > 2 ? std::atoi(argv[2]) : 1;
    try {
        const auto buf = utils::read_test_data(argv[1]);
        std::vector<demod_2400::ModeSMessage> data;
        const auto t0 = std::chrono::steady_clock::now();
        for (int it = 0; it < iters; it++) {
            icao_filter::icao_flush();                       // tests/test.rs:9
            const auto outbuf = utils::to_mag(buf);          // :11
            data = demod_2400::demodulate2400(*outbuf);      // :13
        }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        for (const auto &m : data)
            std::printf("*%s;\n", m.hex().c_str());          // the AVR line main.rs:176 would send
        std::fprintf(stderr, "%zu frames, %.3f ms per routine\n", data.size(), ms / iters);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
