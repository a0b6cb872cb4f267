// CPU-only check of AvrServer (no CUDA, no libb200adsb): two clients receive the AVR lines of two
// reads in order; a client that resets its connection is dropped, the other keeps receiving.
//   g++ -std=c++17 -O2 -o avr_server_test avr_server_test.cpp && ./avr_server_test
#include <cstdio>
#include <cstdlib>

#include "avr_server.hpp"

using dump1090_rs::AvrServer;

static int connect_to(int port)
{
    const int s = ::socket(AF_INET, SOCK_STREAM, 0);
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_port = htons((uint16_t)port);
    ::inet_pton(AF_INET, "127.0.0.1", &a.sin_addr);
    if (::connect(s, reinterpret_cast<sockaddr *>(&a), sizeof a) < 0) {
        std::perror("connect");
        std::exit(1);
    }
    return s;
}
static std::string read_n(int s, std::size_t n)
{
    std::string out;
    char buf[256];
    while (out.size() < n) {
        pollfd pf{s, POLLIN, 0};
        if (::poll(&pf, 1, 2000) <= 0)
            break;
        const ssize_t r = ::recv(s, buf, std::min(sizeof buf, n - out.size()), 0);
        if (r <= 0)
            break;
        out.append(buf, (std::size_t)r);
    }
    return out;
}
#define CHECK(c)                                                      \
    do {                                                              \
        if (!(c)) {                                                   \
            std::fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #c); \
            return 1;                                                 \
        }                                                             \
    } while (0)

int main()
{
    AvrServer srv;
    srv.bind("127.0.0.1", 0);
    CHECK(srv.port() > 0);
    CHECK(!srv.accept_one());                      // non-blocking: nobody is waiting
    const int c1 = connect_to(srv.port()), c2 = connect_to(srv.port());
    for (int k = 0; k < 100 && srv.clients() < 2; k++) {   // one accept per loop iteration
        srv.accept_one();
        ::usleep(1000);
    }
    CHECK(srv.clients() == 2);
    const std::vector<std::string> read1 = {"*8d4840d6202cc371c32ce0576098;\n", "*5d4840d6e6a4b1;\n"};
    CHECK(srv.broadcast(read1) == 0);
    const std::string want1 = read1[0] + read1[1];
    CHECK(read_n(c1, want1.size()) == want1);
    CHECK(read_n(c2, want1.size()) == want1);
    // client 1 resets its connection (SO_LINGER 0 -> RST)
    linger lg{1, 0};
    ::setsockopt(c1, SOL_SOCKET, SO_LINGER, &lg, sizeof lg);
    ::close(c1);
    ::usleep(20000);
    const std::vector<std::string> read2 = {"*02e19cb02512c3;\n"};
    std::size_t dropped = 0;
    for (int k = 0; k < 5 && dropped == 0; k++) {  // the reset surfaces on the first or second write
        dropped += srv.broadcast(read2);
        ::usleep(5000);
    }
    CHECK(dropped == 1);
    CHECK(srv.clients() == 1);
    const std::string got2 = read_n(c2, read2[0].size());
    CHECK(got2 == read2[0]);
    CHECK(srv.broadcast({}) == 0);
    ::close(c2);
    std::puts("avr_server_test ok");
    return 0;
}
