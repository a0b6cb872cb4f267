// dump1090_rs_b200/host/avr_server.hpp -- the TCP side of the reference binary
// (dump1090_rs/src/main.rs:149-200): a non-blocking listener (default 127.0.0.1:30002) whose
// clients receive one AVR line "*{hex};\n" per decoded frame.  Behaviour mirrored from main.rs:
//   * one accept attempt per loop iteration (:154-157), never blocking the receive loop;
//   * every line of a read is written completely to every client (write_all, :183);
//   * a client whose write fails with ConnectionReset is dropped, other errors are ignored
//     (:183-195).  EPIPE is treated like ConnectionReset: it is what a reset peer reports on the
//     following write on Linux (and SIGPIPE is suppressed with MSG_NOSIGNAL).
// Plain POSIX sockets, no CUDA: the formatting is b200adsb_format_avr's (main.rs:174-176).
#pragma once
#include <arpa/inet.h>
#include <cerrno>
#include <cstring>
#include <fcntl.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <stdexcept>
#include <string>
#include <sys/socket.h>
#include <unistd.h>
#include <vector>

namespace dump1090_rs {

class AvrServer {
public:
    AvrServer() = default;
    AvrServer(const AvrServer &) = delete;
    AvrServer &operator=(const AvrServer &) = delete;
    ~AvrServer()
    {
        for (int s : sockets_)
            ::close(s);
        if (listener_ >= 0)
            ::close(listener_);
    }

    // TcpListener::bind((host, port)) + set_nonblocking(true) (main.rs:149-150); port 0 picks a free one
    void bind(const std::string &host = "127.0.0.1", int port = 30002)
    {
        listener_ = ::socket(AF_INET, SOCK_STREAM, 0);
        if (listener_ < 0)
            throw std::runtime_error(std::string("socket: ") + std::strerror(errno));
        int one = 1;
        ::setsockopt(listener_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
        sockaddr_in a{};
        a.sin_family = AF_INET;
        a.sin_port = htons((uint16_t)port);
        if (::inet_pton(AF_INET, host.c_str(), &a.sin_addr) != 1)
            throw std::runtime_error("bad listen address " + host);
        if (::bind(listener_, reinterpret_cast<sockaddr *>(&a), sizeof a) < 0 || ::listen(listener_, 16) < 0)
            throw std::runtime_error(std::string("bind/listen: ") + std::strerror(errno));
        ::fcntl(listener_, F_SETFL, ::fcntl(listener_, F_GETFL, 0) | O_NONBLOCK);
        socklen_t len = sizeof a;
        ::getsockname(listener_, reinterpret_cast<sockaddr *>(&a), &len);
        port_ = ntohs(a.sin_port);
    }
    int port() const { return port_; }
    std::size_t clients() const { return sockets_.size(); }

    // "add more clients": at most one per call, like the reference loop (main.rs:154-157)
    bool accept_one()
    {
        const int s = ::accept(listener_, nullptr, nullptr);
        if (s < 0)
            return false;
        int one = 1;
        ::setsockopt(s, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
        sockets_.push_back(s);
        return true;
    }

    // the lines of one read to every client (main.rs:180-195); returns the number of clients dropped
    std::size_t broadcast(const std::vector<std::string> &lines)
    {
        std::vector<std::size_t> remove;
        for (std::size_t i = 0; i < sockets_.size(); i++)
            for (const auto &msg : lines)
                if (!write_all(sockets_[i], msg.data(), msg.size())) {
                    remove.push_back(i);
                    break;
                }
        for (std::size_t k = remove.size(); k-- > 0;) {
            ::close(sockets_[remove[k]]);
            sockets_.erase(sockets_.begin() + (std::ptrdiff_t)remove[k]);
        }
        return remove.size();
    }

private:
    // false only for a reset connection; other errors end the write silently (as the reference ignores them)
    static bool write_all(int s, const char *p, std::size_t n)
    {
        while (n) {
            const ssize_t w = ::send(s, p, n, MSG_NOSIGNAL);
            if (w < 0) {
                if (errno == EINTR)
                    continue;
                if (errno == EAGAIN || errno == EWOULDBLOCK) {
                    pollfd pf{s, POLLOUT, 0};
                    ::poll(&pf, 1, 1000);
                    continue;
                }
                return !(errno == ECONNRESET || errno == EPIPE);
            }
            p += w;
            n -= (std::size_t)w;
        }
        return true;
    }

    int listener_ = -1, port_ = 0;
    std::vector<int> sockets_;
};

}  // namespace dump1090_rs
