#include "net.h"

#include <stddef.h>
#include <sys/un.h>

#include <atomic>

#include <algorithm>
#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/sendfile.h>
#include <sys/socket.h>
#include <unistd.h>

namespace cv {

static std::atomic<int> g_sock_buf{0};
void set_socket_buffer_bytes(int bytes) { g_sock_buf.store(bytes); }

void set_sock_opts(int fd) {
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));  // fails harmlessly on a unix socket
    setsockopt(fd, SOL_SOCKET, SO_KEEPALIVE, &one, sizeof(one));
    // An explicit SO_RCVBUF/SO_SNDBUF is clamped to net.core.{r,w}mem_max (often 208 KiB) AND switches the kernel's buffer
    // autotuning off (tcp_rmem goes to 6 MiB on its own): only set when asked to.
    if (int buf = g_sock_buf.load()) {
        setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &buf, sizeof(buf));
        setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &buf, sizeof(buf));
    }
}

std::string local_socket_name(int tcp_port) { return str_printf("curvine-b200-worker-%d", tcp_port); }

static socklen_t abstract_addr(const std::string& name, sockaddr_un* sa) {
    memset(sa, 0, sizeof(*sa));
    sa->sun_family = AF_UNIX;
    const size_t n = std::min(name.size(), sizeof(sa->sun_path) - 2);
    memcpy(sa->sun_path + 1, name.data(), n);  // sun_path[0] == 0: abstract namespace
    return static_cast<socklen_t>(offsetof(sockaddr_un, sun_path) + 1 + n);
}

Err unix_listen(const std::string& name, int* fd_out) {
    sockaddr_un sa;
    const socklen_t len = abstract_addr(name, &sa);
    const int fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return Err::io(str_printf("socket(AF_UNIX): %s", strerror(errno)));
    if (bind(fd, reinterpret_cast<sockaddr*>(&sa), len) != 0 || listen(fd, 1024) != 0) {
        const int e = errno;
        ::close(fd);
        return Err::io(str_printf("bind/listen @%s: %s", name.c_str(), strerror(e)));
    }
    *fd_out = fd;
    return Err::ok();
}

Err unix_connect(const std::string& name, int* fd_out, int64_t io_timeout_ms) {
    sockaddr_un sa;
    const socklen_t len = abstract_addr(name, &sa);
    const int fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return Err::io(str_printf("socket(AF_UNIX): %s", strerror(errno)));
    if (connect(fd, reinterpret_cast<sockaddr*>(&sa), len) != 0) {
        const int e = errno;
        ::close(fd);
        return Err::io(str_printf("connect @%s: %s", name.c_str(), strerror(e)));
    }
    set_sock_opts(fd);
    if (io_timeout_ms > 0) {
        timeval tv;
        tv.tv_sec = static_cast<time_t>(io_timeout_ms / 1000), tv.tv_usec = static_cast<suseconds_t>((io_timeout_ms % 1000) * 1000);
        setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
        setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
    }
    *fd_out = fd;
    return Err::ok();
}

void close_fd(int fd) {
    if (fd >= 0) ::close(fd);
}

static Err resolve(const std::string& host, int port, sockaddr_in* sa) {
    memset(sa, 0, sizeof(*sa));
    sa->sin_family = AF_INET;
    sa->sin_port = htons(static_cast<uint16_t>(port));
    const std::string h = (host.empty() || host == "localhost") ? "127.0.0.1" : host;
    if (inet_pton(AF_INET, h.c_str(), &sa->sin_addr) == 1) return Err::ok();
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    if (getaddrinfo(h.c_str(), nullptr, &hints, &res) != 0 || !res) {
        // a same-host worker whose hostname does not resolve (containers): fall back to loopback
        inet_pton(AF_INET, "127.0.0.1", &sa->sin_addr);
        return Err::ok();
    }
    sa->sin_addr = reinterpret_cast<sockaddr_in*>(res->ai_addr)->sin_addr;
    freeaddrinfo(res);
    return Err::ok();
}

Err tcp_connect(const std::string& host, int port, int* fd_out, int64_t conn_timeout_ms, int64_t io_timeout_ms) {
    sockaddr_in sa;
    CV_RETURN_IF_ERR(resolve(host, port, &sa));
    const int fd = socket(AF_INET, SOCK_STREAM, 0);
    if (fd < 0) return Err::io(str_printf("socket: %s", strerror(errno)));
    set_sock_opts(fd);
    int rc;
    if (conn_timeout_ms > 0) {  // non-blocking connect + poll, then back to blocking
        const int fl = fcntl(fd, F_GETFL, 0);
        fcntl(fd, F_SETFL, fl | O_NONBLOCK);
        rc = connect(fd, reinterpret_cast<sockaddr*>(&sa), sizeof(sa));
        if (rc != 0 && errno == EINPROGRESS) {
            pollfd pf{fd, POLLOUT, 0};
            int pr;
            do pr = poll(&pf, 1, static_cast<int>(std::min<int64_t>(conn_timeout_ms, 0x7fffffff)));
            while (pr < 0 && errno == EINTR);
            if (pr == 0) {
                ::close(fd);
                return Err::io(str_printf("connect %s:%d: timed out after %lld ms", host.c_str(), port, (long long)conn_timeout_ms));
            }
            int soerr = 0;
            socklen_t sl = sizeof(soerr);
            getsockopt(fd, SOL_SOCKET, SO_ERROR, &soerr, &sl);
            rc = soerr == 0 ? 0 : -1;
            errno = soerr;
        }
        const int e = errno;
        fcntl(fd, F_SETFL, fl);
        errno = e;
    } else {
        rc = connect(fd, reinterpret_cast<sockaddr*>(&sa), sizeof(sa));
    }
    if (rc != 0) {
        const int e = errno;
        ::close(fd);
        return Err::io(str_printf("connect %s:%d: %s", host.c_str(), port, strerror(e)));
    }
    if (io_timeout_ms > 0) {
        timeval tv;
        tv.tv_sec = static_cast<time_t>(io_timeout_ms / 1000), tv.tv_usec = static_cast<suseconds_t>((io_timeout_ms % 1000) * 1000);
        setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
        setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
    }
    *fd_out = fd;
    return Err::ok();
}

Err tcp_listen(const std::string& host, int port, int* fd_out, int* bound_port) {
    sockaddr_in sa;
    CV_RETURN_IF_ERR(resolve(host.empty() ? "0.0.0.0" : host, port, &sa));
    if (host.empty()) sa.sin_addr.s_addr = htonl(INADDR_ANY);
    const int fd = socket(AF_INET, SOCK_STREAM, 0);
    if (fd < 0) return Err::io(str_printf("socket: %s", strerror(errno)));
    int one = 1;
    setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    if (bind(fd, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) != 0 || listen(fd, 1024) != 0) {
        const int e = errno;
        ::close(fd);
        return Err::io(str_printf("bind/listen :%d: %s", port, strerror(e)));
    }
    socklen_t sl = sizeof(sa);
    getsockname(fd, reinterpret_cast<sockaddr*>(&sa), &sl);
    *bound_port = ntohs(sa.sin_port);
    *fd_out = fd;
    return Err::ok();
}

static Err send_flags(int fd, const void* buf, size_t n, int flags);
Err send_more(int fd, const void* buf, size_t n) { return send_flags(fd, buf, n, MSG_NOSIGNAL | MSG_MORE); }
Err send_all(int fd, const void* buf, size_t n) { return send_flags(fd, buf, n, MSG_NOSIGNAL); }

static Err send_flags(int fd, const void* buf, size_t n, int flags) {
    const uint8_t* p = static_cast<const uint8_t*>(buf);
    while (n) {
        const ssize_t w = ::send(fd, p, n, flags);
        if (w < 0) {
            if (errno == EINTR) continue;
            if (errno == EAGAIN || errno == EWOULDBLOCK) return Err::io("send: timed out");  // SO_SNDTIMEO elapsed
            return Err::io(str_printf("send: %s", strerror(errno)));
        }
        p += w, n -= static_cast<size_t>(w);
    }
    return Err::ok();
}

Err recv_exact(int fd, void* buf, size_t n) {
    uint8_t* p = static_cast<uint8_t*>(buf);
    while (n) {
        const ssize_t r = ::recv(fd, p, n, 0);
        if (r == 0) return Err::io("connection closed");
        if (r < 0) {
            if (errno == EINTR) continue;
            if (errno == EAGAIN || errno == EWOULDBLOCK) return Err::io("recv: timed out");  // SO_RCVTIMEO elapsed
            return Err::io(str_printf("recv: %s", strerror(errno)));
        }
        p += r, n -= static_cast<size_t>(r);
    }
    return Err::ok();
}

Err send_file_full(int sock, int file_fd, int64_t off, size_t n) {
    off_t o = off;
    while (n) {
        const ssize_t w = ::sendfile(sock, file_fd, &o, n);
        if (w < 0) {
            if (errno == EINTR || errno == EAGAIN) continue;
            return Err::io(str_printf("sendfile: %s", strerror(errno)));
        }
        if (w == 0) return Err::io("sendfile: unexpected end of file");
        n -= static_cast<size_t>(w);
    }
    return Err::ok();
}

}  // namespace cv
