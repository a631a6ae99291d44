// Round-2 box probe for the framed path: what is the ceiling of same-host socket transport on this box?
// (tools/ = measurement scaffolding, not product code.)
//   server: one thread per connection; per request (8-byte length) it sends that many bytes from a tmpfs file region
//           either by sendfile(2) (what the reference worker does, rpc_frame.rs:97-121) or by send(2) from a mapping
//   client: one thread per connection, requests 4 MiB chunks, receives into a private buffer with big recv() calls
// Sweeps: connections, TCP vs AF_UNIX, sendfile vs send-from-mapping, CPU placement of the two sides.
#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/sendfile.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static int g_ncpu = 1;
// placement: 0 none, 1 = first half of each socket's threads ("node0": cpus [0,n/4) + [n/2, 3n/4)), 2 = "node1"
static void place(int where) {
    if (where == 0) return;
    cpu_set_t s;
    CPU_ZERO(&s);
    const int q = g_ncpu / 4;
    if (q == 0) return;
    const int base = where == 1 ? 0 : q;
    for (int c = 0; c < q; c++) CPU_SET(base + c, &s), CPU_SET(base + c + g_ncpu / 2, &s);
    sched_setaffinity(0, sizeof(s), &s);
}
static const size_t CHUNK = 4 << 20;
static size_t g_file_bytes = 0;
static int g_file_fd = -1;
static uint8_t* g_map = nullptr;

static void serve_conn(int fd, int mode, int where) {
    place(where);
    uint64_t req[2];
    for (;;) {
        size_t got = 0;
        while (got < sizeof(req)) {
            ssize_t r = recv(fd, (char*)req + got, sizeof(req) - got, 0);
            if (r <= 0) { close(fd); return; }
            got += r;
        }
        size_t off = req[0], len = req[1];
        uint8_t prefix[22] = {0};
        if (send(fd, prefix, 22, MSG_MORE) != 22) { close(fd); return; }
        if (mode == 0) {
            off_t o = off;
            size_t left = len;
            while (left) {
                ssize_t w = sendfile(fd, g_file_fd, &o, left);
                if (w <= 0) { if (errno == EINTR) continue; close(fd); return; }
                left -= w;
            }
        } else {
            size_t done = 0;
            while (done < len) {
                ssize_t w = send(fd, g_map + off + done, len - done, 0);
                if (w <= 0) { if (errno == EINTR) continue; close(fd); return; }
                done += w;
            }
        }
    }
}

int main(int argc, char** argv) {
    g_ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    const size_t GiB = 1ull << 30;
    g_file_bytes = (argc > 1 ? atoi(argv[1]) : 8) * GiB;
    const double per_run_gib = argc > 2 ? atof(argv[2]) : 24;
    const char* path = "/dev/shm/loopback_probe.dat";
    g_file_fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (ftruncate(g_file_fd, g_file_bytes) != 0) { perror("ftruncate"); return 1; }
    g_map = (uint8_t*)mmap(nullptr, g_file_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, g_file_fd, 0);
    {
        std::vector<std::thread> ts;
        for (int t = 0; t < 16; t++) ts.emplace_back([&, t] { place(1); for (size_t o = g_file_bytes * t / 16; o < g_file_bytes * (t + 1) / 16; o += 4096) g_map[o] = (uint8_t)o; });
        for (auto& t : ts) t.join();
    }
    printf("ncpu %d file %zu GiB per-run %.0f GiB\n", g_ncpu, g_file_bytes / GiB, per_run_gib);
    setvbuf(stdout, nullptr, _IOLBF, 0);
    auto run = [&](bool unix_sock, int mode, int conns, int cwhere, int swhere, int rcvbuf) {
        int lfd;
        sockaddr_in a4{};
        sockaddr_un au{};
        socklen_t alen;
        sockaddr* ap;
        if (unix_sock) {
            lfd = socket(AF_UNIX, SOCK_STREAM, 0);
            au.sun_family = AF_UNIX;
            snprintf(au.sun_path, sizeof(au.sun_path), "/tmp/lbprobe.%d.sock", getpid());
            unlink(au.sun_path);
            ap = (sockaddr*)&au, alen = sizeof(au);
        } else {
            lfd = socket(AF_INET, SOCK_STREAM, 0);
            int one = 1;
            setsockopt(lfd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
            a4.sin_family = AF_INET, a4.sin_addr.s_addr = htonl(INADDR_LOOPBACK), a4.sin_port = 0;
            ap = (sockaddr*)&a4, alen = sizeof(a4);
        }
        if (bind(lfd, ap, alen) != 0 || listen(lfd, 256) != 0) { perror("bind/listen"); exit(1); }
        if (!unix_sock) getsockname(lfd, ap, &alen);
        std::vector<std::thread> sthreads;
        std::thread acc([&] {
            for (int i = 0; i < conns; i++) {
                int fd = accept(lfd, nullptr, nullptr);
                if (fd < 0) break;
                if (rcvbuf) setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &rcvbuf, sizeof(rcvbuf));
                sthreads.emplace_back(serve_conn, fd, mode, swhere);
            }
        });
        const size_t total_chunks = (size_t)(per_run_gib * GiB) / CHUNK;
        std::atomic<size_t> next{0};
        std::vector<std::thread> cthreads;
        std::vector<int> cfds(conns);
        for (int i = 0; i < conns; i++) {
            int fd = socket(unix_sock ? AF_UNIX : AF_INET, SOCK_STREAM, 0);
            if (rcvbuf) setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &rcvbuf, sizeof(rcvbuf));
            if (connect(fd, ap, alen) != 0) { perror("connect"); exit(1); }
            if (!unix_sock) { int one = 1; setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one)); }
            cfds[i] = fd;
        }
        acc.join();
        const double t0 = now();
        for (int i = 0; i < conns; i++)
            cthreads.emplace_back([&, i] {
                place(cwhere);
                uint8_t* buf = (uint8_t*)aligned_alloc(4096, CHUNK + 4096);
                memset(buf, 1, CHUNK + 4096);
                const size_t file_chunks = g_file_bytes / CHUNK;
                for (;;) {
                    size_t c = next.fetch_add(1);
                    if (c >= total_chunks) break;
                    uint64_t req[2] = {(c % file_chunks) * CHUNK, CHUNK};
                    if (send(cfds[i], req, sizeof(req), 0) != sizeof(req)) { perror("send"); exit(1); }
                    size_t got = 0;
                    while (got < CHUNK + 22) {
                        ssize_t r = recv(cfds[i], buf + got, CHUNK + 22 - got, 0);
                        if (r <= 0) { perror("recv"); exit(1); }
                        got += r;
                    }
                }
                free(buf);
            });
        for (auto& t : cthreads) t.join();
        const double dt = now() - t0;
        for (int fd : cfds) close(fd);
        for (auto& t : sthreads) t.join();
        close(lfd);
        if (unix_sock) unlink(au.sun_path);
        printf("%-5s %-8s conns=%2d client@%d server@%d buf=%4dK : %6.2f GB/s\n", unix_sock ? "unix" : "tcp", mode == 0 ? "sendfile" : "send-map", conns, cwhere, swhere,
               rcvbuf >> 10, total_chunks * (double)CHUNK / dt / 1e9);
    };
    run(false, 0, 16, 1, 0, 0);  // warm-up + round-1 configuration (client on node0, server unbound)
    for (int unix_sock = 0; unix_sock < 2; unix_sock++)
        for (int mode = 0; mode < 2; mode++) {
            for (int conns : {8, 16, 24, 32, 48}) run(unix_sock, mode, conns, 1, 2, 0);  // client node0, server node1
            for (int conns : {16, 32}) run(unix_sock, mode, conns, 1, 1, 0);            // both on node0
            for (int conns : {16, 32}) run(unix_sock, mode, conns, 1, 0, 0);            // server unbound
            run(unix_sock, mode, 24, 1, 2, 4 << 20);                                      // big socket buffers
        }
    unlink(path);
    return 0;
}
