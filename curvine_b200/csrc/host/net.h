// Blocking TCP helpers (the reference uses tokio sockets with TCP_NODELAY + keepalive:
// orpc/src/client/raw_client.rs:36-39, orpc/src/server/rpc_server.rs:179-181; sendfile loop sys_libc.rs:76-122).
#pragma once
#include "common.h"

namespace cv {

Err tcp_connect(const std::string& host, int port, int* fd_out);
Err tcp_listen(const std::string& host, int port, int* fd_out, int* bound_port);
Err send_all(int fd, const void* buf, size_t n);
Err recv_exact(int fd, void* buf, size_t n);  // kIO "connection closed" on EOF
Err send_file_full(int sock, int file_fd, int64_t off, size_t n);
void set_sock_opts(int fd);
void close_fd(int fd);

}  // namespace cv
