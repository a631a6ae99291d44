// Blocking TCP helpers (the reference uses tokio sockets with TCP_NODELAY + keepalive:
// orpc/src/client/raw_client.rs:36-39, orpc/src/server/rpc_server.rs:179-181; sendfile loop sys_libc.rs:76-122).
#pragma once
#include "common.h"

namespace cv {

// conn_timeout_ms > 0: give up connecting after that long (raw_client.rs:51-55, client_conf.rs conn_timeout_ms);
// io_timeout_ms > 0: every send/recv on the socket fails with kIO "... timed out" after that long without progress
// (the blocking-socket stand-in for RpcClient::timeout_rpc(data_timeout_ms, ..), block_client.rs:56,88-95: an elapsed
// timer becomes io::ErrorKind::TimedOut, i.e. FsError::IO, orpc/src/io/io_error.rs:148-153)
Err tcp_connect(const std::string& host, int port, int* fd_out, int64_t conn_timeout_ms = 0, int64_t io_timeout_ms = 0);
Err tcp_listen(const std::string& host, int port, int* fd_out, int* bound_port);
Err send_all(int fd, const void* buf, size_t n);
Err send_more(int fd, const void* buf, size_t n);  // MSG_MORE: more bytes of the same message follow at once
Err recv_exact(int fd, void* buf, size_t n);  // kIO "connection closed" on EOF
Err send_file_full(int sock, int file_fd, int64_t off, size_t n);
void set_sock_opts(int fd);
void close_fd(int fd);

}  // namespace cv
