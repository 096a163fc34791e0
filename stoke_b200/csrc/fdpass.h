// fdpass.h -- POSIX file descriptors between the rank processes (SCM_RIGHTS over an abstract unix socket).  Pure POSIX, no
// CUDA: the VMM back end (vmm.cu) uses it to hand cuMemExportToShareableHandle descriptors to the peers, and
// tests/test_abi_and_host.py compiles it with g++ and exercises it between two processes on a machine without a GPU.
#pragma once
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <thread>

namespace stk_fd {

inline void sock_name(sockaddr_un& addr, socklen_t& len, int pid, int serial) {
  std::memset(&addr, 0, sizeof(addr));
  addr.sun_family = AF_UNIX;
  char name[64];
  int n = snprintf(name, sizeof(name), "stk_b200.%d.%d", pid, serial);
  addr.sun_path[0] = '\0';  // abstract namespace: no file system entry, vanishes with the process
  std::memcpy(addr.sun_path + 1, name, n);
  len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
}


// ---- descriptor server: serves the descriptors this process exported, by number, to whoever connects ------------------
struct FdServer {
  int listen_fd = -1;
  std::thread th;
  std::atomic<bool> stop{false};
  std::mutex mu;
  std::set<int> exported;  // only descriptors this context exported are served

  bool start(int serial) {
    listen_fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (listen_fd < 0) return false;
    sockaddr_un addr;
    socklen_t len;
    sock_name(addr, len, (int)getpid(), serial);
    if (bind(listen_fd, reinterpret_cast<sockaddr*>(&addr), len) != 0 || listen(listen_fd, 64) != 0) {
      close(listen_fd);
      listen_fd = -1;
      return false;
    }
    th = std::thread([this] { loop(); });
    return true;
  }
  void loop() {
    while (!stop.load()) {
      pollfd pfd{listen_fd, POLLIN, 0};
      int r = poll(&pfd, 1, 100);
      if (r <= 0) continue;
      int cfd = accept4(listen_fd, nullptr, nullptr, SOCK_CLOEXEC);
      if (cfd < 0) continue;
      int want = -1;
      pollfd cp{cfd, POLLIN, 0};
      if (poll(&cp, 1, 5000) > 0 && recv(cfd, &want, sizeof(want), MSG_WAITALL) == (ssize_t)sizeof(want)) {
        bool allowed;
        {
          std::lock_guard<std::mutex> lk(mu);
          allowed = exported.count(want) != 0;
        }
        char payload = allowed ? 'y' : 'n';
        iovec iov{&payload, 1};
        msghdr msg{};
        msg.msg_iov = &iov;
        msg.msg_iovlen = 1;
        alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
        if (allowed) {
          std::memset(ctrl, 0, sizeof(ctrl));
          msg.msg_control = ctrl;
          msg.msg_controllen = sizeof(ctrl);
          cmsghdr* cm = CMSG_FIRSTHDR(&msg);
          cm->cmsg_level = SOL_SOCKET;
          cm->cmsg_type = SCM_RIGHTS;
          cm->cmsg_len = CMSG_LEN(sizeof(int));
          std::memcpy(CMSG_DATA(cm), &want, sizeof(int));
        }
        sendmsg(cfd, &msg, MSG_NOSIGNAL);
      }
      close(cfd);
    }
  }
  void shutdown() {
    stop.store(true);
    if (th.joinable()) th.join();
    if (listen_fd >= 0) close(listen_fd);
    listen_fd = -1;
  }
};


// asks process `pid` / context `serial` for its descriptor number `remote_fd`; returns a local descriptor or -1
inline int fetch_fd(int pid, int serial, int remote_fd, std::string& why) {
  int fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (fd < 0) {
    why = "socket() failed";
    return -1;
  }
  sockaddr_un addr;
  socklen_t len;
  sock_name(addr, len, pid, serial);
  int rc = -1;
  for (int attempt = 0; attempt < 200 && rc != 0; ++attempt) {  // the peer's server may not be listening yet
    rc = connect(fd, reinterpret_cast<sockaddr*>(&addr), len);
    if (rc != 0) usleep(10000);
  }
  if (rc != 0) {
    why = "connect() to the peer's descriptor server failed";
    close(fd);
    return -1;
  }
  if (send(fd, &remote_fd, sizeof(remote_fd), MSG_NOSIGNAL) != (ssize_t)sizeof(remote_fd)) {
    why = "send() failed";
    close(fd);
    return -1;
  }
  char payload = 0;
  iovec iov{&payload, 1};
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  std::memset(ctrl, 0, sizeof(ctrl));
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  pollfd pfd{fd, POLLIN, 0};
  int got = -1;
  if (poll(&pfd, 1, 20000) > 0 && recvmsg(fd, &msg, MSG_CMSG_CLOEXEC) == 1 && payload == 'y') {
    for (cmsghdr* cm = CMSG_FIRSTHDR(&msg); cm; cm = CMSG_NXTHDR(&msg, cm))
      if (cm->cmsg_level == SOL_SOCKET && cm->cmsg_type == SCM_RIGHTS) std::memcpy(&got, CMSG_DATA(cm), sizeof(int));
  }
  if (got < 0) why = "the peer did not hand over the descriptor";
  close(fd);
  return got;
}


}  // namespace stk_fd
