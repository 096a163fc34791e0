// test_fdpass.cpp -- CPU check of stoke_b200/csrc/fdpass.h: a descriptor exported by one process is fetched by another
// (SCM_RIGHTS) and refers to the same open file; descriptors that were not exported are refused.
//   g++ -std=c++17 -O1 -pthread tools/test_fdpass.cpp -o /tmp/test_fdpass && /tmp/test_fdpass
#include <fcntl.h>
#include <sys/wait.h>

#include <cstdlib>

#include "../stoke_b200/csrc/fdpass.h"

int main() {
  int pipefd[2];
  if (pipe(pipefd) != 0) return 2;
  int secret[2];
  if (pipe(secret) != 0) return 2;
  const int parent = (int)getpid();
  pid_t child = fork();
  if (child == 0) {
    // child: fetch the parent's read end by number, read the message through it
    std::string why;
    int fd = stk_fd::fetch_fd(parent, 7, pipefd[0], why);
    if (fd < 0) {
      fprintf(stderr, "fetch failed: %s\n", why.c_str());
      _exit(10);
    }
    char buf[16] = {};
    if (read(fd, buf, 5) != 5 || std::memcmp(buf, "hello", 5) != 0) _exit(11);
    int no = stk_fd::fetch_fd(parent, 7, secret[0], why);  // not exported: must be refused
    if (no >= 0) _exit(12);
    _exit(0);
  }
  stk_fd::FdServer srv;
  {
    std::lock_guard<std::mutex> lk(srv.mu);
    srv.exported.insert(pipefd[0]);  // exported before anybody can ask for it (vmm.cu: before the blob is handed out)
  }
  if (!srv.start(7)) return 3;
  if (write(pipefd[1], "hello", 5) != 5) return 4;
  int status = 0;
  waitpid(child, &status, 0);
  srv.shutdown();
  if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) {
    fprintf(stderr, "child status %d\n", WIFEXITED(status) ? WEXITSTATUS(status) : -1);
    return 5;
  }
  printf("fdpass ok\n");
  return 0;
}
