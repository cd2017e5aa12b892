// CPU test of the descriptor exchange used by the slab-decomposed DCGrid solver (dcgrid_b200/csrc/shard_vmm.h):
// two processes publish three descriptors each on an abstract AF_UNIX socket (vmm::FdServer) and fetch the peer's
// (vmm::fetch_fds) — here the descriptors are pipes carrying a rank-specific message instead of GPU allocations.
// usage: fd_exchange_test   (forks the second rank itself; exit code 0 = every descriptor arrived intact)
#include <sys/wait.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../dcgrid_b200/csrc/shard_vmm.h"

using namespace dcg;

static int run_rank(int rank, int name_pipe_in, int name_pipe_out) {
  // three "allocations": pipes whose read ends travel to the peer; the payload identifies (rank, index)
  std::vector<int> send_fds;
  for (int i = 0; i < 3; i++) {
    int p[2];
    if (pipe(p) != 0) return 10;
    char msg[32];
    std::snprintf(msg, sizeof msg, "rank%d-piece%d", rank, i);
    if (write(p[1], msg, std::strlen(msg) + 1) < 0) return 11;
    close(p[1]);
    send_fds.push_back(p[0]);
  }
  vmm::FdServer server;
  if (!server.start(send_fds, 1)) return 12;
  // the 64-byte handle = the socket name, exchanged through "any host channel" (here: a pair of pipes)
  if (write(name_pipe_out, server.name, vmm::kHandleBytes) != (ssize_t)vmm::kHandleBytes) return 13;
  char peer[vmm::kHandleBytes + 1] = {0};
  if (read(name_pipe_in, peer, vmm::kHandleBytes) != (ssize_t)vmm::kHandleBytes) return 14;
  std::vector<int> got;
  if (!vmm::fetch_fds(peer, 3, got)) return 15;
  server.finish_after_serving(1, 30000);  // (the peer may not have connected yet: finish() alone would cancel the server under it)
  if (server.served.load() != 1) return 16;
  for (int i = 0; i < 3; i++) {
    char buf[32] = {0}, want[32];
    std::snprintf(want, sizeof want, "rank%d-piece%d", 1 - rank, i);
    if (read(got[i], buf, sizeof buf) <= 0 || std::strcmp(buf, want) != 0) {
      std::fprintf(stderr, "rank %d: descriptor %d carried '%s', expected '%s'\n", rank, i, buf, want);
      return 17;
    }
    close(got[i]);
  }
  return 0;
}

int main() {
  int a2b[2], b2a[2];
  if (pipe(a2b) != 0 || pipe(b2a) != 0) return 1;
  const pid_t child = fork();
  if (child < 0) return 2;
  if (child == 0) std::exit(run_rank(1, a2b[0], b2a[1]));
  const int rc0 = run_rank(0, b2a[0], a2b[1]);
  int st = 0;
  waitpid(child, &st, 0);
  const int rc1 = WIFEXITED(st) ? WEXITSTATUS(st) : 99;
  std::printf("rank0 rc=%d rank1 rc=%d\n", rc0, rc1);
  return rc0 == 0 && rc1 == 0 ? 0 : 3;
}
