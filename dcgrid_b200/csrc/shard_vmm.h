// Cross-process unified address space for the slab-decomposed DCGrid solver (host side only).
//
// One process per GPU.  Every rank creates the physical pieces it owns on its device (cuMemCreate: one per field
// and run of consecutive owned units — cuMemMap maps whole allocations only), exports them as POSIX file
// descriptors, and every process stitches the pieces of all ranks into one virtual range per field
// (cuMemAddressReserve + cuMemMap), so that a pool cell id indexes the same array on every GPU: kernels are the
// single-GPU kernels, a load or store of a cell another rank owns simply travels over NVLink / NVSwitch.
// The driver entry points are resolved through cudaGetDriverEntryPoint (no link-time dependency on libcuda, so
// the library still loads on a machine without a driver); descriptors travel over abstract AF_UNIX sockets
// (SCM_RIGHTS), whose names are the 64-byte handles the ranks exchange through any host channel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace dcg {
namespace vmm {

constexpr size_t kHandleBytes = 64;

struct Driver {
  CUresult (*memCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
  CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*memExport)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*memImport)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
  CUresult (*memReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*memFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*memUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
  CUresult (*memGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
  bool ok = false;

  template <class F>
  static bool get(const char *name, F &fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) return false;
    fn = reinterpret_cast<F>(p);
    return true;
  }
  bool load() {
    ok = get("cuMemCreate", memCreate) && get("cuMemRelease", memRelease) && get("cuMemExportToShareableHandle", memExport) &&
         get("cuMemImportFromShareableHandle", memImport) && get("cuMemAddressReserve", memReserve) && get("cuMemAddressFree", memFree) &&
         get("cuMemMap", memMap) && get("cuMemUnmap", memUnmap) && get("cuMemSetAccess", memSetAccess) &&
         get("cuMemGetAllocationGranularity", memGranularity);
    return ok;
  }
};

inline CUmemAllocationProp device_prop(int device) {
  CUmemAllocationProp prop;
  std::memset(&prop, 0, sizeof prop);
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

// ---- descriptor passing -------------------------------------------------------------------------------------
inline sockaddr_un abstract_addr(const char *name, socklen_t &len) {
  sockaddr_un a;
  std::memset(&a, 0, sizeof a);
  a.sun_family = AF_UNIX;
  const size_t n = std::min(std::strlen(name), sizeof(a.sun_path) - 2);
  std::memcpy(a.sun_path + 1, name, n);  // leading NUL = abstract namespace: no file, vanishes with the process
  len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
  return a;
}
inline bool send_fd(int sock, int fd) {
  char byte = 'F';
  iovec io = {&byte, 1};
  char ctl[CMSG_SPACE(sizeof(int))];
  std::memset(ctl, 0, sizeof ctl);
  msghdr msg;
  std::memset(&msg, 0, sizeof msg);
  msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = ctl; msg.msg_controllen = sizeof ctl;
  cmsghdr *c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
  std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
  return sendmsg(sock, &msg, 0) == 1;
}
inline int recv_fd(int sock) {
  char byte = 0;
  iovec io = {&byte, 1};
  char ctl[CMSG_SPACE(sizeof(int))];
  std::memset(ctl, 0, sizeof ctl);
  msghdr msg;
  std::memset(&msg, 0, sizeof msg);
  msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = ctl; msg.msg_controllen = sizeof ctl;
  if (recvmsg(sock, &msg, 0) != 1) return -1;
  cmsghdr *c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) return -1;
  int fd = -1;
  std::memcpy(&fd, CMSG_DATA(c), sizeof(int));
  return fd;
}

// Handle blob (kHandleBytes, exchanged by the ranks through any host channel): bytes [0, 48) = NUL-terminated name
// of the abstract socket, bytes [48, 64) = a random token.  The server only sends descriptors to a peer that
// (a) runs under the same effective uid (SO_PEERCRED) and (b) presents the token: an abstract socket has no
// filesystem permissions, and its name alone is guessable.  Connections that fail either test are dropped
// without consuming one of the `npeers` slots.
constexpr size_t kNameBytes = 48, kTokenBytes = 16;

inline bool random_bytes(unsigned char *out, size_t n) {
  FILE *f = std::fopen("/dev/urandom", "rb");
  if (!f) return false;
  const bool ok = std::fread(out, 1, n, f) == n;
  std::fclose(f);
  return ok;
}
inline bool read_exact(int fd, void *buf, size_t n, int timeout_ms) {
  size_t got = 0;
  while (got < n) {
    pollfd p = {fd, POLLIN, 0};
    if (poll(&p, 1, timeout_ms) <= 0) return false;
    const ssize_t r = read(fd, static_cast<char *>(buf) + got, n - got);
    if (r <= 0) return false;
    got += (size_t)r;
  }
  return true;
}

// serves the descriptors `fds` (all of them, in order) to `npeers` authenticated peers on an abstract socket, in
// a helper thread; finish() cancels it
struct FdServer {
  int lfd = -1;
  std::thread th;
  std::atomic<int> served{0}, rejected{0};
  std::atomic<bool> stop{false};
  char name[kHandleBytes] = {0};  // the whole handle blob: socket name + token

  bool start(std::vector<int> fds, int npeers) {
    static std::atomic<int> counter{0};
    auto close_all = [&fds] { for (int fd : fds) close(fd); };
    std::memset(name, 0, sizeof name);
    std::snprintf(name, kNameBytes, "dcgrid-b200-vmm-%d-%d", (int)getpid(), counter.fetch_add(1));
    if (!random_bytes(reinterpret_cast<unsigned char *>(name) + kNameBytes, kTokenBytes)) { close_all(); return false; }
    lfd = socket(AF_UNIX, SOCK_STREAM, 0);
    if (lfd < 0) { close_all(); return false; }
    socklen_t len;
    sockaddr_un a = abstract_addr(name, len);
    if (bind(lfd, reinterpret_cast<sockaddr *>(&a), len) != 0 || listen(lfd, 16) != 0) {
      close(lfd);
      lfd = -1;
      close_all();
      return false;
    }
    th = std::thread([this, fds, npeers] {
      int waited_ms = 0;
      while (served.load() < npeers && !stop.load() && waited_ms < 120000) {  // a peer never came: give up after two minutes
        pollfd p = {lfd, POLLIN, 0};
        const int pr = poll(&p, 1, 100);
        if (pr < 0) break;
        if (pr == 0) { waited_ms += 100; continue; }
        const int c = accept(lfd, nullptr, nullptr);
        if (c < 0) break;
        ucred cred;
        socklen_t cl = sizeof cred;
        char token[kTokenBytes];
        const bool trusted = getsockopt(c, SOL_SOCKET, SO_PEERCRED, &cred, &cl) == 0 && cred.uid == geteuid() &&
                             read_exact(c, token, kTokenBytes, 5000) && std::memcmp(token, name + kNameBytes, kTokenBytes) == 0;
        if (!trusted) {
          rejected.fetch_add(1);
          close(c);
          continue;
        }
        bool ok = true;
        for (int fd : fds) ok = ok && send_fd(c, fd);
        if (ok) served.fetch_add(1);
        close(c);
      }
      for (int fd : fds) close(fd);  // the allocations stay alive through their handles
    });
    return true;
  }
  void finish() {
    stop.store(true);
    if (th.joinable()) th.join();
    if (lfd >= 0) close(lfd);
    lfd = -1;
  }
  // waits (bounded) until every peer has been served, then stops the thread
  void finish_after_serving(int npeers, int timeout_ms) {
    for (int w = 0; served.load() < npeers && w < timeout_ms; w += 10) usleep(10000);
    finish();
  }
  ~FdServer() { finish(); }
};

// receives `count` descriptors from the peer whose handle blob is `handle`; false if the peer cannot be reached
inline bool fetch_fds(const char *handle, int count, std::vector<int> &out) {
  char sock_name[kNameBytes + 1] = {0};
  std::memcpy(sock_name, handle, kNameBytes);
  for (int attempt = 0; attempt < 1200; attempt++) {  // up to two minutes: the peer may not be listening yet
    const int s = socket(AF_UNIX, SOCK_STREAM, 0);
    if (s < 0) return false;
    socklen_t len;
    sockaddr_un a = abstract_addr(sock_name, len);
    if (connect(s, reinterpret_cast<sockaddr *>(&a), len) == 0) {
      out.clear();
      bool ok = write(s, handle + kNameBytes, kTokenBytes) == (ssize_t)kTokenBytes;
      for (int i = 0; ok && i < count; i++) {
        const int fd = recv_fd(s);
        if (fd < 0) break;
        out.push_back(fd);
      }
      close(s);
      if ((int)out.size() == count) return true;
      for (int fd : out) close(fd);  // partial transfer: do not leak what did arrive
      out.clear();
      return false;
    }
    close(s);
    usleep(100000);
  }
  return false;
}

}  // namespace vmm
}  // namespace dcg
