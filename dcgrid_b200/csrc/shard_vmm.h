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

// serves the descriptors `fds` (all of them, in order) to `npeers` connecting peers on an abstract socket, in a
// helper thread
struct FdServer {
  int lfd = -1;
  std::thread th;
  std::atomic<int> served{0};
  char name[kHandleBytes] = {0};

  bool start(std::vector<int> fds, int npeers) {
    static std::atomic<int> counter{0};
    std::snprintf(name, sizeof name, "dcgrid-b200-vmm-%d-%d", (int)getpid(), counter.fetch_add(1));
    lfd = socket(AF_UNIX, SOCK_STREAM, 0);
    if (lfd < 0) return false;
    socklen_t len;
    sockaddr_un a = abstract_addr(name, len);
    if (bind(lfd, reinterpret_cast<sockaddr *>(&a), len) != 0 || listen(lfd, 16) != 0) return false;
    th = std::thread([this, fds, npeers] {
      for (int i = 0; i < npeers; i++) {
        pollfd p = {lfd, POLLIN, 0};
        if (poll(&p, 1, 120000) <= 0) break;  // a peer never came: give up after two minutes
        const int c = accept(lfd, nullptr, nullptr);
        if (c < 0) break;
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
    if (th.joinable()) th.join();
    if (lfd >= 0) close(lfd);
    lfd = -1;
  }
  ~FdServer() { finish(); }
};

// receives `count` descriptors from the peer listening on `name`; false if the peer cannot be reached
inline bool fetch_fds(const char *name, int count, std::vector<int> &out) {
  for (int attempt = 0; attempt < 1200; attempt++) {  // up to two minutes: the peer may not be listening yet
    const int s = socket(AF_UNIX, SOCK_STREAM, 0);
    if (s < 0) return false;
    socklen_t len;
    sockaddr_un a = abstract_addr(name, len);
    if (connect(s, reinterpret_cast<sockaddr *>(&a), len) == 0) {
      out.clear();
      for (int i = 0; i < count; i++) {
        const int fd = recv_fd(s);
        if (fd < 0) break;
        out.push_back(fd);
      }
      close(s);
      return (int)out.size() == count;
    }
    close(s);
    usleep(100000);
  }
  return false;
}

}  // namespace vmm
}  // namespace dcg
