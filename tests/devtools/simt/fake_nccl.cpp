// fake_nccl.cpp — development aid: the five NCCL entry points libkdnb dlopens (kdnb_api.cu: load_nccl), for "ranks" that
// are THREADS of one process running the kernel sources under the SIMT model (multirank.py).  Built as libnccl.so.2 into
// _build/ and loaded before the library looks for it.  Collectives are synchronous: publish pointers, barrier, copy,
// barrier.  Not product, not shipped, never loaded next to the real NCCL.
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

namespace {
struct Group {
  int n = 0, joined = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  unsigned gen = 0;
  const void* send[64] = {};
  void* recv[64] = {};
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = gen;
    if (++arrived == n) {
      arrived = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
};
struct Comm {
  Group* g;
  int rank;
};
std::mutex g_mu;
std::map<std::string, Group*> g_groups;
unsigned long long g_next_id = 1;
struct Id {
  char internal[128];
};
size_t dtype_bytes(int t) {
  switch (t) {
    case 0: case 1: return 1;   // ncclInt8 / ncclUint8
    case 2: case 3: return 4;   // ncclInt32 / ncclUint32
    case 4: case 5: return 8;   // ncclInt64 / ncclUint64
    case 6: return 2;           // ncclFloat16
    case 7: return 4;           // ncclFloat32
    case 8: return 8;           // ncclFloat64
    default: return 0;
  }
}
}  // namespace

extern "C" {

int ncclGetUniqueId(Id* id) {
  std::lock_guard<std::mutex> lk(g_mu);
  memset(id, 0, sizeof *id);
  const unsigned long long v = g_next_id++;
  memcpy(id->internal, "simt-fake-nccl", 14);
  memcpy(id->internal + 16, &v, sizeof v);
  return 0;
}

int ncclCommInitRank(void** comm, int nranks, Id id, int rank) {
  if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) return 4;  // ncclInvalidArgument
  Group* g = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    Group*& slot = g_groups[std::string(id.internal, sizeof id.internal)];
    if (!slot) {
      slot = new Group();
      slot->n = nranks;
    }
    g = slot;
  }
  if (g->n != nranks) return 4;
  *comm = new Comm{g, rank};
  g->barrier();  // like the real call: returns once every rank has joined
  return 0;
}

int ncclAllGather(const void* send, void* recv, size_t count, int dtype, void* comm, void* /*stream*/) {
  Comm* c = static_cast<Comm*>(comm);
  Group* g = c->g;
  const size_t bytes = count * dtype_bytes(dtype);
  if (!bytes && count) return 4;
  g->send[c->rank] = send;
  g->recv[c->rank] = recv;
  g->barrier();
  for (int r = 0; r < g->n; ++r) {
    char* dst = static_cast<char*>(recv) + (size_t)r * bytes;
    if (dst != g->send[r]) memmove(dst, g->send[r], bytes);
  }
  g->barrier();  // nobody reuses its send buffer before everybody has read it
  return 0;
}

int ncclCommDestroy(void* comm) {
  delete static_cast<Comm*>(comm);  // (groups live until the process ends)
  return 0;
}

const char* ncclGetErrorString(int r) { return r == 0 ? "no error" : "fake nccl: invalid argument"; }

}  // extern "C"
