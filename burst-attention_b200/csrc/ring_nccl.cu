// Ring transport over NCCL point-to-point on a dedicated side stream
// (replaces comm.py Ring single-ring path: _make_ring_ops :148-172, _commit_ops
// :267-283 (the BMTrain side-stream variant), wait :301-321).
//
// NCCL is resolved at run time with dlopen so that the library shares the one
// libnccl.so.2 the host process (PyTorch) already loaded instead of linking a
// second copy.  One ring = one communicator + one high-priority stream + two
// events; no host synchronisation anywhere.  A ring whose receive arena has been connected
// (ba_ring_arena_*, ring_ce.cu) posts its hops over the copy engines instead.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "ring_internal.h"

namespace ba {

// Minimal NCCL ABI (stable since 2.7): opaque comm, 128-byte unique id, result enum.
typedef struct {
  char internal[BA_NCCL_UNIQUE_ID_BYTES];
} ncclUniqueId;
typedef int ncclResult_t;  // ncclSuccess == 0
constexpr int kNcclInt8 = 0;

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, void*);  // optional (NCCL >= 2.14)
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
#define BA_SYM(name)                                                         \
  api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name)); \
  if (!api.name) return;
    BA_SYM(GetUniqueId)
    BA_SYM(CommInitRank)
    BA_SYM(CommDestroy)
    BA_SYM(Send)
    BA_SYM(Recv)
    BA_SYM(GroupStart)
    BA_SYM(GroupEnd)
    BA_SYM(GetErrorString)
#undef BA_SYM
    api.CommInitRankConfig = reinterpret_cast<decltype(api.CommInitRankConfig)>(dlsym(h, "ncclCommInitRankConfig"));
    api.ok = true;
  });
  return api;
}

// ncclConfig_t as of NCCL 2.18 (newer libraries accept an older, shorter struct by its size/version fields and
// default the attributes added later).  Only maxCTAs is set: the ring moves a few hundred MB per round while the
// tile kernels own every SM, and each CTA NCCL's SM-resident send/recv kernel occupies is an SM the tile kernel
// loses for the duration of the hop -- the hop needs bandwidth for < 100 GB/s, not NCCL's default channel count.
struct NcclConfigV21800 {
  size_t size;
  unsigned int magic;
  unsigned int version;
  int blocking;
  int cgaClusterSize;
  int minCTAs;
  int maxCTAs;
  const char* netName;
  int splitShare;
};
constexpr int kNcclUndefInt = -2147483647 - 1;  // NCCL_CONFIG_UNDEF_INT

static int nccl_fail(ncclResult_t r, const char* what) {
  set_error("NCCL error %d (%s) at %s", r, nccl().ok ? nccl().GetErrorString(r) : "?", what);
  return BA_ERR_NCCL;
}
#define BA_CHECK_NCCL(expr)                              \
  do {                                                   \
    ::ba::ncclResult_t _r = (expr);                      \
    if (_r != 0) return ::ba::nccl_fail(_r, #expr);      \
  } while (0)

}  // namespace ba

extern "C" int ba_ring_unique_id(void* out_id128) {
  using namespace ba;
  BA_REQUIRE(out_id128, "ba_ring_unique_id: null output");
  if (!nccl().ok) {
    set_error("libnccl.so.2 could not be loaded");
    return BA_ERR_NCCL;
  }
  ncclUniqueId id;
  BA_CHECK_NCCL(nccl().GetUniqueId(&id));
  memcpy(out_id128, &id, sizeof(id));
  return BA_OK;
}

extern "C" int ba_ring_create(const void* id128, int rank, int world, ba_ring** out) {
  using namespace ba;
  BA_REQUIRE(out && world >= 1 && rank >= 0 && rank < world, "ba_ring_create: bad rank/world %d/%d", rank, world);
  ba_ring* r = new ba_ring();
  r->rank = rank;
  r->world = world;
  BA_CHECK_CUDA(cudaGetDevice(&r->device));
  int lo = 0, hi = 0;
  BA_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  BA_CHECK_CUDA(cudaStreamCreateWithPriority(&r->side, cudaStreamNonBlocking, hi));
  BA_CHECK_CUDA(cudaEventCreateWithFlags(&r->ev_ready, cudaEventDisableTiming));
  BA_CHECK_CUDA(cudaEventCreateWithFlags(&r->ev_done, cudaEventDisableTiming));
  if (world > 1 && id128) {  // id128 == NULL: a copy-engine-only ring (ba_ring_arena_*), no communicator
    if (!nccl().ok) {
      set_error("libnccl.so.2 could not be loaded");
      return BA_ERR_NCCL;
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    // BA_NCCL_MAX_CTAS (default 0 = NCCL's own choice): cap the CTAs of this ring's send/recv kernels
    const char* e = getenv("BA_NCCL_MAX_CTAS");
    const int max_ctas = e ? atoi(e) : 0;
    if (max_ctas > 0 && nccl().CommInitRankConfig) {
      NcclConfigV21800 cfg = {sizeof(NcclConfigV21800), 0xcafebeefu, 21800u, kNcclUndefInt, kNcclUndefInt, 1, max_ctas,
                              nullptr, kNcclUndefInt};
      BA_CHECK_NCCL(nccl().CommInitRankConfig(&r->comm, world, id, rank, &cfg));
    } else {
      BA_CHECK_NCCL(nccl().CommInitRank(&r->comm, world, id, rank));
    }
  }
  *out = r;
  return BA_OK;
}

extern "C" int ba_ring_post(ba_ring* ring, const void* const* src, void* const* dst, const int64_t* nbytes, int n,
                            void* compute_stream) {
  using namespace ba;
  BA_REQUIRE(ring && (n == 0 || (src && dst && nbytes)), "ba_ring_post: bad arguments");
  cudaStream_t cs = static_cast<cudaStream_t>(compute_stream);
  // sources were produced (and destinations last consumed) by work already queued on cs
  BA_CHECK_CUDA(cudaEventRecord(ring->ev_ready, cs));
  BA_CHECK_CUDA(cudaStreamWaitEvent(ring->side, ring->ev_ready, 0));
  if (ring->world == 1) {
    for (int i = 0; i < n; ++i)
      BA_CHECK_CUDA(cudaMemcpyAsync(dst[i], src[i], (size_t)nbytes[i], cudaMemcpyDeviceToDevice, ring->side));
  } else if (ring->ce.connected) {
    int rc = ce_post(ring, src, dst, nbytes, n);
    if (rc != BA_OK) return rc;
  } else {
    BA_REQUIRE(ring->comm, "ba_ring_post: the ring has neither a communicator nor a connected arena");
    const int next = (ring->rank + 1) % ring->world;
    const int prev = (ring->rank + ring->world - 1) % ring->world;
    BA_CHECK_NCCL(nccl().GroupStart());
    for (int i = 0; i < n; ++i) {
      BA_CHECK_NCCL(nccl().Send(src[i], (size_t)nbytes[i], kNcclInt8, next, ring->comm, ring->side));
      BA_CHECK_NCCL(nccl().Recv(dst[i], (size_t)nbytes[i], kNcclInt8, prev, ring->comm, ring->side));
    }
    BA_CHECK_NCCL(nccl().GroupEnd());
  }
  BA_CHECK_CUDA(cudaEventRecord(ring->ev_done, ring->side));
  return BA_OK;
}

extern "C" int ba_ring_wait(ba_ring* ring, void* compute_stream) {
  using namespace ba;
  BA_REQUIRE(ring, "ba_ring_wait: null ring");
  BA_CHECK_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(compute_stream), ring->ev_done, 0));
  return BA_OK;
}

extern "C" int ba_ring_rank(const ba_ring* ring) { return ring ? ring->rank : -1; }
extern "C" int ba_ring_world(const ba_ring* ring) { return ring ? ring->world : -1; }

extern "C" int ba_ring_destroy(ba_ring* ring) {
  using namespace ba;
  if (!ring) return BA_OK;
  if (ring->side) cudaStreamSynchronize(ring->side);
  ce_destroy(ring);
  if (ring->comm && nccl().ok) nccl().CommDestroy(ring->comm);
  if (ring->ev_ready) cudaEventDestroy(ring->ev_ready);
  if (ring->ev_done) cudaEventDestroy(ring->ev_done);
  if (ring->side) cudaStreamDestroy(ring->side);
  delete ring;
  return BA_OK;
}
