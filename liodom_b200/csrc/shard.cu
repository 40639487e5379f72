// Point-sharded mode (BASELINE.json config 5): one scan is split across the GPUs of a
// node.  Extraction shards by ring, registration by edge; the local map and the trust-region
// controller are replicated.  The only data-path collectives are an all-gather of the edge slots
// and one all-reduce of 29 doubles (upper-triangular J'J, J'r, cost, block count) per LM evaluation.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 — the copy torch already loaded, if any), so
// the library loads and the single-GPU path works on machines without NCCL.
#include "common.cuh"

#include <dlfcn.h>
#include <cstdio>
#include <cstring>

namespace liodom {

namespace {
typedef int (*fn_get_uid)(void*);
typedef int (*fn_init_rank)(void**, int, const void*, int);   // ncclUniqueId passed by value: see call below
typedef int (*fn_destroy)(void*);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_group)(void);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
  void* lib = nullptr;
  void* get_uid = nullptr; void* init_rank = nullptr; fn_destroy destroy = nullptr;
  fn_allreduce allreduce = nullptr; fn_allgather allgather = nullptr; fn_group group_start = nullptr, group_end = nullptr;
  fn_errstr errstr = nullptr;
  bool ok = false;
};
NcclApi g_nccl;

struct UniqueId { char internal[128]; };   // ncclUniqueId
typedef int (*fn_get_uid_t)(UniqueId*);
typedef int (*fn_init_rank_t)(void**, int, UniqueId, int);

bool load_nccl() {
  if (g_nccl.ok) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) return false;
  g_nccl.get_uid = dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.init_rank = dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.destroy = (fn_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.allgather = (fn_allgather)dlsym(g_nccl.lib, "ncclAllGather");
  g_nccl.group_start = (fn_group)dlsym(g_nccl.lib, "ncclGroupStart");
  g_nccl.group_end = (fn_group)dlsym(g_nccl.lib, "ncclGroupEnd");
  g_nccl.errstr = (fn_errstr)dlsym(g_nccl.lib, "ncclGetErrorString");
  g_nccl.ok = g_nccl.get_uid && g_nccl.init_rank && g_nccl.destroy && g_nccl.allreduce && g_nccl.allgather && g_nccl.group_start && g_nccl.group_end;
  return g_nccl.ok;
}
constexpr int kNcclFloat64 = 8, kNcclInt8 = 0, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values
}  // namespace

int shard_unique_id(char out[128]) {
  if (!load_nccl()) return -1;
  UniqueId id;
  const int rc = ((fn_get_uid_t)g_nccl.get_uid)(&id);
  if (rc != 0) return rc;
  std::memcpy(out, id.internal, 128);
  return 0;
}

int shard_comm_init(ShardComm* sc, int rank, int world, const char id_bytes[128]) {
  if (!load_nccl()) return -1;
  UniqueId id;
  std::memcpy(id.internal, id_bytes, 128);
  void* comm = nullptr;
  const int rc = ((fn_init_rank_t)g_nccl.init_rank)(&comm, world, id, rank);
  if (rc != 0) return rc;
  sc->comm = comm; sc->rank = rank; sc->world = world;
  return 0;
}

void shard_comm_destroy(ShardComm* sc) {
  if (sc->comm && g_nccl.ok) g_nccl.destroy(sc->comm);
  sc->comm = nullptr; sc->world = 1; sc->rank = 0;
}

const char* shard_error_string(int rc) {
  if (rc == -1) return "NCCL (libnccl.so.2) could not be loaded";
  return g_nccl.errstr ? g_nccl.errstr(rc) : "NCCL error";
}

int shard_allreduce_f64(const ShardComm* sc, double* buf, size_t n, cudaStream_t s) {
  return g_nccl.allreduce(buf, buf, n, kNcclFloat64, kNcclSum, sc->comm, s);
}

// In-place all-gather of `bytes_per_rank` bytes: rank r contributes buf[r * bytes_per_rank ...).
int shard_allgather_bytes(const ShardComm* sc, void* buf, size_t bytes_per_rank, cudaStream_t s) {
  const char* send = static_cast<const char*>(buf) + (size_t)sc->rank * bytes_per_rank;
  return g_nccl.allgather(send, buf, bytes_per_rank, kNcclInt8, sc->comm, s);
}
int shard_group_start() { return g_nccl.group_start(); }
int shard_group_end() { return g_nccl.group_end(); }

}  // namespace liodom
