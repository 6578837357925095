// Multi-GPU exchange for one MSM split by point range (BASELINE configs[3]).
//
// Every rank holds the window tables of its own contiguous range of the base points; an MSM is
// the sum of the per-rank partial results.  The only communication is one all-gather of the
// partial points (128 B XYZZ per MSM of the batch per rank) over NCCL — there is no group-operation
// reduction in NCCL, so the ranks sum the gathered partials themselves (world - 1 additions).
// NCCL is resolved with dlopen so that libzkodst.so loads without it on single-GPU hosts; inside a
// torch process the already-loaded libnccl.so.2 is reused.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <vector>

#include "ec.cuh"
#include "zk_ctx.h"

namespace zkodst {
namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
    a.AllGather = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllGather && a.CommDestroy && a.GetErrorString;
    return a;
  }();
  return api;
}

int32_t nccl_error(zk_ctx* ctx, ncclResult_t r, const char* what) {
  return set_error(ctx, ZK_E_CUDA, std::string(what) + ": " + nccl().GetErrorString(r));
}

}  // namespace

// [lo, hi) of rank r when n points are cut into `world` contiguous ranges
void dist_range(uint64_t n, int rank, int world, uint64_t* lo, uint64_t* hi) {
  *lo = n * (uint64_t)rank / (uint64_t)world;
  *hi = n * (uint64_t)(rank + 1) / (uint64_t)world;
}

// block of witness-column slots transformed by `rank`: [lo, hi) of `count`, per_rank slots per block
void dist_column_block(uint32_t count, int rank, int world, uint32_t* lo, uint32_t* hi, uint32_t* per_rank) {
  const uint32_t per = (count + (uint32_t)world - 1) / (uint32_t)world;
  const uint32_t first = per * (uint32_t)rank;
  *per_rank = per;
  *lo = first < count ? first : count;
  *hi = first + per < count ? first + per : count;
}

// results[m] <- sum over ranks of their results[m]; identical on every rank afterwards
int32_t dist_sum_points(zk_ctx* ctx, XYZZ* results, int nb) {
  if (ctx->dist_world <= 1) return ZK_OK;
  const size_t bytes = (size_t)nb * sizeof(XYZZ);
  const int world = ctx->dist_world;
  int32_t rc = ensure_buf(ctx, ctx->dist_buf, bytes * (world + 1));
  if (rc) return rc;
  char* send = (char*)ctx->dist_buf.ptr;
  char* recv = send + bytes;
  ZK_CUDA(ctx, cudaMemcpyAsync(send, results, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ncclResult_t r = nccl().AllGather(send, recv, bytes, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream);
  if (r != ncclSuccess) return nccl_error(ctx, r, "ncclAllGather");
  std::vector<XYZZ> all((size_t)nb * world);
  ZK_CUDA(ctx, cudaMemcpyAsync(all.data(), recv, bytes * world, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  for (int m = 0; m < nb; m++) {
    XYZZ acc = all[m];
    for (int q = 1; q < world; q++) acc = acc.add(all[(size_t)q * nb + m]);
    results[m] = acc;
  }
  return ZK_OK;
}

int32_t dist_allgather_device(zk_ctx* ctx, const void* send, void* recv, size_t bytes) {
  ncclResult_t r = nccl().AllGather(send, recv, bytes, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream);
  if (r != ncclSuccess) return nccl_error(ctx, r, "ncclAllGather");
  return ZK_OK;
}

void dist_free(zk_ctx* ctx) {
  if (ctx->nccl_comm) {
    nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  cudaFree(ctx->dist_buf.ptr);
  ctx->dist_buf = DevBuf();
  ctx->dist_rank = 0;
  ctx->dist_world = 1;
}

}  // namespace zkodst

using namespace zkodst;

extern "C" int32_t zk_dist_unique_id(uint8_t out[ZK_DIST_ID_BYTES]) {
  if (!out) return ZK_E_INVALID;
  static_assert(sizeof(ncclUniqueId) == ZK_DIST_ID_BYTES, "ncclUniqueId size");
  if (!nccl().ok) return ZK_E_CUDA;
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) return ZK_E_CUDA;
  memcpy(out, &id, sizeof id);
  return ZK_OK;
}

extern "C" int32_t zk_dist_init(zk_ctx* ctx, const uint8_t id_bytes[ZK_DIST_ID_BYTES], int32_t rank, int32_t world) {
  if (!ctx || !id_bytes || world < 1 || rank < 0 || rank >= world) return ZK_E_INVALID;
  if (ctx->prover_state) return set_error(ctx, ZK_E_STATE, "zk_dist_init must precede params and keys");
  if (ctx->nccl_comm) return set_error(ctx, ZK_E_STATE, "context already joined a group");
  if (world == 1) return ZK_OK;
  if (!nccl().ok) return set_error(ctx, ZK_E_CUDA, "libnccl.so.2 not found");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof id);
  ncclComm_t comm = nullptr;
  ncclResult_t r = nccl().CommInitRank(&comm, world, id, rank);
  if (r != ncclSuccess) return nccl_error(ctx, r, "ncclCommInitRank");
  ctx->nccl_comm = comm;
  ctx->dist_rank = rank;
  ctx->dist_world = world;
  return ZK_OK;
}

extern "C" int32_t zk_dist_info(const zk_ctx* ctx, int32_t* rank, int32_t* world) {
  if (!ctx || !rank || !world) return ZK_E_INVALID;
  *rank = ctx->dist_rank;
  *world = ctx->dist_world;
  return ZK_OK;
}

extern "C" int32_t zk_dist_range(uint64_t n_points, int32_t rank, int32_t world, uint64_t* lo, uint64_t* hi) {
  if (!lo || !hi || world < 1 || rank < 0 || rank >= world) return ZK_E_INVALID;
  dist_range(n_points, rank, world, lo, hi);
  return ZK_OK;
}

extern "C" int32_t zk_dist_column_block(int32_t rank, int32_t world, uint32_t* lo, uint32_t* hi, uint32_t* per_rank) {
  if (!lo || !hi || !per_rank || world < 1 || rank < 0 || rank >= world) return ZK_E_INVALID;
  dist_column_block(ZK_NUM_WITNESS_COLUMNS, rank, world, lo, hi, per_rank);
  return ZK_OK;
}
