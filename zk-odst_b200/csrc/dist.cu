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

#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
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
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;                        // optional
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;  // optional
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
    a.Send = (decltype(a.Send))dlsym(a.lib, "ncclSend");
    a.Recv = (decltype(a.Recv))dlsym(a.lib, "ncclRecv");
    a.GroupStart = (decltype(a.GroupStart))dlsym(a.lib, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.lib, "ncclGroupEnd");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
    a.CommAbort = (decltype(a.CommAbort))dlsym(a.lib, "ncclCommAbort");
    a.CommGetAsyncError = (decltype(a.CommGetAsyncError))dlsym(a.lib, "ncclCommGetAsyncError");
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllGather && a.Send && a.Recv && a.GroupStart && a.GroupEnd &&
           a.CommDestroy && a.GetErrorString;
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

// Rows of the coset-major quotient domain (NUM_COSETS cosets of n rows) that evaluating rows [lo, hi) reads:
// a row reads itself and its rotations -1, +1 and -(blinding + 1) = -6 inside its own coset (quotient.cu),
// so per coset the local range [a, b) needs [a - 6, b + 1) modulo n.  Sorted, merged (start, length) pairs.
void dist_quotient_segments(uint64_t n, uint64_t lo, uint64_t hi, std::vector<std::pair<uint64_t, uint64_t>>& out) {
  constexpr int64_t BACK = 6, FWD = 1;
  std::vector<std::pair<uint64_t, uint64_t>> iv;  // [start, end) absolute
  const int64_t N = (int64_t)n;
  // local interval [s, e) of coset c (s may be negative, e may exceed n), wrapped into the coset
  auto add = [&](uint64_t c0, int64_t s_, int64_t e_) {
    if (e_ - s_ >= N) {
      iv.push_back({c0, c0 + n});
      return;
    }
    if (e_ <= 0) {
      s_ += N;
      e_ += N;
    } else if (s_ >= N) {
      s_ -= N;
      e_ -= N;
    }
    if (s_ < 0) {
      iv.push_back({c0 + (uint64_t)(N + s_), c0 + n});
      s_ = 0;
    }
    if (e_ > N) {
      iv.push_back({c0, c0 + (uint64_t)(e_ - N)});
      e_ = N;
    }
    if (s_ < e_) iv.push_back({c0 + (uint64_t)s_, c0 + (uint64_t)e_});
  };
  for (uint64_t c = 0; c * n < hi; c++) {
    const uint64_t c0 = c * n, c1 = c0 + n;
    if (c1 <= lo) continue;
    const int64_t a = (int64_t)((lo > c0 ? lo : c0) - c0), b = (int64_t)((hi < c1 ? hi : c1) - c0);
    if (a >= b) continue;
    add(c0, a - 1, b + FWD);     // the rows themselves and their rotations -1, +1
    add(c0, a - BACK, b - BACK);  // rotation -(blinding + 1)
  }
  std::sort(iv.begin(), iv.end());
  out.clear();
  for (auto& v : iv) {
    if (!out.empty() && v.first <= out.back().first + out.back().second) {
      const uint64_t end = v.second > out.back().first + out.back().second ? v.second : out.back().first + out.back().second;
      out.back().second = end - out.back().first;
    } else {
      out.push_back({v.first, v.second - v.first});
    }
  }
}

// Column-sharded coset arrays -> row-sharded quotient input.  `slots` holds nslots arrays of `en` rows each
// (en = NUM_COSETS * n); this rank has filled the whole arrays of its own column block and evaluates the
// quotient on rows [rank * en / world, (rank + 1) * en / world).  Every rank sends each peer the row segments
// that peer reads, of the columns it owns, and receives its own segments of the peers' columns in place:
// en / world rows (plus a 7-row halo per coset) of every column arrive, instead of every column in full.
int32_t dist_exchange_quotient_rows(zk_ctx* ctx, char* slots, size_t elem_bytes, uint64_t n, uint64_t en,
                                    uint32_t nslots) {
  const int world = ctx->dist_world, me = ctx->dist_rank;
  if (world <= 1) return ZK_OK;
  if (!ctx->nccl_comm) return set_error(ctx, ZK_E_STATE, "the context left its group after an error");
  const uint64_t rows = en / (uint64_t)world;
  std::vector<std::vector<std::pair<uint64_t, uint64_t>>> segs(world);
  for (int q = 0; q < world; q++) dist_quotient_segments(n, rows * q, rows * (q + 1), segs[q]);
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  ncclResult_t r = nccl().GroupStart();
  if (r != ncclSuccess) return nccl_error(ctx, r, "ncclGroupStart");
  for (int q = 0; q < world && r == ncclSuccess; q++) {
    if (q == me) continue;
    uint32_t lo, hi, per;
    dist_column_block(nslots, me, world, &lo, &hi, &per);
    for (uint32_t s = lo; s < hi && r == ncclSuccess; s++)
      for (auto& g : segs[q]) {
        r = nccl().Send(slots + ((size_t)s * en + g.first) * elem_bytes, g.second * elem_bytes, ncclUint8, q, comm,
                        ctx->stream);
        if (r != ncclSuccess) break;
      }
    dist_column_block(nslots, q, world, &lo, &hi, &per);
    for (uint32_t s = lo; s < hi && r == ncclSuccess; s++)
      for (auto& g : segs[me]) {
        r = nccl().Recv(slots + ((size_t)s * en + g.first) * elem_bytes, g.second * elem_bytes, ncclUint8, q, comm,
                        ctx->stream);
        if (r != ncclSuccess) break;
      }
  }
  const ncclResult_t e = nccl().GroupEnd();
  if (r != ncclSuccess) return nccl_error(ctx, r, "ncclSend/ncclRecv");
  if (e != ncclSuccess) return nccl_error(ctx, e, "ncclGroupEnd");
  return ZK_OK;
}

// Column-sharded slot arrays -> range-sharded: `slots` holds nslots arrays of n elements; the rank that owns slot s
// (dist_column_block) holds it in full and sends every peer that peer's range [q n / world, (q + 1) n / world) of it;
// every rank receives its own range of the peers' slots in place.  1 / world of an all-gather's traffic: what the
// evaluations and the multiopen argument read (they work by coefficient range).  Slots [skip_lo, skip_hi) are not sent.
int32_t dist_exchange_ranges(zk_ctx* ctx, char* slots, size_t elem_bytes, uint64_t n, uint32_t nslots, uint32_t skip_lo,
                             uint32_t skip_hi) {
  const int world = ctx->dist_world, me = ctx->dist_rank;
  if (world <= 1) return ZK_OK;
  if (!ctx->nccl_comm) return set_error(ctx, ZK_E_STATE, "the context left its group after an error");
  const uint64_t cnt = n / (uint64_t)world;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  ncclResult_t r = nccl().GroupStart();
  if (r != ncclSuccess) return nccl_error(ctx, r, "ncclGroupStart");
  for (int q = 0; q < world && r == ncclSuccess; q++) {
    if (q == me) continue;
    uint32_t lo, hi, per;
    dist_column_block(nslots, me, world, &lo, &hi, &per);
    for (uint32_t s = lo; s < hi && r == ncclSuccess; s++)
      if (s < skip_lo || s >= skip_hi)
        r = nccl().Send(slots + ((size_t)s * n + (size_t)q * cnt) * elem_bytes, cnt * elem_bytes, ncclUint8, q, comm,
                        ctx->stream);
    dist_column_block(nslots, q, world, &lo, &hi, &per);
    for (uint32_t s = lo; s < hi && r == ncclSuccess; s++)
      if (s < skip_lo || s >= skip_hi)
        r = nccl().Recv(slots + ((size_t)s * n + (size_t)me * cnt) * elem_bytes, cnt * elem_bytes, ncclUint8, q, comm,
                        ctx->stream);
  }
  const ncclResult_t e = nccl().GroupEnd();
  if (r != ncclSuccess) return nccl_error(ctx, r, "ncclSend/ncclRecv");
  if (e != ncclSuccess) return nccl_error(ctx, e, "ncclGroupEnd");
  return ZK_OK;
}

// The quotient h on NUM_COSETS cosets of n rows (coset-major, `h`), evaluated by row range rows_per_rank * rank: the
// rows of coset c are collected on rank c mod world (phase 0: before the inverse transforms, in place in `h`); after
// that rank transformed its cosets into `hc`, every rank receives its coefficient range of every coset (phase 1, in
// place in `hc`).  Only 1 / world of h ever moves twice, and no rank transforms more than ceil(cosets / world) cosets.
int32_t dist_exchange_h(zk_ctx* ctx, char* buf, size_t elem_bytes, uint64_t n, int ncosets, int phase) {
  const int world = ctx->dist_world, me = ctx->dist_rank;
  if (world <= 1) return ZK_OK;
  if (!ctx->nccl_comm) return set_error(ctx, ZK_E_STATE, "the context left its group after an error");
  const uint64_t en = (uint64_t)ncosets * n, rows = en / (uint64_t)world, cnt = n / (uint64_t)world;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  ncclResult_t r = nccl().GroupStart();
  if (r != ncclSuccess) return nccl_error(ctx, r, "ncclGroupStart");
  for (int c = 0; c < ncosets && r == ncclSuccess; c++) {
    const int owner = c % world;
    const uint64_t c0 = (uint64_t)c * n, c1 = c0 + n;
    for (int q = 0; q < world && r == ncclSuccess; q++) {
      if (q == owner) continue;
      if (phase == 0) {  // rows of coset c that rank q evaluated -> owner
        const uint64_t a = std::max<uint64_t>(c0, rows * q), b = std::min<uint64_t>(c1, rows * (q + 1));
        if (a >= b) continue;
        if (me == q) r = nccl().Send(buf + a * elem_bytes, (b - a) * elem_bytes, ncclUint8, owner, comm, ctx->stream);
        else if (me == owner) r = nccl().Recv(buf + a * elem_bytes, (b - a) * elem_bytes, ncclUint8, q, comm, ctx->stream);
      } else {  // coefficient range of rank q: owner -> q
        const uint64_t a = c0 + cnt * q;
        if (me == owner) r = nccl().Send(buf + a * elem_bytes, cnt * elem_bytes, ncclUint8, q, comm, ctx->stream);
        else if (me == q) r = nccl().Recv(buf + a * elem_bytes, cnt * elem_bytes, ncclUint8, owner, comm, ctx->stream);
      }
    }
  }
  const ncclResult_t e = nccl().GroupEnd();
  if (r != ncclSuccess) return nccl_error(ctx, r, "ncclSend/ncclRecv");
  if (e != ncclSuccess) return nccl_error(ctx, e, "ncclGroupEnd");
  return ZK_OK;
}

// results[m] <- sum over ranks of their results[m]; identical on every rank afterwards
int32_t dist_sum_points(zk_ctx* ctx, XYZZ* results, int nb) {
  if (ctx->dist_world <= 1) return ZK_OK;
  if (!ctx->nccl_comm) return set_error(ctx, ZK_E_STATE, "the context left its group after an error");
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

// host_out[i] <- sum over ranks of their d_vals[i] (device, `count` field elements); identical on every rank
int32_t dist_sum_fields(zk_ctx* ctx, const Fp* d_vals, int count, Fp* host_out) {
  const int world = ctx->dist_world;
  if (!ctx->nccl_comm) return set_error(ctx, ZK_E_STATE, "the context left its group after an error");
  const size_t bytes = (size_t)count * sizeof(Fp);
  int32_t rc = ensure_buf(ctx, ctx->dist_buf, bytes * world);
  if (rc) return rc;
  ncclResult_t r = nccl().AllGather(d_vals, ctx->dist_buf.ptr, bytes, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream);
  if (r != ncclSuccess) return nccl_error(ctx, r, "ncclAllGather");
  std::vector<Fp> all((size_t)count * world);
  ZK_CUDA(ctx, cudaMemcpyAsync(all.data(), ctx->dist_buf.ptr, bytes * world, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  for (int i = 0; i < count; i++) {
    Fp acc = all[i];
    for (int q = 1; q < world; q++) acc = acc + all[(size_t)q * count + i];
    host_out[i] = acc;
  }
  return ZK_OK;
}

// host_out[q * count + i] <- rank q's d_vals[i]: one small all-gather, identical on every rank
int32_t dist_gather_fields(zk_ctx* ctx, const Fp* d_vals, int count, Fp* host_out) {
  const int world = ctx->dist_world;
  if (!ctx->nccl_comm) return set_error(ctx, ZK_E_STATE, "the context left its group after an error");
  const size_t bytes = (size_t)count * sizeof(Fp);
  int32_t rc = ensure_buf(ctx, ctx->dist_buf, bytes * world);
  if (rc) return rc;
  ncclResult_t r = nccl().AllGather(d_vals, ctx->dist_buf.ptr, bytes, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream);
  if (r != ncclSuccess) return nccl_error(ctx, r, "ncclAllGather");
  ZK_CUDA(ctx, cudaMemcpyAsync(host_out, ctx->dist_buf.ptr, bytes * world, cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  return ZK_OK;
}

int32_t dist_allgather_device(zk_ctx* ctx, const void* send, void* recv, size_t bytes) {
  if (!ctx->nccl_comm) return set_error(ctx, ZK_E_STATE, "the context left its group after an error");
  ncclResult_t r = nccl().AllGather(send, recv, bytes, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream);
  if (r != ncclSuccess) return nccl_error(ctx, r, "ncclAllGather");
  return ZK_OK;
}

// Leaves the group after a local failure: the communicator is aborted (outstanding collectives of this rank
// are cancelled) so that a rank which returned early cannot sit in the group half-alive.  Its peers, blocked
// in the collectives it skipped, leave through the timed wait below.
void dist_abort(zk_ctx* ctx) {
  if (!ctx->nccl_comm) return;
  if (nccl().CommAbort) nccl().CommAbort((ncclComm_t)ctx->nccl_comm);
  else nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->dist_failed = true;
}

// Host wait of a context that belongs to a group: polls the stream instead of blocking in the driver, watches
// the communicator's asynchronous error state and gives up after ZK_DIST_TIMEOUT_S seconds (default 300), so
// a peer that failed and skipped its collectives ends this rank's call with ZK_E_CUDA instead of a deadlock.
cudaError_t dist_stream_sync(zk_ctx* ctx) {
  static const double limit_s = [] {
    const char* e = getenv("ZK_DIST_TIMEOUT_S");
    const double v = e ? atof(e) : 300.0;
    return v > 0 ? v : 300.0;
  }();
  const auto t0 = std::chrono::steady_clock::now();
  uint32_t spins = 0;
  for (;;) {
    const cudaError_t q = cudaStreamQuery(ctx->stream);
    if (q != cudaErrorNotReady) return q;
    if ((++spins & 0x3ff) == 0) {
      ncclResult_t async = ncclSuccess;
      if (ctx->nccl_comm && nccl().CommGetAsyncError &&
          nccl().CommGetAsyncError((ncclComm_t)ctx->nccl_comm, &async) == ncclSuccess && async != ncclSuccess &&
          async != ncclInProgress) {
        dist_abort(ctx);
        ctx->err = std::string("NCCL asynchronous error: ") + nccl().GetErrorString(async);
        return cudaErrorUnknown;
      }
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit_s) {
        dist_abort(ctx);
        ctx->err = "group wait timed out: a peer rank did not reach the collective (ZK_DIST_TIMEOUT_S)";
        return cudaErrorLaunchTimeout;
      }
      if (ctx->blocking_sync) usleep(50);
    }
  }
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

extern "C" int32_t zk_dist_quotient_rows(uint64_t n, int32_t rank, int32_t world, uint64_t* row_lo, uint64_t* row_hi,
                                         uint64_t* segments, uint32_t* n_segments) {
  if (!row_lo || !row_hi || !n_segments || world < 1 || rank < 0 || rank >= world || n == 0) return ZK_E_INVALID;
  const uint64_t en = 3 * n;  // NUM_COSETS cosets of the n-th roots (prover_state.h)
  if (en % (uint64_t)world) return ZK_E_INVALID;
  const uint64_t rows = en / (uint64_t)world;
  *row_lo = rows * (uint64_t)rank;
  *row_hi = *row_lo + rows;
  std::vector<std::pair<uint64_t, uint64_t>> segs;
  dist_quotient_segments(n, *row_lo, *row_hi, segs);
  const uint32_t cap = *n_segments;
  *n_segments = (uint32_t)segs.size();
  if (!segments || cap < segs.size()) return ZK_E_BUFFER;
  for (size_t i = 0; i < segs.size(); i++) {
    segments[2 * i] = segs[i].first;
    segments[2 * i + 1] = segs[i].second;
  }
  return ZK_OK;
}
