// extern "C" entry points of include/zkodst.h: context management and the K1 witness path.
#include <cstdlib>
#include <cstring>
#include <new>

#include "zk_ctx.h"

namespace zkodst {

int32_t set_error(zk_ctx* ctx, int32_t code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

int32_t check_cuda(zk_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return ZK_OK;
  std::string msg = std::string(what) + ": " + cudaGetErrorString(e);
  if (ctx && ctx->dist_failed && !ctx->err.empty()) msg += " [" + ctx->err + "]";  // reason given by dist_stream_sync
  return set_error(ctx, e == cudaErrorMemoryAllocation ? ZK_E_NOMEM : ZK_E_CUDA, msg);
}

cudaError_t zk_stream_sync(zk_ctx* ctx) {
  if (ctx->nccl_comm) return dist_stream_sync(ctx);
  if (!ctx->blocking_sync) return cudaStreamSynchronize(ctx->stream);
  if (!ctx->sync_event) {
    cudaError_t e = cudaEventCreateWithFlags(&ctx->sync_event, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaEventRecord(ctx->sync_event, ctx->stream);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ctx->sync_event);
}

int32_t ensure_buf(zk_ctx* ctx, DevBuf& b, size_t bytes) {
  if (b.cap >= bytes) return ZK_OK;
  if (b.ptr) {
    ZK_CUDA(ctx, zk_stream_sync(ctx));
    ZK_CUDA(ctx, cudaFree(b.ptr));
    b.ptr = nullptr;
    b.cap = 0;
  }
  ZK_CUDA(ctx, cudaMalloc(&b.ptr, bytes));
  b.cap = bytes;
  return ZK_OK;
}

int32_t get_layout(zk_ctx* ctx, uint32_t rounds, DeviceRegionLayout** out) {
  auto it = ctx->layouts.find(rounds);
  if (it == ctx->layouts.end()) {
    DeviceRegionLayout L;
    try {
      build_region_layout(rounds, L.host);
    } catch (std::exception& e) {
      return set_error(ctx, ZK_E_INVALID, e.what());
    }
    size_t bytes = L.host.desc.size() * sizeof(uint32_t);
    DevTemps tmp;
    tmp.own(&L.d_desc);
    ZK_CUDA(ctx, cudaMalloc((void**)&L.d_desc, bytes));
    ZK_CUDA(ctx, cudaMemcpyAsync(L.d_desc, L.host.desc.data(), bytes, cudaMemcpyHostToDevice,
                                 ctx->stream));
    ZK_CUDA(ctx, zk_stream_sync(ctx));
    tmp.release(&L.d_desc);
    it = ctx->layouts.emplace(rounds, std::move(L)).first;
  }
  *out = &it->second;
  return ZK_OK;
}

}  // namespace zkodst

using namespace zkodst;

extern "C" {

int32_t zk_ctx_create(int32_t device_id, zk_ctx** out) {
  if (!out) return ZK_E_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return ZK_E_CUDA;
  if (device_id < 0 || device_id >= count) return ZK_E_INVALID;
  zk_ctx* ctx = new (std::nothrow) zk_ctx();
  if (!ctx) return ZK_E_NOMEM;
  ctx->device = device_id;
  if (cudaSetDevice(device_id) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return ZK_E_CUDA;
  }
  ctx->stream = ctx->own_stream;
  {
    const char* e = getenv("ZK_BLOCKING_SYNC");
    ctx->blocking_sync = e && e[0] == '1';
  }
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device_id);
  if (cudaMalloc((void**)&ctx->d_status, sizeof(int)) != cudaSuccess ||
      cudaMemset(ctx->d_status, 0, sizeof(int)) != cudaSuccess) {
    zk_ctx_destroy(ctx);
    return ZK_E_CUDA;
  }
  *out = ctx;
  return ZK_OK;
}

void zk_ctx_destroy(zk_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->layouts) cudaFree(kv.second.d_desc);
  cudaFree(ctx->scratch_inputs.ptr);
  cudaFree(ctx->scratch_advice.ptr);
  cudaFree(ctx->scratch_digests.ptr);
  cudaFree(ctx->scratch_a.ptr);
  cudaFree(ctx->scratch_b.ptr);
  cudaFree(ctx->msm_ws.ptr);
  cudaFree(ctx->msm_out.ptr);
  cudaFree(ctx->ntt_tmp.ptr);
  cudaFree(ctx->scan_ws.ptr);
  cudaFree(ctx->eval_ws.ptr);
  cudaFree(ctx->misc_ws.ptr);
  cudaFree(ctx->inv_ws.ptr);
  dist_free(ctx);
  if (ctx->prover_state && ctx->prover_state_free) ctx->prover_state_free(ctx->prover_state);
  for (auto& kv : ctx->ntt_tables) {
    cudaFree(kv.second.tw_fwd);
    cudaFree(kv.second.tw_inv);
  }
  cudaFree(ctx->d_status);
  if (ctx->sync_event) cudaEventDestroy(ctx->sync_event);
  for (auto& e : ctx->ev_pool)
    if (e) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* zk_last_error(const zk_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int32_t zk_ctx_set_stream(zk_ctx* ctx, void* cuda_stream) {
  if (!ctx) return ZK_E_INVALID;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return ZK_OK;
}

int32_t zk_ctx_set_blocking_sync(zk_ctx* ctx, int32_t on) {
  if (!ctx) return ZK_E_INVALID;
  ctx->blocking_sync = on != 0;
  return ZK_OK;
}

int32_t zk_ctx_synchronize(zk_ctx* ctx) {
  if (!ctx) return ZK_E_INVALID;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  int st = 0;
  ZK_CUDA(ctx, cudaMemcpy(&st, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost));
  if (st) {
    cudaMemset(ctx->d_status, 0, sizeof(int));
    return set_error(ctx, ZK_E_INPUT, "a kernel rejected an EIP-152 record (f or rounds)");
  }
  return ZK_OK;
}

uint64_t zk_ctx_launch_count(const zk_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t zk_ctx_enable_timing(zk_ctx* ctx, int32_t on) {
  if (!ctx) return ZK_E_INVALID;
  ctx->timing = on != 0;
  return ZK_OK;
}

// Sums the timed regions recorded since the previous report, per kernel class.
int32_t zk_ctx_timing_report(zk_ctx* ctx, float* ms_per_class, uint32_t* launches_per_class) {
  if (!ctx) return ZK_E_INVALID;
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  float acc[KC_COUNT] = {};
  uint32_t cnt[KC_COUNT] = {};
  for (auto& t : ctx->timed) {
    float ms = 0;
    ZK_CUDA(ctx, cudaEventElapsedTime(&ms, t.a, t.b));
    acc[t.which] += ms;
    cnt[t.which]++;
    ctx->last_ms[t.which] = ms;
    ctx->last_valid[t.which] = true;
  }
  ctx->timed.clear();
  ctx->ev_used = 0;
  for (int i = 0; i < KC_COUNT; i++) {
    if (ms_per_class) ms_per_class[i] = acc[i];
    if (launches_per_class) launches_per_class[i] = cnt[i];
  }
  return ZK_OK;
}

int32_t zk_ctx_last_kernel_ms(zk_ctx* ctx, int32_t which, float* ms) {
  if (!ctx || !ms || which < 0 || which >= KC_COUNT) return ZK_E_INVALID;
  int32_t rc = zk_ctx_timing_report(ctx, nullptr, nullptr);
  if (rc) return rc;
  if (!ctx->last_valid[which]) return set_error(ctx, ZK_E_STATE, "kernel class not timed yet");
  *ms = ctx->last_ms[which];
  return ZK_OK;
}

int32_t zk_blake2f_rows_per_compression(uint32_t rounds, uint64_t* rows) {
  if (!rows) return ZK_E_INVALID;
  *rows = region_rows(rounds);
  return ZK_OK;
}

int32_t zk_blake2f_min_k(uint32_t rounds, uint64_t n_compressions, int32_t* k) {
  if (!k) return ZK_E_INVALID;
  unsigned __int128 need = (unsigned __int128)region_rows(rounds) * n_compressions;
  for (int kk = 17; kk <= 28; kk++) {
    if (need <= (((unsigned __int128)1 << kk) - 6)) {
      *k = kk;
      return ZK_OK;
    }
  }
  return ZK_E_ROWS;
}

static uint64_t fnv1a(uint64_t h, const void* data, size_t len) {
  const uint8_t* p = (const uint8_t*)data;
  for (size_t i = 0; i < len; i++) {
    h ^= p[i];
    h *= 0x100000001b3ULL;
  }
  return h;
}

int32_t zk_blake2f_layout_hash(uint32_t rounds, uint64_t* copies_hash, uint64_t* selectors_hash,
                               uint64_t* n_copies) {
  if (!copies_hash || !selectors_hash || !n_copies) return ZK_E_INVALID;
  RegionLayout L;
  try {
    build_region_layout(rounds, L);
  } catch (std::exception&) {
    return ZK_E_INVALID;
  }
  uint64_t h = 0xcbf29ce484222325ULL;
  for (auto& c : L.copies) {
    uint32_t rec[4] = {c.left_col, c.left_row, c.right_col, c.right_row};
    h = fnv1a(h, rec, sizeof rec);
  }
  *copies_hash = h;
  *selectors_hash = fnv1a(0xcbf29ce484222325ULL, L.selectors.data(), L.selectors.size());
  *n_copies = L.copies.size();
  return ZK_OK;
}

// The region's layout tables themselves, for a host-language `Circuit::synthesize` that must issue the same
// `constrain_equal` / `Selector::enable` / `assign_fixed` calls as the library's keygen assumes.
int32_t zk_blake2f_layout_tables(uint32_t rounds, uint32_t* copies, uint64_t* n_copies, uint8_t* selectors,
                                 uint64_t* constants, uint32_t* chain_rows) {
  if (!n_copies) return ZK_E_INVALID;
  RegionLayout L;
  try {
    build_region_layout(rounds, L);
  } catch (std::exception&) {
    return ZK_E_INVALID;
  }
  const uint64_t cap = *n_copies;
  *n_copies = L.copies.size();
  if (copies) {
    if (cap < L.copies.size()) return ZK_E_BUFFER;
    for (size_t i = 0; i < L.copies.size(); i++) {
      const CopyConstraint& c = L.copies[i];
      copies[4 * i] = c.left_col, copies[4 * i + 1] = c.left_row, copies[4 * i + 2] = c.right_col,
                 copies[4 * i + 3] = c.right_row;
    }
  }
  if (selectors) memcpy(selectors, L.selectors.data(), L.selectors.size());
  if (constants) memcpy(constants, L.constants.data(), L.constants.size() * 8);
  if (chain_rows)
    for (int i = 0; i < 8; i++) chain_rows[i] = L.h_word_row[i], chain_rows[8 + i] = L.out_word_row[i];
  return ZK_OK;
}

int32_t zk_blake2f_witness_batch_device(zk_ctx* ctx, int32_t k, uint32_t rounds,
                                        const uint8_t* d_inputs, uint64_t n_compressions,
                                        void* d_advice, uint64_t* d_digests) {
  if (!ctx) return ZK_E_INVALID;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  return launch_witness(ctx, k, rounds, d_inputs, n_compressions, d_advice, d_digests);
}

int32_t zk_blake2f_witness_batch(zk_ctx* ctx, int32_t k, uint32_t rounds, const uint8_t* inputs,
                                 uint64_t n_compressions, void* advice_out, uint64_t* digests_out) {
  if (!ctx) return ZK_E_INVALID;
  if (k < 17 || k > 28) return set_error(ctx, ZK_E_INVALID, "k out of range [17, 28]");
  if (!inputs && n_compressions) return set_error(ctx, ZK_E_INVALID, "null inputs");
  // the row check of launch_witness, before the records are read or any size is multiplied out
  if (n_compressions > ((1ull << k) - 6) / region_rows(rounds))
    return set_error(ctx, ZK_E_ROWS, "compressions do not fit in 2^k rows");
  // EIP-152 rejection cases, checked before any device work
  for (uint64_t i = 0; i < n_compressions; i++) {
    const uint8_t* r = inputs + i * ZK_BLAKE2F_INPUT_BYTES;
    uint32_t rr = ((uint32_t)r[0] << 24) | ((uint32_t)r[1] << 16) | ((uint32_t)r[2] << 8) | r[3];
    if (r[212] > 1) return set_error(ctx, ZK_E_INPUT, "final-block flag must be 0 or 1");
    if (rr != rounds) return set_error(ctx, ZK_E_INPUT, "record rounds differ from circuit rounds");
  }
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t n = 1ull << k;
  const size_t advice_bytes = (size_t)ZK_NUM_ADVICE * n * ZK_FIELD_BYTES;
  int32_t rc;
  if ((rc = ensure_buf(ctx, ctx->scratch_inputs, n_compressions * ZK_BLAKE2F_INPUT_BYTES + 16))) return rc;
  if ((rc = ensure_buf(ctx, ctx->scratch_advice, advice_bytes))) return rc;
  if ((rc = ensure_buf(ctx, ctx->scratch_digests, n_compressions * 64 + 16))) return rc;
  if (n_compressions)
    ZK_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_inputs.ptr, inputs,
                                 n_compressions * ZK_BLAKE2F_INPUT_BYTES, cudaMemcpyHostToDevice,
                                 ctx->stream));
  rc = launch_witness(ctx, k, rounds, (const uint8_t*)ctx->scratch_inputs.ptr, n_compressions,
                      ctx->scratch_advice.ptr, (uint64_t*)ctx->scratch_digests.ptr);
  if (rc) return rc;
  if (advice_out)
    ZK_CUDA(ctx, cudaMemcpyAsync(advice_out, ctx->scratch_advice.ptr, advice_bytes,
                                 cudaMemcpyDeviceToHost, ctx->stream));
  if (digests_out && n_compressions)
    ZK_CUDA(ctx, cudaMemcpyAsync(digests_out, ctx->scratch_digests.ptr, n_compressions * 64,
                                 cudaMemcpyDeviceToHost, ctx->stream));
  return zk_ctx_synchronize(ctx);
}

}  // extern "C"
