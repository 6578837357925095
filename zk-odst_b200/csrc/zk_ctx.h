// Context object behind the opaque `zk_ctx*` of include/zkodst.h.  Product code.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/zkodst.h"
#include "blake2f_layout.h"
#include "field.cuh"

namespace zkodst {

struct DeviceRegionLayout {
  RegionLayout host;
  uint32_t* d_desc = nullptr;  // [NUM_USED_COLUMNS][rows]
};

struct DevBuf {  // grow-only device scratch buffer
  void* ptr = nullptr;
  size_t cap = 0;
};

struct NttTables {  // per-domain twiddle tables (device) and host constants
  int log_n = 0;
  Fp omega, omega_inv, n_inv;
  Fp* tw_fwd = nullptr;  // omega^i, i < N/2
  Fp* tw_inv = nullptr;  // omega^-i
};

enum KernelClass {
  KC_WITNESS = 0, KC_MSM = 1, KC_NTT = 2, KC_QUOTIENT = 3, KC_COLLAPSE = 4, KC_MSM_ACC = 5, KC_COUNT = 8
};

}  // namespace zkodst

struct zk_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  bool timing = false;
  // timed regions since the last report: events come from a grow-only pool
  struct Timed {
    cudaEvent_t a, b;
    int which;
  };
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::vector<Timed> timed;
  float last_ms[zkodst::KC_COUNT] = {};
  bool last_valid[zkodst::KC_COUNT] = {};
  std::map<uint32_t, zkodst::DeviceRegionLayout> layouts;
  zkodst::DevBuf scratch_inputs, scratch_advice, scratch_digests;
  zkodst::DevBuf scratch_a, scratch_b, msm_ws, msm_out, ntt_tmp, scan_ws, eval_ws, misc_ws, inv_ws;
  void* prover_state = nullptr;  // zkodst::ProverState (params, keys), owned; see prover_state.h
  void (*prover_state_free)(void*) = nullptr;
  std::map<int, zkodst::NttTables> ntt_tables;
  int* d_status = nullptr;  // device-side error flag (bad EIP-152 record seen by a kernel)
  int sm_count = 148;
  int msm_acc_blocks_per_sm = 0;  // resident blocks of the MSM accumulation kernel (msm_fixed.cu), queried once
  // MSM split across GPUs (dist.cu): NCCL communicator (ncclComm_t), this rank, group size
  void* nccl_comm = nullptr;
  int dist_rank = 0, dist_world = 1;
  bool dist_failed = false;  // the context aborted its communicator after an error (dist.cu)
  zkodst::DevBuf dist_buf;
  // host waits: spin (cudaStreamSynchronize) or sleep on a blocking event (many contexts per core)
  bool blocking_sync = false;
  cudaEvent_t sync_event = nullptr;
  bool xs_table_loaded = false;  // XorShift jump matrices resident in misc_ws (prover.cu)
};

namespace zkodst {

int32_t set_error(zk_ctx* ctx, int32_t code, const std::string& msg);
int32_t check_cuda(zk_ctx* ctx, cudaError_t e, const char* what);
int32_t ensure_buf(zk_ctx* ctx, DevBuf& b, size_t bytes);
// waits for the context's stream; sleeps instead of spinning when ctx->blocking_sync is set
cudaError_t zk_stream_sync(zk_ctx* ctx);
int32_t get_layout(zk_ctx* ctx, uint32_t rounds, DeviceRegionLayout** out);

struct KernelTimer {  // CUDA-event bracket on the context's stream, only when timing is on
  zk_ctx* ctx;
  int which;
  cudaEvent_t a = nullptr, b = nullptr;
  KernelTimer(zk_ctx* c, int w) : ctx(c), which(w) {
    if (!ctx->timing) return;
    if (ctx->ev_used + 2 > ctx->ev_pool.size()) {
      ctx->ev_pool.resize(ctx->ev_pool.size() + 64, nullptr);
      for (size_t i = ctx->ev_pool.size() - 64; i < ctx->ev_pool.size(); i++) cudaEventCreate(&ctx->ev_pool[i]);
    }
    a = ctx->ev_pool[ctx->ev_used++];
    b = ctx->ev_pool[ctx->ev_used++];
    cudaEventRecord(a, ctx->stream);
  }
  ~KernelTimer() {
    if (!a) return;
    cudaEventRecord(b, ctx->stream);
    ctx->timed.push_back(zk_ctx::Timed{a, b, which});
  }
};

// Device temporaries of one call: freed when the holder goes out of scope (every early return included)
// unless released to a longer-lived owner.
struct DevTemps {
  std::vector<void**> slots;
  template <class T>
  void own(T** p) { slots.push_back((void**)p); }
  template <class T>
  T* release(T** p) {
    for (auto& s : slots)
      if (s == (void**)p) s = nullptr;
    return *p;
  }
  ~DevTemps() {
    for (auto s : slots)
      if (s && *s) {
        cudaFree(*s);
        *s = nullptr;
      }
  }
};

#define ZK_CUDA(ctx, call)                                        \
  do {                                                            \
    int32_t _rc = zkodst::check_cuda((ctx), (call), #call);       \
    if (_rc) return _rc;                                          \
  } while (0)

// dist.cu
struct XYZZ;
void dist_range(uint64_t n, int rank, int world, uint64_t* lo, uint64_t* hi);
void dist_column_block(uint32_t count, int rank, int world, uint32_t* lo, uint32_t* hi, uint32_t* per_rank);
int32_t dist_sum_points(zk_ctx* ctx, XYZZ* results, int nb);
// recv[r * bytes .. (r + 1) * bytes) <- rank r's send buffer (device pointers), on the context's stream
int32_t dist_allgather_device(zk_ctx* ctx, const void* send, void* recv, size_t bytes);
// column-sharded slot arrays -> the row segments each rank's share of the quotient reads (dist.cu)
int32_t dist_exchange_quotient_rows(zk_ctx* ctx, char* slots, size_t elem_bytes, uint64_t n, uint64_t en,
                                    uint32_t nslots);
// column-sharded slot arrays -> every rank's own element range of every slot (dist.cu)
int32_t dist_exchange_ranges(zk_ctx* ctx, char* slots, size_t elem_bytes, uint64_t n, uint32_t nslots, uint32_t skip_lo,
                             uint32_t skip_hi);
// rows of the quotient -> the rank that transforms their coset (phase 0); coefficient ranges back (phase 1)
int32_t dist_exchange_h(zk_ctx* ctx, char* buf, size_t elem_bytes, uint64_t n, int ncosets, int phase);
// host_out[i] <- sum over ranks of d_vals[i] (one small all-gather, summed on the host)
int32_t dist_sum_fields(zk_ctx* ctx, const Fp* d_vals, int count, Fp* host_out);
// host_out[q * count + i] <- rank q's d_vals[i] (device, `count` field elements)
int32_t dist_gather_fields(zk_ctx* ctx, const Fp* d_vals, int count, Fp* host_out);
void dist_free(zk_ctx* ctx);
// group contexts: leave the group after a local error; timed host wait that cannot deadlock on a failed peer
void dist_abort(zk_ctx* ctx);
cudaError_t dist_stream_sync(zk_ctx* ctx);

// witness.cu
int32_t launch_witness(zk_ctx* ctx, int32_t k, uint32_t rounds, const uint8_t* d_inputs,
                       uint64_t n_compressions, void* d_advice, uint64_t* d_digests);

}  // namespace zkodst
