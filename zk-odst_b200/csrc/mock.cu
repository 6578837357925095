// MockProver-equivalent constraint check on the device.
//
// Replaces halo2_proofs 0.3.0 `dev::MockProver::run(k, &circuit, vec![]).verify()` as the reference
// calls it (blake2f-circuit/src/blake2f/table16/spread_table.rs:759-763) for the BLAKE2f Table16
// circuit: every gate polynomial must vanish on every usable row (gates.cuh), every lookup input
// triple (tag, dense, spread) must be a row of the spread table (spread_table.rs:425-467), and the
// two cells of every copy constraint must be equal.  The first failure in (kind, row, index) order
// is reported, like MockProver's sorted error list.
#include "gates.cuh"
#include "prover_state.h"

namespace zkodst {
namespace {

struct MockArgs {
  const Fp* advice[NUM_ADVICE_COLUMNS];  // values, n rows each
  const Fp* fixed[NUM_FIXED];
  SelectorExpr sel[NUM_SELECTORS];
  GateConsts k;
};

struct Checker {  // accumulator policy for fold_gates: remembers the first non-zero polynomial
  int idx = 0, bad = -1;
  __device__ __forceinline__ void fold(const Fp& v) {
    if (bad < 0 && !v.is_zero()) bad = idx;
    idx++;
  }
};

// failure key: kind << 56 | row << 16 | index  (smaller = reported first)
__device__ __forceinline__ void report(unsigned long long* first, uint64_t kind, uint64_t row, uint64_t index) {
  atomicMin(first, (unsigned long long)((kind << 56) | (row << 16) | (index & 0xffff)));
}

__global__ void __launch_bounds__(128) mock_rows_kernel(const __grid_constant__ MockArgs ma, uint64_t n, uint64_t usable,
                                                        unsigned long long* first) {
  const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (r >= usable) return;
  const uint64_t rp = (r + n - 1) & (n - 1), rn = (r + 1) & (n - 1);
  // a0..a9 -> halo2 columns 7,8,9,1,2,0,3,4,5,6 (A_NUMBER_COLUMN)
  const Fp* A[10] = {ma.advice[7], ma.advice[8], ma.advice[9], ma.advice[1], ma.advice[2],
                     ma.advice[0], ma.advice[3], ma.advice[4], ma.advice[5], ma.advice[6]};
  GateCells v;
  v.a0c = A[0][r], v.a0n = A[0][rn];
  v.a1p = A[1][rp], v.a1c = A[1][r], v.a1n = A[1][rn];
  v.a2p = A[2][rp], v.a2c = A[2][r], v.a2n = A[2][rn];
  v.a3p = A[3][rp], v.a3c = A[3][r], v.a3n = A[3][rn];
  v.a4p = A[4][rp], v.a4c = A[4][r], v.a4n = A[4][rn];
  v.a5p = A[5][rp], v.a5c = A[5][r], v.a5n = A[5][rn];
  v.a6p = A[6][rp], v.a6c = A[6][r];
  v.a7p = A[7][rp], v.a7c = A[7][r];
  v.a8p = A[8][rp], v.a8c = A[8][r];
  v.a9c = A[9][r];
  Fp sel[NUM_SELECTORS];
  bool any = false;
#pragma unroll
  for (int s = 0; s < NUM_SELECTORS; s++) {
    sel[s] = selector_expr(ma.fixed[ma.sel[s].fixed_col][r], ma.sel[s].root, ma.sel[s].len, ma.k.small);
    any |= !sel[s].is_zero();
  }
  if (any) {
    Checker c;
    fold_gates(c, v, sel, ma.k, ma.fixed[FIXED_CONSTANTS][r]);
    if (c.bad >= 0) report(first, 1, r, (uint64_t)c.bad);
  }
  // lookup: (a0, a1, a2) must be table row `dense`
  uint64_t d[4];
  v.a1c.to_canonical(d);
  const bool in_table = !(d[1] | d[2] | d[3]) && d[0] < 65536 && v.a0c == ma.fixed[0][d[0]] &&
                        v.a1c == ma.fixed[1][d[0]] && v.a2c == ma.fixed[2][d[0]];
  if (!in_table) report(first, 2, r, 0);
}

struct DevCopy {
  uint32_t lcol, lrow, rcol, rrow;
};
__global__ void mock_copies_kernel(const __grid_constant__ MockArgs ma, const DevCopy* __restrict__ copies, uint32_t ncopies, uint64_t n_regions,
                                   uint64_t region_rows, unsigned long long* first) {
  const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (t >= n_regions * ncopies) return;
  const uint64_t region = t / ncopies;
  const uint32_t i = (uint32_t)(t % ncopies);
  const DevCopy c = copies[i];
  const uint64_t lr = region * region_rows + c.lrow, rr = region * region_rows + c.rrow;
  if (ma.advice[c.lcol][lr] != ma.advice[c.rcol][rr]) report(first, 3, lr, i);
}

// chaining copies (absolute rows): h_i of a continuing compression == h'_i of its predecessor
__global__ void mock_chain_kernel(const __grid_constant__ MockArgs ma, const DevCopy* __restrict__ copies, uint32_t ncopies,
                                  unsigned long long* first) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncopies) return;
  const DevCopy c = copies[t];
  if (ma.advice[c.lcol][c.lrow] != ma.advice[c.rcol][c.rrow]) report(first, 3, c.lrow, 0xffff);
}

}  // namespace
}  // namespace zkodst

using namespace zkodst;

extern "C" int32_t zk_mock_verify(zk_ctx* ctx, const uint8_t* inputs, uint64_t n_compressions,
                                  const void* advice_override, uint64_t failure[3]) {
  if (!ctx) return ZK_E_INVALID;
  ProverState* S = prover_state(ctx);
  if (!S->has_params || !S->has_keys) return set_error(ctx, ZK_E_STATE, "mock_verify before params/keygen");
  const DeviceKeys& K = S->keys;
  if (n_compressions != K.n_compressions) return set_error(ctx, ZK_E_INVALID, "batch size differs from keygen");
  if (!advice_override && n_compressions && !inputs) return ZK_E_INVALID;
  // the EIP-152 rejection cases zk_create_proof applies (prover.cu): a record with f > 1 or a foreign round
  // count would otherwise yield a self-consistent witness of a different statement
  for (uint64_t i = 0; i < n_compressions && !advice_override; i++) {
    const uint8_t* r = inputs + i * ZK_BLAKE2F_INPUT_BYTES;
    const uint32_t rr = ((uint32_t)r[0] << 24) | ((uint32_t)r[1] << 16) | ((uint32_t)r[2] << 8) | r[3];
    if (r[212] > 1) return set_error(ctx, ZK_E_INPUT, "final-block flag must be 0 or 1");
    if (rr != K.rounds) return set_error(ctx, ZK_E_INPUT, "record rounds differ from circuit rounds");
  }
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  ZK_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(int), st));  // a status bit can only come from this call
  const uint64_t n = K.n, usable = n - (BLINDING + 1);
  int32_t rc = ensure_buf(ctx, ctx->scratch_advice, (size_t)NUM_ADVICE_COLUMNS * n * sizeof(Fp));
  if (rc) return rc;
  Fp* adv = (Fp*)ctx->scratch_advice.ptr;
  if (advice_override) {
    ZK_CUDA(ctx, cudaMemcpyAsync(adv, advice_override, (size_t)NUM_ADVICE_COLUMNS * n * sizeof(Fp),
                                 cudaMemcpyHostToDevice, st));
  } else {
    rc = ensure_buf(ctx, ctx->scratch_inputs, n_compressions * ZK_BLAKE2F_INPUT_BYTES + 16);
    if (rc) return rc;
    if (n_compressions)
      ZK_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_inputs.ptr, inputs, n_compressions * ZK_BLAKE2F_INPUT_BYTES,
                                   cudaMemcpyHostToDevice, st));
    rc = launch_witness(ctx, K.k, K.rounds, (const uint8_t*)ctx->scratch_inputs.ptr, n_compressions, adv, nullptr);
    if (rc) return rc;
  }
  MockArgs args;
  for (int c = 0; c < NUM_ADVICE_COLUMNS; c++) args.advice[c] = adv + (size_t)c * n;
  for (int c = 0; c < NUM_FIXED; c++) args.fixed[c] = K.fixed_values[c];
  for (int s = 0; s < NUM_SELECTORS; s++) args.sel[s] = K.selectors[s];
  for (int i = 0; i < 4; i++) args.k.small[i] = Fp::from_u64(i);
  args.k.pow2[0] = Fp::one();
  for (int e = 1; e < 127; e++) args.k.pow2[e] = args.k.pow2[e - 1].dbl();
  DeviceRegionLayout* L = nullptr;
  if ((rc = get_layout(ctx, K.rounds, &L))) return rc;
  const uint32_t ncopies = (uint32_t)L->host.copies.size();
  std::vector<DevCopy> hc(ncopies);
  for (uint32_t i = 0; i < ncopies; i++) {
    const CopyConstraint& c = L->host.copies[i];
    hc[i] = DevCopy{c.left_col, c.left_row, c.right_col, c.right_row};
  }
  rc = ensure_buf(ctx, ctx->scratch_a, (size_t)ncopies * sizeof(DevCopy) + 64);
  if (rc) return rc;
  unsigned long long* d_first = (unsigned long long*)ctx->scratch_a.ptr;
  DevCopy* d_copies = (DevCopy*)((char*)ctx->scratch_a.ptr + 64);
  ZK_CUDA(ctx, cudaMemsetAsync(d_first, 0xff, 8, st));
  if (ncopies) ZK_CUDA(ctx, cudaMemcpyAsync(d_copies, hc.data(), (size_t)ncopies * sizeof(DevCopy), cudaMemcpyHostToDevice, st));
  mock_rows_kernel<<<(unsigned)((usable + 127) / 128), 128, 0, st>>>(args, n, usable, d_first);
  ctx->launches++;
  const uint64_t total = n_compressions * ncopies;
  if (total) {
    mock_copies_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(args, d_copies, ncopies, n_compressions, K.region_rows,
                                                                       d_first);
    ctx->launches++;
  }
  {  // record chaining (zk_blake2f_keygen_chained): index 0xffff in the failure report
    std::vector<DevCopy> cc;
    for (uint64_t j = 1; j < n_compressions; j++)
      if (K.chain[j])
        for (int i = 0; i < 8; i++)
          cc.push_back(DevCopy{CHAIN_OUT_COLUMN, (uint32_t)((j - 1) * K.region_rows + L->host.out_word_row[i]),
                               CHAIN_H_COLUMN, (uint32_t)(j * K.region_rows + L->host.h_word_row[i])});
    if (!cc.empty()) {
      rc = ensure_buf(ctx, ctx->scratch_b, cc.size() * sizeof(DevCopy));
      if (rc) return rc;
      ZK_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_b.ptr, cc.data(), cc.size() * sizeof(DevCopy), cudaMemcpyHostToDevice, st));
      mock_chain_kernel<<<(unsigned)((cc.size() + 127) / 128), 128, 0, st>>>(args, (const DevCopy*)ctx->scratch_b.ptr,
                                                                           (uint32_t)cc.size(), d_first);
      ctx->launches++;
      ZK_CUDA(ctx, zk_stream_sync(ctx));  // cc leaves scope
    }
  }
  ZK_CUDA(ctx, cudaGetLastError());
  unsigned long long first = 0;
  int status = 0;
  ZK_CUDA(ctx, cudaMemcpyAsync(&first, d_first, 8, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(ctx, cudaMemcpyAsync(&status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  if (status) {
    cudaMemsetAsync(ctx->d_status, 0, sizeof(int), st);
    return set_error(ctx, ZK_E_INPUT, "EIP-152 record rejected by the witness kernel (final flag or round count)");
  }
  if (first == ~0ull) {
    if (failure) failure[0] = failure[1] = failure[2] = 0;
    return ZK_OK;
  }
  const uint64_t kind = first >> 56, row = (first >> 16) & ((1ull << 40) - 1), index = first & 0xffff;
  if (failure) {
    failure[0] = kind;
    failure[1] = row;
    failure[2] = index;
  }
  static const char* names[] = {"", "gate polynomial not satisfied", "lookup input not in the spread table",
                                "copy constraint not satisfied"};
  char msg[160];
  snprintf(msg, sizeof msg, "%s: row %llu, index %llu", names[kind < 4 ? kind : 0], (unsigned long long)row,
           (unsigned long long)index);
  return set_error(ctx, ZK_E_VERIFY, msg);
}
