// Kernel-level entry points of include/zkodst.h over the context's params and keys: the batched
// fixed-base commitment pipeline (msm_fixed.cu), batched transforms and the coefficient -> coset
// evaluation (ntt.cu) exactly as create_proof drives them, so that each can be compared with the CPU
// oracle on its own (tests/test_commit_gpu.py) instead of only through whole proofs.
//
// Replaces halo2_proofs 0.3.0 `Params::commit` / `Params::commit_lagrange` (-> `best_multiexp`),
// `EvaluationDomain::{lagrange_to_coeff, coeff_to_extended}` (-> `best_fft`) as reached from
// `create_proof` (blake2f-circuit/benches/blake2f.rs:125).
#include <vector>

#include "prover_state.h"

using namespace zkodst;

// out[m] = sum_t scalars[m][t] * base_t + blinds[m] * W for m < ncols, one batched pipeline per
// MSM_MAX_BATCH columns.  basis 0 = g (coefficient form, Params::commit), 1 = g_lagrange
// (Params::commit_lagrange).  index_mask != 0 keeps only the terms t with ((t & index_mask) != 0) ==
// (index_select != 0) — the support of the L / R vectors of an IPA round.
extern "C" int32_t zk_commit_batch(zk_ctx* ctx, int32_t basis, const void* scalars, uint32_t ncols,
                                   const void* blinds, uint32_t index_mask, int32_t index_select,
                                   int32_t on_device, void* out_affine) {
  if (!ctx || !out_affine || !blinds || (ncols && !scalars) || basis < 0 || basis > 1) return ZK_E_INVALID;
  ProverState* S = prover_state(ctx);
  if (!S->has_params) return set_error(ctx, ZK_E_STATE, "zk_commit_batch before params");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  const DeviceParams& P = S->params;
  const FixedBase& fb = basis == 0 ? P.fb_g : P.fb_gl;
  const uint64_t n = P.n;
  const Fp* d = (const Fp*)scalars;
  if (!on_device && ncols) {
    const size_t bytes = (size_t)ncols * n * sizeof(Fp);
    int32_t rc = ensure_buf(ctx, ctx->scratch_b, bytes);
    if (rc) return rc;
    ZK_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_b.ptr, scalars, bytes, cudaMemcpyHostToDevice, ctx->stream));
    d = (const Fp*)ctx->scratch_b.ptr;
  }
  const Fp* bl = (const Fp*)blinds;
  Affine* out = (Affine*)out_affine;
  for (uint32_t first = 0; first < ncols; first += MSM_MAX_BATCH) {
    const int nb = (int)(ncols - first < (uint32_t)MSM_MAX_BATCH ? ncols - first : MSM_MAX_BATCH);
    MsmJob jobs[MSM_MAX_BATCH];
    XYZZ r[MSM_MAX_BATCH];
    for (int m = 0; m < nb; m++) {
      jobs[m].scalars = d + (size_t)(first + m) * n;
      jobs[m].n_extra = 1;
      jobs[m].extra[0] = bl[first + m];
      jobs[m].extra_index[0] = (uint32_t)n;  // W follows the n base points in both tables
      jobs[m].side_mask = index_mask;
      jobs[m].side_select = index_select != 0;
    }
    int32_t rc = msm_fixed_batch(ctx, fb, jobs, nb, n, r);
    if (rc) return rc;
    for (int m = 0; m < nb; m++) out[first + m] = r[m].to_affine();
  }
  return ZK_OK;
}

// `batch` transforms of size 2^log_n in one sequence of launches (the 19 witness columns go to
// coefficients this way): transform b reads in + b * 2^log_n and writes out + b * 2^log_n; in != out.
extern "C" int32_t zk_ntt_fp_batch(zk_ctx* ctx, const void* in, void* out, int32_t log_n, uint32_t batch,
                                   int32_t inverse, int32_t on_device) {
  if (!ctx || !in || !out || in == out || log_n < 1 || log_n > 28 || batch < 1 || batch > 65535) return ZK_E_INVALID;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t n = 1ull << log_n;
  const size_t bytes = (size_t)batch * n * sizeof(Fp);
  const Fp* d_in = (const Fp*)in;
  Fp* d_out = (Fp*)out;
  if (!on_device) {
    int32_t rc = ensure_buf(ctx, ctx->scratch_a, bytes);
    if (rc) return rc;
    if ((rc = ensure_buf(ctx, ctx->scratch_b, bytes))) return rc;
    ZK_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_a.ptr, in, bytes, cudaMemcpyHostToDevice, ctx->stream));
    d_in = (const Fp*)ctx->scratch_a.ptr;
    d_out = (Fp*)ctx->scratch_b.ptr;
  }
  NttOptions o;
  o.inverse = inverse != 0;
  o.batch = (int)batch;
  o.in_stride = o.out_stride = n;
  int32_t rc = ntt_run(ctx, d_in, (uint32_t)n, d_out, log_n, o);
  if (rc) return rc;
  if (!on_device) {
    ZK_CUDA(ctx, cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return zk_ctx_synchronize(ctx);
  }
  return ZK_OK;
}

// Coefficients of `ncols` polynomials (n each) -> their evaluations on the quotient domain of the
// context's keys: the three cosets c_j <omega_n>, c_j = zeta * omega_4n^j (three of the four cosets of
// halo2's extended domain; out[col][j][i] = extended[4 i + j]), as one columns x cosets batch.
extern "C" int32_t zk_coeff_to_cosets(zk_ctx* ctx, const void* coeffs, uint32_t ncols, int32_t on_device,
                                      void* out) {
  if (!ctx || !coeffs || !out || ncols < 1 || ncols > 21845) return ZK_E_INVALID;
  ProverState* S = prover_state(ctx);
  if (!S->has_keys) return set_error(ctx, ZK_E_STATE, "zk_coeff_to_cosets before keygen");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  const DeviceKeys& K = S->keys;
  const uint64_t n = K.n, en = K.en;
  const size_t in_bytes = (size_t)ncols * n * sizeof(Fp), out_bytes = (size_t)ncols * en * sizeof(Fp);
  const Fp* d_in = (const Fp*)coeffs;
  Fp* d_out = (Fp*)out;
  if (!on_device) {
    int32_t rc = ensure_buf(ctx, ctx->scratch_a, in_bytes);
    if (rc) return rc;
    if ((rc = ensure_buf(ctx, ctx->scratch_b, out_bytes))) return rc;
    ZK_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_a.ptr, coeffs, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    d_in = (const Fp*)ctx->scratch_a.ptr;
    d_out = (Fp*)ctx->scratch_b.ptr;
  }
  NttOptions c;  // the same options create_proof uses (prover.cu, quotient phase)
  c.batch = NUM_COSETS;
  c.in_stride = 0;
  c.out_stride = n;
  c.scale_in = K.coset_scale;
  c.scale_stride = n;
  c.batch2 = (int)ncols;
  c.in_stride2 = n;
  c.out_stride2 = en;
  int32_t rc = ntt_run(ctx, d_in, (uint32_t)n, d_out, K.k, c);
  if (rc) return rc;
  if (!on_device) {
    ZK_CUDA(ctx, cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return zk_ctx_synchronize(ctx);
  }
  return ZK_OK;
}
