// Device building blocks of the prover: element-wise maps, affine-recurrence scans (grand
// products, synthetic division), batched inversion, polynomial evaluation and inner products.
// Product code.  Together they replace the CPU loops of halo2_proofs 0.3.0
// `plonk/permutation/prover.rs`, `plonk/lookup/prover.rs`, `poly/multiopen/prover.rs` and
// `arithmetic.rs::{eval_polynomial, kate_division, compute_inner_product}` (SURVEY.md §2.5
// K8-K10), all reached from `create_proof` (blake2f-circuit/benches/blake2f.rs:125).
#pragma once
#include "ec.cuh"
#include "zk_ctx.h"

namespace zkodst {

// ---- element-wise map with an extended lambda: f(i) for i in [0, n) ------------------------------
template <class F>
__global__ void map_kernel(uint64_t n, F f) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    f(i);
}
template <class F>
inline void launch_map(zk_ctx* ctx, uint64_t n, F f, int threads = 256) {
  if (n == 0) return;
  uint64_t blocks = (n + threads - 1) / threads;
  uint64_t cap = (uint64_t)ctx->sm_count * 32;
  if (blocks > cap) blocks = cap;
  map_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(n, f);
  ctx->launches++;
}

// the same over the index range [lo, lo + count): f(i) for i in the range (row / coefficient ranges of a group)
template <class F>
__global__ void map_range_kernel(uint64_t lo, uint64_t count, F f) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
    f(lo + i);
}
template <class F>
inline void launch_map_range(zk_ctx* ctx, uint64_t lo, uint64_t count, F f, int threads = 256) {
  if (count == 0) return;
  uint64_t blocks = (count + threads - 1) / threads;
  uint64_t cap = (uint64_t)ctx->sm_count * 32;
  if (blocks > cap) blocks = cap;
  map_range_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(lo, count, f);
  ctx->launches++;
}

// ---- affine recurrence scan ----------------------------------------------------------------------
// y_0 = init;  y_{i+1} = y_i * m_i + a_i      (i = 0 .. n-1), output y_0 .. y_{n-1}  (exclusive)
// With a == nullptr the recurrence is the running product used by the permutation and lookup
// grand products; with a constant multiplier it is Horner / synthetic division.
// Tile-wise scan in three launches (polyops.cu).
struct AffinePair {
  Fp m, a;
};

// Host driver.  m may be nullptr (use m_const); a may be nullptr (zero).  out may alias m or a.
int32_t affine_scan(zk_ctx* ctx, const Fp* m, const Fp& m_const, const Fp* a, uint64_t n, const Fp& init,
                    Fp* out);

// ---- batched inversion (in place); zeros stay zero -------------------------------------------------
int32_t batch_invert(zk_ctx* ctx, Fp* data, uint64_t n);

// ---- evaluations ------------------------------------------------------------------------------------
struct EvalJob {
  const Fp* poly;   // coefficients, length n
  Fp point;
};
// Coefficients per evaluation block: a group shards every evaluation by coefficient range in whole blocks.
constexpr uint64_t EVAL_BLOCK = 16384;
// A group of `world` ranks works on rows / coefficients by contiguous range when every rank gets whole
// evaluation blocks: rank r owns [r n / world, (r + 1) n / world) (the point ranges of the split MSMs as well).
static inline bool dist_ranges_ok(uint64_t n, int world) { return world > 1 && n % ((uint64_t)world * EVAL_BLOCK) == 0; }
// results[j] = poly_j(point_j); all jobs share the length n.  Results are copied to the host.
// In a group (dist_ranges_ok) every rank reads only its own coefficient range of each polynomial.
int32_t poly_eval_batch(zk_ctx* ctx, const EvalJob* jobs, int njobs, uint64_t n, Fp* results_host);
// <a, b> over n elements
int32_t inner_product_pair(zk_ctx* ctx, const Fp* a0, const Fp* b0, const Fp* a1, const Fp* b1, uint64_t n,
                           Fp results_host[2]);

}  // namespace zkodst
