// Implementation of polyops.cuh (see there).
#include "polyops.cuh"

namespace zkodst {

// ---- affine recurrence scan ----------------------------------------------------------------------
// Three launches whatever n: every block composes its tile of SCAN_T x SCAN_E elements into one affine
// map (thread-serial over SCAN_E elements, shuffle scan inside the warp, one more across the warps); one
// block scans the per-tile maps and applies them to `init`; every block then rebuilds its threads'
// exclusive prefixes the same way, starts from its tile's carry and writes the outputs.
namespace {
constexpr int SCAN_T = 256, SCAN_E = 8, SCAN_TILE = SCAN_T * SCAN_E;

__device__ __forceinline__ Fp shfl_up_fp(const Fp& v, int d) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v.l[i], d);
    const uint32_t hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v.l[i] >> 32), d);
    r.l[i] = ((uint64_t)hi << 32) | lo;
  }
  return r;
}
// (M, A) <- first (M1, A1) then (M, A):  y -> (y M1 + A1) M + A
template <bool HAS_A>
__device__ __forceinline__ void compose_before(Fp& M, Fp& A, const Fp& M1, const Fp& A1) {
  if (HAS_A) A = A1 * M + A;
  M = M1 * M;
}
// the thread's own elements as one map
template <bool HAS_A>
__device__ __forceinline__ void thread_map(const Fp* __restrict__ m, const Fp& m_const, const Fp* __restrict__ a,
                                           uint64_t lo, uint64_t hi, Fp& M, Fp& A) {
  M = Fp::one();
  A = Fp::zero();
  for (uint64_t i = lo; i < hi; i++) {
    const Fp mi = m ? m[i] : m_const;
    if (HAS_A) A = A * mi + a[i];
    M = M * mi;
  }
}
// Inclusive scan of the threads' maps over the block, in thread order; afterwards (M, A) of thread t is the
// composition of threads 0..t.  `excl` additionally returns the composition of threads 0..t-1.
template <bool HAS_A>
__device__ __forceinline__ void block_scan_maps(Fp& M, Fp& A, Fp* exclM, Fp* exclA) {
  __shared__ Fp wM[SCAN_T / 32], wA[SCAN_T / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const Fp pM = shfl_up_fp(M, d);
    Fp pA;
    if (HAS_A) pA = shfl_up_fp(A, d);
    if (lane >= d) compose_before<HAS_A>(M, A, pM, pA);
  }
  if (lane == 31) {
    wM[warp] = M;
    if (HAS_A) wA[warp] = A;
  }
  __syncthreads();
  if (warp == 0) {  // inclusive scan of the SCAN_T / 32 warp totals
    Fp tM = lane < SCAN_T / 32 ? wM[lane] : Fp::one(), tA = Fp::zero();
    if (HAS_A && lane < SCAN_T / 32) tA = wA[lane];
#pragma unroll
    for (int d = 1; d < SCAN_T / 32; d <<= 1) {
      const Fp pM = shfl_up_fp(tM, d);
      Fp pA;
      if (HAS_A) pA = shfl_up_fp(tA, d);
      if (lane >= d) compose_before<HAS_A>(tM, tA, pM, pA);
    }
    if (lane < SCAN_T / 32) {
      wM[lane] = tM;
      if (HAS_A) wA[lane] = tA;
    }
  }
  __syncthreads();
  if (warp > 0) compose_before<HAS_A>(M, A, wM[warp - 1], HAS_A ? wA[warp - 1] : Fp::zero());
  if (exclM) {  // exclusive = inclusive of the previous thread
    Fp eM = shfl_up_fp(M, 1), eA = Fp::zero();
    if (HAS_A) eA = shfl_up_fp(A, 1);
    if (lane == 0) {
      eM = warp > 0 ? wM[warp - 1] : Fp::one();
      eA = (HAS_A && warp > 0) ? wA[warp - 1] : Fp::zero();
    }
    *exclM = eM;
    *exclA = eA;
  }
}

template <bool HAS_A>
__global__ void __launch_bounds__(SCAN_T) affine_tile_up_kernel(const Fp* __restrict__ m, Fp m_const,
                                                               const Fp* __restrict__ a, uint64_t n,
                                                               AffinePair* __restrict__ agg) {
  const uint64_t lo = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_E;
  const uint64_t hi = lo + SCAN_E < n ? lo + SCAN_E : n;
  Fp M, A;
  thread_map<HAS_A>(m, m_const, a, lo < n ? lo : n, hi, M, A);
  block_scan_maps<HAS_A>(M, A, nullptr, nullptr);
  if (threadIdx.x == SCAN_T - 1) agg[blockIdx.x] = AffinePair{M, A};
}

// one block: carry[b] = init pushed through tiles 0..b-1
template <bool HAS_A>
__global__ void __launch_bounds__(SCAN_T) affine_tile_carry_kernel(const AffinePair* __restrict__ agg, uint64_t ntiles,
                                                                  const Fp* __restrict__ init,
                                                                  Fp* __restrict__ carry) {
  __shared__ Fp run;  // value entering the current group of SCAN_T tiles
  if (threadIdx.x == 0) run = *init;
  __syncthreads();
  for (uint64_t base = 0; base < ntiles; base += SCAN_T) {
    const uint64_t b = base + threadIdx.x;
    Fp M = Fp::one(), A = Fp::zero();
    if (b < ntiles) {
      M = agg[b].m;
      if (HAS_A) A = agg[b].a;
    }
    Fp eM, eA;
    block_scan_maps<HAS_A>(M, A, &eM, &eA);
    const Fp start = run;
    if (b < ntiles) carry[b] = HAS_A ? start * eM + eA : start * eM;
    __syncthreads();
    if (threadIdx.x == SCAN_T - 1) run = HAS_A ? start * M + A : start * M;
    __syncthreads();
  }
}

template <bool HAS_A>
__global__ void __launch_bounds__(SCAN_T) affine_tile_down_kernel(const Fp* __restrict__ m, Fp m_const,
                                                                 const Fp* __restrict__ a, uint64_t n,
                                                                 const Fp* __restrict__ carry, Fp* __restrict__ out) {
  const uint64_t lo0 = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_E;
  const uint64_t lo = lo0 < n ? lo0 : n, hi = lo0 + SCAN_E < n ? lo0 + SCAN_E : n;
  Fp M, A, eM, eA;
  thread_map<HAS_A>(m, m_const, a, lo, hi, M, A);
  block_scan_maps<HAS_A>(M, A, &eM, &eA);
  const Fp c = carry[blockIdx.x];
  Fp y = HAS_A ? c * eM + eA : c * eM;
  for (uint64_t i = lo; i < hi; i++) {
    const Fp mi = m ? m[i] : m_const;
    Fp ai;
    if (HAS_A) ai = a[i];
    out[i] = y;  // exclusive; out may alias m or a (an element is read, by its own thread, before it is written)
    y = HAS_A ? y * mi + ai : y * mi;
  }
}

template <bool HAS_A>
int32_t affine_scan_impl(zk_ctx* ctx, const Fp* m, const Fp& m_const, const Fp* a, uint64_t n, const Fp& init,
                         Fp* out) {
  const uint64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  int32_t rc = ensure_buf(ctx, ctx->scan_ws, ntiles * (sizeof(AffinePair) + sizeof(Fp)) + sizeof(Fp));
  if (rc) return rc;
  AffinePair* agg = (AffinePair*)ctx->scan_ws.ptr;
  Fp* carry = (Fp*)(agg + ntiles);
  Fp* d_init = carry + ntiles;
  cudaStream_t st = ctx->stream;
  ZK_CUDA(ctx, cudaMemcpyAsync(d_init, &init, sizeof(Fp), cudaMemcpyHostToDevice, st));
  if (ntiles > 1) {
    affine_tile_up_kernel<HAS_A><<<(unsigned)ntiles, SCAN_T, 0, st>>>(m, m_const, a, n, agg);
    affine_tile_carry_kernel<HAS_A><<<1, SCAN_T, 0, st>>>(agg, ntiles, d_init, carry);
    ctx->launches += 2;
  }
  affine_tile_down_kernel<HAS_A><<<(unsigned)ntiles, SCAN_T, 0, st>>>(m, m_const, a, n, ntiles > 1 ? carry : d_init, out);
  ctx->launches++;
  ZK_CUDA(ctx, cudaGetLastError());
  return ZK_OK;
}
}  // namespace

int32_t affine_scan(zk_ctx* ctx, const Fp* m, const Fp& m_const, const Fp* a, uint64_t n, const Fp& init,
                    Fp* out) {
  if (n == 0) return ZK_OK;
  return a ? affine_scan_impl<true>(ctx, m, m_const, a, n, init, out)
           : affine_scan_impl<false>(ctx, m, m_const, nullptr, n, init, out);
}

// ---- batched inversion ---------------------------------------------------------------------------
namespace {
// Montgomery's trick over chunks of INV_CH elements per thread (one Fermat inversion, ~380
// multiplications, per chunk); the prefix products live in a scratch vector so the chunk can be long.
constexpr int INV_CH = 32;
__global__ void __launch_bounds__(128) batch_invert_kernel(Fp* __restrict__ data, Fp* __restrict__ pre, uint64_t n) {
  uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t lo = c * INV_CH;
  if (lo >= n) return;
  const uint64_t hi = lo + INV_CH < n ? lo + INV_CH : n;
  Fp acc = Fp::one();
  for (uint64_t i = lo; i < hi; i++) {
    pre[i] = acc;
    const Fp x = data[i];
    if (!x.is_zero()) acc = acc * x;
  }
  acc = acc.inv();
  for (uint64_t i = hi; i-- > lo;) {
    const Fp x = data[i];
    if (!x.is_zero()) {
      data[i] = acc * pre[i];
      acc = acc * x;
    }
  }
}

// ---- evaluation ----------------------------------------------------------------------------------
constexpr int EV_THREADS = 256, EV_PER = 64, EV_BLOCK = EV_THREADS * EV_PER;
static_assert(EV_BLOCK == EVAL_BLOCK, "evaluation block size");  // 64: the two power computations per thread amortise over more coefficients (2.4 -> 1.3 products per coefficient)

__device__ __forceinline__ Fp block_sum(Fp v, Fp* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + s];
    __syncthreads();
  }
  return sh[0];
}

struct EvalJobs {
  EvalJob j[48];
};
__global__ void __launch_bounds__(EV_THREADS)
poly_eval_kernel(EvalJobs jobs, uint64_t n, Fp* __restrict__ partials, uint32_t nblocks, uint32_t first_block) {
  __shared__ Fp sh[EV_THREADS];
  const EvalJob& job = jobs.j[blockIdx.y];
  const Fp x = job.point;
  const uint32_t blk = first_block + blockIdx.x;  // coefficients [blk * EV_BLOCK, (blk + 1) * EV_BLOCK)
  uint64_t lo = (uint64_t)blk * EV_BLOCK + (uint64_t)threadIdx.x * EV_PER;
  Fp acc = Fp::zero();
  if (lo < n) {
    uint64_t hi = lo + EV_PER < n ? lo + EV_PER : n;
    for (uint64_t i = hi; i-- > lo;) acc = acc * x + job.poly[i];
    Fp xper = x.pow_u64(EV_PER);
    acc = acc * xper.pow_u64(threadIdx.x);
  }
  Fp total = block_sum(acc, sh);
  if (threadIdx.x == 0)
    partials[(size_t)blockIdx.y * nblocks + blockIdx.x] = total * x.pow_u64((uint64_t)blk * EV_BLOCK);
}
__global__ void __launch_bounds__(EV_THREADS)
sum_partials_kernel(const Fp* __restrict__ partials, uint32_t nblocks, Fp* __restrict__ results) {
  __shared__ Fp sh[EV_THREADS];
  Fp acc = Fp::zero();
  for (uint32_t b = threadIdx.x; b < nblocks; b += EV_THREADS) acc = acc + partials[(size_t)blockIdx.x * nblocks + b];
  Fp total = block_sum(acc, sh);
  if (threadIdx.x == 0) results[blockIdx.x] = total;
}
// blockIdx.y selects the pair: <a0, b0> or <a1, b1>
__global__ void __launch_bounds__(EV_THREADS)
inner_product_kernel(const Fp* __restrict__ a0, const Fp* __restrict__ b0, const Fp* __restrict__ a1,
                     const Fp* __restrict__ b1, uint64_t n, Fp* __restrict__ partials) {
  __shared__ Fp sh[EV_THREADS];
  const Fp* a = blockIdx.y ? a1 : a0;
  const Fp* b = blockIdx.y ? b1 : b0;
  Fp acc = Fp::zero();
  for (uint64_t i = blockIdx.x * (uint64_t)EV_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * EV_THREADS)
    acc = acc + a[i] * b[i];
  Fp total = block_sum(acc, sh);
  if (threadIdx.x == 0) partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = total;
}
}  // namespace

int32_t batch_invert(zk_ctx* ctx, Fp* data, uint64_t n) {
  if (!n) return ZK_OK;
  uint64_t chunks = (n + INV_CH - 1) / INV_CH;
  int32_t rc = ensure_buf(ctx, ctx->inv_ws, n * sizeof(Fp));
  if (rc) return rc;
  batch_invert_kernel<<<(unsigned)((chunks + 127) / 128), 128, 0, ctx->stream>>>(data, (Fp*)ctx->inv_ws.ptr, n);
  ctx->launches++;
  ZK_CUDA(ctx, cudaGetLastError());
  return ZK_OK;
}

// In a multi-GPU group the polynomials are replicated, so the sum splits by coefficient range: every rank
// evaluates its own contiguous range of EV_BLOCK-coefficient blocks for all jobs and the per-rank partial
// values are exchanged (dist_sum_fields): balanced whatever the number of jobs.
int32_t poly_eval_batch(zk_ctx* ctx, const EvalJob* jobs, int njobs, uint64_t n, Fp* results_host) {
  cudaStream_t st = ctx->stream;
  const uint32_t all_blocks = (uint32_t)((n + EV_BLOCK - 1) / EV_BLOCK);
  uint64_t b_lo = 0, b_hi = all_blocks;
  const bool split = dist_ranges_ok(n, ctx->dist_world);  // whole blocks per rank: reads stay inside the rank's range
  if (split) dist_range(all_blocks, ctx->dist_rank, ctx->dist_world, &b_lo, &b_hi);
  const uint32_t nblocks = (uint32_t)(b_hi - b_lo);
  for (int base = 0; base < njobs; base += 48) {
    int cnt = njobs - base < 48 ? njobs - base : 48;
    int32_t rc = ensure_buf(ctx, ctx->eval_ws, ((size_t)cnt * nblocks + 64) * sizeof(Fp));
    if (rc) return rc;
    Fp* partials = (Fp*)ctx->eval_ws.ptr;
    Fp* results = partials + (size_t)cnt * nblocks;
    EvalJobs ej;
    for (int i = 0; i < cnt; i++) ej.j[i] = jobs[base + i];
    poly_eval_kernel<<<dim3(nblocks, cnt), EV_THREADS, 0, st>>>(ej, n, partials, nblocks, (uint32_t)b_lo);
    sum_partials_kernel<<<cnt, EV_THREADS, 0, st>>>(partials, nblocks, results);
    ctx->launches += 2;
    ZK_CUDA(ctx, cudaGetLastError());
    if (split) {
      if ((rc = dist_sum_fields(ctx, results, cnt, results_host + base))) return rc;
    } else {
      ZK_CUDA(ctx, cudaMemcpyAsync(results_host + base, results, (size_t)cnt * sizeof(Fp), cudaMemcpyDeviceToHost, st));
      ZK_CUDA(ctx, zk_stream_sync(ctx));
    }
  }
  return ZK_OK;
}

// results_host[0] = <a0, b0>, results_host[1] = <a1, b1>: one launch pair and one wait for both (the two inner
// products of an IPA round)
int32_t inner_product_pair(zk_ctx* ctx, const Fp* a0, const Fp* b0, const Fp* a1, const Fp* b1, uint64_t n,
                           Fp results_host[2]) {
  cudaStream_t st = ctx->stream;
  uint32_t nblocks = (uint32_t)((n + EV_THREADS * 8 - 1) / (EV_THREADS * 8));
  if (nblocks < 1) nblocks = 1;
  if (nblocks > 1024) nblocks = 1024;
  int32_t rc = ensure_buf(ctx, ctx->eval_ws, (2 * (size_t)nblocks + 64) * sizeof(Fp));
  if (rc) return rc;
  Fp* partials = (Fp*)ctx->eval_ws.ptr;
  Fp* result = partials + 2 * (size_t)nblocks;
  inner_product_kernel<<<dim3(nblocks, 2), EV_THREADS, 0, st>>>(a0, b0, a1, b1, n, partials);
  sum_partials_kernel<<<2, EV_THREADS, 0, st>>>(partials, nblocks, result);
  ctx->launches += 2;
  ZK_CUDA(ctx, cudaGetLastError());
  ZK_CUDA(ctx, cudaMemcpyAsync(results_host, result, 2 * sizeof(Fp), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  return ZK_OK;
}

}  // namespace zkodst
