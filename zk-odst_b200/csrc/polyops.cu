// Implementation of polyops.cuh (see there).
#include "polyops.cuh"

namespace zkodst {

// ---- affine recurrence scan ----------------------------------------------------------------------
__global__ void affine_up_kernel(const Fp* __restrict__ m, Fp m_const, const Fp* __restrict__ a, uint64_t n,
                                 AffinePair* __restrict__ agg, uint64_t nchunks) {
  uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  uint64_t lo = c * SCAN_CH, hi = lo + SCAN_CH < n ? lo + SCAN_CH : n;
  Fp M = Fp::one(), A = Fp::zero();
  for (uint64_t i = lo; i < hi; i++) {
    Fp mi = m ? m[i] : m_const;
    M = M * mi;
    if (a) A = A * mi + a[i];
  }
  agg[c] = AffinePair{M, A};
}
__global__ void affine_up_pairs_kernel(const AffinePair* __restrict__ in, uint64_t n,
                                       AffinePair* __restrict__ agg, uint64_t nchunks) {
  uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  uint64_t lo = c * SCAN_CH, hi = lo + SCAN_CH < n ? lo + SCAN_CH : n;
  Fp M = Fp::one(), A = Fp::zero();
  for (uint64_t i = lo; i < hi; i++) {
    AffinePair p = in[i];
    M = M * p.m;
    A = A * p.m + p.a;
  }
  agg[c] = AffinePair{M, A};
}
__global__ void affine_down_kernel(const Fp* __restrict__ m, Fp m_const, const Fp* __restrict__ a, uint64_t n,
                                   const Fp* __restrict__ carry, Fp* __restrict__ out, uint64_t nchunks) {
  uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  uint64_t lo = c * SCAN_CH, hi = lo + SCAN_CH < n ? lo + SCAN_CH : n;
  Fp y = carry[c];
  for (uint64_t i = lo; i < hi; i++) {
    Fp mi = m ? m[i] : m_const;
    Fp ai = a ? a[i] : Fp::zero();
    out[i] = y;  // exclusive; out may alias m or a (each element is read before it is written)
    y = y * mi + ai;
  }
}
__global__ void affine_down_pairs_kernel(const AffinePair* __restrict__ in, uint64_t n,
                                         const Fp* __restrict__ carry, Fp* __restrict__ out, uint64_t nchunks) {
  uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  uint64_t lo = c * SCAN_CH, hi = lo + SCAN_CH < n ? lo + SCAN_CH : n;
  Fp y = carry[c];
  for (uint64_t i = lo; i < hi; i++) {
    AffinePair p = in[i];
    out[i] = y;
    y = y * p.m + p.a;
  }
}

int32_t affine_scan(zk_ctx* ctx, const Fp* m, const Fp& m_const, const Fp* a, uint64_t n, const Fp& init,
                    Fp* out) {
  if (n == 0) return ZK_OK;
  // level sizes: sz[0] = n elements, sz[l] = chunks of level l-1
  uint64_t sz[8];
  int levels = 0;
  sz[0] = n;
  while (sz[levels] > 1) {
    sz[levels + 1] = (sz[levels] + SCAN_CH - 1) / SCAN_CH;
    levels++;
  }
  // workspace: pairs for levels 1..levels, carries for levels 1..levels (+1 slot for init)
  size_t off = 0, pair_off[8], carry_off[8];
  for (int l = 1; l <= levels; l++) {
    pair_off[l] = off;
    off += sz[l] * sizeof(AffinePair);
  }
  for (int l = 1; l <= levels; l++) {
    carry_off[l] = off;
    off += sz[l] * sizeof(Fp);
  }
  size_t init_off = off;
  off += sizeof(Fp);
  int32_t rc = ensure_buf(ctx, ctx->scan_ws, off);
  if (rc) return rc;
  char* ws = (char*)ctx->scan_ws.ptr;
  auto pairs = [&](int l) { return (AffinePair*)(ws + pair_off[l]); };
  auto carry = [&](int l) { return (Fp*)(ws + carry_off[l]); };
  Fp* d_init = (Fp*)(ws + init_off);
  cudaStream_t st = ctx->stream;
  ZK_CUDA(ctx, cudaMemcpyAsync(d_init, &init, sizeof(Fp), cudaMemcpyHostToDevice, st));
  const int T = 128;
  if (levels == 0) {  // n == 1
    affine_down_kernel<<<1, 1, 0, st>>>(m, m_const, a, n, d_init, out, 1);
    ctx->launches++;
    return ZK_OK;
  }
  affine_up_kernel<<<(unsigned)((sz[1] + T - 1) / T), T, 0, st>>>(m, m_const, a, n, pairs(1), sz[1]);
  ctx->launches++;
  for (int l = 2; l <= levels; l++) {
    affine_up_pairs_kernel<<<(unsigned)((sz[l] + T - 1) / T), T, 0, st>>>(pairs(l - 1), sz[l - 1], pairs(l), sz[l]);
    ctx->launches++;
  }
  // top level has one chunk: its carry is init
  const Fp* c_in = d_init;
  for (int l = levels; l >= 2; l--) {
    affine_down_pairs_kernel<<<(unsigned)((sz[l] + T - 1) / T), T, 0, st>>>(pairs(l - 1), sz[l - 1], c_in,
                                                                          carry(l - 1), sz[l]);
    ctx->launches++;
    c_in = carry(l - 1);
  }
  affine_down_kernel<<<(unsigned)((sz[1] + T - 1) / T), T, 0, st>>>(m, m_const, a, n, c_in, out, sz[1]);
  ctx->launches++;
  ZK_CUDA(ctx, cudaGetLastError());
  return ZK_OK;
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
constexpr int EV_THREADS = 256, EV_PER = 16, EV_BLOCK = EV_THREADS * EV_PER;

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
poly_eval_kernel(EvalJobs jobs, uint64_t n, Fp* __restrict__ partials, uint32_t nblocks) {
  __shared__ Fp sh[EV_THREADS];
  const EvalJob& job = jobs.j[blockIdx.y];
  const Fp x = job.point;
  uint64_t lo = (uint64_t)blockIdx.x * EV_BLOCK + (uint64_t)threadIdx.x * EV_PER;
  Fp acc = Fp::zero();
  if (lo < n) {
    uint64_t hi = lo + EV_PER < n ? lo + EV_PER : n;
    for (uint64_t i = hi; i-- > lo;) acc = acc * x + job.poly[i];
    Fp x16 = x.pow_u64(EV_PER);
    acc = acc * x16.pow_u64(threadIdx.x);
  }
  Fp total = block_sum(acc, sh);
  if (threadIdx.x == 0)
    partials[(size_t)blockIdx.y * nblocks + blockIdx.x] = total * x.pow_u64((uint64_t)blockIdx.x * EV_BLOCK);
}
__global__ void __launch_bounds__(EV_THREADS)
sum_partials_kernel(const Fp* __restrict__ partials, uint32_t nblocks, Fp* __restrict__ results) {
  __shared__ Fp sh[EV_THREADS];
  Fp acc = Fp::zero();
  for (uint32_t b = threadIdx.x; b < nblocks; b += EV_THREADS) acc = acc + partials[(size_t)blockIdx.x * nblocks + b];
  Fp total = block_sum(acc, sh);
  if (threadIdx.x == 0) results[blockIdx.x] = total;
}
__global__ void __launch_bounds__(EV_THREADS)
inner_product_kernel(const Fp* __restrict__ a, const Fp* __restrict__ b, uint64_t n, Fp* __restrict__ partials) {
  __shared__ Fp sh[EV_THREADS];
  Fp acc = Fp::zero();
  for (uint64_t i = blockIdx.x * (uint64_t)EV_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * EV_THREADS)
    acc = acc + a[i] * b[i];
  Fp total = block_sum(acc, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = total;
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

int32_t poly_eval_batch(zk_ctx* ctx, const EvalJob* jobs, int njobs, uint64_t n, Fp* results_host) {
  cudaStream_t st = ctx->stream;
  uint32_t nblocks = (uint32_t)((n + EV_BLOCK - 1) / EV_BLOCK);
  for (int base = 0; base < njobs; base += 48) {
    int cnt = njobs - base < 48 ? njobs - base : 48;
    int32_t rc = ensure_buf(ctx, ctx->eval_ws, ((size_t)cnt * nblocks + 64) * sizeof(Fp));
    if (rc) return rc;
    Fp* partials = (Fp*)ctx->eval_ws.ptr;
    Fp* results = partials + (size_t)cnt * nblocks;
    EvalJobs ej;
    for (int i = 0; i < cnt; i++) ej.j[i] = jobs[base + i];
    poly_eval_kernel<<<dim3(nblocks, cnt), EV_THREADS, 0, st>>>(ej, n, partials, nblocks);
    sum_partials_kernel<<<cnt, EV_THREADS, 0, st>>>(partials, nblocks, results);
    ctx->launches += 2;
    ZK_CUDA(ctx, cudaGetLastError());
    ZK_CUDA(ctx, cudaMemcpyAsync(results_host + base, results, (size_t)cnt * sizeof(Fp), cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, zk_stream_sync(ctx));
  }
  return ZK_OK;
}

int32_t inner_product(zk_ctx* ctx, const Fp* a, const Fp* b, uint64_t n, Fp* result_host) {
  cudaStream_t st = ctx->stream;
  uint32_t nblocks = (uint32_t)((n + EV_THREADS * 8 - 1) / (EV_THREADS * 8));
  if (nblocks < 1) nblocks = 1;
  if (nblocks > 1024) nblocks = 1024;
  int32_t rc = ensure_buf(ctx, ctx->eval_ws, ((size_t)nblocks + 64) * sizeof(Fp));
  if (rc) return rc;
  Fp* partials = (Fp*)ctx->eval_ws.ptr;
  Fp* result = partials + nblocks;
  inner_product_kernel<<<nblocks, EV_THREADS, 0, st>>>(a, b, n, partials);
  sum_partials_kernel<<<1, EV_THREADS, 0, st>>>(partials, nblocks, result);
  ctx->launches += 2;
  ZK_CUDA(ctx, cudaGetLastError());
  ZK_CUDA(ctx, cudaMemcpyAsync(result_host, result, sizeof(Fp), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  return ZK_OK;
}

}  // namespace zkodst
