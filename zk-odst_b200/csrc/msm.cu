// K2/K3 — Pippenger multi-scalar multiplication over Vesta.
//
// Replaces halo2_proofs 0.3.0 `best_multiexp`, reached from `Params::commit_lagrange` /
// `Params::commit` inside `create_proof` (blake2f-circuit/benches/blake2f.rs:125) and from the
// IPA rounds.  Scalars are Montgomery-form Fp, bases are affine Montgomery-form Fq pairs.
//
// Pipeline (all kernels on the context's stream):
//   1. digits:     scalar -> canonical -> signed base-2^c digits; histogram of bucket sizes
//   2. scan:       exclusive prefix sum of the histogram (bucket offsets)
//   3. scatter:    (point index, sign) written into its bucket's slot  (counting sort)
//   4. accumulate: one thread per light bucket, mixed XYZZ additions; heavy buckets (skewed
//                  small-value advice columns) are split into block-sized work items
//   5. reduce:     running-sum over segments of 8 buckets, then fan-in-16 merges of
//                  (A = sum (b - lo + 1) B_b, S = sum B_b) pairs until one per window
//   6. the c-bit Horner over <= 26 window sums runs on the host (a 400-step dependent
//      doubling chain is latency-bound on a GPU thread and ~0.1 ms on a CPU core).
// The result is a group element; only its affine normalisation is observable.
//
// Roofline: integer-pipe bound: 16 bucket additions x 11 Fq mults x 136 MAC
// per full-width term; HBM traffic is 96 B per term.
#include <chrono>
#include <cstdlib>

#include "ec.cuh"
#include "zk_ctx.h"

namespace zkodst {
namespace {

constexpr int HEAVY_THRESHOLD = 2048;   // bucket sizes above this are split across blocks
constexpr int HEAVY_ITEM = 8192;        // points per heavy work item
constexpr int HEAVY_THREADS = 128;
constexpr int SEG = 8;                  // buckets per level-1 reduction thread
constexpr int FAN = 16;                 // fan-in of the merge levels

struct HeavyItem {
  uint32_t bucket, start, len, slot;
};

// -------- step 1: digits + histogram ---------------------------------------------------------
// `extra` (optional) supplies one more scalar, logically scalars[n - 1], so that a commitment's
// blinding term [r] W rides in the same MSM (the bases array carries W after the n - 1 points).
__global__ void msm_digits_kernel(const Fp* __restrict__ scalars, const Fp* __restrict__ extra, uint32_t n,
                                  int c, int nwin, int32_t* __restrict__ digits,
                                  uint32_t* __restrict__ counts) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t s[4];
  if (extra && i == n - 1)
    extra->to_canonical(s);
  else
    scalars[i].to_canonical(s);
  const uint32_t B = 1u << (c - 1);
  uint32_t carry = 0;
  for (int w = 0; w < nwin; w++) {
    int bit = w * c, limb = bit >> 6, off = bit & 63;
    uint64_t v = limb < 4 ? s[limb] >> off : 0;
    if (off + c > 64 && limb + 1 < 4) v |= s[limb + 1] << (64 - off);
    uint32_t d = (uint32_t)(v & ((1ull << c) - 1)) + carry;
    int32_t sd;
    if (d > B) {
      sd = (int32_t)d - (int32_t)(1u << c);
      carry = 1;
    } else {
      sd = (int32_t)d;
      carry = 0;
    }
    digits[(size_t)w * n + i] = sd;
    if (sd != 0) {
      uint32_t mag = sd < 0 ? (uint32_t)(-sd) : (uint32_t)sd;
      atomicAdd(&counts[(size_t)w * B + mag - 1], 1u);
    }
  }
}

// -------- step 2: exclusive scan (three small kernels) ---------------------------------------
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 4, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void scan_tiles_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                  uint32_t* __restrict__ tile_sums, uint32_t n) {
  __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
  uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = base + k < n ? in[base + k] : 0;
    sum += v[k];
  }
  uint32_t incl = sum;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  uint32_t warp_off = 0;
  for (int w = 0; w < warp; w++) warp_off += warp_sums[w];
  uint32_t excl = warp_off + incl - sum;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = excl;
    excl += v[k];
  }
  if (threadIdx.x == SCAN_THREADS - 1) tile_sums[blockIdx.x] = warp_off + incl;
}
__global__ void scan_sums_kernel(uint32_t* tile_sums, uint32_t ntiles) {  // single block
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < ntiles; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < ntiles ? tile_sums[i] : 0, incl = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ uint32_t ws[32];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    uint32_t off = carry;
    for (int w = 0; w < warp; w++) off += ws[w];
    if (i < ntiles) tile_sums[i] = off + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = off + incl;
    __syncthreads();
  }
}
__global__ void scan_add_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ tile_sums,
                                uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += tile_sums[i / SCAN_TILE];
}

// -------- step 3: scatter ----------------------------------------------------------------------
__global__ void msm_scatter_kernel(const int32_t* __restrict__ digits, uint32_t n, int c, int nwin,
                                   const uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor,
                                   uint32_t* __restrict__ sorted) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t B = 1u << (c - 1);
  for (int w = 0; w < nwin; w++) {
    int32_t sd = digits[(size_t)w * n + i];
    if (sd == 0) continue;
    uint32_t mag = sd < 0 ? (uint32_t)(-sd) : (uint32_t)sd;
    uint32_t bucket = (uint32_t)w * B + mag - 1;
    uint32_t pos = offsets[bucket] + atomicAdd(&cursor[bucket], 1u);
    sorted[pos] = i | (sd < 0 ? 0x80000000u : 0u);
  }
}

// -------- step 4: accumulate -------------------------------------------------------------------
__device__ __forceinline__ Affine load_point(const Affine* __restrict__ bases, uint32_t v) {
  Affine p = bases[v & 0x7fffffffu];
  if (v & 0x80000000u) p.y = p.y.neg();
  return p;
}

__global__ void msm_find_heavy_kernel(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                                      uint32_t nbuckets, HeavyItem* __restrict__ items,
                                      uint32_t* __restrict__ n_items, uint32_t max_items,
                                      uint32_t* __restrict__ heavy_buckets, uint32_t* __restrict__ n_heavy) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  uint32_t cnt = counts[b];
  if (cnt <= HEAVY_THRESHOLD) return;
  uint32_t pieces = (cnt + HEAVY_ITEM - 1) / HEAVY_ITEM;
  uint32_t first = atomicAdd(n_items, pieces);
  uint32_t hb = atomicAdd(n_heavy, 1u);
  heavy_buckets[3 * hb] = b;
  heavy_buckets[3 * hb + 1] = first;
  heavy_buckets[3 * hb + 2] = pieces;
  for (uint32_t p = 0; p < pieces && first + p < max_items; p++) {
    uint32_t start = p * HEAVY_ITEM;
    uint32_t len = cnt - start < HEAVY_ITEM ? cnt - start : HEAVY_ITEM;
    items[first + p] = HeavyItem{b, offsets[b] + start, len, first + p};
  }
}

__global__ void __launch_bounds__(128)
msm_accumulate_kernel(const Affine* __restrict__ bases, const uint32_t* __restrict__ sorted,
                      const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                      uint32_t nbuckets, XYZZ* __restrict__ buckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  uint32_t cnt = counts[b];
  if (cnt > HEAVY_THRESHOLD) return;  // done by the heavy path
  XYZZ acc = XYZZ::identity();
  const uint32_t* list = sorted + offsets[b];
  for (uint32_t k = 0; k < cnt; k++) acc = acc.add_affine(load_point(bases, list[k]));
  buckets[b] = acc;
}

__global__ void __launch_bounds__(HEAVY_THREADS)
msm_heavy_accumulate_kernel(const Affine* __restrict__ bases, const uint32_t* __restrict__ sorted,
                            const HeavyItem* __restrict__ items, XYZZ* __restrict__ partials) {
  __shared__ XYZZ sh[HEAVY_THREADS];
  HeavyItem it = items[blockIdx.x];
  XYZZ acc = XYZZ::identity();
  for (uint32_t k = threadIdx.x; k < it.len; k += HEAVY_THREADS)
    acc = acc.add_affine(load_point(bases, sorted[it.start + k]));
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int stride = HEAVY_THREADS / 2; stride > 0; stride >>= 1) {
    if (threadIdx.x < stride) sh[threadIdx.x] = sh[threadIdx.x].add(sh[threadIdx.x + stride]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[it.slot] = sh[0];
}

__global__ void msm_heavy_finalize_kernel(const uint32_t* __restrict__ heavy_buckets, uint32_t n_heavy,
                                          const XYZZ* __restrict__ partials, XYZZ* __restrict__ buckets) {
  uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_heavy) return;
  uint32_t b = heavy_buckets[3 * h], first = heavy_buckets[3 * h + 1], pieces = heavy_buckets[3 * h + 2];
  XYZZ acc = partials[first];
  for (uint32_t p = 1; p < pieces; p++) acc = acc.add(partials[first + p]);
  buckets[b] = acc;
}

// -------- step 5: bucket reduction ---------------------------------------------------------------
// level 1: thread per SEG consecutive buckets of one window:
//   S = sum B_b,  A = sum (b - lo + 1) B_b   (running-sum form, no scalar multiplications)
__global__ void __launch_bounds__(128)
msm_reduce_level1_kernel(const XYZZ* __restrict__ buckets, uint32_t nsegs_total, XYZZ* __restrict__ outA,
                         XYZZ* __restrict__ outS) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsegs_total) return;
  const XYZZ* seg = buckets + (size_t)s * SEG;
  XYZZ running = XYZZ::identity(), acc = XYZZ::identity();
#pragma unroll 1
  for (int b = SEG - 1; b >= 0; b--) {
    running = running.add(seg[b]);
    acc = acc.add(running);
  }
  outA[s] = acc;
  outS[s] = running;
}
// merge level: thread per group of `fan` consecutive items, each item covering `len` buckets
// (len = 2^log_len):  S = sum S_i,  A = sum A_i + len * sum_i i * S_i
__global__ void __launch_bounds__(128)
msm_reduce_merge_kernel(const XYZZ* __restrict__ inA, const XYZZ* __restrict__ inS, uint32_t ngroups,
                        int fan, int log_len, XYZZ* __restrict__ outA, XYZZ* __restrict__ outS) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  const XYZZ* a = inA + (size_t)g * fan;
  const XYZZ* s = inS + (size_t)g * fan;
  XYZZ running = XYZZ::identity(), weighted = XYZZ::identity(), sumA = XYZZ::identity();
#pragma unroll 1
  for (int i = fan - 1; i >= 0; i--) {
    sumA = sumA.add(a[i]);
    if (i > 0) {
      running = running.add(s[i]);
      weighted = weighted.add(running);  // sum_{i>=1} i * S_i
    }
  }
  for (int d = 0; d < log_len; d++) weighted = weighted.dbl();
  outA[g] = sumA.add(weighted);
  outS[g] = running.add(s[0]);
}

int ilog2(uint32_t v) {
  int r = 0;
  while ((1u << r) < v) r++;
  return r;
}

}  // namespace

int msm_window_bits(uint64_t n) {
  if (n < (1u << 9)) return 6;
  if (n < (1u << 12)) return 8;
  if (n < (1u << 15)) return 10;
  if (n < (1u << 17)) return 12;
  if (n < (1u << 21)) return 14;
  return 16;
}

// Runs the device part of an MSM; writes nwin window sums (XYZZ) to d_window_sums.
int32_t msm_device(zk_ctx* ctx, const Fp* d_scalars, const Fp* d_extra, const Affine* d_bases, uint64_t n64, int c,
                   int* nwin_out, XYZZ* d_window_sums) {
  const uint32_t n = (uint32_t)n64;
  const int nwin = (255 + c - 1) / c + ((255 % c) == 0 ? 1 : 0);
  // top window: scalars < 2^255, so for c | 255 an extra window absorbs the signed carry
  *nwin_out = nwin;
  const uint32_t B = 1u << (c - 1);
  const uint32_t nbuckets = (uint32_t)nwin * B;
  cudaStream_t st = ctx->stream;
  // workspace layout
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  const uint32_t ntiles = (nbuckets + SCAN_TILE - 1) / SCAN_TILE;
  const uint32_t max_heavy_items = (uint32_t)((uint64_t)nwin * n / HEAVY_THRESHOLD + 1024);
  const uint32_t nsegs = nbuckets / SEG;
  size_t o_digits = take((size_t)nwin * n * 4), o_counts = take((size_t)nbuckets * 4),
         o_offsets = take((size_t)nbuckets * 4), o_cursor = take((size_t)nbuckets * 4),
         o_tiles = take((size_t)ntiles * 4 + 16), o_sorted = take((size_t)nwin * n * 4),
         o_hitems = take((size_t)max_heavy_items * sizeof(HeavyItem)),
         o_hbuckets = take((size_t)max_heavy_items * 12), o_hcount = take(16),
         o_hpartials = take((size_t)max_heavy_items * sizeof(XYZZ)),
         o_buckets = take((size_t)nbuckets * sizeof(XYZZ)), o_A0 = take((size_t)nsegs * sizeof(XYZZ)),
         o_S0 = take((size_t)nsegs * sizeof(XYZZ)), o_A1 = take((size_t)(nsegs / FAN + nwin) * sizeof(XYZZ)),
         o_S1 = take((size_t)(nsegs / FAN + nwin) * sizeof(XYZZ));
  int32_t rc = ensure_buf(ctx, ctx->msm_ws, off);
  if (rc) return rc;
  char* ws = (char*)ctx->msm_ws.ptr;
  int32_t* digits = (int32_t*)(ws + o_digits);
  uint32_t *counts = (uint32_t*)(ws + o_counts), *offsets = (uint32_t*)(ws + o_offsets),
           *cursor = (uint32_t*)(ws + o_cursor), *tiles = (uint32_t*)(ws + o_tiles),
           *sorted = (uint32_t*)(ws + o_sorted), *hbuckets = (uint32_t*)(ws + o_hbuckets),
           *hcount = (uint32_t*)(ws + o_hcount);
  HeavyItem* hitems = (HeavyItem*)(ws + o_hitems);
  XYZZ *hpartials = (XYZZ*)(ws + o_hpartials), *buckets = (XYZZ*)(ws + o_buckets);
  XYZZ *A[2] = {(XYZZ*)(ws + o_A0), (XYZZ*)(ws + o_A1)}, *S[2] = {(XYZZ*)(ws + o_S0), (XYZZ*)(ws + o_S1)};

  ZK_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)nbuckets * 4, st));
  ZK_CUDA(ctx, cudaMemsetAsync(cursor, 0, (size_t)nbuckets * 4, st));
  ZK_CUDA(ctx, cudaMemsetAsync(hcount, 0, 16, st));
  KernelTimer timer(ctx, KC_MSM);
  const int T = 256;
  msm_digits_kernel<<<(n + T - 1) / T, T, 0, st>>>(d_scalars, d_extra, n, c, nwin, digits, counts);
  scan_tiles_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(counts, offsets, tiles, nbuckets);
  scan_sums_kernel<<<1, 1024, 0, st>>>(tiles, ntiles);
  scan_add_kernel<<<(nbuckets + T - 1) / T, T, 0, st>>>(offsets, tiles, nbuckets);
  msm_scatter_kernel<<<(n + T - 1) / T, T, 0, st>>>(digits, n, c, nwin, offsets, cursor, sorted);
  msm_find_heavy_kernel<<<(nbuckets + T - 1) / T, T, 0, st>>>(counts, offsets, nbuckets, hitems, hcount,
                                                               max_heavy_items, hbuckets, hcount + 1);
  {
    KernelTimer acc_timer(ctx, KC_MSM_ACC);
    msm_accumulate_kernel<<<(nbuckets + 127) / 128, 128, 0, st>>>(d_bases, sorted, counts, offsets, nbuckets,
                                                                   buckets);
  }
  ctx->launches += 7;
  // heavy path: sizes are data dependent, so read the two counters back
  uint32_t hc[2];
  ZK_CUDA(ctx, cudaMemcpyAsync(hc, hcount, 8, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  if (hc[0] > max_heavy_items) return set_error(ctx, ZK_E_NOMEM, "msm: heavy work list overflow");
  if (hc[0]) {
    msm_heavy_accumulate_kernel<<<hc[0], HEAVY_THREADS, 0, st>>>(d_bases, sorted, hitems, hpartials);
    msm_heavy_finalize_kernel<<<(hc[1] + 63) / 64, 64, 0, st>>>(hbuckets, hc[1], hpartials, buckets);
    ctx->launches += 2;
  }
  msm_reduce_level1_kernel<<<(nsegs + 127) / 128, 128, 0, st>>>(buckets, nsegs, A[0], S[0]);
  ctx->launches++;
  uint32_t items_per_window = B / SEG;
  int log_len = ilog2(SEG), cur = 0;
  while (items_per_window > 1) {
    int fan = items_per_window >= (uint32_t)FAN ? FAN : (int)items_per_window;
    uint32_t groups = (uint32_t)nwin * (items_per_window / fan);
    msm_reduce_merge_kernel<<<(groups + 127) / 128, 128, 0, st>>>(A[cur], S[cur], groups, fan, log_len,
                                                                  A[cur ^ 1], S[cur ^ 1]);
    ctx->launches++;
    items_per_window /= fan;
    log_len += ilog2(fan);
    cur ^= 1;
  }
  ZK_CUDA(ctx, cudaGetLastError());
  ZK_CUDA(ctx, cudaMemcpyAsync(d_window_sums, A[cur], (size_t)nwin * sizeof(XYZZ),
                               cudaMemcpyDeviceToDevice, st));
  return ZK_OK;
}

// Full MSM: device pipeline + host Horner over the window sums.
// n counts the extra scalar when d_extra != nullptr (bases has n entries either way).
int32_t msm_run(zk_ctx* ctx, const Fp* d_scalars, const Affine* d_bases, uint64_t n, XYZZ* result,
                const Fp* d_extra) {
  if (n == 0) {
    *result = XYZZ::identity();
    return ZK_OK;
  }
  if (n > 0x7fffffffull) return set_error(ctx, ZK_E_INVALID, "msm: n too large");
  const int c = msm_window_bits(n);
  int nwin = 0;
  int32_t rc = ensure_buf(ctx, ctx->msm_out, 64 * sizeof(XYZZ));
  if (rc) return rc;
  static const bool trace = getenv("ZK_MSM_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t0;
  if (trace) {
    zk_stream_sync(ctx);
    t0 = std::chrono::steady_clock::now();
  }
  rc = msm_device(ctx, d_scalars, d_extra, d_bases, n, c, &nwin, (XYZZ*)ctx->msm_out.ptr);
  if (rc) return rc;
  if (trace) {
    zk_stream_sync(ctx);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[msm] n=%llu c=%d nwin=%d device_ms=%.3f\n", (unsigned long long)n, c, nwin, ms);
  }
  XYZZ sums[48];
  ZK_CUDA(ctx, cudaMemcpyAsync(sums, ctx->msm_out.ptr, (size_t)nwin * sizeof(XYZZ),
                               cudaMemcpyDeviceToHost, ctx->stream));
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  XYZZ total = XYZZ::identity();
  for (int w = nwin - 1; w >= 0; w--) {
    for (int i = 0; i < c; i++) total = total.dbl();
    total = total.add(sums[w]);
  }
  *result = total;
  return ZK_OK;
}

}  // namespace zkodst

using namespace zkodst;

extern "C" int32_t zk_msm_vesta(zk_ctx* ctx, const void* scalars, const void* bases, uint64_t n,
                                int32_t on_device, void* out_affine) {
  if (!ctx || !out_affine || (n && (!scalars || !bases))) return ZK_E_INVALID;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  const Fp* d_s = (const Fp*)scalars;
  const Affine* d_b = (const Affine*)bases;
  if (!on_device && n) {
    int32_t rc = ensure_buf(ctx, ctx->scratch_a, n * sizeof(Fp));
    if (rc) return rc;
    rc = ensure_buf(ctx, ctx->scratch_b, n * sizeof(Affine));
    if (rc) return rc;
    ZK_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_a.ptr, scalars, n * sizeof(Fp), cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_b.ptr, bases, n * sizeof(Affine), cudaMemcpyHostToDevice, ctx->stream));
    d_s = (const Fp*)ctx->scratch_a.ptr;
    d_b = (const Affine*)ctx->scratch_b.ptr;
  }
  XYZZ r;
  int32_t rc = msm_run(ctx, d_s, d_b, n, &r, nullptr);
  if (rc) return rc;
  Affine a = r.to_affine();
  memcpy(out_affine, &a, sizeof a);
  return ZK_OK;
}
