// K2/K3/K11 — fixed-base Pippenger MSM with precomputed window tables.
//
// Every commitment of the proof is an MSM over one of two fixed bases, g or g_lagrange
// (`Params::commit` / `Params::commit_lagrange`, halo2_proofs 0.3.0; reached from
// `create_proof`, blake2f-circuit/benches/blake2f.rs:125), and the IPA rounds are rewritten
// over the original g as well (prover.cu).  For a fixed base the multiples 2^(c w) * G_i are
// tabulated once per params (`FixedBase`), which turns the classic per-window bucket sets into
// ONE set of 2^(c-1) buckets shared by all windows:
//     sum_i s_i G_i = sum_b b * ( sum_{(i,w): |d_iw| = b} sgn(d_iw) * T[w][i] )
// so there is a single bucket reduction per MSM and no doubling chain.
//
// Pipeline (async on the context's stream, one D2H of <= 32 partial points at the end):
//   digits -> histogram -> scan -> scatter            counting sort of (w, i) by |digit|
//   accumulate                                         thread per light bucket; buckets above
//                                                      HEAVY_THRESHOLD (skewed small-value
//                                                      advice columns) are cut into block-sized
//                                                      items reduced by a shared-memory tree
//   reduce                                             sum_b b * B_b: running sums over segments
//                                                      of 8 buckets, then per-bit tree sums of the
//                                                      segment totals (shallow dependency chains:
//                                                      one EC addition is ~10 us of latency)
//   host                                               Horner over <= 20 partial sums
//
// Roofline: integer-pipe bound; see DESIGN.md §MSM for the MAC accounting.
#include <chrono>
#include <cstdlib>

#include "msm_fixed.h"

namespace zkodst {
namespace {

constexpr int HEAVY_THRESHOLD = 256;
constexpr int HEAVY_ITEM = 2048;
constexpr int HEAVY_THREADS = 128;
constexpr int SEG = 8;
constexpr int TREE_THREADS = 128;
constexpr int TREE_PER_THREAD = 8;

struct HeavyItem {
  uint32_t bucket, start, len, slot;
};

// ---- table construction --------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
table_next_kernel(const Affine* __restrict__ prev, Affine* __restrict__ next, uint64_t npoints, int c) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= npoints) return;
  Affine p = prev[i];
  if (p.is_identity()) {
    next[i] = p;
    return;
  }
  XYZZ acc = XYZZ::dbl_affine(p);
  for (int d = 1; d < c; d++) acc = acc.dbl();
  next[i] = acc.to_affine();
}

// ---- digits + histogram -----------------------------------------------------------------------------
__global__ void fixed_digits_kernel(const Fp* __restrict__ scalars, uint64_t count, const Fp* __restrict__ extra,
                                    const uint32_t* __restrict__ extra_index, int n_extra, int c, int nwin,
                                    uint64_t npoints, uint32_t* __restrict__ entries_tmp,
                                    uint32_t* __restrict__ counts, uint32_t side_bit_mask, int side_select) {
  // entries_tmp[w * (count + n_extra) + t] = bucket + 1 | sign << 31   (0 = no entry)
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t total = count + n_extra;
  if (t >= total) return;
  uint64_t s[4];
  bool zero = false;
  if (t < count) {
    // optional support mask (IPA rounds): keep index t only if ((t & mask) != 0) == side_select
    if (side_bit_mask && (((t & side_bit_mask) != 0) != (side_select != 0))) zero = true;
    if (!zero) {
      Fp v = scalars[t];
      zero = v.is_zero();
      if (!zero) v.to_canonical(s);
    }
  } else {
    extra[t - count].to_canonical(s);
  }
  const uint32_t B = 1u << (c - 1);
  uint32_t carry = 0;
  for (int w = 0; w < nwin; w++) {
    uint32_t e = 0;
    if (!zero) {
      int bit = w * c, limb = bit >> 6, off = bit & 63;
      uint64_t v = limb < 4 ? s[limb] >> off : 0;
      if (off + c > 64 && limb + 1 < 4) v |= s[limb + 1] << (64 - off);
      uint32_t d = (uint32_t)(v & ((1ull << c) - 1)) + carry;
      if (d > B) {
        carry = 1;
        uint32_t mag = (1u << c) - d;  // 0 when the raw digit 2^c - 1 absorbs a carry: digit 0, carry 1
        if (mag) {
          e = mag | 0x80000000u;
          atomicAdd(&counts[mag - 1], 1u);
        }
      } else {
        carry = 0;
        if (d) {
          e = d;
          atomicAdd(&counts[d - 1], 1u);
        }
      }
    }
    entries_tmp[(size_t)w * total + t] = e;
  }
  (void)npoints;
  (void)extra_index;
}

constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 4, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
__global__ void fscan_tiles_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
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
__global__ void fscan_sums_kernel(uint32_t* tile_sums, uint32_t ntiles) {  // single block
  __shared__ uint32_t carry;
  __shared__ uint32_t ws[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < ntiles; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < ntiles ? tile_sums[i] : 0, incl = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
__global__ void fscan_add_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ tile_sums, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += tile_sums[i / SCAN_TILE];
}

// scatter: sorted[pos] = table entry index (w * npoints + point) | sign << 31
__global__ void fixed_scatter_kernel(const uint32_t* __restrict__ entries_tmp, uint64_t count, int n_extra,
                                     const uint32_t* __restrict__ extra_index, int nwin, uint64_t npoints,
                                     const uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor,
                                     uint32_t* __restrict__ sorted) {
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t total = count + n_extra;
  if (t >= total) return;
  const uint32_t point = t < count ? (uint32_t)t : extra_index[t - count];
  for (int w = 0; w < nwin; w++) {
    uint32_t e = entries_tmp[(size_t)w * total + t];
    if (!e) continue;
    uint32_t bucket = (e & 0x7fffffffu) - 1;
    uint32_t pos = offsets[bucket] + atomicAdd(&cursor[bucket], 1u);
    sorted[pos] = (uint32_t)((uint64_t)w * npoints + point) | (e & 0x80000000u);
  }
}

__device__ __forceinline__ Affine load_entry(const Affine* __restrict__ table, uint32_t v) {
  Affine p = table[v & 0x7fffffffu];
  if (v & 0x80000000u) p.y = p.y.neg();
  return p;
}

__global__ void fixed_find_heavy_kernel(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                                        uint32_t nbuckets, HeavyItem* __restrict__ items,
                                        uint32_t* __restrict__ hcount, uint32_t max_items,
                                        uint32_t* __restrict__ heavy_buckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  uint32_t cnt = counts[b];
  if (cnt <= HEAVY_THRESHOLD) return;
  uint32_t pieces = (cnt + HEAVY_ITEM - 1) / HEAVY_ITEM;
  uint32_t first = atomicAdd(&hcount[0], pieces);
  uint32_t hb = atomicAdd(&hcount[1], 1u);
  if (first + pieces > max_items) {
    hcount[2] = 1;  // overflow flag
    return;
  }
  heavy_buckets[3 * hb] = b;
  heavy_buckets[3 * hb + 1] = first;
  heavy_buckets[3 * hb + 2] = pieces;
  for (uint32_t p = 0; p < pieces; p++) {
    uint32_t start = p * HEAVY_ITEM;
    uint32_t len = cnt - start < HEAVY_ITEM ? cnt - start : HEAVY_ITEM;
    items[first + p] = HeavyItem{b, offsets[b] + start, len, first + p};
  }
}

__global__ void __launch_bounds__(128)
fixed_accumulate_kernel(const Affine* __restrict__ table, const uint32_t* __restrict__ sorted,
                        const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                        uint32_t nbuckets, XYZZ* __restrict__ buckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  uint32_t cnt = counts[b];
  if (cnt > HEAVY_THRESHOLD) return;
  XYZZ acc = XYZZ::identity();
  const uint32_t* list = sorted + offsets[b];
  for (uint32_t k = 0; k < cnt; k++) acc = acc.add_affine(load_entry(table, list[k]));
  buckets[b] = acc;
}

__global__ void __launch_bounds__(HEAVY_THREADS)
fixed_heavy_accumulate_kernel(const Affine* __restrict__ table, const uint32_t* __restrict__ sorted,
                              const HeavyItem* __restrict__ items, const uint32_t* __restrict__ hcount,
                              XYZZ* __restrict__ partials) {
  __shared__ XYZZ sh[HEAVY_THREADS];
  for (uint32_t item = blockIdx.x; item < hcount[0]; item += gridDim.x) {
    HeavyItem it = items[item];
    XYZZ acc = XYZZ::identity();
    for (uint32_t k = threadIdx.x; k < it.len; k += HEAVY_THREADS)
      acc = acc.add_affine(load_entry(table, sorted[it.start + k]));
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int stride = HEAVY_THREADS / 2; stride > 0; stride >>= 1) {
      if ((int)threadIdx.x < stride) sh[threadIdx.x] = sh[threadIdx.x].add(sh[threadIdx.x + stride]);
      __syncthreads();
    }
    if (threadIdx.x == 0) partials[it.slot] = sh[0];
    __syncthreads();
  }
}
__global__ void fixed_heavy_finalize_kernel(const uint32_t* __restrict__ heavy_buckets,
                                            const uint32_t* __restrict__ hcount, const XYZZ* __restrict__ partials,
                                            XYZZ* __restrict__ buckets) {
  uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= hcount[1]) return;
  uint32_t b = heavy_buckets[3 * h], first = heavy_buckets[3 * h + 1], pieces = heavy_buckets[3 * h + 2];
  XYZZ acc = partials[first];
  for (uint32_t p = 1; p < pieces; p++) acc = acc.add(partials[first + p]);
  buckets[b] = acc;
}

// ---- reduction: W = sum_{b=1..B} b * bucket[b-1] -----------------------------------------------------
// level 1: per segment s of SEG buckets:  S_s = sum B,  A_s = sum (b_local + 1) B
__global__ void __launch_bounds__(128)
fixed_reduce_level1_kernel(const XYZZ* __restrict__ buckets, uint32_t nsegs, XYZZ* __restrict__ outA,
                           XYZZ* __restrict__ outS) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsegs) return;
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
// W = sum_s A_s + SEG * sum_s s * S_s,  and  sum_s s * S_s = sum_j 2^j * (sum_{s: bit j} S_s).
// Tree kernel: blockIdx.y = 0 sums all A_s; blockIdx.y = 1 + j sums the S_s with bit j of s set.
__global__ void __launch_bounds__(TREE_THREADS)
fixed_reduce_tree_kernel(const XYZZ* __restrict__ A, const XYZZ* __restrict__ S, uint32_t nsegs,
                         XYZZ* __restrict__ partials, uint32_t blocks_x) {
  __shared__ XYZZ sh[TREE_THREADS];
  const int which = blockIdx.y;
  const XYZZ* src = which == 0 ? A : S;
  const uint32_t bitmask = which == 0 ? 0 : (1u << (which - 1));
  uint32_t base = (blockIdx.x * TREE_THREADS + threadIdx.x) * TREE_PER_THREAD;
  XYZZ acc = XYZZ::identity();
  for (int k = 0; k < TREE_PER_THREAD; k++) {
    uint32_t s = base + k;
    if (s < nsegs && (which == 0 || (s & bitmask))) acc = acc.add(src[s]);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int stride = TREE_THREADS / 2; stride > 0; stride >>= 1) {
    if ((int)threadIdx.x < stride) sh[threadIdx.x] = sh[threadIdx.x].add(sh[threadIdx.x + stride]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[(size_t)which * blocks_x + blockIdx.x] = sh[0];
}
// second stage: one block per output sums its blocks_x partials
__global__ void __launch_bounds__(TREE_THREADS)
fixed_reduce_final_kernel(const XYZZ* __restrict__ partials, uint32_t blocks_x, XYZZ* __restrict__ out) {
  __shared__ XYZZ sh[TREE_THREADS];
  XYZZ acc = XYZZ::identity();
  for (uint32_t k = threadIdx.x; k < blocks_x; k += TREE_THREADS) acc = acc.add(partials[(size_t)blockIdx.x * blocks_x + k]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int stride = TREE_THREADS / 2; stride > 0; stride >>= 1) {
    if ((int)threadIdx.x < stride) sh[threadIdx.x] = sh[threadIdx.x].add(sh[threadIdx.x + stride]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

}  // namespace

int fixed_window_bits(uint64_t npoints) {
  if (npoints <= (1u << 10)) return 8;
  if (npoints <= (1u << 14)) return 12;
  if (npoints <= (1u << 17)) return 16;
  return 18;
}

int32_t fixed_base_build(zk_ctx* ctx, const Affine* d_bases, uint64_t npoints, FixedBase* out) {
  FixedBase fb;
  fb.c = fixed_window_bits(npoints);
  fb.nwin = (255 + fb.c - 1) / fb.c + ((255 % fb.c) == 0 ? 1 : 0);
  fb.npoints = npoints;
  if ((uint64_t)fb.nwin * npoints >= 0x7fffffffull) return set_error(ctx, ZK_E_INVALID, "msm table too large");
  ZK_CUDA(ctx, cudaMalloc((void**)&fb.table, (size_t)fb.nwin * npoints * sizeof(Affine)));
  ZK_CUDA(ctx, cudaMemcpyAsync(fb.table, d_bases, npoints * sizeof(Affine), cudaMemcpyDeviceToDevice, ctx->stream));
  for (int w = 1; w < fb.nwin; w++) {
    table_next_kernel<<<(unsigned)((npoints + 127) / 128), 128, 0, ctx->stream>>>(
        fb.table + (size_t)(w - 1) * npoints, fb.table + (size_t)w * npoints, npoints, fb.c);
    ctx->launches++;
  }
  ZK_CUDA(ctx, cudaGetLastError());
  *out = fb;
  return ZK_OK;
}

void fixed_base_free(FixedBase& fb) {
  cudaFree(fb.table);
  fb = FixedBase();
}

int32_t msm_fixed(zk_ctx* ctx, const FixedBase& fb, const Fp* d_scalars, uint64_t count, const Fp* extra_host,
                  const uint32_t* extra_index_host, int n_extra, XYZZ* result, uint32_t side_bit_mask,
                  int side_select) {
  if (count > fb.npoints || n_extra > 4) return set_error(ctx, ZK_E_INVALID, "msm_fixed: bad sizes");
  const uint64_t total = count + n_extra;
  if (total == 0) {
    *result = XYZZ::identity();
    return ZK_OK;
  }
  cudaStream_t st = ctx->stream;
  const int c = fb.c, nwin = fb.nwin;
  const uint32_t B = 1u << (c - 1);
  const uint32_t nsegs = B / SEG;
  int nbits = 0;
  while ((1u << nbits) < nsegs) nbits++;
  const uint32_t blocks_x = (nsegs + TREE_THREADS * TREE_PER_THREAD - 1) / (TREE_THREADS * TREE_PER_THREAD);
  const uint32_t nout = 1 + nbits;
  const uint32_t max_heavy = (uint32_t)((uint64_t)nwin * total / HEAVY_THRESHOLD + 64);
  const uint32_t ntiles = (B + SCAN_TILE - 1) / SCAN_TILE;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  size_t o_tmp = take((size_t)nwin * total * 4), o_counts = take((size_t)B * 4), o_offsets = take((size_t)B * 4),
         o_cursor = take((size_t)B * 4), o_tiles = take((size_t)ntiles * 4 + 16), o_sorted = take((size_t)nwin * total * 4),
         o_hitems = take((size_t)max_heavy * sizeof(HeavyItem)), o_hb = take((size_t)max_heavy * 12),
         o_hcount = take(16), o_hpart = take((size_t)max_heavy * sizeof(XYZZ)),
         o_buckets = take((size_t)B * sizeof(XYZZ)), o_A = take((size_t)nsegs * sizeof(XYZZ)),
         o_S = take((size_t)nsegs * sizeof(XYZZ)), o_part = take((size_t)nout * blocks_x * sizeof(XYZZ)),
         o_out = take((size_t)nout * sizeof(XYZZ)), o_extra = take(4 * sizeof(Fp)), o_eidx = take(64);
  int32_t rc = ensure_buf(ctx, ctx->msm_ws, off);
  if (rc) return rc;
  char* ws = (char*)ctx->msm_ws.ptr;
  uint32_t *tmp = (uint32_t*)(ws + o_tmp), *counts = (uint32_t*)(ws + o_counts), *offsets = (uint32_t*)(ws + o_offsets),
           *cursor = (uint32_t*)(ws + o_cursor), *tiles = (uint32_t*)(ws + o_tiles), *sorted = (uint32_t*)(ws + o_sorted),
           *hb = (uint32_t*)(ws + o_hb), *hcount = (uint32_t*)(ws + o_hcount), *eidx = (uint32_t*)(ws + o_eidx);
  HeavyItem* hitems = (HeavyItem*)(ws + o_hitems);
  XYZZ *hpart = (XYZZ*)(ws + o_hpart), *buckets = (XYZZ*)(ws + o_buckets), *A = (XYZZ*)(ws + o_A),
       *S = (XYZZ*)(ws + o_S), *part = (XYZZ*)(ws + o_part), *out = (XYZZ*)(ws + o_out);
  Fp* d_extra = (Fp*)(ws + o_extra);
  static const bool trace = getenv("ZK_MSM_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t0;
  if (trace) {
    cudaStreamSynchronize(st);
    t0 = std::chrono::steady_clock::now();
  }
  if (n_extra) {
    ZK_CUDA(ctx, cudaMemcpyAsync(d_extra, extra_host, n_extra * sizeof(Fp), cudaMemcpyHostToDevice, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(eidx, extra_index_host, n_extra * 4, cudaMemcpyHostToDevice, st));
  }
  ZK_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)B * 4, st));
  ZK_CUDA(ctx, cudaMemsetAsync(cursor, 0, (size_t)B * 4, st));
  ZK_CUDA(ctx, cudaMemsetAsync(hcount, 0, 16, st));
  {
    KernelTimer timer(ctx, KC_MSM);
    const int T = 256;
    const unsigned gt = (unsigned)((total + T - 1) / T);
    fixed_digits_kernel<<<gt, T, 0, st>>>(d_scalars, count, d_extra, eidx, n_extra, c, nwin, fb.npoints, tmp, counts,
                                          side_bit_mask, side_select);
    fscan_tiles_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(counts, offsets, tiles, B);
    fscan_sums_kernel<<<1, 1024, 0, st>>>(tiles, ntiles);
    fscan_add_kernel<<<(B + T - 1) / T, T, 0, st>>>(offsets, tiles, B);
    fixed_scatter_kernel<<<gt, T, 0, st>>>(tmp, count, n_extra, eidx, nwin, fb.npoints, offsets, cursor, sorted);
    fixed_find_heavy_kernel<<<(B + T - 1) / T, T, 0, st>>>(counts, offsets, B, hitems, hcount, max_heavy, hb);
    {
      KernelTimer acc_timer(ctx, KC_MSM_ACC);
      fixed_accumulate_kernel<<<(B + 127) / 128, 128, 0, st>>>(fb.table, sorted, counts, offsets, B, buckets);
      fixed_heavy_accumulate_kernel<<<ctx->sm_count * 4, HEAVY_THREADS, 0, st>>>(fb.table, sorted, hitems, hcount, hpart);
      fixed_heavy_finalize_kernel<<<(max_heavy + 63) / 64, 64, 0, st>>>(hb, hcount, hpart, buckets);
    }
    fixed_reduce_level1_kernel<<<(nsegs + 127) / 128, 128, 0, st>>>(buckets, nsegs, A, S);
    fixed_reduce_tree_kernel<<<dim3(blocks_x, nout), TREE_THREADS, 0, st>>>(A, S, nsegs, part, blocks_x);
    fixed_reduce_final_kernel<<<nout, TREE_THREADS, 0, st>>>(part, blocks_x, out);
    ctx->launches += 12;
  }
  ZK_CUDA(ctx, cudaGetLastError());
  XYZZ sums[40];
  uint32_t hc[4];
  ZK_CUDA(ctx, cudaMemcpyAsync(sums, out, (size_t)nout * sizeof(XYZZ), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(ctx, cudaMemcpyAsync(hc, hcount, 16, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(ctx, cudaStreamSynchronize(st));
  if (hc[2]) return set_error(ctx, ZK_E_NOMEM, "msm_fixed: heavy work list overflow");
  // W = sums[0] + SEG * sum_j 2^j sums[1 + j]
  XYZZ weighted = XYZZ::identity();
  for (int j = nbits - 1; j >= 0; j--) {
    weighted = weighted.dbl();
    weighted = weighted.add(sums[1 + j]);
  }
  for (int d = 0; (1 << d) < SEG; d++) weighted = weighted.dbl();
  *result = sums[0].add(weighted);
  if (trace) {
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[msm_fixed] count=%llu c=%d nwin=%d heavy_items=%u ms=%.3f\n", (unsigned long long)total, c, nwin,
            hc[0], ms);
  }
  return ZK_OK;
}

}  // namespace zkodst
