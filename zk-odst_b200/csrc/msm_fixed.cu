// K2/K3/K11 — batched fixed-base Pippenger MSM with precomputed window tables.
//
// Every commitment of the proof is an MSM over one of two fixed bases, g or g_lagrange
// (`Params::commit` / `Params::commit_lagrange`, halo2_proofs 0.3.0; reached from
// `create_proof`, blake2f-circuit/benches/blake2f.rs:125), and the IPA rounds are rewritten
// over the original g as well (prover.cu).  For a fixed base the multiples 2^(c w) * G_i are
// tabulated once per params (`FixedBase`), which turns the classic per-window bucket sets into
// ONE set of 2^(c-1) buckets shared by all windows:
//     sum_i s_i G_i = sum_b b * ( sum_{(i,w): |d_iw| = b} sgn(d_iw) * T[w][i] )
// so there is a single bucket reduction per MSM and no doubling chain.  Several MSMs over the
// same base (the 12 advice columns, the L/R pair of an IPA round, the h pieces) run as ONE
// pipeline over `nb * 2^(c-1)` buckets with one host synchronisation.
//
// Pipeline (async on the context's stream, one D2H of the partial sums at the end):
//   digits -> histogram -> scan -> scatter   counting sort of (job, w, i) by (job, |digit|)
//   plan                                       chunk length L = entries / resident threads: the
//                                              accumulation is ONE wave of equal-sized chunks
//   accumulate                                 thread per chunk of L consecutive sorted entries,
//                                              whatever the bucket sizes: the piece of the bucket
//                                              a chunk starts in goes to heads[chunk], buckets that
//                                              begin inside the chunk are written directly
//   fix-up                                     thread per bucket adds the heads of the chunks it
//                                              spills into; buckets spanning many chunks (skewed
//                                              small-value advice columns) go through a block-wide
//                                              tree reduction
//   reduce                                     sum_b b * B_b: running sums over segments of 8
//                                              buckets, then per-bit tree sums of the segment
//                                              totals (shallow dependency chains)
//   host                                       Horner over <= 16 partial sums per job
//
// Roofline: integer-pipe bound; see DESIGN.md §MSM for the MAC accounting.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <vector>

#include "msm_fixed.h"

namespace zkodst {
namespace {

constexpr int ACC_THREADS = 128;
constexpr int ACC_MIN_BLOCKS = 4;   // registers: 124 per thread
constexpr int MIN_CHUNK = 16;
constexpr int SERIAL_HEADS = 8;     // buckets spilling into more chunks than this take the heavy path
constexpr int PIECE = 2048;         // heads per heavy work item
constexpr int HEAVY_THREADS = 256;
constexpr int TREE_THREADS = 128;
// Bucket reduction shape (ZK_REDUCE_VARIANT): buckets per level-1 segment (serial running sums, 2 SEG additions
// deep) and segment sums per thread of the tree kernel.  Measured at k = 19, MSM ms per proof on one stream
// (profiles/r02_reduce_variants.json): {8, 8} 36.09, {4, 4} 35.51, {4, 8} 35.49 (default), {8, 4} 35.76.
struct ReduceShape {
  int seg, per_thread;
};
constexpr ReduceShape REDUCE_SHAPES[4] = {{8, 8}, {4, 4}, {4, 8}, {8, 4}};

struct HeavyItem {
  uint32_t start, len;
};
struct HeavyBucket {
  uint32_t bucket, first, pieces, own;
};
struct Plan {
  uint32_t entries, chunk, nchunks, pad;
  uint32_t heavy_items, heavy_buckets, overflow, pad2;
};

struct DevJobs {
  const Fp* scalars[MSM_MAX_BATCH];
  uint32_t side_mask[MSM_MAX_BATCH];
  int32_t side_select[MSM_MAX_BATCH];
  int32_t n_extra[MSM_MAX_BATCH];
  // Support mask of one index bit (the L / R vectors of an IPA round): the job's threads enumerate only the kept
  // indices of this rank's range — thread t handles the (kept_before + t)-th kept index of the whole vector —
  // instead of visiting every index and dropping half.  kept = how many there are in [lo, lo + count).
  uint64_t kept_before[MSM_MAX_BATCH];
  uint64_t kept[MSM_MAX_BATCH];
};
// indices with ((index & mask) != 0) == select, mask one bit: how many lie below x, and the j-th of them
__host__ __device__ inline uint64_t kept_below(uint64_t x, uint64_t mask, bool select) {
  const uint64_t r = x & (2 * mask - 1), start = select ? mask : 0;
  const uint64_t in_block = r <= start ? 0 : (r - start < mask ? r - start : mask);
  return (x / (2 * mask)) * mask + in_block;
}
__host__ __device__ inline uint64_t kept_index(uint64_t j, uint64_t mask, bool select) {
  return (j / mask) * 2 * mask + (j & (mask - 1)) + (select ? mask : 0);
}

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
// warp-aggregated increment: lanes with the same key elect a leader that adds the group size
__device__ __forceinline__ uint32_t aggregated_add(uint32_t* counters, uint32_t key, bool active) {
  uint32_t base = 0;
  if (active) {
    const uint32_t peers = __match_any_sync(__activemask(), key);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    if (lane == leader) base = atomicAdd(&counters[key], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    base += __popc(peers & ((1u << lane) - 1));
  }
  return base;
}

// entries_tmp[((job * nwin + w) * stride) + t] = bucket + 1 | sign << 31   (0 = no entry);
// ranks_tmp[same index] = the entry's position inside its bucket: the value the histogram increment returned,
// so the scatter needs no second round of atomics
__global__ void fixed_digits_kernel(DevJobs jobs, uint64_t count, uint64_t lo, const Fp* __restrict__ extra, int c,
                                    int nwin, uint64_t stride, uint32_t B, uint32_t* __restrict__ entries_tmp,
                                    uint32_t* __restrict__ ranks_tmp, uint32_t* __restrict__ counts) {
  const int job = blockIdx.y;
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t kept = jobs.kept[job];
  if (t >= kept + jobs.n_extra[job]) return;
  uint64_t s[4];
  bool zero = false;
  if (t < kept) {
    const uint32_t mask = jobs.side_mask[job];
    const uint64_t index = mask ? kept_index(jobs.kept_before[job] + t, mask, jobs.side_select[job] != 0) : lo + t;
    Fp v = jobs.scalars[job][index];
    zero = v.is_zero();
    if (!zero) v.to_canonical(s);
  } else {
    extra[job * 4 + (t - kept)].to_canonical(s);
  }
  uint32_t* my_counts = counts + (size_t)job * B;
  uint32_t* my_tmp = entries_tmp + (size_t)job * nwin * stride + t;
  uint32_t* my_rank = ranks_tmp + (size_t)job * nwin * stride + t;
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
        if (mag) e = mag | 0x80000000u;
      } else {
        carry = 0;
        e = d;
      }
    }
    my_tmp[(size_t)w * stride] = e;
    // (measured and dropped, profiles/r02_msm_helper_variants.json: plain atomicAdd for jobs whose digits are
    // uniform — the match / shuffle round looks like pure latency in ncu, short_scoreboard 12.5 — is 0.5 ms per
    // proof SLOWER: an atomic that returns its old value pays an L2 round trip per lane either way)
    const uint32_t rank = aggregated_add(my_counts, (e & 0x7fffffffu) - 1, e != 0);
    if (e) my_rank[(size_t)w * stride] = rank;
  }
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
// offsets += tile offsets; the last thread also writes the sentinel offsets[n] and the plan
__global__ void fscan_add_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ tile_sums,
                                 const uint32_t* __restrict__ counts, uint32_t n, uint32_t resident_threads,
                                 uint32_t min_chunk, Plan* __restrict__ plan) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t v = out[i] + tile_sums[i / SCAN_TILE];
  out[i] = v;
  if (i == n - 1) {
    const uint32_t entries = v + counts[i];
    out[n] = entries;
    uint32_t chunk = (entries + resident_threads - 1) / resident_threads;
    // small MSMs (the IPA's second stage: 2^14 points, 2,048 buckets of ~180 entries) would get 16-entry chunks
    // and send every bucket down the heavy path (more than SERIAL_HEADS chunks per bucket, 68 us per launch):
    // keep an average bucket within ~6 chunks instead
    const uint32_t per_bucket = entries / n;
    if (chunk < per_bucket / 6) chunk = per_bucket / 6;
    chunk = (chunk + 3) & ~3u;
    if (chunk < min_chunk) chunk = min_chunk;
    Plan p;
    p.entries = entries;
    p.chunk = chunk;
    p.nchunks = (entries + chunk - 1) / chunk;
    p.pad = 0;
    p.heavy_items = p.heavy_buckets = p.overflow = p.pad2 = 0;
    *plan = p;
  }
}

// scatter: sorted[pos] = table entry index (w * npoints + point) | sign << 31
__global__ void fixed_scatter_kernel(DevJobs jobs, const uint32_t* __restrict__ entries_tmp,
                                     const uint32_t* __restrict__ ranks_tmp, uint64_t lo,
                                     const uint32_t* __restrict__ extra_index, int nwin, uint64_t stride,
                                     uint64_t npoints, uint32_t B, const uint32_t* __restrict__ offsets,
                                     uint32_t* __restrict__ sorted) {
  const int job = blockIdx.y;
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t kept = jobs.kept[job];
  if (t >= kept + jobs.n_extra[job]) return;
  const uint32_t mask = jobs.side_mask[job];
  // local indices (table rows of this rank's range)
  const uint32_t point = t >= kept ? extra_index[job * 4 + (t - kept)]
                         : mask    ? (uint32_t)(kept_index(jobs.kept_before[job] + t, mask, jobs.side_select[job] != 0) - lo)
                                   : (uint32_t)t;
  const uint32_t* my_tmp = entries_tmp + (size_t)job * nwin * stride + t;
  const uint32_t* my_rank = ranks_tmp + (size_t)job * nwin * stride + t;
  const uint32_t* my_offsets = offsets + (size_t)job * B;
  // (measured and dropped: four windows in flight per thread — loads, offset gathers, stores grouped — 0.1 ms
  // per proof slower; the kernel is bound by the 32-byte sector each scattered 4-byte store dirties)
  for (int w = 0; w < nwin; w++) {
    const uint32_t e = my_tmp[(size_t)w * stride];
    if (e)
      sorted[my_offsets[(e & 0x7fffffffu) - 1] + my_rank[(size_t)w * stride]] =
          (uint32_t)((uint64_t)w * npoints + point) | (e & 0x80000000u);
  }
}

__device__ __forceinline__ Affine load_entry(const Affine* __restrict__ table, uint32_t v) {
  Affine p = table[v & 0x7fffffffu];
  if (v & 0x80000000u) p.y = p.y.neg();
  return p;
}

// the bucket containing sorted position `pos`: first b in [lo, hi] with offsets[b + 1] > pos
__device__ __forceinline__ uint32_t bucket_of_position(const uint32_t* __restrict__ offsets, uint32_t lo,
                                                       uint32_t hi, uint32_t pos) {
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (offsets[mid + 1] <= pos) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---- experimental accumulation variants (ZK_ACC_VARIANT, see msm_fixed_batch) -----------------------------
// The next point of a thread's chunk is staged in shared memory instead of registers, which frees 16 registers
// per thread for a fifth resident block per SM.  STAGE = 1: per-thread cp.async (LDGSTS) of the 64-byte table
// row, 16-byte pieces laid out [buffer][piece][thread] (conflict-free reads); STAGE = 2: one bulk asynchronous
// copy per row (cp.async.bulk, the TMA engine; completion on a per-thread mbarrier), rows laid out
// [buffer][thread][64 B].  Both double-buffered: the copy of entry pos + 2 is issued when entry pos has been read.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int STAGE, int MINB>
__global__ void __launch_bounds__(ACC_THREADS, MINB)
fixed_accumulate_staged_kernel(const Affine* __restrict__ table, const uint32_t* __restrict__ sorted,
                               const uint32_t* __restrict__ offsets, uint32_t nbuckets, const Plan* __restrict__ plan,
                               XYZZ* __restrict__ heads, XYZZ* __restrict__ buckets) {
  __shared__ __align__(128) uint4 stage[2 * 4 * ACC_THREADS];
  __shared__ __align__(8) uint64_t bars[2 * ACC_THREADS];
  const uint32_t tid = threadIdx.x;
  const uint32_t entries = plan->entries, L = plan->chunk;
  uint32_t phase[2] = {0, 0};
  if (STAGE == 2) {
    for (int b = 0; b < 2; b++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bars[b * ACC_THREADS + tid])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  auto slot = [&](int buf, int piece) -> uint4* {
    return STAGE == 1 ? &stage[(buf * 4 + piece) * ACC_THREADS + tid] : &stage[(buf * ACC_THREADS + tid) * 4 + piece];
  };
  auto issue = [&](int buf, uint32_t v) {
    const uint4* src = reinterpret_cast<const uint4*>(table + (v & 0x7fffffffu));
    if (STAGE == 1) {
#pragma unroll
      for (int j = 0; j < 4; j++)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(slot(buf, j))), "l"(src + j) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
      const uint32_t bar = smem_addr(&bars[buf * ACC_THREADS + tid]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 64;" ::"r"(bar) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];" ::"r"(
                       smem_addr(slot(buf, 0))),
                   "l"(src), "r"(bar)
                   : "memory");
    }
  };
  auto wait_for = [&](int buf, bool newest) {
    if (STAGE == 1) {
      if (newest) asm volatile("cp.async.wait_group 0;" ::: "memory");
      else asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      const uint32_t bar = smem_addr(&bars[buf * ACC_THREADS + tid]);
      uint32_t done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar), "r"(phase[buf])
                     : "memory");
      phase[buf] ^= 1;
    }
  };
  auto take = [&](int buf, uint32_t v) {
    Affine p;
    uint4 q[4];
#pragma unroll
    for (int j = 0; j < 4; j++) q[j] = *slot(buf, j);
    p.x.l[0] = (uint64_t)q[0].x | ((uint64_t)q[0].y << 32); p.x.l[1] = (uint64_t)q[0].z | ((uint64_t)q[0].w << 32);
    p.x.l[2] = (uint64_t)q[1].x | ((uint64_t)q[1].y << 32); p.x.l[3] = (uint64_t)q[1].z | ((uint64_t)q[1].w << 32);
    p.y.l[0] = (uint64_t)q[2].x | ((uint64_t)q[2].y << 32); p.y.l[1] = (uint64_t)q[2].z | ((uint64_t)q[2].w << 32);
    p.y.l[2] = (uint64_t)q[3].x | ((uint64_t)q[3].y << 32); p.y.l[3] = (uint64_t)q[3].z | ((uint64_t)q[3].w << 32);
    if (v & 0x80000000u) p.y = p.y.neg();
    return p;
  };
  for (uint32_t chunk = blockIdx.x * ACC_THREADS + tid; chunk < plan->nchunks; chunk += gridDim.x * ACC_THREADS) {
    const uint32_t start = chunk * L;
    const uint32_t end = start + L < entries ? start + L : entries;
    uint32_t b = bucket_of_position(offsets, 0, nbuckets - 1, start);
    uint32_t boundary = offsets[b + 1];
    bool first = true;
    XYZZ acc = XYZZ::identity();
    uint32_t v0 = sorted[start], v1 = start + 1 < end ? sorted[start + 1] : 0;
    issue(0, v0);
    if (start + 1 < end) issue(1, v1);
    for (uint32_t pos = start; pos < end; pos++) {
      const int buf = (pos - start) & 1;
      wait_for(buf, pos + 1 >= end);
      const Affine cur = take(buf, v0);
      v0 = v1;
      if (pos + 2 < end) {
        v1 = sorted[pos + 2];
        issue(buf, v1);
      }
      if (pos >= boundary) {
        if (first) heads[chunk] = acc; else buckets[b] = acc;
        first = false;
        acc = XYZZ::identity();
        b++;
        boundary = offsets[b + 1];
        if (pos >= boundary) {
          b = bucket_of_position(offsets, b + 1, nbuckets - 1, pos);
          boundary = offsets[b + 1];
        }
      }
      acc = acc.add_affine(cur);
    }
    if (first) heads[chunk] = acc; else buckets[b] = acc;
  }
}

// ---- accumulation: one chunk of `plan->chunk` consecutive sorted entries per thread -------------------
__global__ void __launch_bounds__(ACC_THREADS, ACC_MIN_BLOCKS)
fixed_accumulate_kernel(const Affine* __restrict__ table, const uint32_t* __restrict__ sorted,
                        const uint32_t* __restrict__ offsets, uint32_t nbuckets, const Plan* __restrict__ plan,
                        XYZZ* __restrict__ heads, XYZZ* __restrict__ buckets) {
  const uint32_t entries = plan->entries, L = plan->chunk;
  for (uint32_t chunk = blockIdx.x * ACC_THREADS + threadIdx.x; chunk < plan->nchunks;
       chunk += gridDim.x * ACC_THREADS) {
    const uint32_t start = chunk * L;
    const uint32_t end = start + L < entries ? start + L : entries;
    uint32_t b = bucket_of_position(offsets, 0, nbuckets - 1, start);
    uint32_t boundary = offsets[b + 1];
    bool first = true;
    XYZZ acc = XYZZ::identity();
    Affine nxt = load_entry(table, sorted[start]);
    for (uint32_t pos = start; pos < end; pos++) {
      const Affine cur = nxt;
      if (pos + 1 < end) nxt = load_entry(table, sorted[pos + 1]);
      if (pos >= boundary) {
        if (first) heads[chunk] = acc; else buckets[b] = acc;
        first = false;
        acc = XYZZ::identity();
        // next non-empty bucket: usually b + 1; small-value columns leave long runs of empty
        // buckets, which a linear walk would cross one dependent load at a time
        b++;
        boundary = offsets[b + 1];
        if (pos >= boundary) {
          b = bucket_of_position(offsets, b + 1, nbuckets - 1, pos);
          boundary = offsets[b + 1];
        }
      }
      acc = acc.add_affine(cur);
    }
    if (first) heads[chunk] = acc; else buckets[b] = acc;
  }
}

// ---- fix-up: bucket = (piece written by the chunk it starts in) + heads of the chunks it spills into ---
__global__ void __launch_bounds__(128)
fixed_fixup_kernel(const uint32_t* __restrict__ offsets, uint32_t nbuckets, Plan* __restrict__ plan,
                   const XYZZ* __restrict__ heads, XYZZ* __restrict__ buckets, HeavyItem* __restrict__ items,
                   HeavyBucket* __restrict__ hbuckets, uint32_t max_items, uint32_t max_hbuckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  const uint32_t s = offsets[b], e = offsets[b + 1];
  if (s == e) {
    buckets[b] = XYZZ::identity();
    return;
  }
  const uint32_t L = plan->chunk;
  const uint32_t t_first = s / L, t_last = (e - 1) / L;
  const bool own = (s % L) != 0;
  const uint32_t h0 = own ? t_first + 1 : t_first;
  const uint32_t nheads = t_last + 1 - h0;
  if (nheads == 0) return;
  if (nheads <= SERIAL_HEADS) {
    XYZZ acc = own ? buckets[b] : heads[h0];
    for (uint32_t t = own ? h0 : h0 + 1; t <= t_last; t++) acc = acc.add(heads[t]);
    buckets[b] = acc;
    return;
  }
  const uint32_t pieces = (nheads + PIECE - 1) / PIECE;
  const uint32_t first = atomicAdd(&plan->heavy_items, pieces);
  const uint32_t hb = atomicAdd(&plan->heavy_buckets, 1u);
  if (first + pieces > max_items || hb >= max_hbuckets) {
    plan->overflow = 1;
    return;
  }
  hbuckets[hb] = HeavyBucket{b, first, pieces, own ? 1u : 0u};
  for (uint32_t p = 0; p < pieces; p++) {
    const uint32_t st = p * PIECE;
    items[first + p] = HeavyItem{h0 + st, nheads - st < PIECE ? nheads - st : PIECE};
  }
}
__global__ void __launch_bounds__(HEAVY_THREADS)
fixed_heavy_kernel(const XYZZ* __restrict__ heads, const HeavyItem* __restrict__ items,
                   const Plan* __restrict__ plan, XYZZ* __restrict__ partials) {
  __shared__ XYZZ sh[HEAVY_THREADS];
  const uint32_t nitems = plan->overflow ? 0 : plan->heavy_items;
  for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const HeavyItem it = items[item];
    XYZZ acc = XYZZ::identity();
    for (uint32_t k = threadIdx.x; k < it.len; k += HEAVY_THREADS) acc = acc.add(heads[it.start + k]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int stride = HEAVY_THREADS / 2; stride > 0; stride >>= 1) {
      if ((int)threadIdx.x < stride) sh[threadIdx.x] = sh[threadIdx.x].add(sh[threadIdx.x + stride]);
      __syncthreads();
    }
    if (threadIdx.x == 0) partials[item] = sh[0];
    __syncthreads();
  }
}
__global__ void fixed_heavy_finalize_kernel(const HeavyBucket* __restrict__ hbuckets, const Plan* __restrict__ plan,
                                            const XYZZ* __restrict__ partials, XYZZ* __restrict__ buckets) {
  uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  if (plan->overflow || h >= plan->heavy_buckets) return;
  const HeavyBucket hb = hbuckets[h];
  XYZZ acc = hb.own ? buckets[hb.bucket] : XYZZ::identity();
  for (uint32_t p = 0; p < hb.pieces; p++) acc = acc.add(partials[hb.first + p]);
  buckets[hb.bucket] = acc;
}

// ---- reduction: W = sum_{b=1..B} b * bucket[b-1], per job (blockIdx.z / .y selects the job) -----------
// level 1: per segment s of SEG buckets:  S_s = sum B,  A_s = sum (b_local + 1) B
template <int SEG>
__global__ void __launch_bounds__(128)
fixed_reduce_level1_kernel(const XYZZ* __restrict__ buckets, uint32_t nsegs_total, XYZZ* __restrict__ outA,
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
// W = sum_s A_s + SEG * sum_s s * S_s,  and  sum_s s * S_s = sum_j 2^j * (sum_{s: bit j} S_s).
// Tree kernel: blockIdx.y = 0 sums all A_s; blockIdx.y = 1 + j sums the S_s with bit j of s set.
template <int TREE_PER_THREAD>
__global__ void __launch_bounds__(TREE_THREADS)
fixed_reduce_tree_kernel(const XYZZ* __restrict__ A, const XYZZ* __restrict__ S, uint32_t nsegs,
                         XYZZ* __restrict__ partials, uint32_t blocks_x, uint32_t nout) {
  __shared__ XYZZ sh[TREE_THREADS];
  const int which = blockIdx.y, job = blockIdx.z;
  const XYZZ* src = (which == 0 ? A : S) + (size_t)job * nsegs;
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
  if (threadIdx.x == 0) partials[((size_t)job * nout + which) * blocks_x + blockIdx.x] = sh[0];
}
// second stage: one warp per (output, job) sums its blocks_x partials (strided, then a 5-level tree)
__global__ void __launch_bounds__(32)
fixed_reduce_final_kernel(const XYZZ* __restrict__ partials, uint32_t blocks_x, XYZZ* __restrict__ out) {
  __shared__ XYZZ sh[32];
  const size_t o = blockIdx.x;
  XYZZ acc = XYZZ::identity();
  for (uint32_t k = threadIdx.x; k < blocks_x; k += 32) acc = acc.add(partials[o * blocks_x + k]);
  sh[threadIdx.x] = acc;
  __syncwarp();
  for (int stride = 16; stride > 0; stride >>= 1) {
    if ((int)threadIdx.x < stride && threadIdx.x + stride < blocks_x) sh[threadIdx.x] = sh[threadIdx.x].add(sh[threadIdx.x + stride]);
    __syncwarp();
  }
  if (threadIdx.x == 0) out[o] = sh[0];
}

int window_bits_override() {
  static const int v = [] {
    const char* e = getenv("ZK_MSM_C");
    return e ? atoi(e) : 0;
  }();
  return v;
}

}  // namespace

int fixed_window_bits(uint64_t npoints) {
  const int o = window_bits_override();
  if (o >= 4 && o <= 20) return o;
  if (npoints <= (1u << 10)) return 8;
  if (npoints <= (1u << 14)) return 12;
  return 16;
}

int32_t fixed_base_build(zk_ctx* ctx, const Affine* d_bases, uint64_t total_main, uint64_t n_extra,
                         FixedBase* out, int force_c) {
  FixedBase fb;
  fb.c = force_c ? force_c : fixed_window_bits(total_main + n_extra);
  fb.nwin = (255 + fb.c - 1) / fb.c + ((255 % fb.c) == 0 ? 1 : 0);
  uint64_t hi = 0;
  dist_range(total_main, ctx->dist_rank, ctx->dist_world, &fb.lo, &hi);
  fb.nmain = hi - fb.lo;
  fb.total_main = total_main;
  fb.nextra = ctx->dist_rank == 0 ? n_extra : 0;
  fb.split = ctx->dist_world > 1;
  fb.npoints = fb.nmain + fb.nextra;
  const uint64_t npoints = fb.npoints;
  if ((uint64_t)fb.nwin * npoints >= 0x7fffffffull) return set_error(ctx, ZK_E_INVALID, "msm table too large");
  ZK_CUDA(ctx, cudaMalloc((void**)&fb.table, (size_t)fb.nwin * npoints * sizeof(Affine)));
  ZK_CUDA(ctx, cudaMemcpyAsync(fb.table, d_bases + fb.lo, fb.nmain * sizeof(Affine), cudaMemcpyDeviceToDevice,
                               ctx->stream));
  if (fb.nextra)
    ZK_CUDA(ctx, cudaMemcpyAsync(fb.table + fb.nmain, d_bases + total_main, fb.nextra * sizeof(Affine),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
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

int32_t msm_fixed_batch(zk_ctx* ctx, const FixedBase& fb, const MsmJob* jobs, int nb, uint64_t global_count,
                        XYZZ* results) {
  if (nb < 1 || nb > MSM_MAX_BATCH || global_count > fb.total_main)
    return set_error(ctx, ZK_E_INVALID, "msm_fixed: bad sizes");
  for (int m = 0; m < nb; m++) {
    if (jobs[m].n_extra < 0 || jobs[m].n_extra > 4) return set_error(ctx, ZK_E_INVALID, "msm_fixed: bad extras");
    if (jobs[m].side_mask & (jobs[m].side_mask - 1))
      return set_error(ctx, ZK_E_INVALID, "msm_fixed: the index mask must be one bit");
    for (int e = 0; e < jobs[m].n_extra; e++)
      if (jobs[m].extra_index[e] < fb.total_main) return set_error(ctx, ZK_E_INVALID, "msm_fixed: extra index");
  }
  // this rank's share of the terms: global indices [lo, lo + count)
  const uint64_t count = global_count <= fb.lo ? 0 : std::min<uint64_t>(global_count - fb.lo, fb.nmain);
  cudaStream_t st = ctx->stream;
  const int c = fb.c, nwin = fb.nwin;
  const uint32_t B = 1u << (c - 1);
  const uint32_t NB = (uint32_t)nb * B;
  const uint64_t stride = count + 4;
  const uint64_t max_entries = (uint64_t)nb * nwin * stride;
  if (max_entries >= 0xffffffffull) return set_error(ctx, ZK_E_INVALID, "msm_fixed: batch too large");
  // Variant switch (measured: profiles/r02_accumulate_variants.json, 12 full-width columns at 2^19):
  // 1 = cp.async staging at 5 resident blocks, the product path (15.6 ms); 0 = register prefetch at 4 blocks
  // (16.1 ms, round 1); 2 = cp.async.bulk (TMA) + mbarrier staging at 5 blocks (17.0 ms: one bulk copy and one
  // mbarrier round trip per 64-byte row cost more than four LDGSTS); 11 / 12 = the staged forms at 4 blocks.
  static const int acc_variant = [] {
    const char* e = getenv("ZK_ACC_VARIANT");
    return e ? atoi(e) : 1;
  }();
  typedef void (*AccKernel)(const Affine*, const uint32_t*, const uint32_t*, uint32_t, const Plan*, XYZZ*, XYZZ*);
  const AccKernel acc_kernel = acc_variant == 1    ? fixed_accumulate_staged_kernel<1, 5>
                               : acc_variant == 2  ? fixed_accumulate_staged_kernel<2, 5>
                               : acc_variant == 11 ? fixed_accumulate_staged_kernel<1, 4>
                               : acc_variant == 12 ? fixed_accumulate_staged_kernel<2, 4>
                                                   : fixed_accumulate_kernel;
  if (!ctx->msm_acc_blocks_per_sm) {  // per context: no state shared between host threads
    int v = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, acc_kernel, ACC_THREADS, 0);
    ctx->msm_acc_blocks_per_sm = v > 0 ? v : ACC_MIN_BLOCKS;
  }
  const uint32_t acc_grid = (uint32_t)ctx->sm_count * (uint32_t)ctx->msm_acc_blocks_per_sm;
  const uint32_t resident = acc_grid * ACC_THREADS;
  // chunk = max(MIN_CHUNK, entries / resident) so there are never more chunks than resident threads
  // Small MSMs (the IPA rounds on the folded generators, < 2^20 entries) are bound by their chains of dependent
  // point additions, not by throughput: 2 segment sums per tree thread instead of 8 (profiles/r02_ipa_stage2_sweep.jsonl:
  // 7.18 -> 6.97 ms over the 14 rounds of a k = 19 proof; an 8-entry minimum chunk changed nothing and stays out).
  // ZK_SMALL_MIN_CHUNK / ZK_SMALL_TREE_PER override for tools/ipa_sweep.py.
  static const int small_min_chunk = [] { const char* e = getenv("ZK_SMALL_MIN_CHUNK"); return e ? atoi(e) : MIN_CHUNK; }();
  static const int small_tree_per = [] { const char* e = getenv("ZK_SMALL_TREE_PER"); return e ? atoi(e) : 2; }();
  const bool small = max_entries < (1u << 20);
  const uint32_t min_chunk = small ? (uint32_t)small_min_chunk : (uint32_t)MIN_CHUNK;
  const uint32_t max_chunks = (uint32_t)std::min<uint64_t>(max_entries / min_chunk + 1, (uint64_t)resident + 1);
  const uint32_t max_hbuckets = max_chunks / (SERIAL_HEADS + 1) + 16;
  const uint32_t max_items = max_chunks / PIECE + max_hbuckets + 16;
  static const int reduce_variant = [] {
    const char* e = getenv("ZK_REDUCE_VARIANT");
    const int v = e ? atoi(e) : 2;   // 4-bucket segments, 8 segment sums per tree thread
    return v >= 0 && v < 4 ? v : 2;
  }();
  const int SEG = REDUCE_SHAPES[reduce_variant].seg;
  const int TREE_PER_THREAD = small && small_tree_per ? small_tree_per : REDUCE_SHAPES[reduce_variant].per_thread;
  const uint32_t nsegs = B / SEG;
  int nbits = 0;
  while ((1u << nbits) < nsegs) nbits++;
  const uint32_t blocks_x = (nsegs + TREE_THREADS * TREE_PER_THREAD - 1) / (TREE_THREADS * TREE_PER_THREAD);
  const uint32_t nout = 1 + nbits;
  const uint32_t ntiles = (NB + SCAN_TILE - 1) / SCAN_TILE;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  const size_t o_tmp = take(max_entries * 4), o_ranks = take(max_entries * 4), o_counts = take((size_t)NB * 4), o_offsets = take((size_t)(NB + 1) * 4),
               o_tiles = take((size_t)ntiles * 4 + 16), o_sorted = take(max_entries * 4 + 16),
               o_plan = take(sizeof(Plan)), o_heads = take((size_t)max_chunks * sizeof(XYZZ)),
               o_hitems = take((size_t)max_items * sizeof(HeavyItem)),
               o_hb = take((size_t)max_hbuckets * sizeof(HeavyBucket)),
               o_hpart = take((size_t)max_items * sizeof(XYZZ)), o_buckets = take((size_t)NB * sizeof(XYZZ)),
               o_A = take((size_t)nb * nsegs * sizeof(XYZZ)), o_S = take((size_t)nb * nsegs * sizeof(XYZZ)),
               o_part = take((size_t)nb * nout * blocks_x * sizeof(XYZZ)), o_out = take((size_t)nb * nout * sizeof(XYZZ)),
               o_extra = take((size_t)MSM_MAX_BATCH * 4 * sizeof(Fp)), o_eidx = take((size_t)MSM_MAX_BATCH * 4 * 4);
  int32_t rc = ensure_buf(ctx, ctx->msm_ws, off);
  if (rc) return rc;
  char* ws = (char*)ctx->msm_ws.ptr;
  uint32_t *tmp = (uint32_t*)(ws + o_tmp), *counts = (uint32_t*)(ws + o_counts), *offsets = (uint32_t*)(ws + o_offsets),
           *ranks = (uint32_t*)(ws + o_ranks), *tiles = (uint32_t*)(ws + o_tiles), *sorted = (uint32_t*)(ws + o_sorted),
           *eidx = (uint32_t*)(ws + o_eidx);
  Plan* plan = (Plan*)(ws + o_plan);
  HeavyItem* hitems = (HeavyItem*)(ws + o_hitems);
  HeavyBucket* hb = (HeavyBucket*)(ws + o_hb);
  XYZZ *heads = (XYZZ*)(ws + o_heads), *hpart = (XYZZ*)(ws + o_hpart), *buckets = (XYZZ*)(ws + o_buckets),
       *A = (XYZZ*)(ws + o_A), *S = (XYZZ*)(ws + o_S), *part = (XYZZ*)(ws + o_part), *out = (XYZZ*)(ws + o_out);
  Fp* d_extra = (Fp*)(ws + o_extra);
  static const bool trace = getenv("ZK_MSM_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t0;
  if (trace) {
    zk_stream_sync(ctx);
    t0 = std::chrono::steady_clock::now();
  }
  DevJobs dj;
  Fp h_extra[MSM_MAX_BATCH * 4];
  uint32_t h_eidx[MSM_MAX_BATCH * 4];
  bool any_extra = false;
  uint64_t most = 4;  // threads of the widest job: digits and scatter grids
  for (int m = 0; m < MSM_MAX_BATCH; m++) {
    const MsmJob& j = jobs[m < nb ? m : 0];
    dj.scalars[m] = j.scalars;
    dj.side_mask[m] = j.side_mask;
    dj.side_select[m] = j.side_select;
    dj.kept_before[m] = j.side_mask ? kept_below(fb.lo, j.side_mask, j.side_select != 0) : 0;
    dj.kept[m] = j.side_mask ? kept_below(fb.lo + count, j.side_mask, j.side_select != 0) - dj.kept_before[m] : count;
    if (m < nb && dj.kept[m] + 4 > most) most = dj.kept[m] + 4;
    const int ne = (m < nb && fb.nextra) ? j.n_extra : 0;  // extras live on rank 0 only
    dj.n_extra[m] = ne;
    for (int e = 0; e < 4; e++) {
      h_extra[m * 4 + e] = e < ne ? j.extra[e] : Fp::zero();
      h_eidx[m * 4 + e] = e < ne ? (uint32_t)(j.extra_index[e] - fb.total_main + fb.nmain) : 0;
    }
    if (ne) any_extra = true;
  }
  if (any_extra) {
    // pageable copies: the runtime stages them before returning, so the stack arrays may go away
    ZK_CUDA(ctx, cudaMemcpyAsync(d_extra, h_extra, sizeof(h_extra), cudaMemcpyHostToDevice, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(eidx, h_eidx, sizeof(h_eidx), cudaMemcpyHostToDevice, st));
  }
  ZK_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)NB * 4, st));
  {
    KernelTimer timer(ctx, KC_MSM);
    const int T = 256;
    const dim3 gt((unsigned)((most + T - 1) / T), (unsigned)nb);
    fixed_digits_kernel<<<gt, T, 0, st>>>(dj, count, fb.lo, d_extra, c, nwin, stride, B, tmp, ranks, counts);
    fscan_tiles_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(counts, offsets, tiles, NB);
    fscan_sums_kernel<<<1, 1024, 0, st>>>(tiles, ntiles);
    fscan_add_kernel<<<(NB + T - 1) / T, T, 0, st>>>(offsets, tiles, counts, NB, resident, min_chunk, plan);
    fixed_scatter_kernel<<<gt, T, 0, st>>>(dj, tmp, ranks, fb.lo, eidx, nwin, stride, fb.npoints, B, offsets, sorted);
    {
      KernelTimer acc_timer(ctx, KC_MSM_ACC);
      acc_kernel<<<acc_grid, ACC_THREADS, 0, st>>>(fb.table, sorted, offsets, NB, plan, heads, buckets);
      fixed_fixup_kernel<<<(NB + 127) / 128, 128, 0, st>>>(offsets, NB, plan, heads, buckets, hitems, hb, max_items,
                                                           max_hbuckets);
      fixed_heavy_kernel<<<ctx->sm_count * 2, HEAVY_THREADS, 0, st>>>(heads, hitems, plan, hpart);
      fixed_heavy_finalize_kernel<<<(max_hbuckets + 63) / 64, 64, 0, st>>>(hb, plan, hpart, buckets);
    }
    const uint32_t nsegs_total = (uint32_t)nb * nsegs;
    if (SEG == 8) fixed_reduce_level1_kernel<8><<<(nsegs_total + 127) / 128, 128, 0, st>>>(buckets, nsegs_total, A, S);
    else fixed_reduce_level1_kernel<4><<<(nsegs_total + 127) / 128, 128, 0, st>>>(buckets, nsegs_total, A, S);
    const dim3 tree_grid(blocks_x, nout, (unsigned)nb);
    if (TREE_PER_THREAD == 8) fixed_reduce_tree_kernel<8><<<tree_grid, TREE_THREADS, 0, st>>>(A, S, nsegs, part, blocks_x, nout);
    else if (TREE_PER_THREAD == 2) fixed_reduce_tree_kernel<2><<<tree_grid, TREE_THREADS, 0, st>>>(A, S, nsegs, part, blocks_x, nout);
    else if (TREE_PER_THREAD == 1) fixed_reduce_tree_kernel<1><<<tree_grid, TREE_THREADS, 0, st>>>(A, S, nsegs, part, blocks_x, nout);
    else fixed_reduce_tree_kernel<4><<<tree_grid, TREE_THREADS, 0, st>>>(A, S, nsegs, part, blocks_x, nout);
    fixed_reduce_final_kernel<<<(unsigned)nb * nout, 32, 0, st>>>(part, blocks_x, out);
    ctx->launches += 12;
  }
  ZK_CUDA(ctx, cudaGetLastError());
  std::vector<XYZZ> sums((size_t)nb * nout);
  Plan hp;
  ZK_CUDA(ctx, cudaMemcpyAsync(sums.data(), out, sums.size() * sizeof(XYZZ), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(ctx, cudaMemcpyAsync(&hp, plan, sizeof(Plan), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  if (hp.overflow) return set_error(ctx, ZK_E_NOMEM, "msm_fixed: heavy work list overflow");
  for (int m = 0; m < nb; m++) {
    const XYZZ* sm = &sums[(size_t)m * nout];
    // W = sums[0] + SEG * sum_j 2^j sums[1 + j]
    XYZZ weighted = XYZZ::identity();
    for (int j = nbits - 1; j >= 0; j--) {
      weighted = weighted.dbl();
      weighted = weighted.add(sm[1 + j]);
    }
    for (int d = 0; (1 << d) < SEG; d++) weighted = weighted.dbl();
    results[m] = sm[0].add(weighted);
  }
  if (ctx->dist_world > 1 && fb.split) {
    int32_t drc = dist_sum_points(ctx, results, nb);
    if (drc) return drc;
  }
  if (trace) {
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[msm_fixed] jobs=%d count=%llu c=%d nwin=%d entries=%u chunk=%u heavy_items=%u ms=%.3f\n", nb,
            (unsigned long long)count, c, nwin, hp.entries, hp.chunk, hp.heavy_items, ms);
  }
  return ZK_OK;
}

int32_t msm_fixed(zk_ctx* ctx, const FixedBase& fb, const Fp* d_scalars, uint64_t count, const Fp* extra_host,
                  const uint32_t* extra_index_host, int n_extra, XYZZ* result, uint32_t side_bit_mask,
                  int side_select) {
  if (n_extra < 0 || n_extra > 4) return set_error(ctx, ZK_E_INVALID, "msm_fixed: bad sizes");
  if (count + n_extra == 0) {
    *result = XYZZ::identity();
    return ZK_OK;
  }
  MsmJob j;
  j.scalars = d_scalars;
  j.n_extra = n_extra;
  for (int e = 0; e < n_extra; e++) {
    j.extra[e] = extra_host[e];
    j.extra_index[e] = extra_index_host[e];
  }
  j.side_mask = side_bit_mask;
  j.side_select = side_select;
  return msm_fixed_batch(ctx, fb, &j, 1, count, result);
}

}  // namespace zkodst
