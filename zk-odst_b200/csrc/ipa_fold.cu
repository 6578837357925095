// K11 — materialising the folded IPA generators after r rounds.
//
// halo2_proofs 0.3.0 `poly::commitment::prover::create_proof` folds the generators every round
// (`parallel_generator_collapse`: G'_i = G_lo_i + [u] G_hi_i, n scalar multiplications per proof;
// reached from `create_proof`, blake2f-circuit/benches/blake2f.rs:125).  The device prover keeps the
// first r rounds on the ORIGINAL generators (prover.cu) and then builds the r-times folded ones
// directly,
//     H_i = sum_{q < 2^r} s_q * g_{q * len + i},      s_q = prod_{t < r} u_t^(bit_{r-1-t}(q)),
// for all len = n / 2^r outputs at once.  The 2^r scalars are shared by every output, so their
// signed 8-bit digits are computed and bucket-sorted ONCE on the host (<= 2048 entries); on the
// device a thread per (output i, bucket b) walks the same short list for a whole warp of adjacent
// outputs — no divergence, table reads T8[w][q * len + i] contiguous across the warp — followed by
// the usual running-sum bucket reduction, per output.  The remaining k - r rounds then run against a
// window table of H, at 2^-r of the cost of a round on the original generators.
#include "msm_fixed.h"
#include "prover_state.h"

namespace zkodst {
namespace {

constexpr int FOLD_C = 8, FOLD_WINDOWS = 32, FOLD_BUCKETS = 128, FOLD_SEG = 8, FOLD_NSEG = FOLD_BUCKETS / FOLD_SEG;

struct FoldLists {  // bucket b (|digit| = b + 1): entries[offset[b] .. offset[b + 1])
  uint32_t offset[FOLD_BUCKETS + 1];
};

// buckets[b][i] = sum over the bucket's entries (q, w, sign) of +-T8[w][q * len + i]
__global__ void __launch_bounds__(128)
fold_accumulate_kernel(const Affine* __restrict__ table, uint64_t table_width, FoldLists lists,
                       const uint32_t* __restrict__ entries, uint32_t len, uint32_t q_lo,
                       XYZZ* __restrict__ buckets) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t b = blockIdx.y;
  if (i >= len) return;
  XYZZ acc = XYZZ::identity();
  for (uint32_t e = lists.offset[b]; e < lists.offset[b + 1]; e++) {
    const uint32_t v = entries[e];  // q << 8 | w << 1 | sign
    const uint32_t q = v >> 8, w = (v >> 1) & 0x7f;
    Affine p = table[(size_t)w * table_width + (size_t)(q - q_lo) * len + i];  // the table starts at q_lo
    if (v & 1) p.y = p.y.neg();
    acc = acc.add_affine(p);
  }
  buckets[(size_t)b * len + i] = acc;
}
// per (output, segment of 8 buckets): S = sum B, A = sum (b_local + 1) B
__global__ void __launch_bounds__(128)
fold_segments_kernel(const XYZZ* __restrict__ buckets, uint32_t len, XYZZ* __restrict__ A, XYZZ* __restrict__ S) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t s = blockIdx.y;
  if (i >= len) return;
  XYZZ running = XYZZ::identity(), acc = XYZZ::identity();
#pragma unroll 1
  for (int b = FOLD_SEG - 1; b >= 0; b--) {
    running = running.add(buckets[(size_t)(s * FOLD_SEG + b) * len + i]);
    acc = acc.add(running);
  }
  A[(size_t)s * len + i] = acc;
  S[(size_t)s * len + i] = running;
}
// partial_i = sum_s A_s + 8 * sum_s s * S_s  (this rank's share of H_i)
__global__ void __launch_bounds__(128)
fold_finish_kernel(const XYZZ* __restrict__ A, const XYZZ* __restrict__ S, uint32_t len, XYZZ* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  XYZZ running = XYZZ::identity(), weighted = XYZZ::identity();
#pragma unroll 1
  for (int s = FOLD_NSEG - 1; s >= 1; s--) {
    running = running.add(S[(size_t)s * len + i]);
    weighted = weighted.add(running);
  }
  for (int d = 0; (1 << d) < FOLD_SEG; d++) weighted = weighted.dbl();
#pragma unroll 1
  for (int s = 0; s < FOLD_NSEG; s++) weighted = weighted.add(A[(size_t)s * len + i]);
  out[i] = weighted;
}
// H_i = sum over ranks of their partials, normalised to affine
__global__ void __launch_bounds__(128)
fold_combine_kernel(const XYZZ* __restrict__ parts, uint32_t nparts, uint32_t len, Affine* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  XYZZ acc = parts[i];
  for (uint32_t r = 1; r < nparts; r++) acc = acc.add(parts[(size_t)r * len + i]);
  out[i] = acc.to_affine();
}

// window table of `npoints` affine points in one launch: the doubling chain per point, then ONE
// inversion per point for all its windows (Montgomery's trick inside the thread)
__global__ void __launch_bounds__(128)
table_build_kernel(Affine* __restrict__ table, uint64_t npoints, int c, int nwin, XYZZ* __restrict__ tmp,
                   Fq* __restrict__ prefix) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= npoints) return;
  const Affine p = table[i];
  if (p.is_identity()) {
    for (int w = 1; w < nwin; w++) table[(size_t)w * npoints + i] = p;
    return;
  }
  XYZZ acc = XYZZ::from_affine(p);
  Fq run = Fq::one();
  for (int w = 1; w < nwin; w++) {
    for (int d = 0; d < c; d++) acc = acc.dbl();
    tmp[(size_t)w * npoints + i] = acc;
    run = run * acc.zzz;  // no point of the chain is the identity: the group order is odd
    prefix[(size_t)w * npoints + i] = run;
  }
  Fq inv = run.inv();
  for (int w = nwin - 1; w >= 1; w--) {
    const XYZZ q = tmp[(size_t)w * npoints + i];
    const Fq before = w > 1 ? prefix[(size_t)(w - 1) * npoints + i] : Fq::one();
    const Fq zi = inv * before;                 // 1 / ZZZ_w
    inv = inv * q.zzz;
    const Fq zz_inv = zi.sqr() * q.zz * q.zz;   // ZZ^2 / ZZZ^2 = 1 / ZZ
    table[(size_t)w * npoints + i] = Affine{q.x * zz_inv, q.y * zi};
  }
}

}  // namespace

// buckets, segment sums, the entry list, this rank's partials and (multi-GPU) every rank's partials
size_t ipa_fold_workspace_bytes(uint64_t len, int world) {
  return (size_t)(FOLD_BUCKETS + 2 * FOLD_NSEG + 1 + world) * len * sizeof(XYZZ) + (size_t)(1u << 16) * 4;
}

// out[0 .. len): the generators after r folds with challenges u[0 .. r)  (len = fb8.total_main >> r)
int32_t ipa_fold_generators(zk_ctx* ctx, const FixedBase& fb8, const Fp* u, int r, void* workspace, Affine* out) {
  if (fb8.c != FOLD_C || fb8.nwin != FOLD_WINDOWS || r < 1 || r > 8)
    return set_error(ctx, ZK_E_INVALID, "ipa_fold_generators: unsupported table or round count");
  const uint64_t n = fb8.total_main;
  const uint32_t len = (uint32_t)(n >> r);
  const uint32_t nq = 1u << r;
  const int world = ctx->dist_world;
  // this rank's table covers points [lo, lo + nmain) = q in [q_lo, q_hi)
  if (fb8.lo % len || fb8.nmain % len || (world == 1 && fb8.nmain != n))
    return set_error(ctx, ZK_E_INVALID, "ipa_fold_generators: point range is not a whole number of blocks");
  const uint32_t q_lo = (uint32_t)(fb8.lo / len), q_hi = (uint32_t)((fb8.lo + fb8.nmain) / len);
  // digits of the 2^r shared scalars, bucket-sorted on the host
  std::vector<uint32_t> bucket_of((size_t)nq * FOLD_WINDOWS), packed((size_t)nq * FOLD_WINDOWS);
  FoldLists lists;
  uint32_t counts[FOLD_BUCKETS] = {};
  for (uint32_t q = 0; q < nq; q++) {
    Fp s = Fp::one();
    for (int t = 0; t < r; t++)
      if ((q >> (r - 1 - t)) & 1) s = s * u[t];
    uint64_t cs[4];
    s.to_canonical(cs);
    uint32_t carry = 0;
    for (int w = 0; w < FOLD_WINDOWS; w++) {
      uint32_t d = (uint32_t)((cs[w >> 3] >> ((w & 7) * 8)) & 0xff) + carry;
      uint32_t e = 0xffffffffu;
      if (d > FOLD_BUCKETS) {
        carry = 1;
        const uint32_t mag = 256 - d;
        if (mag) e = ((mag - 1) << 1) | 1;
      } else {
        carry = 0;
        if (d) e = (d - 1) << 1;
      }
      if (q < q_lo || q >= q_hi) e = 0xffffffffu;  // another rank's points
      bucket_of[(size_t)q * FOLD_WINDOWS + w] = e;
      if (e != 0xffffffffu) counts[e >> 1]++;
    }
  }
  lists.offset[0] = 0;
  for (int b = 0; b < FOLD_BUCKETS; b++) lists.offset[b + 1] = lists.offset[b] + counts[b];
  uint32_t cursor[FOLD_BUCKETS];
  for (int b = 0; b < FOLD_BUCKETS; b++) cursor[b] = lists.offset[b];
  for (uint32_t q = 0; q < nq; q++)
    for (uint32_t w = 0; w < FOLD_WINDOWS; w++) {
      const uint32_t e = bucket_of[(size_t)q * FOLD_WINDOWS + w];
      if (e == 0xffffffffu) continue;
      packed[cursor[e >> 1]++] = (q << 8) | (w << 1) | (e & 1);
    }
  cudaStream_t st = ctx->stream;
  XYZZ* buckets = (XYZZ*)workspace;
  XYZZ* A = buckets + (size_t)FOLD_BUCKETS * len;
  XYZZ* S = A + (size_t)FOLD_NSEG * len;
  XYZZ* mine = S + (size_t)FOLD_NSEG * len;
  XYZZ* all = mine + len;
  uint32_t* d_entries = (uint32_t*)(all + (size_t)world * len);
  if (lists.offset[FOLD_BUCKETS])
    ZK_CUDA(ctx, cudaMemcpyAsync(d_entries, packed.data(), (size_t)lists.offset[FOLD_BUCKETS] * 4,
                                 cudaMemcpyHostToDevice, st));
  const unsigned bx = (len + 127) / 128;
  {
    KernelTimer timer(ctx, KC_COLLAPSE);
    fold_accumulate_kernel<<<dim3(bx, FOLD_BUCKETS), 128, 0, st>>>(fb8.table, fb8.npoints, lists, d_entries, len, q_lo,
                                                                  buckets);
    fold_segments_kernel<<<dim3(bx, FOLD_NSEG), 128, 0, st>>>(buckets, len, A, S);
    fold_finish_kernel<<<bx, 128, 0, st>>>(A, S, len, world > 1 ? mine : all);
    if (world > 1) {
      int32_t rc = dist_allgather_device(ctx, mine, all, (size_t)len * sizeof(XYZZ));
      if (rc) return rc;
    }
    fold_combine_kernel<<<bx, 128, 0, st>>>(all, (uint32_t)world, len, out);
    ctx->launches += 4;
  }
  ZK_CUDA(ctx, cudaGetLastError());
  return ZK_OK;
}

// Window table over `npoints` points already stored in storage[0 .. npoints); storage holds
// nwin * npoints points, tmp nwin * npoints XYZZ + nwin * npoints Fq.
int32_t fixed_base_build_inplace(zk_ctx* ctx, uint64_t total_main, uint64_t n_extra, int c, Affine* storage,
                                 void* tmp, FixedBase* out) {
  FixedBase fb;
  fb.c = c;
  fb.nwin = (255 + c - 1) / c + ((255 % c) == 0 ? 1 : 0);
  fb.lo = 0;
  fb.nmain = fb.total_main = total_main;
  fb.nextra = n_extra;
  fb.split = false;  // every rank holds the whole table: results are complete, not partial sums
  fb.npoints = total_main + n_extra;
  fb.table = storage;
  XYZZ* t = (XYZZ*)tmp;
  Fq* prefix = (Fq*)(t + (size_t)fb.nwin * fb.npoints);
  table_build_kernel<<<(unsigned)((fb.npoints + 127) / 128), 128, 0, ctx->stream>>>(storage, fb.npoints, c, fb.nwin, t,
                                                                                 prefix);
  ctx->launches++;
  ZK_CUDA(ctx, cudaGetLastError());
  *out = fb;
  return ZK_OK;
}

}  // namespace zkodst
