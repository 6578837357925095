// create_proof on the device.
//
// Replaces halo2_proofs 0.3.0 `plonk::create_proof` (Pasta / IPA) as the reference calls it:
//   create_proof(&params, &pk, &[circuit], &[&[]], rng, &mut Blake2bWrite<_, _, Challenge255<_>>)
// (blake2f-circuit/benches/blake2f.rs:124-127).  The host keeps only the byte-serial parts —
// the BLAKE2b transcript, challenge derivation and the seeded RNG stream — and drives kernels
// for everything else: witness (K1), commitments (K2/K3), NTTs (K4/K5), lookup permutation
// (K7), grand products (K8), the quotient (K6), evaluations (K9), multiopen (K10) and the inner
// product argument (K11).  The order of transcript writes, challenge squeezes and RNG draws is
// halo2's; proof bytes are compared with the CPU oracle byte for byte in tests/.
#include <nvtx3/nvToolsExt.h>

#include <chrono>
#include <cstdlib>

#include "polyops.cuh"
#include "prover_state.h"
#include "quotient.h"
#include "transcript.h"
#include "xorshift_jump.h"

namespace zkodst {

// ---- device buffers of one proof ---------------------------------------------------------------------
struct ProofWorkspace {
  uint64_t n = 0, en = 0;
  uint8_t* inputs = nullptr;
  uint64_t* digests = nullptr;
  // The 19 witness columns live in three slot arrays (values, coefficients, coset values) so that their
  // transforms run as batches and a multi-GPU group can all-gather the results in place: slots 0..11
  // advice, 12 permuted input, 13 permuted table, 14..17 permutation products, 18 lookup product
  // (coefficient and coset arrays padded to a multiple of the group size).
  Fp* values_all = nullptr;     // [19][n]
  Fp* advice_values = nullptr;  // = values_all ([12][n])
  Fp* polys_all = nullptr;      // [slots][n]
  Fp* cosets_all = nullptr;     // [slots][en]
  Fp* advice_polys = nullptr;   // = polys_all   ([12][n])
  Fp* advice_cosets = nullptr;  // = cosets_all  ([12][en])
  Fp *cin = nullptr, *ctab = nullptr, *pin = nullptr, *ptab = nullptr;  // lookup, n each
  Fp *pin_poly = nullptr, *ptab_poly = nullptr, *zl_poly = nullptr;
  Fp *pin_coset = nullptr, *ptab_coset = nullptr, *zl_coset = nullptr;
  Fp* z_poly[NUM_SETS] = {};
  Fp* z_vals[NUM_SETS + 1] = {};  // grand products (values) of the permutation sets, then the lookup
  Fp* z_coset[NUM_SETS] = {};
  Fp *tmp_a = nullptr, *tmp_b = nullptr, *tmp_c = nullptr;  // n each
  Fp *h = nullptr, *h_coeffs = nullptr;                      // en each
  Fp *random_poly = nullptr, *s_poly = nullptr, *q_prime = nullptr, *p_poly = nullptr, *b_vec = nullptr;
  Fp* q_polys[8] = {};
  Fp* h_poly = nullptr;
  // IPA second stage (allocated on first use): folding scratch, table over the folded generators
  void* fold_ws = nullptr;
  Affine* h_table = nullptr;
  void* h_tmp = nullptr;
  // lookup permutation tables
  Fp* table_vals = nullptr;      // 65536 compressed table values
  Fp* table_sorted = nullptr;    // ascending
  uint32_t* rank_of = nullptr;   // dense index -> rank
  uint32_t* counts = nullptr;    // per rank: number of input rows
  uint32_t* offsets = nullptr;   // exclusive scan of counts
  uint32_t* left_cnt = nullptr;  // per rank: leftover table multiplicity
  uint32_t* left_off = nullptr;  // exclusive scan of left_cnt
  uint32_t* first_flag_scan = nullptr;  // n: number of first-occurrence positions before p
  uint32_t* left_rank = nullptr;        // n: leftover list (rank per entry), ascending
  Fp* pow_tabs = nullptr;               // group only: [4][n / world] powers of the opening points (multiopen)
  std::vector<void*> all;
};

constexpr int SLOT_PIN = 12, SLOT_PTAB = 13, SLOT_Z0 = 14, SLOT_ZL = 18, NUM_WITNESS_POLYS = ZK_NUM_WITNESS_COLUMNS;
// slots of the coefficient / coset arrays when the transforms are sharded by column over `world` ranks
static inline uint64_t witness_slots_padded(int world) {
  uint32_t lo, hi, per;
  dist_column_block(NUM_WITNESS_POLYS, 0, world, &lo, &hi, &per);
  return (uint64_t)per * world;
}

void free_workspace(void* p) {
  ProofWorkspace* W = (ProofWorkspace*)p;
  for (void* q : W->all) cudaFree(q);
  delete W;
}

namespace {

template <class T>
int32_t dalloc(zk_ctx* ctx, ProofWorkspace* W, T** p, size_t count) {
  ZK_CUDA(ctx, cudaMalloc((void**)p, count * sizeof(T)));
  W->all.push_back(*p);
  return ZK_OK;
}

int32_t get_workspace(zk_ctx* ctx, DeviceKeys& K, ProofWorkspace** out) {
  if (!K.workspace) {
    ProofWorkspace* W = new ProofWorkspace();
    K.workspace = W;
    const uint64_t n = K.n, en = K.en;
    W->n = n;
    W->en = en;
    int32_t rc = 0;
#define A(p, cnt) if ((rc = dalloc(ctx, W, &W->p, (cnt)))) return rc
    A(inputs, K.n_compressions * 213 + 16);
    A(digests, K.n_compressions * 8 + 8);
    A(values_all, NUM_WITNESS_POLYS * n);
    W->advice_values = W->values_all;
    W->pin = W->values_all + SLOT_PIN * n;
    W->ptab = W->values_all + SLOT_PTAB * n;
    for (int s = 0; s <= NUM_SETS; s++) W->z_vals[s] = W->values_all + (SLOT_Z0 + s) * n;
    const uint64_t slots = witness_slots_padded(ctx->dist_world);
    A(polys_all, slots * n);
    A(cosets_all, slots * en);
    W->advice_polys = W->polys_all;
    W->advice_cosets = W->cosets_all;
    W->pin_poly = W->polys_all + SLOT_PIN * n;
    W->ptab_poly = W->polys_all + SLOT_PTAB * n;
    W->zl_poly = W->polys_all + SLOT_ZL * n;
    W->pin_coset = W->cosets_all + SLOT_PIN * en;
    W->ptab_coset = W->cosets_all + SLOT_PTAB * en;
    W->zl_coset = W->cosets_all + SLOT_ZL * en;
    for (int s = 0; s < NUM_SETS; s++) {
      W->z_poly[s] = W->polys_all + (SLOT_Z0 + s) * n;
      W->z_coset[s] = W->cosets_all + (SLOT_Z0 + s) * en;
    }
    A(cin, n); A(ctab, n);
    A(tmp_a, n); A(tmp_b, n); A(tmp_c, n);
    A(h, en); A(h_coeffs, en);
    A(random_poly, n); A(s_poly, n); A(q_prime, n); A(p_poly, n); A(b_vec, n); A(h_poly, n);
    for (int s = 0; s < 8; s++) A(q_polys[s], n);
    A(table_vals, 65536); A(table_sorted, 65536);
    A(rank_of, 65536); A(counts, 65536 + 8); A(offsets, 65536 + 8); A(left_cnt, 65536 + 8); A(left_off, 65536 + 8);
    A(first_flag_scan, n + 8); A(left_rank, n + 8);
    if (dist_ranges_ok(n, ctx->dist_world)) A(pow_tabs, 4 * (n / (uint64_t)ctx->dist_world));
#undef A
  }
  *out = (ProofWorkspace*)K.workspace;
  return ZK_OK;
}

int32_t ensure_fold_workspace(zk_ctx* ctx, ProofWorkspace* W, uint64_t len) {
  if (W->fold_ws) return ZK_OK;
  const int c2 = ipa_stage2_c();
  const int nwin = (255 + c2 - 1) / c2 + ((255 % c2) == 0 ? 1 : 0);
  const size_t pts = (size_t)nwin * (len + 2);
  ZK_CUDA(ctx, cudaMalloc(&W->fold_ws, ipa_fold_workspace_bytes(len, ctx->dist_world)));
  W->all.push_back(W->fold_ws);
  ZK_CUDA(ctx, cudaMalloc((void**)&W->h_table, pts * sizeof(Affine)));
  W->all.push_back(W->h_table);
  ZK_CUDA(ctx, cudaMalloc(&W->h_tmp, pts * (sizeof(XYZZ) + sizeof(Fq))));
  W->all.push_back(W->h_tmp);
  return ZK_OK;
}

// ---- prover randomness: the seeded XorShift stream, sequential on the host for the ~170 scalar draws of
// a proof, generated on the device (jump-ahead, xorshift_jump.h) for the two n-element random polynomials
constexpr int XS_FIELDS_PER_THREAD = 64;        // 1024 stream words per thread
constexpr int XS_THREAD_SHIFT = 10;             // log2(16 * XS_FIELDS_PER_THREAD)

__global__ void __launch_bounds__(128)
xorshift_fields_kernel(XsState base, const XsMatrix* __restrict__ pow2, uint64_t t_lo, uint64_t n,
                       Fp* __restrict__ out) {
  const uint64_t t = t_lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;  // stream position / 1024 words
  const uint64_t first = t * XS_FIELDS_PER_THREAD;
  if (first >= n) return;
  XsState st = base;
  for (int b = 0; (t >> b) != 0; b++)
    if ((t >> b) & 1) st = xs_apply(pow2[XS_THREAD_SHIFT + b], st);
  const uint64_t last = first + XS_FIELDS_PER_THREAD < n ? first + XS_FIELDS_PER_THREAD : n;
  for (uint64_t i = first; i < last; i++) {
    uint64_t w[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint64_t lo = xs_step(st);
      w[j] = lo | ((uint64_t)xs_step(st) << 32);
    }
    out[i] = Fp::from_u512(w);
  }
}

class ProofRng {
 public:
  ProofRng(zk_ctx* ctx, const uint8_t seed[16]) : ctx_(ctx) {
    memcpy(st_.s, seed, 16);
    if (!(st_.s[0] | st_.s[1] | st_.s[2] | st_.s[3])) st_.s[0] = st_.s[1] = st_.s[2] = st_.s[3] = 0x0BAD5EED;
  }
  Fp next() {  // Field::random: eight next_u64 (low word first) -> from_u512
    uint64_t w[8];
    for (int j = 0; j < 8; j++) {
      const uint64_t lo = xs_step(st_);
      w[j] = lo | ((uint64_t)xs_step(st_) << 32);
    }
    return Fp::from_u512(w);
  }
  // the next n field elements of the stream, produced on the device: out[lo .. lo + cnt) only (a rank of a group
  // that works by coefficient range needs no more; lo and cnt whole threads' worth), the stream advances by n
  int32_t fill_device(uint64_t n, Fp* out, uint64_t lo = 0, uint64_t cnt = ~0ull) {
    if (cnt == ~0ull) cnt = n - lo;
    if (lo % XS_FIELDS_PER_THREAD || (lo + cnt != n && cnt % XS_FIELDS_PER_THREAD))
      return set_error(ctx_, ZK_E_INVALID, "fill_device: range not aligned");
    int32_t rc = ensure_buf(ctx_, ctx_->misc_ws, sizeof(XsMatrix) * XS_JUMP_POWERS);
    if (rc) return rc;
    if (!ctx_->xs_table_loaded) {
      ZK_CUDA(ctx_, cudaMemcpyAsync(ctx_->misc_ws.ptr, xs_jump_table(), sizeof(XsMatrix) * XS_JUMP_POWERS,
                                    cudaMemcpyHostToDevice, ctx_->stream));
      ctx_->xs_table_loaded = true;
    }
    const uint64_t threads = (cnt + XS_FIELDS_PER_THREAD - 1) / XS_FIELDS_PER_THREAD;
    if (threads)
      xorshift_fields_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx_->stream>>>(
          st_, (const XsMatrix*)ctx_->misc_ws.ptr, lo / XS_FIELDS_PER_THREAD, lo + cnt, out);
    ctx_->launches++;
    ZK_CUDA(ctx_, cudaGetLastError());
    st_ = xs_jump(st_, n * 16);
    return ZK_OK;
  }

 private:
  zk_ctx* ctx_;
  XsState st_;
};

int32_t upload_fp(zk_ctx* ctx, Fp* dst, const Fp* src, size_t count) {
  ZK_CUDA(ctx, cudaMemcpyAsync(dst, src, count * sizeof(Fp), cudaMemcpyHostToDevice, ctx->stream));
  return ZK_OK;
}

// ---- phases of one proof: NVTX ranges (always; a no-op without a profiler attached) and an optional
// wall-clock trace (ZK_PHASE_TRACE=1: syncs at phase boundaries, stderr only) -----------------------------------
static const char* const PROOF_PHASES[] = {"witness", "advice commit + iNTT", "lookup permute + commit",
                                           "permutation products", "lookup product", "random poly commit",
                                           "cosets + quotient + iNTT", "h commit", "evaluations", "multiopen", "ipa"};
struct PhaseTrace {
  bool on;
  cudaStream_t st;
  std::chrono::steady_clock::time_point t;
  size_t next = 0;  // index of the phase that is running
  explicit PhaseTrace(cudaStream_t s) : on(getenv("ZK_PHASE_TRACE") != nullptr), st(s) {
    if (on) {
      cudaStreamSynchronize(st);
      t = std::chrono::steady_clock::now();
    }
    nvtxRangePushA(PROOF_PHASES[0]);
  }
  ~PhaseTrace() {
    if (next < sizeof PROOF_PHASES / sizeof *PROOF_PHASES) nvtxRangePop();  // error return inside a phase
  }
  // trace only: elapsed time of a step inside the running phase
  void note(const char* name) {
    if (!on) return;
    cudaStreamSynchronize(st);
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[phase]   . %-26s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
  // ends the running phase (whose name is passed for the trace) and starts the next one
  void mark(const char* name) {
    nvtxRangePop();
    if (++next < sizeof PROOF_PHASES / sizeof *PROOF_PHASES) nvtxRangePushA(PROOF_PHASES[next]);
    if (!on) return;
    cudaStreamSynchronize(st);
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[phase] %-28s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

// ---- lookup argument helpers (A.5 permute_expression_pair, specialised to this circuit) ------------
// The lookup inputs are (tag, dense, spread) triples; a valid row equals table row `dense`, so the
// theta-compressed input of a row is table_vals[dense].  Sorting the inputs therefore reduces to a
// counting sort on the rank of that table value, and the permuted table follows from per-rank
// multiplicities.  Rows whose triple is not a table row raise the error flag
// (halo2: Error::ConstraintSystemFailure).
__global__ void lookup_compress_kernel(const Fp* __restrict__ c0, const Fp* __restrict__ c1,
                                       const Fp* __restrict__ c2, Fp theta, uint64_t n, Fp* __restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (c0[i] * theta + c1[i]) * theta + c2[i];
}
// Ascending order of the 65536 compressed table rows (field `Ord` = canonical integer order).
// The values are uniform, so a bucket by the top 14 bits of the canonical integer (< 2^254) holds
// ~4 of them: histogram, scan, then each value ranks itself against its bucket with a full
// 256-bit comparison.  Equal values raise status bit 4 (theta collision).
constexpr uint32_t RANK_BUCKETS = 1u << 14;
__global__ void table_keys_kernel(const Fp* __restrict__ vals, uint64_t* __restrict__ keys,
                                  uint32_t* __restrict__ hist) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 65536) return;
  uint64_t c[4];
  vals[i].to_canonical(c);
#pragma unroll
  for (int l = 0; l < 4; l++) keys[(size_t)l * 65536 + i] = c[l];  // limb-major
  atomicAdd(&hist[(uint32_t)(c[3] >> 48)], 1u);
}
__global__ void table_members_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ offsets,
                                     uint32_t* __restrict__ cursor, uint32_t* __restrict__ members) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 65536) return;
  const uint32_t bkt = (uint32_t)(keys[(size_t)3 * 65536 + i] >> 48);
  members[offsets[bkt] + atomicAdd(&cursor[bkt], 1u)] = i;
}
__global__ void table_rank_kernel(const uint64_t* __restrict__ keys, const Fp* __restrict__ vals,
                                  const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ hist,
                                  const uint32_t* __restrict__ members, uint32_t* __restrict__ rank_of,
                                  Fp* __restrict__ sorted, int* __restrict__ status) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 65536) return;
  uint64_t mine[4];
#pragma unroll
  for (int l = 0; l < 4; l++) mine[l] = keys[(size_t)l * 65536 + i];
  const uint32_t bkt = (uint32_t)(mine[3] >> 48);
  const uint32_t first = offsets[bkt], cnt = hist[bkt];
  uint32_t rank = first;
  for (uint32_t j = 0; j < cnt; j++) {
    const uint32_t o = members[first + j];
    if (o == i) continue;
    bool less = false, equal = true;
    for (int l = 3; l >= 0 && equal; l--) {
      const uint64_t a = keys[(size_t)l * 65536 + o];
      if (a != mine[l]) {
        equal = false;
        less = a < mine[l];
      }
    }
    if (equal) atomicOr(status, 4);
    rank += less ? 1u : 0u;
  }
  rank_of[i] = rank;
  sorted[rank < 65536 ? rank : 0] = vals[i];
}
__global__ void lookup_count_kernel(const Fp* __restrict__ dense_col, const Fp* __restrict__ cin,
                                    const Fp* __restrict__ table_vals, const uint32_t* __restrict__ rank_of,
                                    uint64_t usable, uint32_t* __restrict__ counts, int* __restrict__ status) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= usable) return;
  uint64_t d[4];
  dense_col[i].to_canonical(d);
  if ((d[1] | d[2] | d[3]) || d[0] >= 65536 || cin[i] != table_vals[d[0]]) {
    atomicOr(status, 2);
    return;
  }
  atomicAdd(&counts[rank_of[d[0]]], 1u);
}
// per rank: leftover multiplicity = table multiplicity - [rank used by an input]
__global__ void lookup_leftover_kernel(const uint32_t* __restrict__ counts, uint32_t zero_rank,
                                       uint32_t zero_mult, uint32_t* __restrict__ left_cnt) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= 65536) return;
  uint32_t mult = r == zero_rank ? zero_mult : 1;
  left_cnt[r] = mult - (counts[r] ? 1 : 0);
}
// permuted input: run of counts[r] copies of table_sorted[r]; first flags for the table pass
__global__ void lookup_fill_input_kernel(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                                         const Fp* __restrict__ table_sorted, Fp* __restrict__ pin,
                                         uint32_t* __restrict__ first_flag) {
  uint32_t r = blockIdx.x;
  uint32_t cnt = counts[r], off = offsets[r];
  Fp v = table_sorted[r];
  for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) {
    pin[off + j] = v;
    first_flag[off + j] = j == 0 ? 1 : 0;
  }
}
// leftover list: rank r occupies [left_off[r], left_off[r] + left_cnt[r])
__global__ void lookup_fill_leftover_kernel(const uint32_t* __restrict__ left_cnt,
                                            const uint32_t* __restrict__ left_off,
                                            uint32_t* __restrict__ left_rank) {
  uint32_t r = blockIdx.x;
  uint32_t cnt = left_cnt[r], off = left_off[r];
  for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) left_rank[off + j] = r;
}
// permuted table: first occurrences take the input value; the m-th repeated row (ascending)
// takes leftover[total - 1 - m]  (BTreeMap ascending order popped onto the highest rows first)
__global__ void lookup_fill_table_kernel(const Fp* __restrict__ pin, const uint32_t* __restrict__ first_flag,
                                         const uint32_t* __restrict__ first_scan,
                                         const uint32_t* __restrict__ left_rank,
                                         const Fp* __restrict__ table_sorted, uint64_t usable,
                                         uint32_t n_repeated, Fp* __restrict__ ptab) {
  uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (p >= usable) return;
  if (first_flag[p]) {
    ptab[p] = pin[p];
  } else {
    uint32_t m = (uint32_t)p - first_scan[p];  // repeated rows before p
    ptab[p] = table_sorted[left_rank[n_repeated - 1 - m]];
  }
}

// u32 exclusive scan (single block, up to a few million entries) used by the lookup pass
__global__ void scan_u32_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n,
                                uint32_t* __restrict__ total) {
  __shared__ uint32_t ws[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < n; base += blockDim.x * 4) {
    uint32_t i0 = base + threadIdx.x * 4;
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      v[k] = i0 + k < n ? in[i0 + k] : 0;
      sum += v[k];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    uint32_t off = carry;
    for (int w = 0; w < warp; w++) off += ws[w];
    uint32_t excl = off + incl - sum;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i0 + k < n) out[i0 + k] = excl;
      excl += v[k];
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

// z[0] = init, z[i + 1] = z[i] * ratio[i] over n rows.  In a group working by row range (polyops.cuh
// dist_ranges_ok) every rank scans its own rows [lo, lo + cnt) from 1, the range products are exchanged (one
// field element per rank), the rows are scaled by the product of the ranges before them and the ranges are
// all-gathered in place: the maps, the inversion and the scan cost 1 / world of the replicated form.
int32_t grand_product(zk_ctx* ctx, const Fp* ratio, uint64_t n, uint64_t lo, uint64_t cnt, const Fp& init, Fp* z) {
  if (cnt == n) return affine_scan(ctx, ratio, Fp::zero(), nullptr, n, init, z);
  int32_t rc = affine_scan(ctx, ratio + lo, Fp::zero(), nullptr, cnt, Fp::one(), z + lo);
  if (rc) return rc;
  if ((rc = ensure_buf(ctx, ctx->eval_ws, 64 * sizeof(Fp)))) return rc;
  Fp* d_tot = (Fp*)ctx->eval_ws.ptr;
  const uint64_t last = lo + cnt - 1;
  launch_map(ctx, 1, [=] __device__(uint64_t) { d_tot[0] = z[last] * ratio[last]; });
  std::vector<Fp> tot((size_t)ctx->dist_world);
  if ((rc = dist_gather_fields(ctx, d_tot, 1, tot.data()))) return rc;
  Fp prefix = init;
  for (int q = 0; q < ctx->dist_rank; q++) prefix = prefix * tot[q];
  launch_map_range(ctx, lo, cnt, [=] __device__(uint64_t i) { z[i] = z[i] * prefix; });
  return dist_allgather_device(ctx, z + lo, z, cnt * sizeof(Fp));
}

}  // namespace

// ---- the prover ------------------------------------------------------------------------------------------
static int32_t create_proof_impl(zk_ctx* ctx, const uint8_t* inputs, bool inputs_on_device,
                                 uint64_t n_compressions, const uint8_t seed[16],
                                 std::vector<uint8_t>& proof_out) {
  ProverState* S = prover_state(ctx);
  if (!S->has_params || !S->has_keys) return set_error(ctx, ZK_E_STATE, "create_proof before params/keygen");
  DeviceKeys& K = S->keys;
  const DeviceParams& P = S->params;
  if (n_compressions != K.n_compressions) return set_error(ctx, ZK_E_INVALID, "batch size differs from keygen");
  for (uint64_t i = 0; i < n_compressions && !inputs_on_device; i++) {
    const uint8_t* r = inputs + i * ZK_BLAKE2F_INPUT_BYTES;
    uint32_t rr = ((uint32_t)r[0] << 24) | ((uint32_t)r[1] << 16) | ((uint32_t)r[2] << 8) | r[3];
    if (r[212] > 1) return set_error(ctx, ZK_E_INPUT, "final-block flag must be 0 or 1");
    if (rr != K.rounds) return set_error(ctx, ZK_E_INPUT, "record rounds differ from circuit rounds");
  }
  ProofWorkspace* W = nullptr;
  int32_t rc = get_workspace(ctx, K, &W);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const uint64_t n = K.n, en = K.en;
  const int k = K.k;
  const uint64_t usable = n - (BLINDING + 1);
  // A group shards the row- and coefficient-wise steps that are not MSMs or transforms by contiguous range
  // (grand products, multiopen, the evaluations): rank r works on [r_lo, r_lo + r_cnt), the same range its
  // share of every MSM covers; a single GPU (or a group size that does not divide the blocks) works on [0, n).
  const bool by_range = dist_ranges_ok(n, ctx->dist_world);
  const uint64_t r_cnt = by_range ? n / (uint64_t)ctx->dist_world : n;
  const uint64_t r_lo = by_range ? r_cnt * (uint64_t)ctx->dist_rank : 0;
  NttOptions inv;
  inv.inverse = true;
  const unsigned T = 256;
  auto blocks = [&](uint64_t cnt) { return (unsigned)((cnt + T - 1) / T); };

  ProofRng tape(ctx, seed);
  TranscriptWriter tr;
  tr.common_scalar(K.transcript_repr);

  PhaseTrace phase(st);
  // a status bit may only come from this call (an earlier asynchronous zk_blake2f_witness_batch_device that was
  // never followed by zk_ctx_synchronize must not fail this proof)
  ZK_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(int), st));
  // ---- witness (K1) ---------------------------------------------------------------------------------
  if (n_compressions)
    ZK_CUDA(ctx, cudaMemcpyAsync(W->inputs, inputs, n_compressions * 213,
                                 inputs_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  if ((rc = launch_witness(ctx, k, K.rounds, W->inputs, n_compressions, W->advice_values, W->digests))) return rc;
  auto adv = [&](int c) { return W->advice_values + (size_t)c * n; };
  auto adv_poly = [&](int c) { return W->advice_polys + (size_t)c * n; };
  auto adv_coset = [&](int c) { return W->advice_cosets + (size_t)c * en; };

  phase.mark("witness");
  // ---- advice blinding rows, blinds, commitments -----------------------------------------------------
  {
    Fp tails[12 * 6];
    for (int i = 0; i < 12 * 6; i++) tails[i] = tape.next();
    for (int c = 0; c < 12; c++)
      if ((rc = upload_fp(ctx, adv(c) + usable, tails + 6 * c, 6))) return rc;
  }
  Fp advice_blinds[12];
  for (auto& b : advice_blinds) b = tape.next();
  {
    const Fp* cols[12];
    Affine cms[12];
    for (int c = 0; c < 12; c++) cols[c] = adv(c);
    if ((rc = commit_batch(ctx, cols, P.fb_gl, n, advice_blinds, 12, cms))) return rc;
    for (int c = 0; c < 12; c++) tr.write_point(cms[c]);
  }
  const Fp theta = tr.squeeze_challenge();

  phase.mark("advice commit + iNTT");
  // ---- lookup: compress, permute, commit (K7) ----------------------------------------------------------
  Fp pin_blind, ptab_blind, zl_blind;
  {
    lookup_compress_kernel<<<blocks(n), T, 0, st>>>(adv(7), adv(8), adv(9), theta, n, W->cin);
    lookup_compress_kernel<<<blocks(n), T, 0, st>>>(K.fixed_values[0], K.fixed_values[1], K.fixed_values[2], theta,
                                                    n, W->ctab);
    ctx->launches += 2;
    // table values = compressed table rows 0..65535, ranked on the device
    uint64_t* keys = (uint64_t*)W->tmp_a;  // 65536 x 4 limbs, limb-major (tmp_a holds n >= 2^17 elements)
    uint32_t* hist = (uint32_t*)(keys + 4 * 65536);  // RANK_BUCKETS counts, offsets, cursors; 65536 members
    uint32_t *boff = hist + RANK_BUCKETS, *bcur = boff + RANK_BUCKETS, *members = bcur + RANK_BUCKETS;
    ZK_CUDA(ctx, cudaMemsetAsync(hist, 0, 3 * RANK_BUCKETS * 4, st));
    table_keys_kernel<<<256, 256, 0, st>>>(W->ctab, keys, hist);
    scan_u32_kernel<<<1, 1024, 0, st>>>(hist, boff, RANK_BUCKETS, nullptr);
    table_members_kernel<<<256, 256, 0, st>>>(keys, boff, bcur, members);
    table_rank_kernel<<<256, 256, 0, st>>>(keys, W->ctab, boff, hist, members, W->rank_of, W->table_sorted,
                                           ctx->d_status);
    ctx->launches += 2;
    ZK_CUDA(ctx, cudaMemcpyAsync(W->table_vals, W->ctab, 65536 * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    ctx->launches += 2;
    const uint32_t zero_rank = 0;  // table row 0 = (0,0,0) compresses to 0, the minimum
    ZK_CUDA(ctx, cudaMemsetAsync(W->counts, 0, 65536 * 4, st));
    lookup_count_kernel<<<blocks(usable), T, 0, st>>>(adv(8), W->cin, W->table_vals, W->rank_of, usable, W->counts,
                                                      ctx->d_status);
    scan_u32_kernel<<<1, 1024, 0, st>>>(W->counts, W->offsets, 65536, nullptr);
    // table multiplicity: each of the 65536 rows once, the padding rows (usable - 65536) repeat row 0
    lookup_leftover_kernel<<<256, 256, 0, st>>>(W->counts, zero_rank, (uint32_t)(usable - 65536 + 1), W->left_cnt);
    uint32_t* d_total = W->left_off + 65536;
    scan_u32_kernel<<<1, 1024, 0, st>>>(W->left_cnt, W->left_off, 65536, d_total);
    ctx->launches += 4;
    // first-occurrence flags: tmp_c (n field elements) doubles as u32 scratch here
    uint32_t* first_flag = (uint32_t*)W->tmp_c;
    lookup_fill_input_kernel<<<65536, 64, 0, st>>>(W->counts, W->offsets, W->table_sorted, W->pin, first_flag);
    scan_u32_kernel<<<1, 1024, 0, st>>>(first_flag, W->first_flag_scan, (uint32_t)usable, nullptr);
    lookup_fill_leftover_kernel<<<65536, 64, 0, st>>>(W->left_cnt, W->left_off, W->left_rank);
    uint32_t n_repeated = 0;
    ZK_CUDA(ctx, cudaMemcpyAsync(&n_repeated, d_total, 4, cudaMemcpyDeviceToHost, st));
    int status = 0;
    ZK_CUDA(ctx, cudaMemcpyAsync(&status, ctx->d_status, 4, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, zk_stream_sync(ctx));
    if (status) {  // bit 1: the witness kernel rejected a record; 2: lookup input not a table row; 4: collision
      cudaMemsetAsync(ctx->d_status, 0, 4, st);
      if (status & 1) return set_error(ctx, ZK_E_INPUT, "EIP-152 record rejected (final flag or round count)");
      return set_error(ctx, ZK_E_VERIFY, (status & 4) ? "theta collision in the spread table"
                                                      : "lookup input not in the spread table (ConstraintSystemFailure)");
    }
    lookup_fill_table_kernel<<<blocks(usable), T, 0, st>>>(W->pin, first_flag, W->first_flag_scan, W->left_rank,
                                                           W->table_sorted, usable, n_repeated, W->ptab);
    ctx->launches += 4;
    ZK_CUDA(ctx, cudaGetLastError());
    Fp tails[12];
    for (auto& t : tails) t = tape.next();
    if ((rc = upload_fp(ctx, W->pin + usable, tails, 6))) return rc;
    if ((rc = upload_fp(ctx, W->ptab + usable, tails + 6, 6))) return rc;
    pin_blind = tape.next();
    ptab_blind = tape.next();
    const Fp* cols[2] = {W->pin, W->ptab};
    const Fp blinds[2] = {pin_blind, ptab_blind};
    Affine cms[2];
    if ((rc = commit_batch(ctx, cols, P.fb_gl, n, blinds, 2, cms))) return rc;
    tr.write_point(cms[0]);
    tr.write_point(cms[1]);
  }

  const Fp beta = tr.squeeze_challenge();
  const Fp gamma = tr.squeeze_challenge();

  phase.mark("lookup permute + commit");
  // ---- permutation grand products (K8) ------------------------------------------------------------------
  Fp z_blinds[NUM_SETS];
  {
    NttTables* TN = nullptr;
    if ((rc = ntt_tables(ctx, k, &TN))) return rc;
    const Fp* tw = TN->tw_fwd;
    Fp delta_pow[NUM_PERM];
    delta_pow[0] = Fp::one();
    for (int i = 1; i < NUM_PERM; i++) delta_pow[i] = delta_pow[i - 1] * Fp::delta();
    Fp last_z = Fp::one();
    for (int s = 0; s < NUM_SETS; s++) {
      const int c0 = 2 * s, c1 = 2 * s + 1;
      const Fp *v0 = adv(PERM_COLUMNS[c0]), *v1 = adv(PERM_COLUMNS[c1]);
      const Fp *s0 = K.sigma_values[c0], *s1 = K.sigma_values[c1];
      Fp* den = W->tmp_a;
      Fp* num = W->tmp_b;
      const Fp d0 = delta_pow[c0] * beta, d1 = delta_pow[c1] * beta;
      launch_map_range(ctx, r_lo, r_cnt, [=] __device__(uint64_t i) {
        den[i] = (beta * s0[i] + gamma + v0[i]) * (beta * s1[i] + gamma + v1[i]);
        Fp w = i < n / 2 ? tw[i] : tw[i - n / 2].neg();
        num[i] = (d0 * w + gamma + v0[i]) * (d1 * w + gamma + v1[i]);
      });
      if ((rc = batch_invert(ctx, den + r_lo, r_cnt))) return rc;
      launch_map_range(ctx, r_lo, r_cnt, [=] __device__(uint64_t i) { num[i] = num[i] * den[i]; });
      Fp* z = W->z_vals[s];
      if ((rc = grand_product(ctx, num, n, r_lo, r_cnt, last_z, z))) return rc;
      Fp tails[5];
      for (auto& t : tails) t = tape.next();
      if ((rc = upload_fp(ctx, z + n - BLINDING, tails, BLINDING))) return rc;
      if (s + 1 < NUM_SETS) {  // the next set starts from this set's last usable value
        ZK_CUDA(ctx, cudaMemcpyAsync(&last_z, z + n - (BLINDING + 1), sizeof(Fp), cudaMemcpyDeviceToHost, st));
        ZK_CUDA(ctx, zk_stream_sync(ctx));
      }
      z_blinds[s] = tape.next();
    }
  }
  phase.mark("permutation products");
  // ---- lookup grand product ------------------------------------------------------------------------------
  {
    Fp *den = W->tmp_a, *num = W->tmp_b, *z = W->z_vals[NUM_SETS];
    const Fp *pin = W->pin, *ptab = W->ptab, *cin = W->cin, *ctab = W->ctab;
    launch_map_range(ctx, r_lo, r_cnt, [=] __device__(uint64_t i) { den[i] = (beta + pin[i]) * (gamma + ptab[i]); });
    if ((rc = batch_invert(ctx, den + r_lo, r_cnt))) return rc;
    launch_map_range(ctx, r_lo, r_cnt,
                     [=] __device__(uint64_t i) { num[i] = den[i] * (cin[i] + beta) * (ctab[i] + gamma); });
    if ((rc = grand_product(ctx, num, n, r_lo, r_cnt, Fp::one(), z))) return rc;
    Fp tails[5];
    for (auto& t : tails) t = tape.next();
    if ((rc = upload_fp(ctx, z + n - BLINDING, tails, BLINDING))) return rc;
    zl_blind = tape.next();
    // one MSM pipeline for the four permutation products and the lookup product
    const Fp* cols[NUM_SETS + 1];
    Fp blinds[NUM_SETS + 1];
    Affine cms[NUM_SETS + 1];
    for (int s = 0; s < NUM_SETS; s++) {
      cols[s] = W->z_vals[s];
      blinds[s] = z_blinds[s];
    }
    cols[NUM_SETS] = z;
    blinds[NUM_SETS] = zl_blind;
    if ((rc = commit_batch(ctx, cols, P.fb_gl, n, blinds, NUM_SETS + 1, cms))) return rc;
    for (int s = 0; s <= NUM_SETS; s++) tr.write_point(cms[s]);
  }
  phase.mark("lookup product");
  // ---- vanishing argument: random polynomial ----------------------------------------------------------------
  // committed by point range, evaluated and opened by coefficient range: a rank of a group needs only its range
  if ((rc = tape.fill_device(n, W->random_poly, r_lo, r_cnt))) return rc;
  const Fp random_blind = tape.next();
  {
    Affine cm;
    if ((rc = commit(ctx, W->random_poly, P.fb_g, n,random_blind, &cm))) return rc;
    tr.write_point(cm);
  }
  const Fp y = tr.squeeze_challenge();

  phase.mark("random poly commit");
  // ---- quotient (K5 + K6) ---------------------------------------------------------------------------------------
  {
    // The 19 witness columns go to coefficients and to the three cosets here, after the last of them is
    // committed (neither form is needed earlier).  Transforms shard by column: in a
    // group every rank transforms its own block of slots and the two slot arrays are all-gathered in
    // place; the column values themselves are already replicated.
    const int world = ctx->dist_world;
    uint32_t blk_lo, blk_hi, blk_per;
    dist_column_block(NUM_WITNESS_POLYS, ctx->dist_rank, world, &blk_lo, &blk_hi, &blk_per);
    const uint64_t spr = blk_per, slot_lo = blk_lo, slot_hi = blk_hi;
    const uint64_t send_slot = spr * (uint64_t)ctx->dist_rank;  // this rank's block of the padded arrays
    // Advice columns 10 and 11 are allocated by SpreadTableChip::configure and never queried
    // (spread_table.rs:435-441, docs/CIRCUIT.md): they are committed, but no gate, evaluation or opening
    // reads their coefficient or coset form, so slots 10 and 11 are not transformed.
    const uint64_t sub[2][2] = {{slot_lo, std::min<uint64_t>(slot_hi, NUM_USED_COLUMNS)},
                                {std::max<uint64_t>(slot_lo, NUM_ADVICE_COLUMNS), slot_hi}};
    for (auto& sb : sub) {  // per sub-range: one batch of inverse transforms, then one of columns x cosets
      if (sb[1] <= sb[0]) continue;
      NttOptions o = inv;
      o.batch = (int)(sb[1] - sb[0]);
      o.in_stride = o.out_stride = n;
      if ((rc = ntt_run(ctx, W->values_all + sb[0] * n, (uint32_t)n, W->polys_all + sb[0] * n, k, o))) return rc;
      NttOptions c;
      c.batch = NUM_COSETS;
      c.in_stride = 0;
      c.out_stride = n;
      c.scale_in = K.coset_scale;
      c.scale_stride = n;
      c.batch2 = o.batch;
      c.in_stride2 = n;
      c.out_stride2 = en;
      if ((rc = ntt_run(ctx, W->polys_all + sb[0] * n, (uint32_t)n, W->cosets_all + sb[0] * en, k, c))) return rc;
    }
    phase.note("quotient: column transforms");
    // coefficients: every rank needs every column in full (evaluations, multiopen) -> in-place all-gather.
    // coset values: every rank needs only the rows of its share of the quotient -> row segments exchanged
    // point to point (world x less traffic than gathering the columns in full)
    const bool shard_rows = world > 1 && en % (uint64_t)world == 0;
    if (world > 1) {
      // coefficients: with the evaluations and the multiopen argument working by coefficient range, a rank needs
      // only its own range of the columns it did not transform (1 / world of an all-gather)
      if (by_range) {
        if ((rc = dist_exchange_ranges(ctx, (char*)W->polys_all, sizeof(Fp), n, NUM_WITNESS_POLYS, NUM_USED_COLUMNS,
                                       NUM_ADVICE_COLUMNS)))
          return rc;
      } else if ((rc = dist_allgather_device(ctx, W->polys_all + send_slot * n, W->polys_all, spr * n * sizeof(Fp)))) {
        return rc;
      }
      if (shard_rows) {
        if ((rc = dist_exchange_quotient_rows(ctx, (char*)W->cosets_all, sizeof(Fp), n, en, NUM_WITNESS_POLYS)))
          return rc;
      } else if ((rc = dist_allgather_device(ctx, W->cosets_all + send_slot * en, W->cosets_all,
                                             spr * en * sizeof(Fp)))) {
        return rc;
      }
    }
    phase.note("quotient: exchanges");
    NttTables* TNq = nullptr;
    if ((rc = ntt_tables(ctx, k, &TNq))) return rc;
    QuotientArgs qa;
    for (int c = 0; c < 12; c++) qa.advice[c] = adv_coset(c);
    for (int c = 0; c < NUM_FIXED; c++) qa.fixed[c] = K.fixed_cosets[c];
    for (int c = 0; c < NUM_PERM; c++) qa.sigma[c] = K.sigma_cosets[c];
    for (int s = 0; s < NUM_SETS; s++) qa.perm_z[s] = W->z_coset[s];
    qa.lookup_z = W->zl_coset;
    qa.lookup_in = W->pin_coset;
    qa.lookup_tab = W->ptab_coset;
    qa.l0 = K.l0;
    qa.l_last = K.l_last;
    qa.l_active = K.l_active;
    qa.tw_n = TNq->tw_fwd;
    qa.h = W->h;
    for (int s = 0; s < NUM_SELECTORS; s++) qa.sel[s] = K.selectors[s];
    qa.theta = theta;
    qa.beta = beta;
    qa.gamma = gamma;
    qa.y = y;
    for (int j = 0; j < NUM_COSETS; j++) {
      qa.coset_gen[j] = K.coset_gen[j];
      qa.t_inv[j] = K.t_inv[j];
    }
    qa.delta_pow[0] = Fp::one();
    for (int i = 1; i < NUM_PERM; i++) qa.delta_pow[i] = qa.delta_pow[i - 1] * Fp::delta();
    for (int i = 0; i < 4; i++) qa.k.small[i] = Fp::from_u64(i);
    qa.ypow[NUM_GATE_POLYS - 1] = Fp::one();
    for (int e = NUM_GATE_POLYS - 2; e >= 0; e--) qa.ypow[e] = qa.ypow[e + 1] * y;
    qa.k.pow2[0] = Fp::one();
    for (int e = 1; e < 127; e++) qa.k.pow2[e] = qa.k.pow2[e - 1].dbl();
    // rows shard too: every rank now holds all coset columns, evaluates its own range of the 3n rows and
    // the ranges of h are all-gathered in place
    const bool h_by_coset = shard_rows && by_range;  // cosets transformed by their owners, pieces by range
    if (shard_rows) {
      const uint64_t rows = en / (uint64_t)world, row_lo = rows * (uint64_t)ctx->dist_rank;
      if ((rc = quotient_run(ctx, qa, n, row_lo, row_lo + rows))) return rc;
      if (h_by_coset) {
        if ((rc = dist_exchange_h(ctx, (char*)W->h, sizeof(Fp), n, NUM_COSETS, 0))) return rc;
      } else if ((rc = dist_allgather_device(ctx, W->h + row_lo, W->h, rows * sizeof(Fp)))) {
        return rc;
      }
    } else {
      if ((rc = quotient_run(ctx, qa, n, 0, en))) return rc;
    }
    phase.note("quotient: rows (+ h exchange)");
    // back to coefficients.  On coset j, h(c_j w^i) = sum_p (c_j^n)^p h_p(c_j w^i): the size-n inverse
    // transform of the coset's values, unscaled by c_j^-i, is e_j = sum_p gamma_j^p h_p coefficient-wise,
    // and the 3 x 3 Vandermonde system gives the three pieces h_p (K.h_solve = V^-1).
    if (h_by_coset) {
      for (int c = 0; c < NUM_COSETS; c++) {
        if (c % world != ctx->dist_rank) continue;
        NttOptions o;
        o.inverse = true;
        if ((rc = ntt_run(ctx, W->h + (size_t)c * n, (uint32_t)n, W->h_coeffs + (size_t)c * n, k, o))) return rc;
      }
      if ((rc = dist_exchange_h(ctx, (char*)W->h_coeffs, sizeof(Fp), n, NUM_COSETS, 1))) return rc;
    } else {
      NttOptions o;
      o.inverse = true;
      o.batch = NUM_COSETS;
      o.in_stride = n;
      o.out_stride = n;
      if ((rc = ntt_run(ctx, W->h, (uint32_t)n, W->h_coeffs, k, o))) return rc;
    }
    {
      Fp* hc = W->h_coeffs;
      const Fp* un = K.coset_unscale;
      Fp m[NUM_COSETS][NUM_COSETS];
      for (int p = 0; p < NUM_COSETS; p++)
        for (int j = 0; j < NUM_COSETS; j++) m[p][j] = K.h_solve[p][j];
      const Fp m00 = m[0][0], m01 = m[0][1], m02 = m[0][2], m10 = m[1][0], m11 = m[1][1], m12 = m[1][2],
               m20 = m[2][0], m21 = m[2][1], m22 = m[2][2];
      // (only the commitments of the pieces and h_poly read h_coeffs: by coefficient range in a group)
      launch_map_range(ctx, r_lo, r_cnt, [=] __device__(uint64_t i) {
        const Fp e0 = hc[i] * un[i], e1 = hc[n + i] * un[n + i], e2 = hc[2 * n + i] * un[2 * n + i];
        hc[i] = m00 * e0 + m01 * e1 + m02 * e2;
        hc[n + i] = m10 * e0 + m11 * e1 + m12 * e2;
        hc[2 * n + i] = m20 * e0 + m21 * e1 + m22 * e2;
      });
    }
  }
  phase.mark("cosets + quotient + iNTT");
  Fp h_blinds[3];
  for (auto& b : h_blinds) b = tape.next();
  {
    const Fp* cols[3] = {W->h_coeffs, W->h_coeffs + n, W->h_coeffs + 2 * n};
    Affine cms[3];
    if ((rc = commit_batch(ctx, cols, P.fb_g, n, h_blinds, 3, cms))) return rc;
    for (int p = 0; p < 3; p++) tr.write_point(cms[p]);
  }

  const Fp x = tr.squeeze_challenge();
  const Fp xn = x.pow_u64(n);
  NttTables* TN = nullptr;
  if ((rc = ntt_tables(ctx, k, &TN))) return rc;
  auto rotate = [&](const Fp& v, int rot) {
    return rot >= 0 ? v * TN->omega.pow_u64((uint64_t)rot) : v * TN->omega_inv.pow_u64((uint64_t)(-rot));
  };
  const Fp x_next = rotate(x, 1), x_prev = rotate(x, -1), x_last = rotate(x, -(BLINDING + 1));

  phase.mark("h commit");
  // ---- evaluations (K9) -----------------------------------------------------------------------------------------
  // advice_queries / fixed_queries in first-use order of `configure` (docs/CIRCUIT.md §Queries)
  auto point_of = [&](int rot) { return rot == 0 ? x : (rot == 1 ? x_next : x_prev); };
  // h_poly = sum_p x^(n p) h_piece_p, h_blind likewise
  {
    const Fp* hc = W->h_coeffs;
    Fp* hp = W->h_poly;
    const Fp xn2 = xn * xn;
    // (only the multiopen argument reads h_poly: a rank of a group needs its own coefficient range)
    launch_map_range(ctx, r_lo, r_cnt,
                     [=] __device__(uint64_t i) { hp[i] = hc[i] + hc[n + i] * xn + hc[2 * n + i] * xn2; });
  }
  const Fp h_blind = h_blinds[0] + h_blinds[1] * xn + h_blinds[2] * xn * xn;
  std::vector<EvalJob> jobs;
  for (auto& q : ADVICE_QUERIES) jobs.push_back(EvalJob{adv_poly(q[0]), point_of(q[1])});
  for (int c = 0; c < NUM_FIXED; c++) jobs.push_back(EvalJob{K.fixed_polys[c], x});
  jobs.push_back(EvalJob{W->random_poly, x});
  for (int c = 0; c < NUM_PERM; c++) jobs.push_back(EvalJob{K.sigma_polys[c], x});
  for (int s = 0; s < NUM_SETS; s++) {
    jobs.push_back(EvalJob{W->z_poly[s], x});
    jobs.push_back(EvalJob{W->z_poly[s], x_next});
    if (s + 1 != NUM_SETS) jobs.push_back(EvalJob{W->z_poly[s], x_last});
  }
  jobs.push_back(EvalJob{W->zl_poly, x});
  jobs.push_back(EvalJob{W->zl_poly, x_next});
  jobs.push_back(EvalJob{W->pin_poly, x});
  jobs.push_back(EvalJob{W->pin_poly, x_prev});
  jobs.push_back(EvalJob{W->ptab_poly, x});
  {
    std::vector<Fp> evals(jobs.size());
    if ((rc = poly_eval_batch(ctx, jobs.data(), (int)jobs.size(), n, evals.data()))) return rc;
    for (auto& e : evals) tr.write_scalar(e);
  }

  phase.mark("evaluations");
  // ---- multiopen (K10) --------------------------------------------------------------------------------------------
  struct OpenPoly {
    const Fp* poly;
    Fp blind;
    std::vector<int> points;  // indices into `points`
    int set = -1;
  };
  std::vector<Fp> points;
  std::vector<OpenPoly> polys;
  auto point_index = [&](const Fp& p) {
    for (size_t i = 0; i < points.size(); i++)
      if (points[i] == p) return (int)i;
    points.push_back(p);
    return (int)points.size() - 1;
  };
  auto add_query = [&](const Fp* poly, const Fp& blind, const Fp& point) {
    int pi = point_index(point);
    for (auto& op : polys)
      if (op.poly == poly) {
        op.points.push_back(pi);
        return;
      }
    polys.push_back(OpenPoly{poly, blind, {pi}, -1});
  };
  for (auto& q : ADVICE_QUERIES) add_query(adv_poly(q[0]), advice_blinds[q[0]], point_of(q[1]));
  for (int s = 0; s < NUM_SETS; s++) {
    add_query(W->z_poly[s], z_blinds[s], x);
    add_query(W->z_poly[s], z_blinds[s], x_next);
  }
  for (int s = NUM_SETS - 2; s >= 0; s--) add_query(W->z_poly[s], z_blinds[s], x_last);
  add_query(W->zl_poly, zl_blind, x);
  add_query(W->pin_poly, pin_blind, x);
  add_query(W->ptab_poly, ptab_blind, x);
  add_query(W->pin_poly, pin_blind, x_prev);
  add_query(W->zl_poly, zl_blind, x_next);
  for (int c = 0; c < NUM_FIXED; c++) add_query(K.fixed_polys[c], Fp::one(), x);
  for (int c = 0; c < NUM_PERM; c++) add_query(K.sigma_polys[c], Fp::one(), x);
  add_query(W->h_poly, h_blind, x);
  add_query(W->random_poly, random_blind, x);

  const Fp x1 = tr.squeeze_challenge();
  const Fp x2 = tr.squeeze_challenge();
  // point sets: ordered sets of point indices, numbered by first appearance over `polys`
  std::vector<std::vector<int>> sets;
  for (auto& op : polys) {
    std::vector<int> s = op.points;
    std::sort(s.begin(), s.end());
    s.erase(std::unique(s.begin(), s.end()), s.end());
    int found = -1;
    for (size_t i = 0; i < sets.size(); i++)
      if (sets[i] == s) found = (int)i;
    if (found < 0) {
      sets.push_back(s);
      found = (int)sets.size() - 1;
    }
    op.set = found;
  }
  if (sets.size() > 8) return set_error(ctx, ZK_E_INVALID, "too many multiopen point sets");
  // halo2 numbers the sets through a BTreeMap keyed by the index set: insertion order defines the
  // set index (`or_insert(num_sets)`), iteration order does not matter for the prover.
  std::vector<Fp> q_blinds(sets.size(), Fp::zero());
  std::vector<bool> started(sets.size(), false);
  // Everything below works on the coefficient range [r_lo, r_lo + r_cnt) (all of [0, n) on one GPU): the
  // linear combinations are coefficient-wise, the commitment of q' is an MSM over the same range of g, the
  // evaluations at x3 are sums over ranges; only the division by (X - p) couples ranges, through one carry.
  for (auto& op : polys) {
    Fp* acc = W->q_polys[op.set];
    const Fp* np = op.poly;
    if (!started[op.set]) {
      ZK_CUDA(ctx, cudaMemcpyAsync(acc + r_lo, np + r_lo, r_cnt * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
      started[op.set] = true;
    } else {
      launch_map_range(ctx, r_lo, r_cnt, [=] __device__(uint64_t i) { acc[i] = acc[i] * x1 + np[i]; });
    }
    q_blinds[op.set] = q_blinds[op.set] * x1 + op.blind;
  }
  // q'(X) = sum_sets x2^.. * q_set(X) / prod (X - p).  Division by (X - p) is the recurrence y_0 = 0,
  // y_{j+1} = y_j p + c_{n-1-j} over the coefficients from the top down (quotient coefficient i = y_{n-1-i}; the
  // vectors keep length n, a zero top coefficient costs nothing) — an affine scan.  By range: rank r scans its
  // own coefficients from 0, which leaves out carry_r p^t at distance t from the top of its range, where carry_r
  // is the value that enters the range from the ranks above; the per-range end values are exchanged, the carries
  // follow on the host, and the correction uses a table of powers of p (built once per opening point).
  const uint64_t r_hi = r_lo + r_cnt;
  std::vector<Fp*> pow_tab(points.size(), nullptr);
  if (by_range) {
    if (points.size() > 4) return set_error(ctx, ZK_E_INVALID, "too many opening points");
    for (size_t pi = 0; pi < points.size(); pi++) {
      pow_tab[pi] = W->pow_tabs + pi * r_cnt;
      if ((rc = affine_scan(ctx, nullptr, points[pi], nullptr, r_cnt, Fp::one(), pow_tab[pi]))) return rc;
    }
  }
  for (size_t s = 0; s < sets.size(); s++) {
    Fp* cur = W->tmp_a;
    ZK_CUDA(ctx, cudaMemcpyAsync(cur + r_lo, W->q_polys[s] + r_lo, r_cnt * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    for (int pi : sets[s]) {
      const Fp pt = points[pi];
      Fp* rev = W->tmp_b;      // rev[t] = cur[r_hi - 1 - t]: the range's coefficients from the top down
      Fp* scanned = W->tmp_c;  // y at distance t from the top of the range, without the carry
      launch_map(ctx, r_cnt, [=] __device__(uint64_t t) { rev[t] = cur[r_hi - 1 - t]; });
      if ((rc = affine_scan(ctx, nullptr, pt, rev, r_cnt, Fp::zero(), scanned))) return rc;
      if (!by_range) {
        launch_map(ctx, r_cnt, [=] __device__(uint64_t t) { cur[r_hi - 1 - t] = scanned[t]; });
        continue;
      }
      // the value leaving this range at its lower end, were nothing entering it: scanned[cnt - 1] p + rev[cnt - 1]
      if ((rc = ensure_buf(ctx, ctx->eval_ws, 64 * sizeof(Fp)))) return rc;
      Fp* d_end = (Fp*)ctx->eval_ws.ptr;
      const uint64_t lastt = r_cnt - 1;
      launch_map(ctx, 1, [=] __device__(uint64_t) { d_end[0] = scanned[lastt] * pt + rev[lastt]; });
      std::vector<Fp> ends((size_t)ctx->dist_world);
      if ((rc = dist_gather_fields(ctx, d_end, 1, ends.data()))) return rc;
      // carries from the top rank down: carry_{world-1} = 0, carry_{q-1} = carry_q p^cnt + end_q
      const Fp p_cnt = pt.pow_u64(r_cnt);
      Fp carry = Fp::zero();
      for (int q = ctx->dist_world - 1; q > ctx->dist_rank; q--) carry = carry * p_cnt + ends[q];
      const Fp* pw = pow_tab[pi];
      launch_map(ctx, r_cnt, [=] __device__(uint64_t t) { cur[r_hi - 1 - t] = scanned[t] + carry * pw[t]; });
    }
    Fp* qp = W->q_prime;
    if (s == 0) {
      ZK_CUDA(ctx, cudaMemcpyAsync(qp + r_lo, cur + r_lo, r_cnt * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    } else {
      launch_map_range(ctx, r_lo, r_cnt, [=] __device__(uint64_t i) { qp[i] = qp[i] * x2 + cur[i]; });
    }
  }
  const Fp q_prime_blind = tape.next();
  {
    Affine cm;
    if ((rc = commit(ctx, W->q_prime, P.fb_g, n,q_prime_blind, &cm))) return rc;
    tr.write_point(cm);
  }
  const Fp x3 = tr.squeeze_challenge();
  {
    std::vector<EvalJob> qj;
    for (size_t s = 0; s < sets.size(); s++) qj.push_back(EvalJob{W->q_polys[s], x3});
    std::vector<Fp> ev(qj.size());
    if ((rc = poly_eval_batch(ctx, qj.data(), (int)qj.size(), n, ev.data()))) return rc;
    for (auto& e : ev) tr.write_scalar(e);
  }
  const Fp x4 = tr.squeeze_challenge();
  Fp p_blind = q_prime_blind;
  {
    Fp* pp = W->p_poly;
    ZK_CUDA(ctx, cudaMemcpyAsync(pp + r_lo, W->q_prime + r_lo, r_cnt * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    for (size_t s = 0; s < sets.size(); s++) {
      const Fp* q = W->q_polys[s];
      launch_map_range(ctx, r_lo, r_cnt, [=] __device__(uint64_t i) { pp[i] = pp[i] * x4 + q[i]; });
      p_blind = p_blind * x4 + q_blinds[s];
    }
  }

  phase.mark("multiopen");
  // ---- inner product argument (K11) ------------------------------------------------------------------------------
  {
    if ((rc = tape.fill_device(n, W->s_poly, r_lo, r_cnt))) return rc;
    Fp* sp = W->s_poly;
    {
      EvalJob j{sp, x3};
      Fp s_at_x3;
      if ((rc = poly_eval_batch(ctx, &j, 1, n, &s_at_x3))) return rc;
      launch_map(ctx, 1, [=] __device__(uint64_t) { sp[0] = sp[0] - s_at_x3; });
    }
    const Fp s_blind = tape.next();
    Affine cm;
    if ((rc = commit(ctx, sp, P.fb_g, n,s_blind, &cm))) return rc;
    tr.write_point(cm);
    phase.note("ipa: s poly + commit");
    const Fp xi = tr.squeeze_challenge();
    const Fp z = tr.squeeze_challenge();
    Fp* pp = W->p_poly;  // becomes p'
    launch_map_range(ctx, r_lo, r_cnt, [=] __device__(uint64_t i) { pp[i] = sp[i] * xi + pp[i]; });
    {
      EvalJob j{pp, x3};
      Fp v;
      if ((rc = poly_eval_batch(ctx, &j, 1, n, &v))) return rc;
      // the folding rounds pair coefficient i with i + half: from here on every rank holds p' in full
      if (by_range && (rc = dist_allgather_device(ctx, pp + r_lo, pp, r_cnt * sizeof(Fp)))) return rc;
      launch_map(ctx, 1, [=] __device__(uint64_t) { pp[0] = pp[0] - v; });
    }
    Fp f = s_blind * xi + p_blind;
    // b = powers of x3
    Fp* b = W->b_vec;
    if ((rc = affine_scan(ctx, nullptr, x3, nullptr, n, Fp::one(), b))) return rc;
    // Two stages.  For the first r rounds the folded generators are not materialised: after j rounds
    //     G'_i = sum_m [m = i mod len] s_m g_m,   s_m = prod_{t < j} u_t^(bit_{k-1-t}(m)),
    // so L_j = <p'_hi, G'_lo> and R_j = <p'_lo, G'_hi> are MSMs over the ORIGINAL g with scalars
    // c_m = p'[(m mod len) +- half] * s_m, supported on bit_{k-1-j}(m) = 0 (L) or 1 (R): every MSM stays
    // on the precomputed fixed-base tables.  After r rounds the generators are folded once
    // (ipa_fold.cu: H_i = sum_q s_q g_{q len + i}, the shared scalars' digits sorted on the host), a
    // window table is built over H, and the remaining k - r rounds run the same scheme against H at
    // 2^-r of the cost.  halo2's `parallel_generator_collapse` (n scalar multiplications per proof,
    // latency-bound on a GPU) never runs.
    Fp* svec = W->tmp_a;
    Fp* cvec = W->tmp_b;
    const int fold_rounds = ipa_fold_rounds(ctx->dist_world);
    const bool two_stage = P.fb_g8.table != nullptr && k > fold_rounds + 8 && (1 << fold_rounds) % ctx->dist_world == 0;
    const int fold_at = two_stage ? fold_rounds : k;
    const FixedBase* fb = &P.fb_g;
    FixedBase fb_h;
    uint64_t base_n = n;  // size of the current generator vector (n, then n >> r)
    // the rounds on g read s and c only at the points of this rank's share of the split MSM
    uint64_t m_lo = 0, m_hi = n;
    if (ctx->dist_world > 1 && P.fb_g.split) {
      m_lo = P.fb_g.lo;
      m_hi = P.fb_g.lo + P.fb_g.nmain;
    }
    launch_map_range(ctx, m_lo, m_hi - m_lo, [=] __device__(uint64_t m) { svec[m] = Fp::one(); });
    std::vector<Fp> us;
    phase.note("ipa: p', b, setup");
    for (int j = 0; j < k; j++) {
      const uint64_t half = 1ull << (k - j - 1);
      if (j == fold_at) phase.note("ipa: rounds on g");
      if (j == fold_at) {
        base_n = n >> fold_at;
        if ((rc = ensure_fold_workspace(ctx, W, base_n))) return rc;
        if ((rc = ipa_fold_generators(ctx, P.fb_g8, us.data(), fold_at, W->fold_ws, W->h_table))) return rc;
        ZK_CUDA(ctx, cudaMemcpyAsync(W->h_table + base_n, P.g + n, 2 * sizeof(Affine), cudaMemcpyDeviceToDevice, st));
        if ((rc = fixed_base_build_inplace(ctx, base_n, 2, ipa_stage2_c(), W->h_table, W->h_tmp, &fb_h))) return rc;
        fb = &fb_h;
        m_lo = 0;  // the table over the folded generators is whole on every rank
        m_hi = base_n;
        launch_map(ctx, base_n, [=] __device__(uint64_t m) { svec[m] = Fp::one(); });
        phase.note("ipa: fold + table");
      }
      const uint64_t bn = base_n, lo = m_lo, hi = m_hi;
      launch_map_range(ctx, lo, hi - lo, [=] __device__(uint64_t m) {
        uint64_t i = m & (2 * half - 1);
        cvec[m] = (i < half ? pp[i + half] : pp[i - half]) * svec[m];
      });
      Fp lr_values[2];
      if ((rc = inner_product_pair(ctx, pp + half, b, pp, b + half, half, lr_values))) return rc;
      const Fp vl = lr_values[0], vr = lr_values[1];
      const Fp l_rand = tape.next(), r_rand = tape.next();
      // L and R share the scalar vector and differ in the index bit they keep: one pipeline, 2 jobs
      MsmJob lr[2];
      for (int side = 0; side < 2; side++) {
        lr[side].scalars = cvec;
        lr[side].n_extra = 2;
        lr[side].extra[0] = (side == 0 ? vl : vr) * z;
        lr[side].extra[1] = side == 0 ? l_rand : r_rand;
        lr[side].extra_index[0] = (uint32_t)bn + 1;  // U
        lr[side].extra_index[1] = (uint32_t)bn;      // W
        lr[side].side_mask = (uint32_t)half;
        lr[side].side_select = side;
      }
      XYZZ lrj[2];
      if ((rc = msm_fixed_batch(ctx, *fb, lr, 2, bn, lrj))) return rc;
      tr.write_point(lrj[0].to_affine());
      tr.write_point(lrj[1].to_affine());
      const Fp u = tr.squeeze_challenge();
      const Fp u_inv = u.inv();
      us.push_back(u);
      launch_map(ctx, hi > half ? hi : half, [=] __device__(uint64_t m) {
        if (m < half) {
          pp[m] = pp[m] + pp[m + half] * u_inv;
          b[m] = b[m] + b[m + half] * u;
        }
        if (m >= lo && m < hi && (m & half)) svec[m] = svec[m] * u;
      });
      f = f + l_rand * u_inv + r_rand * u;
    }
    phase.note("ipa: rounds on H");
    Fp c;
    ZK_CUDA(ctx, cudaMemcpyAsync(&c, pp, sizeof(Fp), cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, zk_stream_sync(ctx));
    tr.write_scalar(c);
    tr.write_scalar(f);
  }
  phase.mark("ipa");
  ZK_CUDA(ctx, cudaGetLastError());
  proof_out = tr.proof();
  return ZK_OK;
}

}  // namespace zkodst

using namespace zkodst;

static int32_t create_proof_entry(zk_ctx* ctx, const uint8_t* inputs, bool inputs_on_device,
                                  uint64_t n_compressions, const uint8_t seed[16], uint8_t* proof_out,
                                  uint64_t* proof_len);

extern "C" int32_t zk_create_proof(zk_ctx* ctx, const uint8_t* inputs, uint64_t n_compressions,
                                   const uint8_t seed[16], uint8_t* proof_out, uint64_t* proof_len) {
  return create_proof_entry(ctx, inputs, false, n_compressions, seed, proof_out, proof_len);
}
extern "C" int32_t zk_create_proof_device_inputs(zk_ctx* ctx, const uint8_t* d_inputs, uint64_t n_compressions,
                                                 const uint8_t seed[16], uint8_t* proof_out,
                                                 uint64_t* proof_len) {
  return create_proof_entry(ctx, d_inputs, true, n_compressions, seed, proof_out, proof_len);
}

static int32_t create_proof_entry(zk_ctx* ctx, const uint8_t* inputs, bool inputs_on_device,
                                  uint64_t n_compressions, const uint8_t seed[16], uint8_t* proof_out,
                                  uint64_t* proof_len) {
  if (!ctx || !seed || !proof_len || (n_compressions && !inputs)) return ZK_E_INVALID;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<uint8_t> proof;
  int32_t rc;
  try {
    rc = create_proof_impl(ctx, inputs, inputs_on_device, n_compressions, seed, proof);
  } catch (std::exception& e) {
    if (ctx->dist_world > 1) dist_abort(ctx);
    return set_error(ctx, ZK_E_VERIFY, e.what());
  }
  if (rc) {
    // a rank of a group that fails has skipped collectives its peers are waiting in: leave the group so that
    // they end with an error (dist_stream_sync) instead of waiting for this rank forever
    // (input / constraint failures are decided identically by every rank from the replicated records: no abort)
    if (ctx->dist_world > 1 && (rc == ZK_E_CUDA || rc == ZK_E_NOMEM)) dist_abort(ctx);
    return rc;
  }
  if (!proof_out || *proof_len < proof.size()) {
    *proof_len = proof.size();
    return set_error(ctx, ZK_E_BUFFER, "proof buffer too small");
  }
  memcpy(proof_out, proof.data(), proof.size());
  *proof_len = proof.size();
  return ZK_OK;
}
