// Params and key generation on the device (amortised, not timed in
// proofs/sec).
//
// Replaces halo2_proofs 0.3.0 `Params::<EqAffine>::{new, read, write}`, `keygen_vk` and
// `keygen_pk` as the reference calls them (blake2f-circuit/benches/blake2f.rs:83-103).
//  * `zk_params_load` / `zk_params_write` speak halo2's params file format (k as u32 LE, n
//    compressed g, n compressed g_lagrange, w, u), so a genuine halo2 params file drops in.
//  * `zk_params_generate_substitute` builds a URS with known discrete logs for benchmarking and
//    parity work, because `Params::new` hashes to the curve with constants that cannot be
//    reproduced offline: g_i = [s_i] G, g_lagrange = [ifft(s)_j] G.
//  * keygen lays out the fixed columns (spread table + compressed selectors,
//    spread_table.rs:470-508, compression.rs:561-577), derives the permutation sigma columns with
//    halo2's cycle-merging `copy` (permutation/keygen.rs), commits to them and precomputes their
//    coefficient and extended-coset forms.
#include <map>
#include <string>

#include "polyops.cuh"
#include "prover_state.h"
#include "transcript.h"

namespace zkodst {

static void free_state(void* p) {
  ProverState* s = (ProverState*)p;
  cudaFree(s->params.g);
  cudaFree(s->params.g_lagrange);
  fixed_base_free(s->params.fb_g);
  fixed_base_free(s->params.fb_gl);
  fixed_base_free(s->params.fb_g8);
  free_keys(s->keys);
  delete s;
}

ProverState* prover_state(zk_ctx* ctx) {
  if (!ctx->prover_state) {
    ctx->prover_state = new ProverState();
    ctx->prover_state_free = free_state;
  }
  return (ProverState*)ctx->prover_state;
}

void free_keys(DeviceKeys& K) {
  for (int i = 0; i < NUM_FIXED; i++) {
    cudaFree(K.fixed_values[i]);
    cudaFree(K.fixed_polys[i]);
    cudaFree(K.fixed_cosets[i]);
  }
  for (int i = 0; i < NUM_PERM; i++) {
    cudaFree(K.sigma_values[i]);
    cudaFree(K.sigma_polys[i]);
    cudaFree(K.sigma_cosets[i]);
  }
  cudaFree(K.l0);
  cudaFree(K.l_last);
  cudaFree(K.l_active);
  cudaFree(K.coset_scale);
  cudaFree(K.coset_unscale);
  if (K.workspace) free_workspace(K.workspace);
  K = DeviceKeys();
}

namespace {

// ---- point helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ Affine xyzz_to_affine_dev(const XYZZ& p) { return p.to_affine(); }

// out[i] = [s_i] G via a table of d * 256^w * G (w < 32, 1 <= d <= 255)
__global__ void __launch_bounds__(128)
fixed_base_mul_kernel(const Fp* __restrict__ scalars, const Affine* __restrict__ table, uint64_t n,
                      Affine* __restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t s[4];
  scalars[i].to_canonical(s);
  XYZZ acc = XYZZ::identity();
  for (int w = 0; w < 32; w++) {
    uint32_t d = (uint32_t)(s[w >> 3] >> ((w & 7) * 8)) & 0xff;
    if (d) acc = acc.add_affine(table[w * 255 + d - 1]);
  }
  out[i] = xyzz_to_affine_dev(acc);
}

struct U256 {
  uint64_t v[4];
};
__global__ void __launch_bounds__(128)
decompress_kernel(const uint8_t* __restrict__ bytes, uint64_t n, U256 tm1o2, Affine* __restrict__ out,
                  int* __restrict__ bad) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine p;
  if (!decompress_point(bytes + 32 * i, tm1o2.v, p)) {
    *bad = 1;
    return;
  }
  out[i] = p;
}
__global__ void compress_kernel(const Affine* __restrict__ pts, uint64_t n, uint8_t* __restrict__ bytes) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine p = pts[i];
  uint64_t c[4] = {0, 0, 0, 0};
  if (!p.is_identity()) {
    p.x.to_canonical(c);
    if (p.y.is_odd()) c[3] |= 0x8000000000000000ULL;
  }
  uint8_t* o = bytes + 32 * i;
  for (int l = 0; l < 4; l++)
    for (int b = 0; b < 8; b++) o[8 * l + b] = (uint8_t)(c[l] >> (8 * b));
}

__global__ void from_u512_kernel(const uint64_t* __restrict__ raw, uint64_t n, Fp* __restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t w[8];
  for (int j = 0; j < 8; j++) w[j] = raw[8 * i + j];
  out[i] = Fp::from_u512(w);
}

Affine host_scalar_mul(const Affine& p, const Fp& s) {
  uint64_t e[4];
  s.to_canonical(e);
  XYZZ acc = XYZZ::identity();
  for (int i = 255; i >= 0; i--) {
    acc = acc.dbl();
    if ((e[i >> 6] >> (i & 63)) & 1) acc = acc.add_affine(p);
  }
  return acc.to_affine();
}
Affine vesta_generator() {
  return Affine{Fq::one().neg(), Fq::from_u64(2)};
}

int32_t install_params(zk_ctx* ctx, int k, Affine* g, Affine* g_lagrange, const Affine& w, const Affine& u) {
  ProverState* S = prover_state(ctx);
  if (S->has_keys) {
    free_keys(S->keys);
    S->has_keys = false;
  }
  cudaFree(S->params.g);
  cudaFree(S->params.g_lagrange);
  fixed_base_free(S->params.fb_g);
  fixed_base_free(S->params.fb_gl);
  fixed_base_free(S->params.fb_g8);
  S->has_params = false;
  S->params.k = k;
  S->params.n = 1ull << k;
  S->params.g = g;
  S->params.g_lagrange = g_lagrange;
  S->params.w = w;
  S->params.u = u;
  const uint64_t n = S->params.n;
  // bases are stored as [g_0 .. g_{n-1}, w, u] and [gl_0 .. gl_{n-1}, w] so that a commitment's
  // blinding term (and the IPA's U term) ride in the same MSM
  ZK_CUDA(ctx, cudaMemcpyAsync(g + n, &S->params.w, sizeof(Affine), cudaMemcpyHostToDevice, ctx->stream));
  ZK_CUDA(ctx, cudaMemcpyAsync(g + n + 1, &S->params.u, sizeof(Affine), cudaMemcpyHostToDevice, ctx->stream));
  ZK_CUDA(ctx, cudaMemcpyAsync(g_lagrange + n, &S->params.w, sizeof(Affine), cudaMemcpyHostToDevice, ctx->stream));
  int32_t rc = fixed_base_build(ctx, g, n, 2, &S->params.fb_g);
  if (rc) return rc;
  rc = fixed_base_build(ctx, g_lagrange, n, 1, &S->params.fb_gl);
  if (rc) return rc;
  // 8-bit windows for the IPA generator fold (ipa_fold.cu); in a multi-GPU group the fold needs every
  // rank's range to be a whole number of n / 2^r blocks
  if (k >= 15 && (1 << ipa_fold_rounds(ctx->dist_world)) % ctx->dist_world == 0) {
    rc = fixed_base_build(ctx, g, n, 2, &S->params.fb_g8, 8);
    if (rc) return rc;
  }
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  S->has_params = true;
  return ZK_OK;
}

}  // namespace

int32_t commit(zk_ctx* ctx, const Fp* d_scalars, const FixedBase& fb, uint64_t n, const Fp& blind, Affine* out) {
  const uint32_t w_index = (uint32_t)n;  // W follows the n base points in both tables
  XYZZ r;
  int32_t rc = msm_fixed(ctx, fb, d_scalars, n, &blind, &w_index, 1, &r);
  if (rc) return rc;
  *out = r.to_affine();
  return ZK_OK;
}

int32_t commit_batch(zk_ctx* ctx, const Fp* const* d_scalars, const FixedBase& fb, uint64_t n, const Fp* blinds,
                     int nb, Affine* out) {
  MsmJob jobs[MSM_MAX_BATCH];
  XYZZ r[MSM_MAX_BATCH];
  if (nb > MSM_MAX_BATCH) return set_error(ctx, ZK_E_INVALID, "commit_batch: too many columns");
  for (int m = 0; m < nb; m++) {
    jobs[m].scalars = d_scalars[m];
    jobs[m].n_extra = 1;
    jobs[m].extra[0] = blinds[m];
    jobs[m].extra_index[0] = (uint32_t)n;  // W follows the n base points in both tables
  }
  int32_t rc = msm_fixed_batch(ctx, fb, jobs, nb, n, r);
  if (rc) return rc;
  for (int m = 0; m < nb; m++) out[m] = r[m].to_affine();
  return ZK_OK;
}

int32_t coeff_to_extended(zk_ctx* ctx, const DeviceKeys& K, const Fp* coeffs, Fp* out) {
  // three size-n transforms of the coefficients scaled by c_j^i, in one batch of launches
  NttOptions opt;
  opt.batch = NUM_COSETS;
  opt.in_stride = 0;
  opt.out_stride = K.n;
  opt.scale_in = K.coset_scale;
  opt.scale_stride = K.n;
  return ntt_run(ctx, coeffs, (uint32_t)K.n, out, K.k, opt);
}

}  // namespace zkodst

using namespace zkodst;

extern "C" int32_t zk_params_generate_substitute(zk_ctx* ctx, int32_t k, const uint8_t seed[16]) {
  if (!ctx || !seed || k < 1 || k > 28) return ZK_E_INVALID;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint64_t n = 1ull << k;
  // scalars s_i (in RNG order), then s_w, s_u
  std::vector<uint64_t> raw(8 * n);
  XorShift rng(seed);
  for (uint64_t i = 0; i < n; i++) rng.next_wide(&raw[8 * i]);
  Fp sw = rng.random_fp(), su = rng.random_fp();
  // fixed-base table of d * 256^w * G
  std::vector<Affine> table(32 * 255);
  {
    XYZZ base = XYZZ::from_affine(vesta_generator());
    for (int w = 0; w < 32; w++) {
      Affine base_aff = base.to_affine();
      XYZZ cur = base;
      for (int d = 1; d <= 255; d++) {
        table[w * 255 + d - 1] = cur.to_affine();
        cur = cur.add_affine(base_aff);
      }
      base = cur;
    }
  }
  uint64_t* d_raw = nullptr;
  Fp *d_s = nullptr, *d_sl = nullptr;
  Affine *d_table = nullptr, *g = nullptr, *gl = nullptr;
  DevTemps tmp;  // frees whatever was allocated on every early return
  tmp.own(&d_raw); tmp.own(&d_s); tmp.own(&d_sl); tmp.own(&d_table); tmp.own(&g); tmp.own(&gl);
  ZK_CUDA(ctx, cudaMalloc((void**)&d_raw, raw.size() * 8));
  ZK_CUDA(ctx, cudaMalloc((void**)&d_s, n * sizeof(Fp)));
  ZK_CUDA(ctx, cudaMalloc((void**)&d_sl, n * sizeof(Fp)));
  ZK_CUDA(ctx, cudaMalloc((void**)&d_table, table.size() * sizeof(Affine)));
  ZK_CUDA(ctx, cudaMalloc((void**)&g, (n + 2) * sizeof(Affine)));
  ZK_CUDA(ctx, cudaMalloc((void**)&gl, (n + 2) * sizeof(Affine)));
  ZK_CUDA(ctx, cudaMemcpyAsync(d_raw, raw.data(), raw.size() * 8, cudaMemcpyHostToDevice, st));
  ZK_CUDA(ctx, cudaMemcpyAsync(d_table, table.data(), table.size() * sizeof(Affine), cudaMemcpyHostToDevice, st));
  const unsigned T = 128, blocks = (unsigned)((n + T - 1) / T);
  from_u512_kernel<<<blocks, T, 0, st>>>(d_raw, n, d_s);
  fixed_base_mul_kernel<<<blocks, T, 0, st>>>(d_s, d_table, n, g);
  ctx->launches += 2;
  NttOptions inv;
  inv.inverse = true;
  int32_t rc = ntt_run(ctx, d_s, (uint32_t)n, d_sl, k, inv);
  if (rc) return rc;
  fixed_base_mul_kernel<<<blocks, T, 0, st>>>(d_sl, d_table, n, gl);
  ctx->launches++;
  ZK_CUDA(ctx, cudaGetLastError());
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  Affine G = vesta_generator();
  return install_params(ctx, k, tmp.release(&g), tmp.release(&gl), host_scalar_mul(G, sw), host_scalar_mul(G, su));
}

extern "C" int32_t zk_params_load(zk_ctx* ctx, const uint8_t* bytes, uint64_t len) {
  if (!ctx || !bytes || len < 4) return ZK_E_INVALID;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  uint32_t k;
  memcpy(&k, bytes, 4);
  if (k < 1 || k > 28) return set_error(ctx, ZK_E_INVALID, "params: k out of range");
  const uint64_t n = 1ull << k;
  if (len != 4 + (2 * n + 2) * 32) return set_error(ctx, ZK_E_INVALID, "params: bad length");
  cudaStream_t st = ctx->stream;
  uint8_t* d_bytes = nullptr;
  Affine *g = nullptr, *gl = nullptr, *d_wu = nullptr;
  int* d_bad = nullptr;
  DevTemps tmp;
  tmp.own(&d_bytes); tmp.own(&g); tmp.own(&gl); tmp.own(&d_wu); tmp.own(&d_bad);
  ZK_CUDA(ctx, cudaMalloc((void**)&d_bytes, len - 4));
  ZK_CUDA(ctx, cudaMalloc((void**)&g, (n + 2) * sizeof(Affine)));
  ZK_CUDA(ctx, cudaMalloc((void**)&gl, (n + 2) * sizeof(Affine)));
  ZK_CUDA(ctx, cudaMalloc((void**)&d_wu, 2 * sizeof(Affine)));
  ZK_CUDA(ctx, cudaMalloc((void**)&d_bad, sizeof(int)));
  ZK_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), st));
  ZK_CUDA(ctx, cudaMemcpyAsync(d_bytes, bytes + 4, len - 4, cudaMemcpyHostToDevice, st));
  U256 e;
  fq_sqrt_exponent(e.v);
  const unsigned T = 128;
  decompress_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(d_bytes, n, e, g, d_bad);
  decompress_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(d_bytes + 32 * n, n, e, gl, d_bad);
  decompress_kernel<<<1, T, 0, st>>>(d_bytes + 64 * n, 2, e, d_wu, d_bad);
  ctx->launches += 3;
  int bad = 0;
  Affine wu[2];
  ZK_CUDA(ctx, cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  ZK_CUDA(ctx, cudaMemcpyAsync(wu, d_wu, sizeof wu, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  if (bad) return set_error(ctx, ZK_E_INVALID, "params: invalid point encoding");
  return install_params(ctx, (int)k, tmp.release(&g), tmp.release(&gl), wu[0], wu[1]);
}

extern "C" int32_t zk_params_write(zk_ctx* ctx, uint8_t* out, uint64_t* len) {
  if (!ctx || !len) return ZK_E_INVALID;
  ProverState* S = prover_state(ctx);
  if (!S->has_params) return set_error(ctx, ZK_E_STATE, "no params loaded");
  const uint64_t n = S->params.n, need = 4 + (2 * n + 2) * 32;
  if (!out || *len < need) {
    *len = need;
    return ZK_E_BUFFER;
  }
  *len = need;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  uint8_t* d_bytes = nullptr;
  DevTemps tmp;
  tmp.own(&d_bytes);
  ZK_CUDA(ctx, cudaMalloc((void**)&d_bytes, 64 * n));
  const unsigned T = 256;
  compress_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(S->params.g, n, d_bytes);
  compress_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(S->params.g_lagrange, n, d_bytes + 32 * n);
  ctx->launches += 2;
  uint32_t k = (uint32_t)S->params.k;
  memcpy(out, &k, 4);
  ZK_CUDA(ctx, cudaMemcpyAsync(out + 4, d_bytes, 64 * n, cudaMemcpyDeviceToHost, st));
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  point_to_bytes(S->params.w, out + 4 + 64 * n);
  point_to_bytes(S->params.u, out + 4 + 64 * n + 32);
  return ZK_OK;
}

// ---- keygen -------------------------------------------------------------------------------------------
namespace zkodst {
namespace {

// halo2 compress_selectors `process` for this circuit's 12 simple selectors.  Gate degrees
// (selector included) follow docs/CIRCUIT.md; max_degree is the constraint-system degree 4.
void combine_selectors(const RegionLayout& L, SelectorExpr out[NUM_SELECTORS], int* n_cols,
                       std::vector<std::vector<uint8_t>>& columns) {
  static const int degree[NUM_SELECTORS] = {4, 2, 3, 2, 4, 2, 3, 2, 2, 2, 3, 2, 2, 3};
  const int max_degree = CS_DEGREE;
  const uint32_t R = L.rows;
  auto active = [&](int s, uint32_t r) { return L.selectors[(size_t)s * R + r] != 0; };
  bool excl[NUM_SELECTORS][NUM_SELECTORS] = {};
  for (int i = 0; i < NUM_SELECTORS; i++)
    for (int j = 0; j < i; j++)
      for (uint32_t r = 0; r < R; r++)
        if (active(i, r) && active(j, r)) {
          excl[i][j] = excl[j][i] = true;
          break;
        }
  bool added[NUM_SELECTORS] = {};
  int col = 0;
  columns.clear();
  for (int i = 0; i < NUM_SELECTORS; i++) {
    if (added[i]) continue;
    added[i] = true;
    int d = degree[i] - 1;
    std::vector<int> comb = {i};
    for (int j = i + 1; j < NUM_SELECTORS; j++) {
      if (d + (int)comb.size() == max_degree) break;
      if (added[j]) continue;
      bool bad = false;
      for (int c : comb) bad |= excl[j][c];
      if (bad) continue;
      int nd = d > degree[j] - 1 ? d : degree[j] - 1;
      if (nd + (int)comb.size() + 1 > max_degree) continue;
      d = nd;
      comb.push_back(j);
      added[j] = true;
    }
    std::vector<uint8_t> vals(R, 0);
    for (size_t c = 0; c < comb.size(); c++) {
      out[comb[c]] = SelectorExpr{FIXED_SELECTOR_BASE + col, (int)c + 1, (int)comb.size()};
      for (uint32_t r = 0; r < R; r++)
        if (active(comb[c], r)) vals[r] = (uint8_t)(c + 1);
    }
    columns.push_back(vals);
    col++;
  }
  *n_cols = col;
}

// permutation::keygen::Assembly over one region (cells outside regions keep the identity)
struct RegionPermutation {
  uint32_t R;
  std::vector<uint64_t> mapping, aux;  // [NUM_PERM][R], value = col << 32 | row
  std::vector<uint32_t> sizes;
  explicit RegionPermutation(uint32_t R_) : R(R_), mapping((size_t)NUM_PERM * R_), aux((size_t)NUM_PERM * R_),
                                            sizes((size_t)NUM_PERM * R_, 1) {
    for (int c = 0; c < NUM_PERM; c++)
      for (uint32_t r = 0; r < R; r++) mapping[(size_t)c * R + r] = aux[(size_t)c * R + r] = ((uint64_t)c << 32) | r;
  }
  size_t at(uint64_t cell) const { return (size_t)(cell >> 32) * R + (uint32_t)cell; }
  static int perm_index(int advice_col) {
    for (int i = 0; i < NUM_PERM; i++)
      if (PERM_COLUMNS[i] == advice_col) return i;
    return -1;
  }
  void copy(const CopyConstraint& cc) {
    uint64_t left = ((uint64_t)perm_index(cc.left_col) << 32) | cc.left_row;
    uint64_t right = ((uint64_t)perm_index(cc.right_col) << 32) | cc.right_row;
    uint64_t lc = aux[at(left)], rc = aux[at(right)];
    if (lc == rc) return;
    if (sizes[at(lc)] < sizes[at(rc)]) std::swap(lc, rc);
    sizes[at(lc)] += sizes[at(rc)];
    uint64_t i = rc;
    do {
      aux[at(i)] = lc;
      i = mapping[at(i)];
    } while (i != rc);
    std::swap(mapping[at(left)], mapping[at(right)]);
  }
};

__global__ void fixed_columns_kernel(Fp* tag, Fp* dense, Fp* spread, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t d = i < 65536 ? (uint32_t)i : 0;
  uint32_t x = d;
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  tag[i] = Fp::from_u64(d < 256 ? 0 : (d < 32768 ? 1 : 2));
  dense[i] = Fp::from_u64(d);
  spread[i] = Fp::from_u64(x);
}
__global__ void selector_column_kernel(const uint8_t* __restrict__ tmpl, uint32_t R, uint64_t n_comp, uint64_t n,
                                       Fp* __restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t comp = i / R;
  uint32_t v = comp < n_comp ? tmpl[i % R] : 0;
  out[i] = Fp::from_u64(v);
}
// the constants column: region row r of every compression holds tmpl[r]
__global__ void constants_column_kernel(const uint64_t* __restrict__ tmpl, uint32_t R, uint64_t n_comp, uint64_t n,
                                        Fp* __restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t v = i / R < n_comp ? tmpl[i % R] : 0;
  out[i] = v ? Fp::from_u64(v) : Fp::zero();
}
// Chaining copies: the two cells of copy t point at each other (both are in no other copy, so halo2's
// cycle-merging `copy` leaves exactly this 2-cycle).  cells[4 t ..] = {perm column a, row a, perm column b, row b}.
__global__ void sigma_chain_kernel(const uint32_t* __restrict__ cells, uint32_t ncopies, uint64_t n,
                                   const Fp* __restrict__ tw, const Fp* __restrict__ delta_pows,
                                   Fp* const* __restrict__ sigma) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncopies) return;
  const uint32_t ca = cells[4 * t], ra = cells[4 * t + 1], cb = cells[4 * t + 2], rb = cells[4 * t + 3];
  auto id = [&](uint32_t c, uint64_t r) { return delta_pows[c] * (r < n / 2 ? tw[r] : tw[r - n / 2].neg()); };
  sigma[ca][ra] = id(cb, rb);
  sigma[cb][rb] = id(ca, ra);
}
// sigma_c[row] = delta^(mapped col) * omega^(mapped row)
__global__ void sigma_kernel(const uint64_t* __restrict__ tmpl, uint32_t R, uint64_t n_comp, uint64_t n, int col,
                             const Fp* __restrict__ tw, Fp delta_pows0, Fp delta_pows1, Fp delta_pows2,
                             Fp delta_pows3, Fp delta_pows4, Fp delta_pows5, Fp delta_pows6, Fp delta_pows7,
                             Fp* __restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t comp = i / R;
  uint32_t mcol = col;
  uint64_t mrow = i;
  if (comp < n_comp) {
    uint64_t m = tmpl[(size_t)col * R + (i % R)];
    mcol = (uint32_t)(m >> 32);
    mrow = comp * R + (uint32_t)m;
  }
  Fp w = mrow < n / 2 ? tw[mrow] : tw[mrow - n / 2].neg();
  Fp d;
  switch (mcol) {
    case 0: d = delta_pows0; break;
    case 1: d = delta_pows1; break;
    case 2: d = delta_pows2; break;
    case 3: d = delta_pows3; break;
    case 4: d = delta_pows4; break;
    case 5: d = delta_pows5; break;
    case 6: d = delta_pows6; break;
    default: d = delta_pows7; break;
  }
  out[i] = d * w;
}

// indicator column (kind 0: l_0, 1: l_last, 2: l_blind) -> extended coset
int32_t build_indicator(zk_ctx* ctx, const DeviceKeys& K, int kind, Fp* tmp, Fp* tmp2, Fp* out_coset) {
  const uint64_t n = K.n;
  launch_map(ctx, n, [=] __device__(uint64_t i) {
    bool one = kind == 0 ? i == 0 : (kind == 1 ? i == n - BLINDING - 1 : i >= n - BLINDING);
    tmp[i] = one ? Fp::one() : Fp::zero();
  });
  NttOptions inv;
  inv.inverse = true;
  int32_t r = ntt_run(ctx, tmp, (uint32_t)n, tmp2, K.k, inv);
  if (r) return r;
  return coeff_to_extended(ctx, K, tmp2, out_coset);
}
void combine_active(zk_ctx* ctx, const DeviceKeys& K, const Fp* lblind) {
  const Fp* l_last = K.l_last;
  Fp* l_active = K.l_active;
  launch_map(ctx, K.en, [=] __device__(uint64_t i) { l_active[i] = Fp::one() - (l_last[i] + lblind[i]); });
}

std::string hex32(const uint8_t b[32]) {
  static const char* d = "0123456789abcdef";
  std::string s;
  for (int i = 0; i < 32; i++) {
    s += d[b[i] >> 4];
    s += d[b[i] & 15];
  }
  return s;
}

}  // namespace
}  // namespace zkodst

extern "C" int32_t zk_blake2f_keygen(zk_ctx* ctx, uint32_t rounds, uint64_t n_compressions) {
  return zk_blake2f_keygen_chained(ctx, rounds, n_compressions, nullptr);
}

extern "C" int32_t zk_blake2f_keygen_chained(zk_ctx* ctx, uint32_t rounds, uint64_t n_compressions,
                                             const uint8_t* chain) {
  if (!ctx) return ZK_E_INVALID;
  if (chain && n_compressions && chain[0])
    return set_error(ctx, ZK_E_INVALID, "the first compression cannot continue another");
  ProverState* S = prover_state(ctx);
  if (!S->has_params) return set_error(ctx, ZK_E_STATE, "keygen before params");
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  free_keys(S->keys);  // unconditionally: a keygen that failed midway leaves allocations behind has_keys == false
  S->has_keys = false;
  DeviceKeys& K = S->keys;
  const int k = S->params.k;
  const uint64_t n = S->params.n;
  if (k < 17) return set_error(ctx, ZK_E_ROWS, "k < 17 cannot hold the 2^16-row spread table");
  DeviceRegionLayout* DL = nullptr;
  int32_t rc = get_layout(ctx, rounds, &DL);
  if (rc) return rc;
  const RegionLayout& L = DL->host;
  if ((unsigned __int128)L.rows * n_compressions > n - (BLINDING + 1))
    return set_error(ctx, ZK_E_ROWS, "compressions do not fit in 2^k rows");
  K.k = k;
  K.n = n;
  K.en = NUM_COSETS * n;
  K.rounds = rounds;
  K.n_compressions = n_compressions;
  K.region_rows = L.rows;
  K.chain.assign(n_compressions, 0);
  if (chain)
    for (uint64_t j = 0; j < n_compressions; j++) K.chain[j] = chain[j] ? 1 : 0;
  {
    // c_j = zeta * omega_4n^j (the first three cosets of halo2's extended domain, degree 4 => 4n)
    Fp omega_ext = Fp::root_of_unity();
    for (int i = k + 2; i < 32; i++) omega_ext = omega_ext.sqr();
    Fp gamma[NUM_COSETS];
    Fp c = Fp::zeta();
    cudaFree(K.coset_scale);
    cudaFree(K.coset_unscale);
    ZK_CUDA(ctx, cudaMalloc((void**)&K.coset_scale, (size_t)NUM_COSETS * n * sizeof(Fp)));
    ZK_CUDA(ctx, cudaMalloc((void**)&K.coset_unscale, (size_t)NUM_COSETS * n * sizeof(Fp)));
    for (int j = 0; j < NUM_COSETS; j++) {
      K.coset_gen[j] = c;
      gamma[j] = c.pow_u64(n);
      K.t_inv[j] = (gamma[j] - Fp::one()).inv();
      if ((rc = affine_scan(ctx, nullptr, c, nullptr, n, Fp::one(), K.coset_scale + (size_t)j * n))) return rc;
      if ((rc = affine_scan(ctx, nullptr, c.inv(), nullptr, n, Fp::one(), K.coset_unscale + (size_t)j * n))) return rc;
      c = c * omega_ext;
    }
    // inverse of the Vandermonde matrix V[j][p] = gamma_j^p (3 x 3, by cofactors)
    Fp V[3][3], cof[3][3];
    for (int j = 0; j < 3; j++) {
      V[j][0] = Fp::one();
      V[j][1] = gamma[j];
      V[j][2] = gamma[j].sqr();
    }
    for (int r = 0; r < 3; r++)
      for (int q = 0; q < 3; q++) {
        const int r1 = (r + 1) % 3, r2 = (r + 2) % 3, q1 = (q + 1) % 3, q2 = (q + 2) % 3;
        cof[r][q] = V[r1][q1] * V[r2][q2] - V[r1][q2] * V[r2][q1];  // cyclic indexing: sign included
      }
    const Fp det = V[0][0] * cof[0][0] + V[0][1] * cof[0][1] + V[0][2] * cof[0][2];
    const Fp det_inv = det.inv();
    for (int p = 0; p < 3; p++)
      for (int j = 0; j < 3; j++) K.h_solve[p][j] = cof[j][p] * det_inv;  // inverse = adjugate^T / det
  }
  // selectors -> fixed columns
  int n_sel_cols = 0;
  std::vector<std::vector<uint8_t>> sel_cols;
  combine_selectors(L, K.selectors, &n_sel_cols, sel_cols);
  if (FIXED_SELECTOR_BASE + n_sel_cols != NUM_FIXED) return set_error(ctx, ZK_E_INVALID, "unexpected selector combination");
  for (int c = 0; c < NUM_FIXED; c++) {
    ZK_CUDA(ctx, cudaMalloc((void**)&K.fixed_values[c], n * sizeof(Fp)));
    ZK_CUDA(ctx, cudaMalloc((void**)&K.fixed_polys[c], n * sizeof(Fp)));
    ZK_CUDA(ctx, cudaMalloc((void**)&K.fixed_cosets[c], K.en * sizeof(Fp)));
  }
  for (int c = 0; c < NUM_PERM; c++) {
    ZK_CUDA(ctx, cudaMalloc((void**)&K.sigma_values[c], n * sizeof(Fp)));
    ZK_CUDA(ctx, cudaMalloc((void**)&K.sigma_polys[c], n * sizeof(Fp)));
    ZK_CUDA(ctx, cudaMalloc((void**)&K.sigma_cosets[c], K.en * sizeof(Fp)));
  }
  ZK_CUDA(ctx, cudaMalloc((void**)&K.l0, K.en * sizeof(Fp)));
  ZK_CUDA(ctx, cudaMalloc((void**)&K.l_last, K.en * sizeof(Fp)));
  ZK_CUDA(ctx, cudaMalloc((void**)&K.l_active, K.en * sizeof(Fp)));
  const unsigned T = 256, blocks = (unsigned)((n + T - 1) / T);
  fixed_columns_kernel<<<blocks, T, 0, st>>>(K.fixed_values[0], K.fixed_values[1], K.fixed_values[2], n);
  ctx->launches++;
  uint8_t* d_tmpl = nullptr;
  uint64_t* d_map = nullptr;
  DevTemps keygen_tmp;  // keygen's device temporaries: freed on every return path
  keygen_tmp.own(&d_tmpl);
  keygen_tmp.own(&d_map);
  ZK_CUDA(ctx, cudaMalloc((void**)&d_tmpl, (size_t)n_sel_cols * L.rows));
  for (int c = 0; c < n_sel_cols; c++) {
    ZK_CUDA(ctx, cudaMemcpyAsync(d_tmpl + (size_t)c * L.rows, sel_cols[c].data(), L.rows, cudaMemcpyHostToDevice, st));
    selector_column_kernel<<<blocks, T, 0, st>>>(d_tmpl + (size_t)c * L.rows, L.rows, n_compressions, n,
                                                 K.fixed_values[FIXED_SELECTOR_BASE + c]);
    ctx->launches++;
  }
  {  // constants column (fixed column 3): the IV words on the rows SEL_CONST pins
    uint64_t* d_const = nullptr;
    DevTemps tmp;
    tmp.own(&d_const);
    ZK_CUDA(ctx, cudaMalloc((void**)&d_const, (size_t)L.rows * 8));
    ZK_CUDA(ctx, cudaMemcpyAsync(d_const, L.constants.data(), (size_t)L.rows * 8, cudaMemcpyHostToDevice, st));
    constants_column_kernel<<<blocks, T, 0, st>>>(d_const, L.rows, n_compressions, n, K.fixed_values[FIXED_CONSTANTS]);
    ctx->launches++;
    ZK_CUDA(ctx, zk_stream_sync(ctx));
  }
  // permutation
  RegionPermutation perm(L.rows);
  for (auto& cc : L.copies) perm.copy(cc);
  ZK_CUDA(ctx, cudaMalloc((void**)&d_map, perm.mapping.size() * 8));
  ZK_CUDA(ctx, cudaMemcpyAsync(d_map, perm.mapping.data(), perm.mapping.size() * 8, cudaMemcpyHostToDevice, st));
  NttTables* TN = nullptr;
  if ((rc = ntt_tables(ctx, k, &TN))) return rc;
  Fp dp[NUM_PERM];
  dp[0] = Fp::one();
  for (int i = 1; i < NUM_PERM; i++) dp[i] = dp[i - 1] * Fp::delta();
  for (int c = 0; c < NUM_PERM; c++) {
    sigma_kernel<<<blocks, T, 0, st>>>(d_map, L.rows, n_compressions, n, c, TN->tw_fwd, dp[0], dp[1], dp[2], dp[3],
                                       dp[4], dp[5], dp[6], dp[7], K.sigma_values[c]);
    ctx->launches++;
  }
  // record chaining: h_i of a continuing compression and h'_i of its predecessor form a 2-cycle
  {
    std::vector<uint32_t> cells;
    const uint32_t pa = (uint32_t)RegionPermutation::perm_index(CHAIN_OUT_COLUMN),
                   pb = (uint32_t)RegionPermutation::perm_index(CHAIN_H_COLUMN);
    for (uint64_t j = 1; j < n_compressions; j++) {
      if (!K.chain[j]) continue;
      for (int i = 0; i < 8; i++) {
        const uint32_t ra = L.out_word_row[i], rb = L.h_word_row[i];
        // the patch below is halo2's `copy` only if both cells are untouched by the region's own copies
        if (perm.mapping[perm.at(((uint64_t)pa << 32) | ra)] != (((uint64_t)pa << 32) | ra) ||
            perm.mapping[perm.at(((uint64_t)pb << 32) | rb)] != (((uint64_t)pb << 32) | rb))
          return set_error(ctx, ZK_E_INVALID, "chaining cells take part in other copy constraints");
        cells.insert(cells.end(), {pa, (uint32_t)((j - 1) * L.rows + ra), pb, (uint32_t)(j * L.rows + rb)});
      }
    }
    if (!cells.empty()) {
      uint32_t* d_cells = nullptr;
      Fp* d_dp = nullptr;
      Fp** d_sigma = nullptr;
      DevTemps tmp;
      tmp.own(&d_cells); tmp.own(&d_dp); tmp.own(&d_sigma);
      ZK_CUDA(ctx, cudaMalloc((void**)&d_cells, cells.size() * 4));
      ZK_CUDA(ctx, cudaMalloc((void**)&d_dp, sizeof dp));
      ZK_CUDA(ctx, cudaMalloc((void**)&d_sigma, sizeof(Fp*) * NUM_PERM));
      ZK_CUDA(ctx, cudaMemcpyAsync(d_cells, cells.data(), cells.size() * 4, cudaMemcpyHostToDevice, st));
      ZK_CUDA(ctx, cudaMemcpyAsync(d_dp, dp, sizeof dp, cudaMemcpyHostToDevice, st));
      ZK_CUDA(ctx, cudaMemcpyAsync(d_sigma, K.sigma_values, sizeof(Fp*) * NUM_PERM, cudaMemcpyHostToDevice, st));
      const uint32_t nc = (uint32_t)(cells.size() / 4);
      sigma_chain_kernel<<<(nc + 127) / 128, 128, 0, st>>>(d_cells, nc, n, TN->tw_fwd, d_dp, d_sigma);
      ctx->launches++;
      ZK_CUDA(ctx, zk_stream_sync(ctx));
    }
  }
  ZK_CUDA(ctx, cudaGetLastError());
  // commitments, coefficient forms, extended cosets
  NttOptions inv;
  inv.inverse = true;
  K.fixed_commitments.resize(NUM_FIXED);
  K.sigma_commitments.resize(NUM_PERM);
  for (int c = 0; c < NUM_FIXED; c++) {
    if ((rc = commit(ctx, K.fixed_values[c], S->params.fb_gl, n, Fp::one(), &K.fixed_commitments[c]))) return rc;
    if ((rc = ntt_run(ctx, K.fixed_values[c], (uint32_t)n, K.fixed_polys[c], k, inv))) return rc;
    if ((rc = coeff_to_extended(ctx, K, K.fixed_polys[c], K.fixed_cosets[c]))) return rc;
  }
  for (int c = 0; c < NUM_PERM; c++) {
    if ((rc = commit(ctx, K.sigma_values[c], S->params.fb_gl, n, Fp::one(), &K.sigma_commitments[c]))) return rc;
    if ((rc = ntt_run(ctx, K.sigma_values[c], (uint32_t)n, K.sigma_polys[c], k, inv))) return rc;
    if ((rc = coeff_to_extended(ctx, K, K.sigma_polys[c], K.sigma_cosets[c]))) return rc;
  }
  // l_0, l_last, l_blind -> l_active = 1 - (l_last + l_blind)
  {
    Fp *tmp = nullptr, *tmp2 = nullptr, *lblind = nullptr;
    DevTemps ltmp;
    ltmp.own(&tmp); ltmp.own(&tmp2); ltmp.own(&lblind);
    ZK_CUDA(ctx, cudaMalloc((void**)&tmp, n * sizeof(Fp)));
    ZK_CUDA(ctx, cudaMalloc((void**)&tmp2, n * sizeof(Fp)));
    ZK_CUDA(ctx, cudaMalloc((void**)&lblind, K.en * sizeof(Fp)));
    auto build = [&](int kind, Fp* out_coset) -> int32_t { return build_indicator(ctx, K, kind, tmp, tmp2, out_coset); };
    if ((rc = build(0, K.l0))) return rc;
    if ((rc = build(1, K.l_last))) return rc;
    if ((rc = build(2, lblind))) return rc;
    combine_active(ctx, K, lblind);
    ZK_CUDA(ctx, zk_stream_sync(ctx));
  }
  ZK_CUDA(ctx, zk_stream_sync(ctx));
  // vk.transcript_repr as VerifyingKey::from_parts derives it: hash of the `{:?}` rendering of vk.pinned()
  // (vk_repr.cpp); a value obtained from halo2 itself can still be injected with zk_vk_repr_override
  K.pinned_debug = vk_pinned_debug(k, K.selectors, K.fixed_commitments, K.sigma_commitments);
  K.transcript_repr = vk_transcript_repr(K.pinned_debug);
  S->has_keys = true;
  return ZK_OK;
}

// fixed commitments, sigma commitments (32 B compressed each), then transcript_repr (32 B)
extern "C" int32_t zk_vk_bytes(zk_ctx* ctx, uint8_t* out, uint64_t* len) {
  if (!ctx || !len) return ZK_E_INVALID;
  ProverState* S = prover_state(ctx);
  if (!S->has_keys) return set_error(ctx, ZK_E_STATE, "no keys");
  uint64_t need = (NUM_FIXED + NUM_PERM + 1) * 32;
  if (!out || *len < need) {
    *len = need;
    return ZK_E_BUFFER;
  }
  *len = need;
  size_t off = 0;
  for (auto& c : S->keys.fixed_commitments) {
    point_to_bytes(c, out + off);
    off += 32;
  }
  for (auto& c : S->keys.sigma_commitments) {
    point_to_bytes(c, out + off);
    off += 32;
  }
  fe_to_repr(S->keys.transcript_repr, out + off);
  return ZK_OK;
}

// Host-only: the same rendering for given commitments (12 fixed then 8 permutation commitments, 64-byte affine
// Montgomery x, y each), without a device — what a Rust caller with its own keygen_vk output would compare.
extern "C" int32_t zk_blake2f_pinned_debug(int32_t k, uint32_t rounds, const void* commitments, char* out,
                                           uint64_t* len) {
  if (!commitments || !len || k < 17 || k > 28) return ZK_E_INVALID;
  RegionLayout L;
  try {
    build_region_layout(rounds, L);
  } catch (std::exception&) {
    return ZK_E_INVALID;
  }
  SelectorExpr sel[NUM_SELECTORS];
  int n_sel_cols = 0;
  std::vector<std::vector<uint8_t>> sel_cols;
  combine_selectors(L, sel, &n_sel_cols, sel_cols);
  if (FIXED_SELECTOR_BASE + n_sel_cols != NUM_FIXED) return ZK_E_INVALID;
  const Affine* pts = (const Affine*)commitments;
  const std::vector<Affine> fixed(pts, pts + NUM_FIXED), sigma(pts + NUM_FIXED, pts + NUM_FIXED + NUM_PERM);
  const std::string d = vk_pinned_debug(k, sel, fixed, sigma);
  if (!out || *len < d.size()) {
    *len = d.size();
    return ZK_E_BUFFER;
  }
  memcpy(out, d.data(), d.size());
  *len = d.size();
  return ZK_OK;
}

// the Rust `{:?}` rendering of vk.pinned() the transcript_repr was hashed from (not NUL-terminated)
extern "C" int32_t zk_vk_pinned_debug(zk_ctx* ctx, char* out, uint64_t* len) {
  if (!ctx || !len) return ZK_E_INVALID;
  ProverState* S = prover_state(ctx);
  if (!S->has_keys) return set_error(ctx, ZK_E_STATE, "no keys");
  const std::string& d = S->keys.pinned_debug;
  if (!out || *len < d.size()) {
    *len = d.size();
    return ZK_E_BUFFER;
  }
  memcpy(out, d.data(), d.size());
  *len = d.size();
  return ZK_OK;
}

extern "C" int32_t zk_vk_repr_override(zk_ctx* ctx, const uint8_t repr[32]) {
  if (!ctx || !repr) return ZK_E_INVALID;
  ProverState* S = prover_state(ctx);
  if (!S->has_keys) return set_error(ctx, ZK_E_STATE, "no keys");
  uint64_t c[4];
  memcpy(c, repr, 32);
  if (Fp::geq_mod(c)) return set_error(ctx, ZK_E_INVALID, "non-canonical field element");
  S->keys.transcript_repr = Fp::from_canonical(c);
  return ZK_OK;
}
