// Device-resident parameters and keys of the prover (product code).
//
// `DeviceParams` replaces halo2_proofs 0.3.0 `poly::commitment::Params<EqAffine>`
// (blake2f-circuit/benches/blake2f.rs:83-97); `DeviceKeys` replaces `ProvingKey`/`VerifyingKey`
// from `keygen_vk` / `keygen_pk` (benches/blake2f.rs:102-103) for the BLAKE2f Table16 circuit.
#pragma once
#include <cstdlib>
#include <string>
#include <vector>

#include "ec.cuh"
#include "msm_fixed.h"
#include "prover.h"

namespace zkodst {

constexpr int NUM_FIXED = 12;   // 3 table columns, the constants column, 8 compressed-selector columns
constexpr int FIXED_CONSTANTS = 3;       // fixed column of the pinned constants (IV words)
constexpr int FIXED_SELECTOR_BASE = 4;   // first column `compress_selectors` allocates
constexpr int NUM_PERM = 8;     // equality-enabled columns a_1..a_8 (table16.rs:312-314)
constexpr int NUM_SETS = 4;     // permutation grand products: chunks of degree - 2 = 2 columns
constexpr int BLINDING = 5;     // ConstraintSystem::blinding_factors() for this circuit
constexpr int CS_DEGREE = 4;
constexpr int NUM_COSETS = 3;        // quotient degree: h = h_0 + X^n h_1 + X^2n h_2
// IPA rounds on the original generators before they are folded once.  The rounds on g are MSMs split over a
// multi-GPU group by point range; the fold's bucket reduction, the window table over the folded generators and
// the rounds on them are not, and each halves with every extra round on g: a group of 2^e ranks folds e rounds
// later (n = 2^23 on 8 GPUs: the argument takes 89 ms at 5 rounds).  ZK_IPA_FOLD_ROUNDS
// overrides (1..8, for measurements).
inline int ipa_fold_rounds(int world) {
  int r = 5;
  for (int w = world; w > 1 && r < 8; w >>= 1) r++;
  if (const char* e = getenv("ZK_IPA_FOLD_ROUNDS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 8) r = v;
  }
  return r;
}
// Window bits of the table over the folded generators.  13, not 12: 255 = 21 * 12 + 3 leaves a 3-bit top window
// whose 8,192 digits land in 7 buckets per job, which then take the heavy path of the accumulation in every round
// (67 us per launch); 255 = 19 * 13 + 8 spreads the top window over 128 buckets.  Rounds on the folded generators of
// a k = 19 proof: 6.97 -> 6.47 ms (profiles/r02_ipa_stage2_sweep.jsonl).  ZK_IPA_STAGE2_C overrides (8..16).
inline int ipa_stage2_c() {
  if (const char* e = getenv("ZK_IPA_STAGE2_C")) {
    const int v = atoi(e);
    if (v >= 8 && v <= 16) return v;
  }
  return 13;
}

// halo2 advice column index of each permutation column, in enable_equality order
static const int PERM_COLUMNS[NUM_PERM] = {8, 9, 1, 2, 0, 3, 4, 5};

// advice_queries in first-use order of `configure` (docs/CIRCUIT.md §Queries): (advice column, rotation)
static const int ADVICE_QUERIES[24][2] = {{7, 0}, {8, 0},  {9, 0},  {1, 0},  {8, -1}, {8, 1},  {2, 0}, {7, 1},
                                          {9, 1}, {0, 0},  {1, -1}, {2, -1}, {0, -1}, {3, -1}, {4, -1}, {5, -1},
                                          {3, 0}, {4, 0},  {5, 0},  {1, 1},  {6, 0},  {9, -1}, {2, 1}, {0, 1}};

struct SelectorExpr {  // compress_selectors result: selector = q * prod_{r != root} (r - q)
  int fixed_col, root, len;
};

struct DeviceParams {
  int k = 0;
  uint64_t n = 0;
  Affine* g = nullptr;           // n + 2 entries: g[0..n), then w, u
  Affine* g_lagrange = nullptr;  // n + 1 entries: g_lagrange[0..n), then w
  Affine w, u;
  FixedBase fb_g, fb_gl;         // window tables over the two arrays above
  FixedBase fb_g8;               // 8-bit windows over g: folds the IPA generators (ipa_fold.cu); single GPU only
};

struct DeviceKeys {
  int k = 0;
  uint64_t n = 0, en = 0;   // en = NUM_COSETS * n: the quotient is evaluated on three cosets of the n-th roots
  uint32_t rounds = 0;
  uint64_t n_compressions = 0, region_rows = 0;
  std::vector<uint8_t> chain;  // [n_compressions]: compression j continues compression j - 1 (keygen_chained)
  SelectorExpr selectors[NUM_SELECTORS];
  Fp* fixed_values[NUM_FIXED] = {};
  Fp* fixed_polys[NUM_FIXED] = {};
  Fp* fixed_cosets[NUM_FIXED] = {};
  Fp* sigma_values[NUM_PERM] = {};
  Fp* sigma_polys[NUM_PERM] = {};
  Fp* sigma_cosets[NUM_PERM] = {};
  Fp *l0 = nullptr, *l_last = nullptr, *l_active = nullptr;  // extended cosets
  std::vector<Affine> fixed_commitments, sigma_commitments;
  Fp transcript_repr;
  std::string pinned_debug;  // the `{:?}` rendering of vk.pinned() transcript_repr hashes (vk_repr.cpp)
  // Quotient domain: h has degree < 3n, so three cosets c_j <omega_n> (c_j = zeta omega_4n^j, j = 0..2:
  // three of the four cosets halo2's extended domain consists of) determine it; stored coset-major.
  Fp coset_gen[NUM_COSETS];     // c_j
  Fp t_inv[NUM_COSETS];         // 1 / (c_j^n - 1): X^n - 1 is constant on a coset
  Fp h_solve[NUM_COSETS][NUM_COSETS];  // inverse of V[j][p] = (c_j^n)^p: coset coefficient vectors -> h pieces
  Fp* coset_scale = nullptr;    // [NUM_COSETS][n]: c_j^i
  Fp* coset_unscale = nullptr;  // [NUM_COSETS][n]: c_j^-i
  void* workspace = nullptr;  // ProofWorkspace (prover.cu), allocated lazily
};

struct ProverState {
  bool has_params = false, has_keys = false;
  DeviceParams params;
  DeviceKeys keys;
};

// vk_repr.cpp: the Debug string of halo2's vk.pinned() for this circuit, and VerifyingKey::from_parts' hash of it
std::string vk_pinned_debug(int k, const SelectorExpr sel[NUM_SELECTORS], const std::vector<Affine>& fixed_commitments,
                            const std::vector<Affine>& sigma_commitments);
Fp vk_transcript_repr(const std::string& pinned);

ProverState* prover_state(zk_ctx* ctx);
void free_keys(DeviceKeys& k);
void free_workspace(void* ws);

// commit: MSM(scalars || blind, bases || w) -> affine (host)
int32_t commit(zk_ctx* ctx, const Fp* d_scalars, const FixedBase& fb, uint64_t n, const Fp& blind, Affine* out);
// the same for nb columns over one base in a single MSM pipeline (nb <= MSM_MAX_BATCH)
int32_t commit_batch(zk_ctx* ctx, const Fp* const* d_scalars, const FixedBase& fb, uint64_t n, const Fp* blinds,
                     int nb, Affine* out);
// coefficients (n) -> evaluations on the three cosets c_j <omega_n>, coset-major (en = 3n)
int32_t coeff_to_extended(zk_ctx* ctx, const DeviceKeys& K, const Fp* coeffs, Fp* out);

}  // namespace zkodst
