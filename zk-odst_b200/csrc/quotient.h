// Arguments of the fused quotient kernel (quotient.cu).  Product code.
#pragma once
#include "gates.cuh"
#include "prover_state.h"

namespace zkodst {

struct QuotientArgs {
  const Fp* advice[NUM_ADVICE_COLUMNS];  // extended cosets, by halo2 advice column index
  const Fp* fixed[NUM_FIXED];
  const Fp* sigma[NUM_PERM];
  const Fp* perm_z[NUM_SETS];
  const Fp *lookup_z, *lookup_in, *lookup_tab;
  const Fp *l0, *l_last, *l_active;
  const Fp* tw_n;    // omega_n^i, i < n / 2
  Fp* h;
  SelectorExpr sel[NUM_SELECTORS];
  Fp theta, beta, gamma, y;
  Fp coset_gen[NUM_COSETS];  // c_j: the point of extended row (j, i) is c_j omega_n^i
  Fp delta_pow[NUM_PERM];
  Fp t_inv[NUM_COSETS];
  GateConsts k;    // small integers and powers of two
  Fp ypow[NUM_GATE_POLYS];  // ypow[k] = y^(NUM_GATE_POLYS - 1 - k)
};

// evaluates h on rows [lo, hi) of the NUM_COSETS cosets (coset-major arrays of NUM_COSETS * n elements);
// a row reads its own and its rotated neighbours' column values and writes only h[row], so disjoint row
// ranges may run on different GPUs that hold the same columns
int32_t quotient_run(zk_ctx* ctx, const QuotientArgs& args, uint64_t n, uint64_t lo, uint64_t hi);

}  // namespace zkodst
