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
  const Fp* tw_ext;  // omega_ext^i, i < en / 2
  Fp* h;
  SelectorExpr sel[NUM_SELECTORS];
  Fp theta, beta, gamma, y, zeta;
  Fp delta_pow[NUM_PERM];
  Fp t_inv[4];
  GateConsts k;    // small integers and powers of two
  Fp ypow[NUM_GATE_POLYS];  // ypow[k] = y^(NUM_GATE_POLYS - 1 - k)
};

int32_t quotient_run(zk_ctx* ctx, const QuotientArgs& args, uint64_t en);

}  // namespace zkodst
