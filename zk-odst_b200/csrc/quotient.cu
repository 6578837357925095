// K6 — fused quotient evaluation h(X) on the extended coset.
//
// Replaces the `poly::Evaluator` AST walk and `divide_by_vanishing_poly` of halo2_proofs 0.3.0
// (`vanishing::Argument::construct`, reached from `create_proof`,
// blake2f-circuit/benches/blake2f.rs:125) for this circuit: one kernel evaluates, per extended
// row, the 23 custom-gate polynomials of docs/CIRCUIT.md (gate names and order follow
// compression.rs:605-1056 / compression_gate.rs), the 9 permutation-argument terms and the 5
// lookup-argument terms, folds them with Horner in y in halo2's order, and multiplies by
// 1 / (X^n - 1) (four distinct values on the coset).
//
// Roofline: ~60 coset values read per row (32 B each, rotations hit L2) and ~200 Fp
// multiplications per row: integer-pipe bound; both fractions are reported by bench.py.
#include "polyops.cuh"
#include "prover_state.h"
#include "quotient.h"

namespace zkodst {
namespace {

__constant__ QuotientArgs qa;

struct Horner {
  Fp h, y;
  __device__ __forceinline__ void fold(const Fp& v) { h = h * y + v; }
};

__global__ void __launch_bounds__(128) quotient_kernel(uint64_t en, uint64_t mask) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= en) return;
  const uint64_t ip = (i - 4) & mask, in = (i + 4) & mask;  // rotation by one row = 4 steps
  // advice by a-number: a0..a9 -> halo2 columns 7,8,9,1,2,0,3,4,5,6
  const Fp* A0 = qa.advice[7];
  const Fp* A1 = qa.advice[8];
  const Fp* A2 = qa.advice[9];
  const Fp* A3 = qa.advice[1];
  const Fp* A4 = qa.advice[2];
  const Fp* A5 = qa.advice[0];
  const Fp* A6 = qa.advice[3];
  const Fp* A7 = qa.advice[4];
  const Fp* A8 = qa.advice[5];
  const Fp* A9 = qa.advice[6];
  GateCells v;
  v.a0c = A0[i], v.a0n = A0[in];
  v.a1p = A1[ip], v.a1c = A1[i], v.a1n = A1[in];
  v.a2p = A2[ip], v.a2c = A2[i], v.a2n = A2[in];
  v.a3p = A3[ip], v.a3c = A3[i], v.a3n = A3[in];
  v.a4p = A4[ip], v.a4c = A4[i], v.a4n = A4[in];
  v.a5p = A5[ip], v.a5c = A5[i], v.a5n = A5[in];
  v.a6p = A6[ip], v.a6c = A6[i];
  v.a7p = A7[ip], v.a7c = A7[i];
  v.a8p = A8[ip], v.a8c = A8[i];
  v.a9c = A9[i];
  Fp sel[NUM_SELECTORS];
  {
    Fp q[NUM_FIXED];
#pragma unroll
    for (int c = 3; c < NUM_FIXED; c++) q[c] = qa.fixed[c][i];
#pragma unroll
    for (int s = 0; s < NUM_SELECTORS; s++)
      sel[s] = selector_expr(q[qa.sel[s].fixed_col], qa.sel[s].root, qa.sel[s].len, qa.k.small);
  }
  const Fp one = Fp::one();
  Horner H{Fp::zero(), qa.y};

  // ---- gates, in declaration order (gates.cuh) --------------------------------------------------
  fold_gates(H, v, sel, qa.k);

  // ---- permutation argument -------------------------------------------------------------------
  const Fp l0 = qa.l0[i], l_last = qa.l_last[i], l_active = qa.l_active[i];
  const uint64_t ilast = (i - 4 * (BLINDING + 1)) & mask;
  Fp z[NUM_SETS];
#pragma unroll
  for (int s = 0; s < NUM_SETS; s++) z[s] = qa.perm_z[s][i];
  H.fold((one - z[0]) * l0);
  H.fold((z[NUM_SETS - 1] * z[NUM_SETS - 1] - z[NUM_SETS - 1]) * l_last);
#pragma unroll
  for (int s = 1; s < NUM_SETS; s++) H.fold((z[s] - qa.perm_z[s - 1][ilast]) * l0);
  {
    // coset point X = zeta * omega_ext^i
    const uint64_t half = en >> 1;
    Fp w = i < half ? qa.tw_ext[i] : qa.tw_ext[i - half].neg();
    const Fp bx = qa.beta * qa.zeta * w;
    // permutation columns in enable_equality order: a1,a2 | a3,a4 | a5,a6 | a7,a8 (all cur)
    const Fp vals[NUM_PERM] = {v.a1c, v.a2c, v.a3c, v.a4c, v.a5c, v.a6c, v.a7c, v.a8c};
#pragma unroll
    for (int s = 0; s < NUM_SETS; s++) {
      Fp left = qa.perm_z[s][in];
      Fp right = z[s];
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int ci = 2 * s + j;
        left = left * (vals[ci] + qa.beta * qa.sigma[ci][i] + qa.gamma);
        right = right * (vals[ci] + bx * qa.delta_pow[ci] + qa.gamma);
      }
      H.fold((left - right) * l_active);
    }
  }
  // ---- lookup argument ---------------------------------------------------------------------------
  {
    const Fp zl = qa.lookup_z[i], zl_next = qa.lookup_z[in];
    const Fp pin = qa.lookup_in[i], pin_prev = qa.lookup_in[ip], ptab = qa.lookup_tab[i];
    const Fp cin = (v.a0c * qa.theta + v.a1c) * qa.theta + v.a2c;
    const Fp ctab = (qa.fixed[0][i] * qa.theta + qa.fixed[1][i]) * qa.theta + qa.fixed[2][i];
    H.fold((one - zl) * l0);
    H.fold((zl * zl - zl) * l_last);
    H.fold((zl_next * (pin + qa.beta) * (ptab + qa.gamma) - zl * (cin + qa.beta) * (ctab + qa.gamma)) * l_active);
    H.fold((pin - ptab) * l0);
    H.fold((pin - ptab) * (pin - pin_prev) * l_active);
  }
  qa.h[i] = H.h * qa.t_inv[i & 3];
}

}  // namespace

int32_t quotient_run(zk_ctx* ctx, const QuotientArgs& args, uint64_t en) {
  ZK_CUDA(ctx, cudaMemcpyToSymbolAsync(qa, &args, sizeof(QuotientArgs), 0, cudaMemcpyHostToDevice, ctx->stream));
  KernelTimer timer(ctx, KC_QUOTIENT);
  quotient_kernel<<<(unsigned)((en + 127) / 128), 128, 0, ctx->stream>>>(en, en - 1);
  ctx->launches++;
  ZK_CUDA(ctx, cudaGetLastError());
  return ZK_OK;
}

}  // namespace zkodst
