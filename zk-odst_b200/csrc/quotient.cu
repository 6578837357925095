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

__device__ __forceinline__ Fp selector_expr(const Fp& q, int root, int len) {
  // q * prod_{r = 1..len, r != root} (r - q)   (compress_selectors substitution)
  Fp e = q;
  for (int r = 1; r <= len; r++)
    if (r != root) e = e * (qa.small[r] - q);
  return e;
}

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
  const Fp a0c = A0[i], a0n = A0[in];
  const Fp a1p = A1[ip], a1c = A1[i], a1n = A1[in];
  const Fp a2p = A2[ip], a2c = A2[i], a2n = A2[in];
  const Fp a3p = A3[ip], a3c = A3[i], a3n = A3[in];
  const Fp a4p = A4[ip], a4c = A4[i], a4n = A4[in];
  const Fp a5p = A5[ip], a5c = A5[i], a5n = A5[in];
  const Fp a6p = A6[ip], a6c = A6[i];
  const Fp a7p = A7[ip], a7c = A7[i];
  const Fp a8p = A8[ip], a8c = A8[i];
  const Fp a9c = A9[i];
  Fp sel[NUM_SELECTORS];
  {
    Fp q[NUM_FIXED];
#pragma unroll
    for (int c = 3; c < NUM_FIXED; c++) q[c] = qa.fixed[c][i];
#pragma unroll
    for (int s = 0; s < NUM_SELECTORS; s++)
      sel[s] = selector_expr(q[qa.sel[s].fixed_col], qa.sel[s].root, qa.sel[s].len);
  }
  const Fp one = Fp::one();
  const Fp &P1 = qa.pow2[1], &P2 = qa.pow2[2], &P8 = qa.pow2[8], &P16 = qa.pow2[16], &P30 = qa.pow2[30],
           &P32 = qa.pow2[32], &P48 = qa.pow2[48], &P62 = qa.pow2[62], &P64 = qa.pow2[64],
           &P80 = qa.pow2[80], &P94 = qa.pow2[94], &P96 = qa.pow2[96], &P112 = qa.pow2[112],
           &P126 = qa.pow2[126];
  Horner H{Fp::zero(), qa.y};

  // ---- gates, in declaration order ----------------------------------------------------------
  // decompose ABCD
  H.fold(sel[SEL_ABCD] * (a3c - a1p - a1c * P16 - a1n * P32 - a4c * P48));
  // Decompose EFGH: tag_p0, tag_p4, dense, spread
  H.fold(sel[SEL_EFGH] * a0c);
  H.fold(sel[SEL_EFGH] * a0n);
  H.fold(sel[SEL_EFGH] * (a3c - a1n - a1c * P8));
  H.fold(sel[SEL_EFGH] * (a4c - a2n - a2c * P16));
  // Decompose IJKL: tag_q0, bit, dense, spread
  H.fold(sel[SEL_IJKL] * (a0c * (a0c - one)));
  H.fold(sel[SEL_IJKL] * (a5c * (a5c - one)));
  H.fold(sel[SEL_IJKL] * (a3c - a5c - a1c * P1));
  H.fold(sel[SEL_IJKL] * (a4c - a5c - a2c * P2));
  // shared linear forms: 8 window inputs X = a3..a6[prev], Y = a7,a8[prev], a3,a4[cur]
  const Fp s0 = a3p + a7p, s1 = a4p + a8p, s2 = a5p + a3c, s3 = a6p + a4c;
  const Fp sum_tail = a1p + a1c * P16 + a1n * P32 + a3n * P48 + a9c * P64;
  const Fp add2_lin = s0 + s1 * P16 + s2 * P32 + s3 * P48 - sum_tail;
  const Fp add3_lin = add2_lin + a5c + a6c * P16 + a7c * P32 + a8c * P48;
  const Fp carry3 = a9c * (a9c - one) * (a9c - qa.small[2]);
  const Fp carry2 = a9c * (a9c - one);
  const Fp xor_limb = a3c + a4c - a2c - a2n * P1;
  const Fp acc32 = s0 + s1 * P32 + s2 * P64 + s3 * P96;
  const Fp odd_w = a2p + a2c * P32 + a2n * P64 + a4n * P96;
  // s_spread_a1: sum, carry
  H.fold(sel[SEL_A1] * add3_lin);
  H.fold(sel[SEL_A1] * carry3);
  // s_spread_d1
  H.fold(sel[SEL_D1] * xor_limb);
  // s_spread_c1: sum, carry
  H.fold(sel[SEL_C1] * add2_lin);
  H.fold(sel[SEL_C1] * carry2);
  // s_spread_b1: even pieces at bit offsets 0, 8, 24, 40, 56
  H.fold(sel[SEL_B1] * (acc32 - (a5c + a6c * P16 + a7c * P48 + a8c * P80 + a3n * P112) - odd_w * P1));
  // s_spread_a2
  H.fold(sel[SEL_A2] * add3_lin);
  H.fold(sel[SEL_A2] * carry3);
  // s_spread_d2
  H.fold(sel[SEL_D2] * xor_limb);
  // s_spread_c2
  H.fold(sel[SEL_C2] * add2_lin);
  H.fold(sel[SEL_C2] * carry2);
  // s_spread_b2: even pieces at bit offsets 0, 15, 31, 47, 63
  H.fold(sel[SEL_B2] * (acc32 - (a5c + a6c * P30 + a7c * P62 + a8c * P94 + a3n * P126) - odd_w * P1));
  // s_digest: xor, word
  H.fold(sel[SEL_DIGEST] * (acc32 - (a2p + a2c * P32 + a2n * P64 + a5c * P96) -
                            (a6c + a7c * P32 + a8c * P64 + a3n * P96) * P1));
  H.fold(sel[SEL_DIGEST] * (a5n - a1p - a1c * P16 - a1n * P32 - a4n * P48));

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
    const Fp vals[NUM_PERM] = {a1c, a2c, a3c, a4c, a5c, a6c, a7c, a8c};
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
    const Fp cin = (a0c * qa.theta + a1c) * qa.theta + a2c;
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
