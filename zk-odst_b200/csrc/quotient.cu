// K6 — fused quotient evaluation h(X) on three cosets of the n-th roots of unity.
//
// Replaces the `poly::Evaluator` AST walk and `divide_by_vanishing_poly` of halo2_proofs 0.3.0
// (`vanishing::Argument::construct`, reached from `create_proof`,
// blake2f-circuit/benches/blake2f.rs:125) for this circuit: one kernel evaluates, per extended
// row, the 26 custom-gate polynomials of docs/CIRCUIT.md (gate names and order follow
// compression.rs:605-1056 / compression_gate.rs), the 9 permutation-argument terms and the 5
// lookup-argument terms, folds them with Horner in y in halo2's order, and multiplies by
// 1 / (X^n - 1), which is constant on a coset.  halo2 evaluates on all four cosets of its extended
// domain (4n points); h has degree < 3n, so three of them (3n points) determine the same h.
//
// Roofline: ~60 coset values read per row (32 B each, rotations hit L2) and ~200 Fp
// multiplications per row: integer-pipe bound; both fractions are reported by bench.py.
#include <cstdlib>

#include "polyops.cuh"
#include "prover_state.h"
#include "quotient.h"

namespace zkodst {
namespace {

// The arguments travel as a (6 KB) __grid_constant__ kernel parameter rather than a __constant__
// symbol: contexts of different host threads may run quotient kernels on the same device concurrently.
// halo2 advice column of permutation column ci (PERM_COLUMNS, prover_state.h), usable in device code
#define PERM_COLUMNS_DEV(ci) ((ci) == 0 ? 8 : (ci) == 1 ? 9 : (ci) == 2 ? 1 : (ci) == 3 ? 2 : (ci) == 4 ? 0 : (ci) == 5 ? 3 : (ci) == 6 ? 4 : 5)

struct Horner {
  Fp h, y;
  __device__ __forceinline__ void fold(const Fp& v) { h = h * y + v; }
};

// The quotient is evaluated in three launches that hand the Horner accumulator through h[]: one kernel
// holding all 61 coset values of a row needs 255 registers (2 warps per scheduler, latency-bound);
// split, each part fits 4 blocks per SM.
constexpr int Q_MINB = 4;

// part 1: the 26 gate polynomials.  sum_k y^(25-k) sel_k e_k (what Horner over the gate list yields) is
// regrouped by expression — gates that share a polynomial (a1/a2, c1/c2, d1/d2) share its evaluation —
// and every cell and selector is fetched where it is used, which keeps the live set small.  The order
// of evaluation does not matter: field arithmetic is exact, the value equals fold_gates (gates.cuh).
__global__ void __launch_bounds__(128, Q_MINB) quotient_gates_kernel(const __grid_constant__ QuotientArgs qa, uint64_t n, uint64_t mask, uint64_t lo, uint64_t hi) {
  const uint64_t i = lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const uint64_t row = i & mask, base = i - row;  // coset-major: i = coset * n + row
  const uint64_t ip = base + ((row - 1) & mask), in = base + ((row + 1) & mask);  // rotations stay in the coset
  // advice by a-number: a0..a9 -> halo2 columns 7,8,9,1,2,0,3,4,5,6
  const Fp* const A0 = qa.advice[7];
  const Fp* const A1 = qa.advice[8];
  const Fp* const A2 = qa.advice[9];
  const Fp* const A3 = qa.advice[1];
  const Fp* const A4 = qa.advice[2];
  const Fp* const A5 = qa.advice[0];
  const Fp* const A6 = qa.advice[3];
  const Fp* const A7 = qa.advice[4];
  const Fp* const A8 = qa.advice[5];
  const Fp* const A9 = qa.advice[6];
  auto SEL = [&](int s) {
    return selector_expr(qa.fixed[qa.sel[s].fixed_col][i], qa.sel[s].root, qa.sel[s].len, qa.k.small);
  };
  const Fp* const YP = qa.ypow;  // YP[k] = y^(NUM_GATE_POLYS - 1 - k)
  const Fp* const P = qa.k.pow2;
  const Fp one = Fp::one();
  Fp acc;
  {  // decompose ABCD (k = 0)
    acc = SEL(SEL_ABCD) * YP[0] * (A3[i] - A1[ip] - A1[i] * P[16] - A1[in] * P[32] - A4[i] * P[48]);
  }
  {  // Decompose EFGH: tag_p0, tag_p4, dense, spread (k = 1..4)
    Fp t = A0[i] * YP[1] + A0[in] * YP[2];
    t = t + (A3[i] - A1[in] - A1[i] * P[8]) * YP[3];
    t = t + (A4[i] - A2[in] - A2[i] * P[16]) * YP[4];
    acc = acc + SEL(SEL_EFGH) * t;
  }
  {  // Decompose IJKL: tag_q0, bit, dense, spread (k = 5..8)
    const Fp a0c = A0[i], a5c = A5[i];
    Fp t = a0c * (a0c - one) * YP[5] + a5c * (a5c - one) * YP[6];
    t = t + (A3[i] - a5c - A1[i] * P[1]) * YP[7];
    t = t + (A4[i] - a5c - A2[i] * P[2]) * YP[8];
    acc = acc + SEL(SEL_IJKL) * t;
  }
  // shared window sums: X = a3..a6[prev], Y = a7,a8[prev], a3,a4[cur]
  const Fp s0 = A3[ip] + A7[ip], s1 = A4[ip] + A8[ip], s2 = A5[ip] + A3[i], s3 = A6[ip] + A4[i];
  {  // add2_lin (c1: k = 12, c2: k = 18), add3_lin (a1: k = 9, a2: k = 15)
    const Fp sum_tail = A1[ip] + A1[i] * P[16] + A1[in] * P[32] + A3[in] * P[48] + A9[i] * P[64];
    const Fp add2_lin = s0 + s1 * P[16] + s2 * P[32] + s3 * P[48] - sum_tail;
    acc = acc + add2_lin * (SEL(SEL_C1) * YP[12] + SEL(SEL_C2) * YP[18]);
    const Fp add3_lin = add2_lin + A5[i] + A6[i] * P[16] + A7[i] * P[32] + A8[i] * P[48];
    acc = acc + add3_lin * (SEL(SEL_A1) * YP[9] + SEL(SEL_A2) * YP[15]);
  }
  {  // carries: carry3 (a1: 10, a2: 16), carry2 (c1: 13, c2: 19)
    const Fp a9c = A9[i];
    const Fp carry2 = a9c * (a9c - one);
    acc = acc + carry2 * (SEL(SEL_C1) * YP[13] + SEL(SEL_C2) * YP[19]);
    acc = acc + carry2 * (a9c - qa.k.small[2]) * (SEL(SEL_A1) * YP[10] + SEL(SEL_A2) * YP[16]);
  }
  {  // xor_limb (d1: 11, d2: 17)
    const Fp xor_limb = A3[i] + A4[i] - A2[i] - A2[in] * P[1];
    acc = acc + xor_limb * (SEL(SEL_D1) * YP[11] + SEL(SEL_D2) * YP[17]);
  }
  {  // b1 (14), b2 (20), digest xor (21): all start from the 32-bit-limb accumulation of the window
    const Fp acc32 = s0 + s1 * P[32] + s2 * P[64] + s3 * P[96];
    const Fp odd_w = (A2[ip] + A2[i] * P[32] + A2[in] * P[64] + A4[in] * P[96]) * P[1];
    const Fp a5c = A5[i], a6c = A6[i], a7c = A7[i], a8c = A8[i], a3n = A3[in];
    // even pieces at bit offsets 0, 8, 24, 40, 56
    acc = acc + (acc32 - (a5c + a6c * P[16] + a7c * P[48] + a8c * P[80] + a3n * P[112]) - odd_w) *
                    (SEL(SEL_B1) * YP[14]);
    // even pieces at bit offsets 0, 15, 31, 47, 63
    acc = acc + (acc32 - (a5c + a6c * P[30] + a7c * P[62] + a8c * P[94] + a3n * P[126]) - odd_w) *
                    (SEL(SEL_B2) * YP[20]);
    // s_digest: xor (21), word (22)
    Fp t = (acc32 - (A2[ip] + A2[i] * P[32] + A2[in] * P[64] + a5c * P[96]) -
            (a6c + a7c * P[32] + a8c * P[64] + a3n * P[96]) * P[1]) * YP[21];
    t = t + (A5[in] - A1[ip] - A1[i] * P[16] - A1[in] * P[32] - A4[in] * P[48]) * YP[22];
    acc = acc + SEL(SEL_DIGEST) * t;
  }
  {  // pinned inputs: pin constant (23); final flag mask (24), bit (25)
    const Fp a3c = A3[i], a9c = A9[i];
    acc = acc + SEL(SEL_CONST) * YP[23] * (a3c - qa.fixed[FIXED_CONSTANTS][i]);
    acc = acc + SEL(SEL_FMASK) * ((a3c - a9c * (P[64] - one)) * YP[24] + a9c * (a9c - one));  // YP[25] = 1
  }
  qa.h[i] = acc;
}
// part 2: the permutation argument (columns in enable_equality order: a1,a2 | a3,a4 | a5,a6 | a7,a8)
__global__ void __launch_bounds__(128, Q_MINB) quotient_perm_kernel(const __grid_constant__ QuotientArgs qa, uint64_t n, uint64_t mask, uint64_t lo, uint64_t hi) {
  const uint64_t i = lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const uint64_t row = i & mask, base = i - row;  // coset-major: i = coset * n + row
  const uint64_t in = base + ((row + 1) & mask);
  const Fp one = Fp::one();
  Horner H{qa.h[i], qa.y};
  const Fp l0 = qa.l0[i], l_last = qa.l_last[i], l_active = qa.l_active[i];
  const uint64_t ilast = base + ((row - (BLINDING + 1)) & mask);
  Fp z[NUM_SETS];
#pragma unroll
  for (int s = 0; s < NUM_SETS; s++) z[s] = qa.perm_z[s][i];
  H.fold((one - z[0]) * l0);
  H.fold((z[NUM_SETS - 1] * z[NUM_SETS - 1] - z[NUM_SETS - 1]) * l_last);
#pragma unroll
  for (int s = 1; s < NUM_SETS; s++) H.fold((z[s] - qa.perm_z[s - 1][ilast]) * l0);
  // point of this row: X = c_coset * omega_n^row
  const uint32_t coset = (uint32_t)(base / n);
  const uint64_t half = n >> 1;
  const Fp w = row < half ? qa.tw_n[row] : qa.tw_n[row - half].neg();
  const Fp bx = qa.beta * qa.coset_gen[coset] * w;
#pragma unroll
  for (int s = 0; s < NUM_SETS; s++) {
    Fp left = qa.perm_z[s][in];
    Fp right = z[s];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int ci = 2 * s + j;
      const Fp val = qa.advice[PERM_COLUMNS_DEV(ci)][i];
      left = left * (val + qa.beta * qa.sigma[ci][i] + qa.gamma);
      right = right * (val + bx * qa.delta_pow[ci] + qa.gamma);
    }
    H.fold((left - right) * l_active);
  }
  qa.h[i] = H.h;
}

// part 3: the lookup argument, then the division by X^n - 1
__global__ void __launch_bounds__(128, Q_MINB) quotient_lookup_kernel(const __grid_constant__ QuotientArgs qa, uint64_t n, uint64_t mask, uint64_t lo, uint64_t hi) {
  const uint64_t i = lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const uint64_t row = i & mask, base = i - row;  // coset-major: i = coset * n + row
  const uint64_t ip = base + ((row - 1) & mask), in = base + ((row + 1) & mask);
  const Fp one = Fp::one();
  Horner H{qa.h[i], qa.y};
  const Fp l0 = qa.l0[i], l_last = qa.l_last[i], l_active = qa.l_active[i];
  const Fp zl = qa.lookup_z[i], zl_next = qa.lookup_z[in];
  const Fp pin = qa.lookup_in[i], pin_prev = qa.lookup_in[ip], ptab = qa.lookup_tab[i];
  const Fp cin = (qa.advice[7][i] * qa.theta + qa.advice[8][i]) * qa.theta + qa.advice[9][i];
  const Fp ctab = (qa.fixed[0][i] * qa.theta + qa.fixed[1][i]) * qa.theta + qa.fixed[2][i];
  H.fold((one - zl) * l0);
  H.fold((zl * zl - zl) * l_last);
  H.fold((zl_next * (pin + qa.beta) * (ptab + qa.gamma) - zl * (cin + qa.beta) * (ctab + qa.gamma)) * l_active);
  H.fold((pin - ptab) * l0);
  H.fold((pin - ptab) * (pin - pin_prev) * l_active);
  qa.h[i] = H.h * qa.t_inv[base / n];
}

}  // namespace

int32_t quotient_run(zk_ctx* ctx, const QuotientArgs& args, uint64_t n, uint64_t lo, uint64_t hi) {
  if (lo > hi || hi > NUM_COSETS * n) return set_error(ctx, ZK_E_INVALID, "quotient row range");
  if (lo == hi) return ZK_OK;
  KernelTimer timer(ctx, KC_QUOTIENT);
  const unsigned grid = (unsigned)((hi - lo + 127) / 128);
  // (measured and dropped, profiles/r02_quotient_variants.json: the gate kernel at 3 blocks per SM / 168 registers,
  // 4.68 ms per proof, and with the pinned-input terms in a launch of their own, 4.64 ms, against 4.75 ms as is —
  // within run-to-run noise, and the extra launch lengthens the proof)
  quotient_gates_kernel<<<grid, 128, 0, ctx->stream>>>(args, n, n - 1, lo, hi);
  quotient_perm_kernel<<<grid, 128, 0, ctx->stream>>>(args, n, n - 1, lo, hi);
  quotient_lookup_kernel<<<grid, 128, 0, ctx->stream>>>(args, n, n - 1, lo, hi);
  ctx->launches += 3;
  ZK_CUDA(ctx, cudaGetLastError());
  return ZK_OK;
}

}  // namespace zkodst
