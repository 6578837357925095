// The 26 custom-gate polynomials of the BLAKE2f Table16 circuit (docs/CIRCUIT.md), written once for
// the three places that evaluate them: the fused quotient kernel (values on the extended coset,
// quotient.cu), the MockProver-equivalent row checker (raw cell values, mock.cu) and the verifier
// (evaluations at the challenge x read from the proof, verifier.cu).  Product code, host + device.
//
// Gate names and order follow the reference's `CompressionConfig::configure`
// (blake2f-circuit/src/blake2f/table16/compression.rs:605-1056) and compression_gate.rs:22-557
// with the defects listed in SURVEY.md §2.3 fixed; the order is the order `create_gate` is called
// in, which is the order halo2 folds the polynomials with powers of y.
#pragma once
#include "blake2f_layout.h"
#include "field.cuh"

namespace zkodst {

// cells by a-number (a0..a9 = halo2 advice columns 7,8,9,1,2,0,3,4,5,6) and rotation p/c/n
struct GateCells {
  Fp a0c, a0n;
  Fp a1p, a1c, a1n;
  Fp a2p, a2c, a2n;
  Fp a3p, a3c, a3n;
  Fp a4p, a4c, a4n;
  Fp a5p, a5c, a5n;
  Fp a6p, a6c;
  Fp a7p, a7c;
  Fp a8p, a8c;
  Fp a9c;
};
static const int A_NUMBER_COLUMN[10] = {7, 8, 9, 1, 2, 0, 3, 4, 5, 6};
constexpr int NUM_GATE_POLYS = 26;

struct GateConsts {
  Fp small[4];   // 0, 1, 2, 3
  Fp pow2[127];  // 2^e
};

// compress_selectors substitution: selector = q * prod_{r = 1..len, r != root} (r - q)
ZK_HD Fp selector_expr(const Fp& q, int root, int len, const Fp* small) {
  Fp e = q;
  for (int r = 1; r <= len; r++)
    if (r != root) e = e * (small[r] - q);
  return e;
}

// Calls acc.fold(value) once per gate polynomial, in declaration order.  cfix = the constants fixed column.
template <class Acc>
ZK_HD void fold_gates(Acc& H, const GateCells& v, const Fp* sel, const GateConsts& k, const Fp& cfix) {
  const Fp one = Fp::one();
  const Fp &P1 = k.pow2[1], &P2 = k.pow2[2], &P8 = k.pow2[8], &P16 = k.pow2[16], &P30 = k.pow2[30],
           &P32 = k.pow2[32], &P48 = k.pow2[48], &P62 = k.pow2[62], &P64 = k.pow2[64], &P80 = k.pow2[80],
           &P94 = k.pow2[94], &P96 = k.pow2[96], &P112 = k.pow2[112], &P126 = k.pow2[126];
  // decompose ABCD
  H.fold(sel[SEL_ABCD] * (v.a3c - v.a1p - v.a1c * P16 - v.a1n * P32 - v.a4c * P48));
  // Decompose EFGH: tag_p0, tag_p4, dense, spread
  H.fold(sel[SEL_EFGH] * v.a0c);
  H.fold(sel[SEL_EFGH] * v.a0n);
  H.fold(sel[SEL_EFGH] * (v.a3c - v.a1n - v.a1c * P8));
  H.fold(sel[SEL_EFGH] * (v.a4c - v.a2n - v.a2c * P16));
  // Decompose IJKL: tag_q0, bit, dense, spread
  H.fold(sel[SEL_IJKL] * (v.a0c * (v.a0c - one)));
  H.fold(sel[SEL_IJKL] * (v.a5c * (v.a5c - one)));
  H.fold(sel[SEL_IJKL] * (v.a3c - v.a5c - v.a1c * P1));
  H.fold(sel[SEL_IJKL] * (v.a4c - v.a5c - v.a2c * P2));
  // shared linear forms: 8 window inputs X = a3..a6[prev], Y = a7,a8[prev], a3,a4[cur]
  const Fp s0 = v.a3p + v.a7p, s1 = v.a4p + v.a8p, s2 = v.a5p + v.a3c, s3 = v.a6p + v.a4c;
  const Fp sum_tail = v.a1p + v.a1c * P16 + v.a1n * P32 + v.a3n * P48 + v.a9c * P64;
  const Fp add2_lin = s0 + s1 * P16 + s2 * P32 + s3 * P48 - sum_tail;
  const Fp add3_lin = add2_lin + v.a5c + v.a6c * P16 + v.a7c * P32 + v.a8c * P48;
  const Fp carry3 = v.a9c * (v.a9c - one) * (v.a9c - k.small[2]);
  const Fp carry2 = v.a9c * (v.a9c - one);
  const Fp xor_limb = v.a3c + v.a4c - v.a2c - v.a2n * P1;
  const Fp acc32 = s0 + s1 * P32 + s2 * P64 + s3 * P96;
  const Fp odd_w = v.a2p + v.a2c * P32 + v.a2n * P64 + v.a4n * P96;
  // s_spread_a1: sum, carry
  H.fold(sel[SEL_A1] * add3_lin);
  H.fold(sel[SEL_A1] * carry3);
  // s_spread_d1
  H.fold(sel[SEL_D1] * xor_limb);
  // s_spread_c1: sum, carry
  H.fold(sel[SEL_C1] * add2_lin);
  H.fold(sel[SEL_C1] * carry2);
  // s_spread_b1: even pieces at bit offsets 0, 8, 24, 40, 56
  H.fold(sel[SEL_B1] *
         (acc32 - (v.a5c + v.a6c * P16 + v.a7c * P48 + v.a8c * P80 + v.a3n * P112) - odd_w * P1));
  // s_spread_a2
  H.fold(sel[SEL_A2] * add3_lin);
  H.fold(sel[SEL_A2] * carry3);
  // s_spread_d2
  H.fold(sel[SEL_D2] * xor_limb);
  // s_spread_c2
  H.fold(sel[SEL_C2] * add2_lin);
  H.fold(sel[SEL_C2] * carry2);
  // s_spread_b2: even pieces at bit offsets 0, 15, 31, 47, 63
  H.fold(sel[SEL_B2] *
         (acc32 - (v.a5c + v.a6c * P30 + v.a7c * P62 + v.a8c * P94 + v.a3n * P126) - odd_w * P1));
  // s_digest: xor, word
  H.fold(sel[SEL_DIGEST] * (acc32 - (v.a2p + v.a2c * P32 + v.a2n * P64 + v.a5c * P96) -
                            (v.a6c + v.a7c * P32 + v.a8c * P64 + v.a3n * P96) * P1));
  H.fold(sel[SEL_DIGEST] * (v.a5n - v.a1p - v.a1c * P16 - v.a1n * P32 - v.a4n * P48));
  // pin constant: the word cell of an IV slot equals the constants column
  H.fold(sel[SEL_CONST] * (v.a3c - cfix));
  // final flag: mask = bit * (2^64 - 1), bit boolean
  H.fold(sel[SEL_FMASK] * (v.a3c - v.a9c * (P64 - one)));
  H.fold(sel[SEL_FMASK] * (v.a9c * (v.a9c - one)));
}

}  // namespace zkodst
