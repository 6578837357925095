// vk.transcript_repr as halo2_proofs 0.3.0 derives it: the BLAKE2b-512
// hash (personalisation "Halo2-Verify-Key") of the length-prefixed Rust `{:?}` rendering of `vk.pinned()`,
// reduced with `from_uniform_bytes`.  Host code, product.
//
// Replaces `VerifyingKey::from_parts` / `hash_into` as reached from `keygen_vk`
// (blake2f-circuit/benches/blake2f.rs:102).  The Debug string nests
//   PinnedVerificationKey { base_modulus, scalar_modulus, domain: PinnedEvaluationDomain { k, extended_k, omega },
//     cs: PinnedConstraintSystem { num_fixed_columns, num_advice_columns, num_instance_columns, num_selectors,
//       gates: [flat list of the gate polynomials' Expression trees], advice_queries, instance_queries,
//       fixed_queries, permutation: Argument { columns }, lookups: [Argument { input_expressions,
//       table_expressions }], constants, minimum_degree }, fixed_commitments, permutation: VerifyingKey {
//       commitments } }
// and therefore depends on the exact shape of every Expression `configure` builds.  The builder below states
// `Table16Chip::configure` for the completed circuit (docs/CIRCUIT.md) once more, as strings: every operator
// appends the Debug form halo2's `Add` / `Sub` / `Neg` / `Mul` impls would produce (`a - b` = Sum(a, Negated(b)),
// `e * constant` = Scaled(e, c), `a * b` = Product(a, b)), and selectors are rendered as `compress_selectors`
// substitutes them.  rust/zkodst-backend/src/lib.rs holds the same `configure` in Rust; rust/xcheck prints
// halo2's own string for a character-by-character comparison (unverified here: no Rust toolchain in the image).
#include <string>
#include <vector>

#include "prover_state.h"
#include "transcript.h"

namespace zkodst {
namespace {

std::string hex_be(const uint64_t c[4]) {  // pasta_curves Debug of a field element: 0x + 64 hex digits, big-endian
  static const char* d = "0123456789abcdef";
  std::string s = "0x";
  for (int l = 3; l >= 0; l--)
    for (int b = 60; b >= 0; b -= 4) s += d[(c[l] >> b) & 15];
  return s;
}
template <class F>
std::string fe(const F& v) {
  uint64_t c[4];
  v.to_canonical(c);
  return hex_be(c);
}
std::string point(const Affine& p) {  // pasta_curves Debug of an affine point
  if (p.is_identity()) return "Infinity";
  return "(" + fe(p.x) + ", " + fe(p.y) + ")";
}

typedef std::string X;  // the Debug rendering of an Expression
X constant(const Fp& c) { return "Constant(" + fe(c) + ")"; }
X neg(const X& a) { return "Negated(" + a + ")"; }
X sum(const X& a, const X& b) { return "Sum(" + a + ", " + b + ")"; }
X sub(const X& a, const X& b) { return sum(a, neg(b)); }
X mul(const X& a, const X& b) { return "Product(" + a + ", " + b + ")"; }
X scaled(const X& a, const Fp& c) { return "Scaled(" + a + ", " + fe(c) + ")"; }
X rotation(int r) { return "Rotation(" + std::to_string(r) + ")"; }
X fixed_query(int column) {  // fixed query index == column index: every fixed column is queried once, at cur, in order
  return "Fixed { query_index: " + std::to_string(column) + ", column_index: " + std::to_string(column) +
         ", rotation: " + rotation(0) + " }";
}
X advice_query(int column, int rot) {
  for (int i = 0; i < 24; i++)
    if (ADVICE_QUERIES[i][0] == column && ADVICE_QUERIES[i][1] == rot)
      return "Advice { query_index: " + std::to_string(i) + ", column_index: " + std::to_string(column) +
             ", rotation: " + rotation(rot) + " }";
  return "Advice { unqueried }";  // unreachable: ADVICE_QUERIES lists every query of `configure`
}
Fp pow2(int e) {
  Fp r = Fp::one();
  for (int i = 0; i < e; i++) r = r.dbl();
  return r;
}

const int A_COLUMN[10] = {7, 8, 9, 1, 2, 0, 3, 4, 5, 6};  // a-number -> halo2 advice column (table16.rs:281-310)

struct Builder {
  const SelectorExpr* sel;
  std::vector<X> polys;
  X A(int an, int rot) const { return advice_query(A_COLUMN[an], rot); }
  // selector s after compress_selectors: q * prod_{root = 1..len, root != assigned} (Constant(root) - q)
  X S(int s) const {
    const X q = fixed_query(sel[s].fixed_col);
    X e = q;
    for (int root = 1; root <= sel[s].len; root++)
      if (root != sel[s].root) e = mul(e, sub(constant(Fp::from_u64((uint64_t)root)), q));
    return e;
  }
  void gate(const X& poly) { polys.push_back(poly); }
};

// sum of window inputs with limb weights 2^(w i): (in[0] + in[4] (+ in[8])) + sum_i (..) * 2^(w i)
X window_acc(const Builder& b, int w, bool three) {
  const int P = -1, C = 0;
  const X in[12] = {b.A(3, P), b.A(4, P), b.A(5, P), b.A(6, P), b.A(7, P), b.A(8, P),
                    b.A(3, C), b.A(4, C), b.A(5, C), b.A(6, C), b.A(7, C), b.A(8, C)};
  X acc = three ? sum(sum(in[0], in[4]), in[8]) : sum(in[0], in[4]);
  for (int i = 1; i < 4; i++)
    acc = sum(acc, scaled(three ? sum(sum(in[i], in[4 + i]), in[8 + i]) : sum(in[i], in[4 + i]), pow2(w * i)));
  return acc;
}

}  // namespace

// the Debug string of vk.pinned() for the BLAKE2f circuit at 2^k rows
std::string vk_pinned_debug(int k, const SelectorExpr sel[NUM_SELECTORS], const std::vector<Affine>& fixed_commitments,
                            const std::vector<Affine>& sigma_commitments) {
  const int P = -1, C = 0, N = 1;
  Builder b{sel, {}};
  const Fp one = Fp::one();
  // decompose ABCD
  b.gate(mul(b.S(SEL_ABCD), sub(sub(sub(sub(b.A(3, C), b.A(1, P)), scaled(b.A(1, C), pow2(16))), scaled(b.A(1, N), pow2(32))),
                                scaled(b.A(4, C), pow2(48)))));
  // Decompose EFGH: tag_p0, tag_p4, dense, spread
  b.gate(mul(b.S(SEL_EFGH), b.A(0, C)));
  b.gate(mul(b.S(SEL_EFGH), b.A(0, N)));
  b.gate(mul(b.S(SEL_EFGH), sub(sub(b.A(3, C), b.A(1, N)), scaled(b.A(1, C), pow2(8)))));
  b.gate(mul(b.S(SEL_EFGH), sub(sub(b.A(4, C), b.A(2, N)), scaled(b.A(2, C), pow2(16)))));
  // Decompose IJKL: tag_q0, bit, dense, spread
  b.gate(mul(b.S(SEL_IJKL), mul(b.A(0, C), sub(b.A(0, C), constant(one)))));
  b.gate(mul(b.S(SEL_IJKL), mul(b.A(5, C), sub(b.A(5, C), constant(one)))));
  b.gate(mul(b.S(SEL_IJKL), sub(sub(b.A(3, C), b.A(5, C)), scaled(b.A(1, C), pow2(1)))));
  b.gate(mul(b.S(SEL_IJKL), sub(sub(b.A(4, C), b.A(5, C)), scaled(b.A(2, C), pow2(2)))));
  auto add_gate = [&](int s, bool three) {
    const X acc = window_acc(b, 16, three);
    const X carry = b.A(9, C);
    const X lin = sub(sub(sub(sub(sub(acc, b.A(1, P)), scaled(b.A(1, C), pow2(16))), scaled(b.A(1, N), pow2(32))),
                          scaled(b.A(3, N), pow2(48))),
                      scaled(carry, pow2(64)));
    X rng = mul(carry, sub(carry, constant(one)));
    if (three) rng = mul(rng, sub(carry, constant(Fp::from_u64(2))));
    b.gate(mul(b.S(s), lin));
    b.gate(mul(b.S(s), rng));
  };
  auto xor_limb_gate = [&](int s) {
    b.gate(mul(b.S(s), sub(sub(sum(b.A(3, C), b.A(4, C)), b.A(2, C)), scaled(b.A(2, N), pow2(1)))));
  };
  auto xor_word_gate = [&](int s, const int offs[5]) {
    const X acc = window_acc(b, 32, false);
    const X p[5] = {b.A(5, C), b.A(6, C), b.A(7, C), b.A(8, C), b.A(3, N)};
    X even = scaled(p[0], pow2(2 * offs[0]));
    for (int i = 1; i < 5; i++) even = sum(even, scaled(p[i], pow2(2 * offs[i])));
    const X odd = sum(sum(sum(b.A(2, P), scaled(b.A(2, C), pow2(32))), scaled(b.A(2, N), pow2(64))),
                      scaled(b.A(4, N), pow2(96)));
    b.gate(mul(b.S(s), sub(sub(acc, even), scaled(odd, pow2(1)))));
  };
  const int offs_b1[5] = {0, 8, 24, 40, 56}, offs_b2[5] = {0, 15, 31, 47, 63};
  add_gate(SEL_A1, true);
  xor_limb_gate(SEL_D1);
  add_gate(SEL_C1, false);
  xor_word_gate(SEL_B1, offs_b1);
  add_gate(SEL_A2, true);
  xor_limb_gate(SEL_D2);
  add_gate(SEL_C2, false);
  xor_word_gate(SEL_B2, offs_b2);
  {  // s_digest: xor, word
    const X acc = window_acc(b, 32, false);
    const X even = sum(sum(sum(b.A(2, P), scaled(b.A(2, C), pow2(32))), scaled(b.A(2, N), pow2(64))),
                       scaled(b.A(5, C), pow2(96)));
    const X odd = sum(sum(sum(b.A(6, C), scaled(b.A(7, C), pow2(32))), scaled(b.A(8, C), pow2(64))),
                      scaled(b.A(3, N), pow2(96)));
    b.gate(mul(b.S(SEL_DIGEST), sub(sub(acc, even), scaled(odd, pow2(1)))));
    b.gate(mul(b.S(SEL_DIGEST), sub(sub(sub(sub(b.A(5, N), b.A(1, P)), scaled(b.A(1, C), pow2(16))),
                                        scaled(b.A(1, N), pow2(32))),
                                    scaled(b.A(4, N), pow2(48)))));
  }
  // pin constant; final flag: mask, bit
  b.gate(mul(b.S(SEL_CONST), sub(b.A(3, C), fixed_query(FIXED_CONSTANTS))));
  b.gate(mul(b.S(SEL_FMASK), sub(b.A(3, C), scaled(b.A(9, C), pow2(64) - one))));
  b.gate(mul(b.S(SEL_FMASK), mul(b.A(9, C), sub(b.A(9, C), constant(one)))));

  auto list = [](const std::vector<std::string>& v) {
    std::string s = "[";
    for (size_t i = 0; i < v.size(); i++) s += (i ? ", " : "") + v[i];
    return s + "]";
  };
  auto column = [](int index, const char* type) {
    return "Column { index: " + std::to_string(index) + ", column_type: " + type + " }";
  };
  std::vector<std::string> advice_queries, fixed_queries, perm_columns, fixed_cm, sigma_cm;
  for (auto& q : ADVICE_QUERIES) advice_queries.push_back("(" + column(q[0], "Advice") + ", " + rotation(q[1]) + ")");
  for (int c = 0; c < NUM_FIXED; c++) fixed_queries.push_back("(" + column(c, "Fixed") + ", " + rotation(0) + ")");
  for (int c : PERM_COLUMNS) perm_columns.push_back(column(c, "Advice"));
  for (auto& p : fixed_commitments) fixed_cm.push_back(point(p));
  for (auto& p : sigma_commitments) sigma_cm.push_back(point(p));
  const std::vector<std::string> lookup_in = {b.A(0, C), b.A(1, C), b.A(2, C)},
                                 lookup_tab = {fixed_query(0), fixed_query(1), fixed_query(2)};
  Fp omega = Fp::root_of_unity();
  for (int i = k; i < 32; i++) omega = omega.sqr();
  const uint64_t* qm = FqParams::MOD;
  const uint64_t* pm = FpParams::MOD;
  std::string s = "PinnedVerificationKey { base_modulus: \"" + hex_be(qm) + "\", scalar_modulus: \"" + hex_be(pm) +
                  "\", domain: PinnedEvaluationDomain { k: " + std::to_string(k) +
                  ", extended_k: " + std::to_string(k + 2) + ", omega: " + fe(omega) + " }, cs: PinnedConstraintSystem { " +
                  "num_fixed_columns: " + std::to_string(NUM_FIXED) + ", num_advice_columns: " +
                  std::to_string(NUM_ADVICE_COLUMNS) + ", num_instance_columns: 0, num_selectors: " +
                  std::to_string(NUM_SELECTORS) + ", gates: " + list(b.polys) + ", advice_queries: " + list(advice_queries) +
                  ", instance_queries: [], fixed_queries: " + list(fixed_queries) + ", permutation: Argument { columns: " +
                  list(perm_columns) + " }, lookups: [Argument { input_expressions: " + list(lookup_in) +
                  ", table_expressions: " + list(lookup_tab) + " }], constants: [], minimum_degree: None }, " +
                  "fixed_commitments: " + list(fixed_cm) + ", permutation: VerifyingKey { commitments: " + list(sigma_cm) +
                  " } }";
  return s;
}

// VerifyingKey::from_parts: transcript_repr = from_uniform_bytes(BLAKE2b-512("Halo2-Verify-Key", len || string))
Fp vk_transcript_repr(const std::string& pinned) {
  Blake2bState h("Halo2-Verify-Key");
  const uint64_t len = pinned.size();
  h.update(&len, 8);
  h.update(pinned.data(), pinned.size());
  uint8_t out[64];
  h.finalize(out);
  uint64_t w[8];
  memcpy(w, out, 64);
  return Fp::from_u512(w);
}

}  // namespace zkodst
