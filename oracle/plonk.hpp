// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of halo2_proofs 0.3.0 (Pasta / IPA) keygen_vk, keygen_pk, create_proof and
// verify_proof for the BLAKE2f circuit — the call sequence the reference writes down at
// blake2f-circuit/benches/blake2f.rs:83-142:
//     Params::<EqAffine>::new(k) -> keygen_vk -> keygen_pk ->
//     Blake2bWrite::<_,_,Challenge255<_>>::init(vec![]) -> create_proof(.., rng, &mut transcript)
//     -> transcript.finalize();   SingleVerifier::new(&params) -> verify_proof(..).
// halo2_proofs is an un-vendored dependency (Cargo.lock:842-857); everything here is restated
// from its published algorithm (plonk/{keygen,prover,verifier}.rs, plonk/{permutation,lookup,
// vanishing}/*, poly/{multiopen,commitment}/*, transcript.rs), following SURVEY.md
// Appendix A.4-A.6 for the order of transcript writes, challenges and RNG draws.
//
// PARITY UNPINNED: the reference contains no advice value, commitment, challenge or proof
// byte, and no Rust toolchain exists in this image to run halo2 itself.
// Known deliberate substitutions, both isolated behind inputs a genuine value can replace:
//   * URS: params.hpp `generate_substitute` (Params::new is not reproducible offline);
//   * vk.transcript_repr: hash of a fixed circuit-version string plus the fixed/permutation
//     commitments, instead of the Rust `{:?}` rendering of `vk.pinned()`.
#pragma once
#include <algorithm>
#include <map>
#include <set>
#include "mock_prover.hpp"
#include "params.hpp"

namespace zko {

// ---- transcript.rs: Blake2bWrite / Blake2bRead with Challenge255 ----------------------------
struct Transcript {
  Blake2b st;
  std::vector<uint8_t> proof;       // write mode
  const uint8_t* in = nullptr;      // read mode
  size_t in_len = 0, pos = 0;
  Transcript() : st("Halo2-Transcript") {}
  Transcript(const uint8_t* p, size_t len) : st("Halo2-Transcript"), in(p), in_len(len) {}

  void common_point(const Affine& p) {
    if (p.is_identity()) throw std::runtime_error("cannot write points at infinity to the transcript");
    uint8_t b[65];
    b[0] = 1;
    p.x.to_repr(b + 1);
    p.y.to_repr(b + 33);
    st.update(b, 65);
  }
  void common_scalar(const Fp& s) {
    uint8_t b[33];
    b[0] = 2;
    s.to_repr(b + 1);
    st.update(b, 33);
  }
  Fp squeeze_challenge() {
    uint8_t z = 0;
    st.update(&z, 1);
    uint8_t out[64];
    st.finalize(out);
    return Fp::from_uniform_bytes(out);
  }
  void write_point(const Affine& p) {
    common_point(p);
    uint8_t b[32];
    p.to_bytes(b);
    proof.insert(proof.end(), b, b + 32);
  }
  void write_scalar(const Fp& s) {
    common_scalar(s);
    uint8_t b[32];
    s.to_repr(b);
    proof.insert(proof.end(), b, b + 32);
  }
  Affine read_point() {
    if (pos + 32 > in_len) throw std::runtime_error("proof truncated");
    Affine p;
    if (!Affine::from_bytes(in + pos, p)) throw std::runtime_error("invalid point encoding in proof");
    pos += 32;
    common_point(p);
    return p;
  }
  Fp read_scalar() {
    if (pos + 32 > in_len) throw std::runtime_error("proof truncated");
    Fp s;
    if (!Fp::from_repr(in + pos, s)) throw std::runtime_error("invalid field element encoding in proof");
    pos += 32;
    common_scalar(s);
    return s;
  }
};

struct VerifyingKey {
  Domain domain;
  CircuitShape shape;  // cs after selector compression, layout parameters
  std::vector<Affine> fixed_commitments, permutation_commitments;
  int cs_degree = 0;
  Fp transcript_repr;
};

struct ProvingKey {
  VerifyingKey vk;
  Poly l0, l_last, l_active_row;                     // extended cosets
  std::vector<Poly> fixed_values, fixed_polys, fixed_cosets;
  std::vector<Poly> perm_values, perm_polys, perm_cosets;  // sigma columns
};

// ---- vk.pinned() Debug rendering and VerifyingKey::from_parts (plonk.rs, plonk/circuit.rs) ------------------
// halo2_proofs 0.3.0 hashes `format!("{:?}", vk.pinned())` (length-prefixed, BLAKE2b-512 personalised
// "Halo2-Verify-Key") into vk.transcript_repr.  Rendered here generically from the constraint system:
// Expression's Debug prints Constant(c) / Fixed { query_index, column_index, rotation } / Advice { .. } /
// Negated(e) / Sum(a, b) / Product(a, b) / Scaled(e, c); field elements print as 0x + 64 hex digits
// big-endian, points as (x, y).  Restated from the published sources; not run against halo2 here
// (rust/xcheck prints halo2's own string for comparison).  Parity unpinned.
static inline std::string debug_fp_hex(const u64 raw[4]) {
  char buf[67];
  snprintf(buf, sizeof buf, "0x%016llx%016llx%016llx%016llx", (unsigned long long)raw[3],
           (unsigned long long)raw[2], (unsigned long long)raw[1], (unsigned long long)raw[0]);
  return buf;
}
template <class F>
static inline std::string debug_field(const F& v) {
  u64 raw[4];
  v.to_raw(raw);
  return debug_fp_hex(raw);
}
static inline std::string debug_point(const Affine& p) {
  if (p.is_identity()) return "Infinity";
  return "(" + debug_field(p.x) + ", " + debug_field(p.y) + ")";
}
static inline std::string debug_expr(const E& e) {
  auto rot = [](int r) { return "Rotation(" + std::to_string(r) + ")"; };
  switch (e->kind) {
    case Expr::Constant: return "Constant(" + debug_field(e->c) + ")";
    case Expr::Selector: return "Selector(Selector(" + std::to_string(e->index) + ", true))";
    case Expr::Fixed:
      return "Fixed { query_index: " + std::to_string(e->index) + ", column_index: " + std::to_string(e->column) +
             ", rotation: " + rot(e->rotation) + " }";
    case Expr::Advice:
      return "Advice { query_index: " + std::to_string(e->index) + ", column_index: " + std::to_string(e->column) +
             ", rotation: " + rot(e->rotation) + " }";
    case Expr::Negated: return "Negated(" + debug_expr(e->a) + ")";
    case Expr::Sum: return "Sum(" + debug_expr(e->a) + ", " + debug_expr(e->b) + ")";
    case Expr::Product: return "Product(" + debug_expr(e->a) + ", " + debug_expr(e->b) + ")";
    case Expr::Scaled: return "Scaled(" + debug_expr(e->a) + ", " + debug_field(e->c) + ")";
  }
  return "";
}
static inline std::string vk_pinned_debug(const VerifyingKey& vk) {
  const ConstraintSystem& cs = vk.shape.cs;
  auto list = [](const std::vector<std::string>& v) {
    std::string s = "[";
    for (size_t i = 0; i < v.size(); i++) s += (i ? ", " : "") + v[i];
    return s + "]";
  };
  auto column = [](int index, const char* type) {
    return "Column { index: " + std::to_string(index) + ", column_type: " + type + " }";
  };
  auto queries = [&](const std::vector<Query>& qs, const char* type) {
    std::vector<std::string> v;
    for (auto& q : qs) v.push_back("(" + column(q.column, type) + ", Rotation(" + std::to_string(q.rotation) + "))");
    return list(v);
  };
  std::vector<std::string> gates, perm, lookups, fixed_cm, perm_cm;
  for (auto& g : cs.gates)
    for (auto& p : g.polys) gates.push_back(debug_expr(p));
  for (int c : cs.permutation_columns) perm.push_back(column(c, "Advice"));
  for (auto& l : cs.lookups) {
    std::vector<std::string> in, tab;
    for (auto& e : l.input_expressions) in.push_back(debug_expr(e));
    for (auto& e : l.table_expressions) tab.push_back(debug_expr(e));
    lookups.push_back("Argument { input_expressions: " + list(in) + ", table_expressions: " + list(tab) + " }");
  }
  for (auto& c : vk.fixed_commitments) fixed_cm.push_back(debug_point(c));
  for (auto& c : vk.permutation_commitments) perm_cm.push_back(debug_point(c));
  const std::string min_degree = cs.minimum_degree < 0 ? "None" : "Some(" + std::to_string(cs.minimum_degree) + ")";
  return "PinnedVerificationKey { base_modulus: \"" + debug_fp_hex(FqParams::MOD) + "\", scalar_modulus: \"" +
         debug_fp_hex(FpParams::MOD) + "\", domain: PinnedEvaluationDomain { k: " + std::to_string(vk.domain.k) +
         ", extended_k: " + std::to_string(vk.domain.extended_k) + ", omega: " + debug_field(vk.domain.omega) +
         " }, cs: PinnedConstraintSystem { num_fixed_columns: " + std::to_string(cs.num_fixed_columns) +
         ", num_advice_columns: " + std::to_string(cs.num_advice_columns) +
         ", num_instance_columns: " + std::to_string(cs.num_instance_columns) +
         ", num_selectors: " + std::to_string(cs.num_selectors) + ", gates: " + list(gates) +
         ", advice_queries: " + queries(cs.advice_queries, "Advice") + ", instance_queries: [], fixed_queries: " +
         queries(cs.fixed_queries, "Fixed") + ", permutation: Argument { columns: " + list(perm) + " }, lookups: " +
         list(lookups) + ", constants: [], minimum_degree: " + min_degree + " }, fixed_commitments: " + list(fixed_cm) +
         ", permutation: VerifyingKey { commitments: " + list(perm_cm) + " } }";
}

static inline Fp vk_transcript_repr(const VerifyingKey& vk) {
  const std::string s = vk_pinned_debug(vk);
  Blake2b h("Halo2-Verify-Key");
  uint64_t len = s.size();
  h.update(&len, 8);
  h.update(s.data(), s.size());
  uint8_t out[64];
  h.finalize(out);
  return Fp::from_uniform_bytes(out);
}

static inline Poly small_to_poly(const std::vector<uint64_t>& v) {
  Poly p(v.size());
  std::map<uint64_t, Fp> cache;
  for (size_t i = 0; i < v.size(); i++) {
    if (v[i] < 4) {
      auto it = cache.find(v[i]);
      if (it == cache.end()) it = cache.emplace(v[i], Fp::from_u64(v[i])).first;
      p[i] = it->second;
    } else {
      p[i] = Fp::from_u64(v[i]);
    }
  }
  return p;
}

// permutation::keygen::Assembly (mapping / aux / sizes with the cycle-merging `copy`)
struct PermutationAssembly {
  size_t n;
  std::vector<int> columns;  // advice column indices
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> mapping, aux;
  std::vector<std::vector<uint32_t>> sizes;
  PermutationAssembly(size_t n_, const std::vector<int>& cols) : n(n_), columns(cols) {
    size_t m = cols.size();
    mapping.assign(m, {});
    for (size_t i = 0; i < m; i++) {
      mapping[i].resize(n);
      for (size_t j = 0; j < n; j++) mapping[i][j] = {(uint32_t)i, (uint32_t)j};
    }
    aux = mapping;
    sizes.assign(m, std::vector<uint32_t>(n, 1));
  }
  int col_pos(int advice_col) const {
    for (size_t i = 0; i < columns.size(); i++)
      if (columns[i] == advice_col) return (int)i;
    throw std::runtime_error("ColumnNotInPermutation");
  }
  void copy(int lcol, size_t lrow, int rcol, size_t rrow) {
    int lc = col_pos(lcol), rc = col_pos(rcol);
    auto left_cycle = aux[lc][lrow], right_cycle = aux[rc][rrow];
    if (left_cycle == right_cycle) return;
    if (sizes[left_cycle.first][left_cycle.second] < sizes[right_cycle.first][right_cycle.second])
      std::swap(left_cycle, right_cycle);
    sizes[left_cycle.first][left_cycle.second] += sizes[right_cycle.first][right_cycle.second];
    auto i = right_cycle;
    for (;;) {
      aux[i.first][i.second] = left_cycle;
      i = mapping[i.first][i.second];
      if (i == right_cycle) break;
    }
    std::swap(mapping[lc][lrow], mapping[rc][rrow]);
  }
};

// vk_only: stop after `keygen_vk` (commitments + transcript_repr), which is all `verify_proof` needs
static inline void keygen(const Params& params, uint32_t rounds, size_t n_compressions,
                          ProvingKey& pk, bool vk_only = false, const uint8_t* chain = nullptr) {
  VerifyingKey& vk = pk.vk;
  build_shape(vk.shape, params.k, rounds, n_compressions, chain);
  const ConstraintSystem& cs = vk.shape.cs;
  vk.cs_degree = cs.degree();
  vk.domain = Domain(vk.cs_degree, params.k);
  const Domain& d = vk.domain;
  size_t n = d.n;
  if (n < (size_t)cs.minimum_rows()) throw std::runtime_error("NotEnoughRowsAvailable");
  // fixed columns (table + compressed selectors)
  pk.fixed_values.clear();
  for (auto& col : vk.shape.fixed) pk.fixed_values.push_back(small_to_poly(col));
  // permutation
  PermutationAssembly as(n, cs.permutation_columns);
  for (auto& c : vk.shape.copies) as.copy(c.lc, c.lr, c.rc, c.rr);
  size_t m = cs.permutation_columns.size();
  std::vector<Fp> omega_powers(n);
  omega_powers[0] = Fp::one();
  for (size_t i = 1; i < n; i++) omega_powers[i] = omega_powers[i - 1] * d.omega;
  std::vector<Fp> delta_pow(m);
  delta_pow[0] = Fp::one();
  for (size_t i = 1; i < m; i++) delta_pow[i] = delta_pow[i - 1] * Fp::C().delta;
  pk.perm_values.assign(m, Poly(n));
  for (size_t i = 0; i < m; i++)
    parallel_for(n, [&](size_t b, size_t e) {
      for (size_t j = b; j < e; j++) {
        auto mp = as.mapping[i][j];
        pk.perm_values[i][j] = delta_pow[mp.first] * omega_powers[mp.second];
      }
    });
  // commitments (Blind::default() == 1)
  vk.fixed_commitments.clear();
  for (auto& f : pk.fixed_values) vk.fixed_commitments.push_back(params.commit_lagrange(f, Fp::one()).to_affine());
  vk.permutation_commitments.clear();
  for (auto& p : pk.perm_values) vk.permutation_commitments.push_back(params.commit_lagrange(p, Fp::one()).to_affine());
  vk.transcript_repr = vk_transcript_repr(vk);
  if (vk_only) return;
  // pk
  pk.fixed_polys.clear();
  pk.fixed_cosets.clear();
  for (auto& f : pk.fixed_values) {
    pk.fixed_polys.push_back(d.lagrange_to_coeff(f));
    pk.fixed_cosets.push_back(d.coeff_to_extended(pk.fixed_polys.back()));
  }
  pk.perm_polys.clear();
  pk.perm_cosets.clear();
  for (auto& p : pk.perm_values) {
    pk.perm_polys.push_back(d.lagrange_to_coeff(p));
    pk.perm_cosets.push_back(d.coeff_to_extended(pk.perm_polys.back()));
  }
  int bf = cs.blinding_factors();
  Poly l0(n, Fp::zero()), l_blind(n, Fp::zero()), l_last(n, Fp::zero());
  l0[0] = Fp::one();
  for (int i = 0; i < bf; i++) l_blind[n - 1 - i] = Fp::one();
  l_last[n - bf - 1] = Fp::one();
  pk.l0 = d.coeff_to_extended(d.lagrange_to_coeff(l0));
  Poly l_blind_e = d.coeff_to_extended(d.lagrange_to_coeff(l_blind));
  pk.l_last = d.coeff_to_extended(d.lagrange_to_coeff(l_last));
  pk.l_active_row.resize(d.extended_n);
  for (size_t i = 0; i < d.extended_n; i++)
    pk.l_active_row[i] = Fp::one() - (pk.l_last[i] + l_blind_e[i]);
}

// ---- multiopen: query bookkeeping (poly/multiopen.rs construct_intermediate_sets) ------------
struct QuerySets {
  struct Commitment {
    int id;                        // caller-side identity of the polynomial / commitment
    std::vector<int> point_indices;
    int set_index = -1;
    std::vector<Fp> evals;          // ordered by the set's point order (verifier only)
  };
  std::vector<Commitment> commitment_map;  // unique commitments, first-appearance order
  std::vector<std::vector<Fp>> point_sets;  // [set] -> points
};
struct OpenQuery {
  int id;
  Fp point;
  Fp eval;  // verifier only
};
static inline QuerySets construct_intermediate_sets(const std::vector<OpenQuery>& queries) {
  QuerySets qs;
  struct FpLess {
    bool operator()(const Fp& a, const Fp& b) const { return Fp::cmp(a, b) < 0; }
  };
  std::map<Fp, int, FpLess> point_index_map;
  for (auto& q : queries) {
    int num = (int)point_index_map.size();
    auto it = point_index_map.emplace(q.point, num).first;
    int pidx = it->second;
    auto pos = std::find_if(qs.commitment_map.begin(), qs.commitment_map.end(),
                            [&](const QuerySets::Commitment& c) { return c.id == q.id; });
    if (pos != qs.commitment_map.end()) {
      pos->point_indices.push_back(pidx);
    } else {
      QuerySets::Commitment c;
      c.id = q.id;
      c.point_indices.push_back(pidx);
      qs.commitment_map.push_back(c);
    }
  }
  std::map<int, Fp> inverse_point_index_map;
  for (auto& kv : point_index_map) inverse_point_index_map[kv.second] = kv.first;
  std::map<std::set<int>, int> point_idx_sets;
  std::vector<std::set<int>> commitment_sets;
  for (auto& c : qs.commitment_map) {
    std::set<int> s(c.point_indices.begin(), c.point_indices.end());
    commitment_sets.push_back(s);
    int num = (int)point_idx_sets.size();
    point_idx_sets.emplace(s, num);
  }
  for (size_t ci = 0; ci < qs.commitment_map.size(); ci++) {
    auto& c = qs.commitment_map[ci];
    c.set_index = point_idx_sets[commitment_sets[ci]];
    c.evals.assign(c.point_indices.size(), Fp::zero());
  }
  for (auto& q : queries) {
    int pidx = point_index_map[q.point];
    for (size_t ci = 0; ci < qs.commitment_map.size(); ci++) {
      auto& c = qs.commitment_map[ci];
      if (c.id != q.id) continue;
      std::vector<int> ordered(commitment_sets[ci].begin(), commitment_sets[ci].end());
      size_t off = std::find(ordered.begin(), ordered.end(), pidx) - ordered.begin();
      c.evals[off] = q.eval;
    }
  }
  qs.point_sets.assign(point_idx_sets.size(), {});
  for (auto& kv : point_idx_sets)
    for (int pidx : kv.first) qs.point_sets[kv.second].push_back(inverse_point_index_map[pidx]);
  return qs;
}

// Values of a query expression over all n rows (lookup::Argument::commit_permuted).
static inline Poly eval_expr_rows(const E& e, const std::vector<Poly>& fixed,
                                  const std::vector<Poly>& advice, size_t n) {
  Poly out(n);
  parallel_for(n, [&](size_t b, size_t en) {
    for (size_t row = b; row < en; row++)
      out[row] = eval_fp(
          e,
          [&](const Expr& q) { return fixed[q.column][(row + n + (size_t)((long)q.rotation)) % n]; },
          [&](const Expr& q) { return advice[q.column][(size_t)(((long)row + q.rotation + (long)n) % (long)n)]; });
  });
  return out;
}

struct ProverTrace {  // intermediate values exposed so the CUDA path can be compared phase by phase
  std::vector<Affine> advice_commitments, lookup_commitments, perm_commitments, h_commitments;
  Affine lookup_product_commitment, random_commitment;
  Fp theta, beta, gamma, y, x;
};

// ---- create_proof ---------------------------------------------------------------------------
// advice: NUM_ADVICE columns of n integers (the witness before blinding).
static inline std::vector<uint8_t> create_proof(const Params& params, const ProvingKey& pk,
                                                const std::vector<std::vector<uint64_t>>& advice_raw,
                                                XorShiftRng& rng, ProverTrace* trace = nullptr) {
  const VerifyingKey& vk = pk.vk;
  const ConstraintSystem& cs = vk.shape.cs;
  const Domain& d = vk.domain;
  const size_t n = d.n, en = d.extended_n;
  const int bf = cs.blinding_factors();
  const size_t unusable_start = n - (bf + 1);
  auto random = [&]() { return rng.random_field<Fp>(); };
  Transcript tr;
  tr.common_scalar(vk.transcript_repr);

  // -- advice: blinding rows, blinds, commitments
  size_t na = cs.num_advice_columns;
  std::vector<Poly> advice(na, Poly(n));
  for (size_t c = 0; c < na; c++) {
    parallel_for(n, [&](size_t b, size_t e) {
      for (size_t r = b; r < e; r++) advice[c][r] = Fp::from_u64(advice_raw[c][r]);
    });
    for (size_t r = unusable_start; r < n; r++) advice[c][r] = random();
  }
  std::vector<Fp> advice_blinds(na);
  for (auto& b : advice_blinds) b = random();
  {
    std::vector<Jac> cj(na);
    for (size_t c = 0; c < na; c++) cj[c] = params.commit_lagrange(advice[c], advice_blinds[c]);
    std::vector<Affine> ca(na);
    batch_normalize(cj.data(), ca.data(), na);
    for (auto& p : ca) tr.write_point(p);
    if (trace) trace->advice_commitments = ca;
  }
  std::vector<Poly> advice_polys(na);
  for (size_t c = 0; c < na; c++) advice_polys[c] = d.lagrange_to_coeff(advice[c]);

  Fp theta = tr.squeeze_challenge();

  // -- lookups: compress, permute, commit
  struct LookupState {
    Poly compressed_input, compressed_table, permuted_input, permuted_table;
    Poly permuted_input_poly, permuted_table_poly, product_poly;
    Fp permuted_input_blind, permuted_table_blind, product_blind;
  };
  std::vector<LookupState> lookups(cs.lookups.size());
  for (size_t li = 0; li < cs.lookups.size(); li++) {
    const LookupArg& arg = cs.lookups[li];
    LookupState& L = lookups[li];
    auto compress = [&](const std::vector<E>& exprs) {
      Poly acc(n, Fp::zero());
      for (auto& e : exprs) {
        Poly v = eval_expr_rows(e, pk.fixed_values, advice, n);
        parallel_for(n, [&](size_t b, size_t en2) {
          for (size_t i = b; i < en2; i++) acc[i] = acc[i] * theta + v[i];
        });
      }
      return acc;
    };
    L.compressed_input = compress(arg.input_expressions);
    L.compressed_table = compress(arg.table_expressions);
    // permute_expression_pair
    size_t usable = unusable_start;
    Poly pin(L.compressed_input.begin(), L.compressed_input.begin() + usable);
    std::sort(pin.begin(), pin.end(), [](const Fp& a, const Fp& b) { return Fp::cmp(a, b) < 0; });
    struct FpLess {
      bool operator()(const Fp& a, const Fp& b) const { return Fp::cmp(a, b) < 0; }
    };
    std::map<Fp, uint32_t, FpLess> leftover;
    for (size_t i = 0; i < usable; i++) leftover[L.compressed_table[i]]++;
    Poly ptab(usable, Fp::zero());
    std::vector<size_t> repeated;
    for (size_t row = 0; row < usable; row++) {
      if (row == 0 || pin[row] != pin[row - 1]) {
        ptab[row] = pin[row];
        auto it = leftover.find(pin[row]);
        if (it == leftover.end() || it->second == 0) throw std::runtime_error("ConstraintSystemFailure: lookup input not in table");
        it->second--;
      } else {
        repeated.push_back(row);
      }
    }
    for (auto& kv : leftover)
      for (uint32_t c = 0; c < kv.second; c++) {
        ptab[repeated.back()] = kv.first;
        repeated.pop_back();
      }
    if (!repeated.empty()) throw std::logic_error("lookup permutation: rows left over");
    for (int i = 0; i < bf + 1; i++) pin.push_back(random());
    for (int i = 0; i < bf + 1; i++) ptab.push_back(random());
    L.permuted_input = pin;
    L.permuted_table = ptab;
    L.permuted_input_poly = d.lagrange_to_coeff(pin);
    L.permuted_input_blind = random();
    Affine ci = params.commit_lagrange(pin, L.permuted_input_blind).to_affine();
    L.permuted_table_poly = d.lagrange_to_coeff(ptab);
    L.permuted_table_blind = random();
    Affine ct = params.commit_lagrange(ptab, L.permuted_table_blind).to_affine();
    tr.write_point(ci);
    tr.write_point(ct);
    if (trace) {
      trace->lookup_commitments.push_back(ci);
      trace->lookup_commitments.push_back(ct);
    }
  }

  Fp beta = tr.squeeze_challenge();
  Fp gamma = tr.squeeze_challenge();

  // -- permutation argument: grand products per chunk of (degree - 2) columns
  const size_t chunk_len = vk.cs_degree - 2;
  const std::vector<int>& pcols = cs.permutation_columns;
  struct PermSet {
    Poly z_poly;
    Fp blind;
  };
  std::vector<PermSet> perm_sets;
  {
    Fp deltaomega = Fp::one(), last_z = Fp::one();
    std::vector<Fp> omega_powers(n);
    omega_powers[0] = Fp::one();
    for (size_t i = 1; i < n; i++) omega_powers[i] = omega_powers[i - 1] * d.omega;
    for (size_t start = 0; start < pcols.size(); start += chunk_len) {
      size_t end = std::min(pcols.size(), start + chunk_len);
      Poly modified(n, Fp::one());
      for (size_t ci = start; ci < end; ci++) {
        const Poly& values = advice[pcols[ci]];
        const Poly& perm = pk.perm_values[ci];
        parallel_for(n, [&](size_t b, size_t e) {
          for (size_t i = b; i < e; i++) modified[i] *= beta * perm[i] + gamma + values[i];
        });
      }
      batch_invert(modified.data(), n);
      for (size_t ci = start; ci < end; ci++) {
        const Poly& values = advice[pcols[ci]];
        parallel_for(n, [&](size_t b, size_t e) {
          for (size_t i = b; i < e; i++)
            modified[i] *= deltaomega * omega_powers[i] * beta + gamma + values[i];
        });
        deltaomega *= Fp::C().delta;
      }
      Poly z(n);
      z[0] = last_z;
      for (size_t row = 1; row < n; row++) z[row] = z[row - 1] * modified[row - 1];
      for (size_t row = n - bf; row < n; row++) z[row] = random();
      last_z = z[n - (bf + 1)];
      PermSet s;
      s.blind = random();
      Affine c = params.commit_lagrange(z, s.blind).to_affine();
      s.z_poly = d.lagrange_to_coeff(z);
      tr.write_point(c);
      if (trace) trace->perm_commitments.push_back(c);
      perm_sets.push_back(std::move(s));
    }
  }

  // -- lookup grand product
  for (auto& L : lookups) {
    Poly prod(n);
    parallel_for(n, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; i++)
        prod[i] = (beta + L.permuted_input[i]) * (gamma + L.permuted_table[i]);
    });
    batch_invert(prod.data(), n);
    parallel_for(n, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; i++) {
        prod[i] *= L.compressed_input[i] + beta;
        prod[i] *= L.compressed_table[i] + gamma;
      }
    });
    Poly z(n);
    z[0] = Fp::one();
    for (size_t i = 1; i < n - bf; i++) z[i] = z[i - 1] * prod[i - 1];
    for (size_t i = n - bf; i < n; i++) z[i] = random();
    L.product_blind = random();
    Affine c = params.commit_lagrange(z, L.product_blind).to_affine();
    L.product_poly = d.lagrange_to_coeff(z);
    tr.write_point(c);
    if (trace) trace->lookup_product_commitment = c;
  }

  // -- vanishing argument: random polynomial
  Poly random_poly(n);
  for (auto& c : random_poly) c = random();
  Fp random_blind = random();
  {
    Affine c = params.commit(random_poly, random_blind).to_affine();
    tr.write_point(c);
    if (trace) trace->random_commitment = c;
  }

  Fp y = tr.squeeze_challenge();

  // -- h(X): all expressions folded with y on the extended coset
  Poly h(en, Fp::zero());
  {
    std::vector<Poly> advice_cosets(na);
    for (size_t c = 0; c < na; c++) advice_cosets[c] = d.coeff_to_extended(advice_polys[c]);
    const size_t ext_step = en / n;  // rotation by 1 on the extended domain = this many steps
    auto rot_idx = [&](size_t i, int rot) {
      long j = (long)i + (long)rot * (long)ext_step;
      long m = (long)en;
      return (size_t)(((j % m) + m) % m);
    };
    // gates
    for (auto& g : cs.gates)
      for (auto& poly : g.polys)
        parallel_for(en, [&](size_t b, size_t e) {
          for (size_t i = b; i < e; i++) {
            Fp v = eval_fp(
                poly, [&](const Expr& q) { return pk.fixed_cosets[q.column][rot_idx(i, q.rotation)]; },
                [&](const Expr& q) { return advice_cosets[q.column][rot_idx(i, q.rotation)]; });
            h[i] = h[i] * y + v;
          }
        });
    // permutation
    std::vector<Poly> z_cosets;
    for (auto& s : perm_sets) z_cosets.push_back(d.coeff_to_extended(s.z_poly));
    auto fold = [&](const std::function<Fp(size_t)>& f) {
      parallel_for(en, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; i++) h[i] = h[i] * y + f(i);
      });
    };
    const int last_rot = -(bf + 1);
    if (!perm_sets.empty()) {
      fold([&](size_t i) { return (Fp::one() - z_cosets[0][i]) * pk.l0[i]; });
      const Poly& zl = z_cosets.back();
      fold([&](size_t i) { return (zl[i] * zl[i] - zl[i]) * pk.l_last[i]; });
      for (size_t s = 1; s < perm_sets.size(); s++)
        fold([&](size_t i) { return (z_cosets[s][i] - z_cosets[s - 1][rot_idx(i, last_rot)]) * pk.l0[i]; });
      // coset points zeta * extended_omega^i
      std::vector<Fp> xs(en);
      xs[0] = d.g_coset;
      for (size_t i = 1; i < en; i++) xs[i] = xs[i - 1] * d.extended_omega;
      for (size_t s = 0; s < perm_sets.size(); s++) {
        size_t start = s * chunk_len, end = std::min(pcols.size(), start + chunk_len);
        Fp delta_start = beta * Fp::C().delta.pow_u64(s * chunk_len);
        fold([&](size_t i) {
          Fp left = z_cosets[s][rot_idx(i, 1)];
          for (size_t ci = start; ci < end; ci++)
            left *= advice_cosets[pcols[ci]][i] + beta * pk.perm_cosets[ci][i] + gamma;
          Fp right = z_cosets[s][i];
          Fp cur = delta_start * xs[i];
          for (size_t ci = start; ci < end; ci++) {
            right *= advice_cosets[pcols[ci]][i] + cur + gamma;
            cur *= Fp::C().delta;
          }
          return (left - right) * pk.l_active_row[i];
        });
      }
    }
    // lookups
    for (size_t li = 0; li < lookups.size(); li++) {
      auto& L = lookups[li];
      const LookupArg& arg = cs.lookups[li];
      Poly zc = d.coeff_to_extended(L.product_poly);
      Poly ic = d.coeff_to_extended(L.permuted_input_poly);
      Poly tc = d.coeff_to_extended(L.permuted_table_poly);
      auto compress_at = [&](const std::vector<E>& exprs, size_t i) {
        Fp acc = Fp::zero();
        for (auto& e : exprs)
          acc = acc * theta +
                eval_fp(
                    e, [&](const Expr& q) { return pk.fixed_cosets[q.column][rot_idx(i, q.rotation)]; },
                    [&](const Expr& q) { return advice_cosets[q.column][rot_idx(i, q.rotation)]; });
        return acc;
      };
      fold([&](size_t i) { return (Fp::one() - zc[i]) * pk.l0[i]; });
      fold([&](size_t i) { return (zc[i] * zc[i] - zc[i]) * pk.l_last[i]; });
      fold([&](size_t i) {
        Fp left = zc[rot_idx(i, 1)] * (ic[i] + beta) * (tc[i] + gamma);
        Fp right = zc[i] * (compress_at(arg.input_expressions, i) + beta) *
                   (compress_at(arg.table_expressions, i) + gamma);
        return (left - right) * pk.l_active_row[i];
      });
      fold([&](size_t i) { return (ic[i] - tc[i]) * pk.l0[i]; });
      fold([&](size_t i) {
        return (ic[i] - tc[i]) * (ic[i] - ic[rot_idx(i, -1)]) * pk.l_active_row[i];
      });
    }
    // divide by the vanishing polynomial
    size_t period = d.t_evaluations_inv.size();
    parallel_for(en, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; i++) h[i] *= d.t_evaluations_inv[i % period];
    });
  }
  Poly h_coeffs = d.extended_to_coeff(h);
  size_t pieces = h_coeffs.size() / n;
  std::vector<Poly> h_pieces(pieces);
  for (size_t p = 0; p < pieces; p++) h_pieces[p].assign(h_coeffs.begin() + p * n, h_coeffs.begin() + (p + 1) * n);
  std::vector<Fp> h_blinds(pieces);
  for (auto& b : h_blinds) b = random();
  {
    std::vector<Jac> cj(pieces);
    for (size_t p = 0; p < pieces; p++) cj[p] = params.commit(h_pieces[p], h_blinds[p]);
    std::vector<Affine> ca(pieces);
    batch_normalize(cj.data(), ca.data(), pieces);
    for (auto& c : ca) tr.write_point(c);
    if (trace) trace->h_commitments = ca;
  }

  Fp x = tr.squeeze_challenge();
  Fp xn = x.pow_u64(n);
  if (trace) {
    trace->theta = theta; trace->beta = beta; trace->gamma = gamma; trace->y = y; trace->x = x;
  }

  // -- evaluations
  for (auto& q : cs.advice_queries) tr.write_scalar(eval_polynomial(advice_polys[q.column], d.rotate_omega(x, q.rotation)));
  for (auto& q : cs.fixed_queries) tr.write_scalar(eval_polynomial(pk.fixed_polys[q.column], d.rotate_omega(x, q.rotation)));
  // vanishing.evaluate
  Poly h_poly(n, Fp::zero());
  Fp h_blind = Fp::zero();
  for (size_t p = pieces; p-- > 0;) {
    parallel_for(n, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; i++) h_poly[i] = h_poly[i] * xn + h_pieces[p][i];
    });
    h_blind = h_blind * xn + h_blinds[p];
  }
  tr.write_scalar(eval_polynomial(random_poly, x));
  // permutation common evals
  for (auto& p : pk.perm_polys) tr.write_scalar(eval_polynomial(p, x));
  // permutation product evals
  Fp x_next = d.rotate_omega(x, 1), x_last = d.rotate_omega(x, -(bf + 1)), x_inv = d.rotate_omega(x, -1);
  for (size_t s = 0; s < perm_sets.size(); s++) {
    tr.write_scalar(eval_polynomial(perm_sets[s].z_poly, x));
    tr.write_scalar(eval_polynomial(perm_sets[s].z_poly, x_next));
    if (s + 1 != perm_sets.size()) tr.write_scalar(eval_polynomial(perm_sets[s].z_poly, x_last));
  }
  for (auto& L : lookups) {
    tr.write_scalar(eval_polynomial(L.product_poly, x));
    tr.write_scalar(eval_polynomial(L.product_poly, x_next));
    tr.write_scalar(eval_polynomial(L.permuted_input_poly, x));
    tr.write_scalar(eval_polynomial(L.permuted_input_poly, x_inv));
    tr.write_scalar(eval_polynomial(L.permuted_table_poly, x));
  }

  // -- multiopen
  std::vector<const Poly*> polys;
  std::vector<Fp> blinds;
  std::vector<OpenQuery> queries;
  auto poly_id = [&](const Poly* p, const Fp& blind) {
    for (size_t i = 0; i < polys.size(); i++)
      if (polys[i] == p) return (int)i;
    polys.push_back(p);
    blinds.push_back(blind);
    return (int)polys.size() - 1;
  };
  auto add_query = [&](const Poly* p, const Fp& blind, const Fp& point) {
    queries.push_back(OpenQuery{poly_id(p, blind), point, Fp::zero()});
  };
  for (auto& q : cs.advice_queries) add_query(&advice_polys[q.column], advice_blinds[q.column], d.rotate_omega(x, q.rotation));
  for (auto& s : perm_sets) {
    add_query(&s.z_poly, s.blind, x);
    add_query(&s.z_poly, s.blind, x_next);
  }
  for (size_t s = perm_sets.size(); s-- > 0;) {
    if (s + 1 == perm_sets.size()) continue;
    add_query(&perm_sets[s].z_poly, perm_sets[s].blind, x_last);
  }
  for (auto& L : lookups) {
    add_query(&L.product_poly, L.product_blind, x);
    add_query(&L.permuted_input_poly, L.permuted_input_blind, x);
    add_query(&L.permuted_table_poly, L.permuted_table_blind, x);
    add_query(&L.permuted_input_poly, L.permuted_input_blind, x_inv);
    add_query(&L.product_poly, L.product_blind, x_next);
  }
  for (auto& q : cs.fixed_queries) add_query(&pk.fixed_polys[q.column], Fp::one(), d.rotate_omega(x, q.rotation));
  for (auto& p : pk.perm_polys) add_query(&p, Fp::one(), x);
  add_query(&h_poly, h_blind, x);
  add_query(&random_poly, random_blind, x);

  Fp x1 = tr.squeeze_challenge();
  Fp x2 = tr.squeeze_challenge();
  QuerySets qs = construct_intermediate_sets(queries);
  size_t nsets = qs.point_sets.size();
  std::vector<Poly> q_polys(nsets);
  std::vector<Fp> q_blinds(nsets, Fp::zero());
  for (auto& c : qs.commitment_map) {
    Poly& acc = q_polys[c.set_index];
    const Poly& np = *polys[c.id];
    if (acc.empty()) {
      acc = np;
    } else {
      parallel_for(n, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; i++) acc[i] = acc[i] * x1 + np[i];
      });
    }
    q_blinds[c.set_index] = q_blinds[c.set_index] * x1 + blinds[c.id];
  }
  Poly q_prime;
  for (size_t s = 0; s < nsets; s++) {
    Poly p = q_polys[s];
    for (auto& pt : qs.point_sets[s]) p = kate_division(p, pt);
    p.resize(n, Fp::zero());
    if (q_prime.empty()) {
      q_prime = p;
    } else {
      for (size_t i = 0; i < n; i++) q_prime[i] = q_prime[i] * x2 + p[i];
    }
  }
  Fp q_prime_blind = random();
  tr.write_point(params.commit(q_prime, q_prime_blind).to_affine());
  Fp x3 = tr.squeeze_challenge();
  for (size_t s = 0; s < nsets; s++) tr.write_scalar(eval_polynomial(q_polys[s], x3));
  Fp x4 = tr.squeeze_challenge();
  Poly p_poly = q_prime;
  Fp p_blind = q_prime_blind;
  for (size_t s = 0; s < nsets; s++) {
    parallel_for(n, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; i++) p_poly[i] = p_poly[i] * x4 + q_polys[s][i];
    });
    p_blind = p_blind * x4 + q_blinds[s];
  }

  // -- inner product argument (poly/commitment/prover.rs)
  {
    Poly s_poly(n);
    for (auto& c : s_poly) c = random();
    Fp s_at_x3 = eval_polynomial(s_poly, x3);
    s_poly[0] -= s_at_x3;
    Fp s_blind = random();
    tr.write_point(params.commit(s_poly, s_blind).to_affine());
    Fp xi = tr.squeeze_challenge();
    Fp z = tr.squeeze_challenge();
    Poly p_prime(n);
    for (size_t i = 0; i < n; i++) p_prime[i] = s_poly[i] * xi + p_poly[i];
    Fp v = eval_polynomial(p_prime, x3);
    p_prime[0] -= v;
    Fp f = s_blind * xi + p_blind;
    std::vector<Fp> b(n);
    b[0] = Fp::one();
    for (size_t i = 1; i < n; i++) b[i] = b[i - 1] * x3;
    std::vector<Affine> g = params.g;
    for (int j = 0; j < params.k; j++) {
      size_t half = (size_t)1 << (params.k - j - 1);
      Jac l = msm(p_prime.data() + half, g.data(), half);
      Jac r = msm(p_prime.data(), g.data() + half, half);
      Fp vl = inner_product(p_prime.data() + half, b.data(), half);
      Fp vr = inner_product(p_prime.data(), b.data() + half, half);
      Fp l_rand = random(), r_rand = random();
      l = l.add(Jac::from_affine(params.u).mul(vl * z)).add(Jac::from_affine(params.w).mul(l_rand));
      r = r.add(Jac::from_affine(params.u).mul(vr * z)).add(Jac::from_affine(params.w).mul(r_rand));
      tr.write_point(l.to_affine());
      tr.write_point(r.to_affine());
      Fp u = tr.squeeze_challenge();
      Fp u_inv = u.invert();
      parallel_for(half, [&](size_t bb, size_t ee) {
        for (size_t i = bb; i < ee; i++) {
          p_prime[i] = p_prime[i] + p_prime[i + half] * u_inv;
          b[i] = b[i] + b[i + half] * u;
        }
      });
      p_prime.resize(half);
      b.resize(half);
      // parallel_generator_collapse
      u64 ur[4];
      u.to_raw(ur);
      std::vector<Jac> gj(half);
      parallel_for(half, [&](size_t bb, size_t ee) {
        for (size_t i = bb; i < ee; i++)
          gj[i] = Jac::from_affine(g[i + half]).mul_raw(ur).add_affine(g[i]);
      }, 64);
      size_t chunk = 1 << 10;
      parallel_for((half + chunk - 1) / chunk, [&](size_t bb, size_t ee) {
        for (size_t c = bb; c < ee; c++) {
          size_t lo = c * chunk, hi = std::min(half, lo + chunk);
          batch_normalize(gj.data() + lo, g.data() + lo, hi - lo);
        }
      }, 1);
      g.resize(half);
      f += l_rand * u_inv;
      f += r_rand * u;
    }
    tr.write_scalar(p_prime[0]);
    tr.write_scalar(f);
  }
  return tr.proof;
}

// ---- verify_proof with SingleVerifier -------------------------------------------------------
struct MsmAcc {  // poly/commitment/msm.rs `MSM`
  const Params* params;
  std::vector<Fp> g_scalars;
  bool has_g = false;
  Fp w_scalar = Fp::zero(), u_scalar = Fp::zero();
  std::vector<Fp> other_scalars;
  std::vector<Affine> other_bases;
  explicit MsmAcc(const Params* p) : params(p) {}
  void append_term(const Fp& s, const Affine& p) {
    other_scalars.push_back(s);
    other_bases.push_back(p);
  }
  void add_constant_term(const Fp& c) {
    if (!has_g) {
      g_scalars.assign(params->n, Fp::zero());
      has_g = true;
    }
    g_scalars[0] += c;
  }
  void add_to_g_scalars(const std::vector<Fp>& s) {
    if (!has_g) {
      g_scalars.assign(params->n, Fp::zero());
      has_g = true;
    }
    for (size_t i = 0; i < s.size(); i++) g_scalars[i] += s[i];
  }
  void scale(const Fp& f) {
    if (has_g)
      for (auto& s : g_scalars) s *= f;
    for (auto& s : other_scalars) s *= f;
    w_scalar *= f;
    u_scalar *= f;
  }
  void add_msm(const MsmAcc& o) {
    other_scalars.insert(other_scalars.end(), o.other_scalars.begin(), o.other_scalars.end());
    other_bases.insert(other_bases.end(), o.other_bases.begin(), o.other_bases.end());
    if (o.has_g) add_to_g_scalars(o.g_scalars);
    w_scalar += o.w_scalar;
    u_scalar += o.u_scalar;
  }
  bool eval() const {
    Jac acc = msm(other_scalars.data(), other_bases.data(), other_scalars.size());
    acc = acc.add(Jac::from_affine(params->w).mul(w_scalar));
    acc = acc.add(Jac::from_affine(params->u).mul(u_scalar));
    if (has_g) acc = acc.add(msm(g_scalars.data(), params->g.data(), params->n));
    return acc.is_identity();
  }
};

static inline bool verify_proof(const Params& params, const VerifyingKey& vk, const uint8_t* proof,
                                size_t proof_len, std::string* why = nullptr) {
  try {
    const ConstraintSystem& cs = vk.shape.cs;
    const Domain& d = vk.domain;
    const size_t n = d.n;
    const int bf = cs.blinding_factors();
    Transcript tr(proof, proof_len);
    tr.common_scalar(vk.transcript_repr);
    size_t na = cs.num_advice_columns;
    std::vector<Affine> advice_commitments(na);
    for (auto& c : advice_commitments) c = tr.read_point();
    Fp theta = tr.squeeze_challenge();
    struct LookupV {
      Affine permuted_input_commitment, permuted_table_commitment, product_commitment;
      Fp product_eval, product_next_eval, permuted_input_eval, permuted_input_inv_eval, permuted_table_eval;
    };
    std::vector<LookupV> lookups(cs.lookups.size());
    for (auto& L : lookups) {
      L.permuted_input_commitment = tr.read_point();
      L.permuted_table_commitment = tr.read_point();
    }
    Fp beta = tr.squeeze_challenge();
    Fp gamma = tr.squeeze_challenge();
    const size_t chunk_len = vk.cs_degree - 2;
    const std::vector<int>& pcols = cs.permutation_columns;
    size_t nsets_perm = (pcols.size() + chunk_len - 1) / chunk_len;
    std::vector<Affine> perm_commitments(nsets_perm);
    for (auto& c : perm_commitments) c = tr.read_point();
    for (auto& L : lookups) L.product_commitment = tr.read_point();
    Affine random_commitment = tr.read_point();
    Fp y = tr.squeeze_challenge();
    size_t pieces = d.quotient_poly_degree;
    std::vector<Affine> h_commitments(pieces);
    for (auto& c : h_commitments) c = tr.read_point();
    Fp x = tr.squeeze_challenge();
    std::vector<Fp> advice_evals(cs.advice_queries.size()), fixed_evals(cs.fixed_queries.size());
    for (auto& e : advice_evals) e = tr.read_scalar();
    for (auto& e : fixed_evals) e = tr.read_scalar();
    Fp random_eval = tr.read_scalar();
    std::vector<Fp> perm_common_evals(pcols.size());
    for (auto& e : perm_common_evals) e = tr.read_scalar();
    struct PermEval {
      Fp eval, next_eval, last_eval;
      bool has_last;
    };
    std::vector<PermEval> perm_evals(nsets_perm);
    for (size_t s = 0; s < nsets_perm; s++) {
      perm_evals[s].eval = tr.read_scalar();
      perm_evals[s].next_eval = tr.read_scalar();
      perm_evals[s].has_last = s + 1 != nsets_perm;
      if (perm_evals[s].has_last) perm_evals[s].last_eval = tr.read_scalar();
    }
    for (auto& L : lookups) {
      L.product_eval = tr.read_scalar();
      L.product_next_eval = tr.read_scalar();
      L.permuted_input_eval = tr.read_scalar();
      L.permuted_input_inv_eval = tr.read_scalar();
      L.permuted_table_eval = tr.read_scalar();
    }
    // expected h(x)
    Fp xn = x.pow_u64(n);
    std::vector<Fp> l_evals = d.l_i_range(x, xn, -(bf + 1), 0);
    Fp l_last = l_evals[0], l_blind = Fp::zero(), l_0 = l_evals[1 + bf];
    for (int i = 1; i < 1 + bf; i++) l_blind += l_evals[i];
    Fp active = Fp::one() - (l_last + l_blind);
    Fp h_eval = Fp::zero();
    auto fold = [&](const Fp& v) { h_eval = h_eval * y + v; };
    auto ev = [&](const E& e) {
      return eval_fp(
          e, [&](const Expr& q) { return fixed_evals[q.index]; },
          [&](const Expr& q) { return advice_evals[q.index]; });
    };
    for (auto& g : cs.gates)
      for (auto& p : g.polys) fold(ev(p));
    auto advice_eval_of = [&](int column, int rot) {
      for (size_t i = 0; i < cs.advice_queries.size(); i++)
        if (cs.advice_queries[i].column == column && cs.advice_queries[i].rotation == rot) return advice_evals[i];
      throw std::runtime_error("permutation column is not queried at Rotation::cur()");
    };
    if (nsets_perm) {
      fold(l_0 * (Fp::one() - perm_evals[0].eval));
      const PermEval& last = perm_evals.back();
      fold(l_last * (last.eval.square() - last.eval));
      for (size_t s = 1; s < nsets_perm; s++) fold((perm_evals[s].eval - perm_evals[s - 1].last_eval) * l_0);
      for (size_t s = 0; s < nsets_perm; s++) {
        size_t start = s * chunk_len, end = std::min(pcols.size(), start + chunk_len);
        Fp left = perm_evals[s].next_eval;
        for (size_t ci = start; ci < end; ci++)
          left *= advice_eval_of(pcols[ci], 0) + beta * perm_common_evals[ci] + gamma;
        Fp right = perm_evals[s].eval;
        Fp cur = beta * x * Fp::C().delta.pow_u64(s * chunk_len);
        for (size_t ci = start; ci < end; ci++) {
          right *= advice_eval_of(pcols[ci], 0) + cur + gamma;
          cur *= Fp::C().delta;
        }
        fold((left - right) * active);
      }
    }
    for (size_t li = 0; li < lookups.size(); li++) {
      auto& L = lookups[li];
      const LookupArg& arg = cs.lookups[li];
      auto compress = [&](const std::vector<E>& ex) {
        Fp acc = Fp::zero();
        for (auto& e : ex) acc = acc * theta + ev(e);
        return acc;
      };
      fold(l_0 * (Fp::one() - L.product_eval));
      fold(l_last * (L.product_eval.square() - L.product_eval));
      Fp left = L.product_next_eval * (L.permuted_input_eval + beta) * (L.permuted_table_eval + gamma);
      Fp right = L.product_eval * (compress(arg.input_expressions) + beta) * (compress(arg.table_expressions) + gamma);
      fold((left - right) * active);
      fold(l_0 * (L.permuted_input_eval - L.permuted_table_eval));
      fold((L.permuted_input_eval - L.permuted_table_eval) * (L.permuted_input_eval - L.permuted_input_inv_eval) * active);
    }
    Fp expected_h_eval = h_eval * (xn - Fp::one()).invert();

    // queries; commitments are MSM accumulators so that h can be a combination
    std::vector<MsmAcc> commitments;
    std::vector<OpenQuery> queries;
    std::map<std::pair<int, int>, int> ids;  // (kind, index) -> commitment id
    auto commit_id = [&](int kind, int index, const Affine* pt) {
      auto key = std::make_pair(kind, index);
      auto it = ids.find(key);
      if (it != ids.end()) return it->second;
      MsmAcc m(&params);
      if (pt) m.append_term(Fp::one(), *pt);
      commitments.push_back(m);
      ids[key] = (int)commitments.size() - 1;
      return (int)commitments.size() - 1;
    };
    Fp x_next = d.rotate_omega(x, 1), x_last = d.rotate_omega(x, -(bf + 1)), x_inv = d.rotate_omega(x, -1);
    for (size_t i = 0; i < cs.advice_queries.size(); i++) {
      auto& q = cs.advice_queries[i];
      queries.push_back(OpenQuery{commit_id(0, q.column, &advice_commitments[q.column]), d.rotate_omega(x, q.rotation), advice_evals[i]});
    }
    for (size_t s = 0; s < nsets_perm; s++) {
      int id = commit_id(1, (int)s, &perm_commitments[s]);
      queries.push_back(OpenQuery{id, x, perm_evals[s].eval});
      queries.push_back(OpenQuery{id, x_next, perm_evals[s].next_eval});
    }
    for (size_t s = nsets_perm; s-- > 0;) {
      if (s + 1 == nsets_perm) continue;
      queries.push_back(OpenQuery{commit_id(1, (int)s, &perm_commitments[s]), x_last, perm_evals[s].last_eval});
    }
    for (size_t li = 0; li < lookups.size(); li++) {
      auto& L = lookups[li];
      int idz = commit_id(2, (int)li * 3, &L.product_commitment);
      int idi = commit_id(2, (int)li * 3 + 1, &L.permuted_input_commitment);
      int idt = commit_id(2, (int)li * 3 + 2, &L.permuted_table_commitment);
      queries.push_back(OpenQuery{idz, x, L.product_eval});
      queries.push_back(OpenQuery{idi, x, L.permuted_input_eval});
      queries.push_back(OpenQuery{idt, x, L.permuted_table_eval});
      queries.push_back(OpenQuery{idi, x_inv, L.permuted_input_inv_eval});
      queries.push_back(OpenQuery{idz, x_next, L.product_next_eval});
    }
    for (size_t i = 0; i < cs.fixed_queries.size(); i++) {
      auto& q = cs.fixed_queries[i];
      queries.push_back(OpenQuery{commit_id(3, q.column, &vk.fixed_commitments[q.column]), d.rotate_omega(x, q.rotation), fixed_evals[i]});
    }
    for (size_t i = 0; i < pcols.size(); i++)
      queries.push_back(OpenQuery{commit_id(4, (int)i, &vk.permutation_commitments[i]), x, perm_common_evals[i]});
    {
      int id = commit_id(5, 0, nullptr);
      MsmAcc& hm = commitments[id];
      for (size_t p = pieces; p-- > 0;) {
        hm.scale(xn);
        hm.append_term(Fp::one(), h_commitments[p]);
      }
      queries.push_back(OpenQuery{id, x, expected_h_eval});
      queries.push_back(OpenQuery{commit_id(6, 0, &random_commitment), x, random_eval});
    }

    // multiopen verifier
    Fp x1 = tr.squeeze_challenge();
    Fp x2 = tr.squeeze_challenge();
    QuerySets qs = construct_intermediate_sets(queries);
    size_t nsets = qs.point_sets.size();
    std::vector<MsmAcc> q_commitments(nsets, MsmAcc(&params));
    std::vector<Fp> x1_power(nsets, Fp::one());
    std::vector<std::vector<Fp>> q_eval_sets(nsets);
    for (size_t s = 0; s < nsets; s++) q_eval_sets[s].assign(qs.point_sets[s].size(), Fp::zero());
    for (size_t ci = qs.commitment_map.size(); ci-- > 0;) {
      auto& c = qs.commitment_map[ci];
      MsmAcc m = commitments[c.id];
      m.scale(x1_power[c.set_index]);
      q_commitments[c.set_index].add_msm(m);
      for (size_t i = 0; i < c.evals.size(); i++) q_eval_sets[c.set_index][i] += c.evals[i] * x1_power[c.set_index];
      x1_power[c.set_index] *= x1;
    }
    Affine q_prime_commitment = tr.read_point();
    Fp x3 = tr.squeeze_challenge();
    std::vector<Fp> u(nsets);
    for (auto& e : u) e = tr.read_scalar();
    Fp msm_eval = Fp::zero();
    for (size_t s = 0; s < nsets; s++) {
      Poly r_poly = lagrange_interpolate(qs.point_sets[s], q_eval_sets[s]);
      Fp r_eval = eval_polynomial(r_poly, x3);
      Fp e = u[s] - r_eval;
      for (auto& pt : qs.point_sets[s]) e *= (x3 - pt).invert();
      msm_eval = msm_eval * x2 + e;
    }
    Fp x4 = tr.squeeze_challenge();
    MsmAcc msm_acc(&params);
    msm_acc.append_term(Fp::one(), q_prime_commitment);
    Fp v = msm_eval;
    for (size_t s = 0; s < nsets; s++) {
      msm_acc.scale(x4);
      msm_acc.add_msm(q_commitments[s]);
      v = v * x4 + u[s];
    }
    // IPA verifier
    msm_acc.add_constant_term(-v);
    Affine s_commitment = tr.read_point();
    Fp xi = tr.squeeze_challenge();
    msm_acc.append_term(xi, s_commitment);
    Fp z = tr.squeeze_challenge();
    int k = params.k;
    std::vector<Affine> ls(k), rs(k);
    std::vector<Fp> us(k), us_inv(k);
    for (int j = 0; j < k; j++) {
      ls[j] = tr.read_point();
      rs[j] = tr.read_point();
      us[j] = tr.squeeze_challenge();
      us_inv[j] = us[j];
    }
    batch_invert(us_inv.data(), k);
    for (int j = 0; j < k; j++) {
      msm_acc.append_term(us_inv[j], ls[j]);
      msm_acc.append_term(us[j], rs[j]);
    }
    Fp c = tr.read_scalar();
    Fp neg_c = -c;
    Fp f = tr.read_scalar();
    Fp b = Fp::one(), cur = x3;
    for (int j = k - 1; j >= 0; j--) {
      b *= Fp::one() + us[j] * cur;
      cur = cur.square();
    }
    msm_acc.u_scalar += neg_c * b * z;
    msm_acc.w_scalar += -f;
    // SingleVerifier: use_challenges -> compute_s
    std::vector<Fp> s(n, Fp::zero());
    s[0] = neg_c;
    {
      size_t len = 1;
      for (int j = k - 1; j >= 0; j--, len <<= 1)
        for (size_t i = 0; i < len; i++) s[len + i] = s[i] * us[j];
    }
    msm_acc.add_to_g_scalars(s);
    if (tr.pos != proof_len) throw std::runtime_error("trailing bytes in proof");
    if (!msm_acc.eval()) {
      if (why) *why = "final MSM is not the identity";
      return false;
    }
    return true;
  } catch (std::exception& e) {
    if (why) *why = e.what();
    return false;
  }
}

}  // namespace zko
