// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CircuitShape: everything keygen derives from `configure` + a witness-less `synthesize`
// (fixed columns, compressed selectors, copy constraints), and a MockProver-equivalent
// checker — the only thing the reference's single live test does
// (spread_table.rs:759-763 `MockProver::<Fp>::run(17, &circuit, vec![])` / `prover.verify()`).
// Semantics restated from halo2_proofs 0.3.0 dev.rs (`MockProver::verify`: every gate
// polynomial on every usable row, every lookup input tuple present in the table on usable
// rows, every permutation cycle value-consistent).  Parity unpinned.
#pragma once
#include <map>
#include <set>
#include <sstream>
#include "blake2f_circuit.hpp"

namespace zko {

struct CircuitShape {
  int k = 0;
  size_t n = 0;
  uint32_t rounds = 12;
  size_t n_compressions = 0;
  ConstraintSystem cs;
  Blake2fConfig cfg;
  std::vector<std::vector<uint64_t>> fixed;  // [fixed column][row] integers (< 2^64)
  std::vector<uint8_t> chain;                // [compression]: continues the previous one
  std::vector<SelectorAssignment> selector_assignments;
  std::vector<Blake2fAssignment::Copy> copies;
  int blinding_factors = 0;
  size_t usable_rows = 0;
};

// keygen's witness-less pass: configure, synthesize shape, load table, compress selectors
static inline void build_shape(CircuitShape& sh, int k, uint32_t rounds, size_t n_compressions,
                               const uint8_t* chain = nullptr) {
  sh.chain.assign(n_compressions, 0);
  if (chain) sh.chain.assign(chain, chain + n_compressions);
  sh.k = k;
  sh.n = (size_t)1 << k;
  sh.rounds = rounds;
  sh.n_compressions = n_compressions;
  sh.cs = ConstraintSystem();
  sh.cfg = blake2f_configure(sh.cs);
  sh.blinding_factors = sh.cs.blinding_factors();
  sh.usable_rows = sh.n - (sh.blinding_factors + 1);
  if (sh.n < (1u << 16) + (size_t)sh.blinding_factors + 1) throw std::runtime_error("k too small");
  Blake2fAssignment as;
  as.want_witness = false;
  as.want_shape = true;
  blake2f_synthesize(as, k, rounds, nullptr, n_compressions, sh.blinding_factors, chain);
  sh.copies = as.copies;
  // SpreadTableChip::load (spread_table.rs:470-508); tail rows take the first row's value
  sh.fixed.assign(sh.cs.num_fixed_columns, std::vector<uint64_t>(sh.n, 0));
  for (uint32_t i = 0; i < (1u << 16); i++) {
    uint32_t tag, dense, spread;
    spread_table_row(i, tag, dense, spread);
    sh.fixed[sh.cfg.table_tag][i] = tag;
    sh.fixed[sh.cfg.table_dense][i] = dense;
    sh.fixed[sh.cfg.table_spread][i] = spread;
  }
  sh.fixed[sh.cfg.constants] = as.constants;
  auto combos = compress_selectors(sh.cs, as.selectors, sh.selector_assignments);
  for (auto& c : combos) {
    std::vector<uint64_t> col(sh.n);
    for (size_t r = 0; r < sh.n; r++) col[r] = c[r];
    sh.fixed.push_back(col);
  }
}

struct MockFailure {
  bool ok = true;
  std::string what;
};

static inline MockFailure mock_verify(const CircuitShape& sh,
                                      const std::vector<std::vector<uint64_t>>& advice) {
  MockFailure f;
  size_t n = sh.n;
  auto fail = [&](const std::string& s) {
    if (f.ok) {
      f.ok = false;
      f.what = s;
    }
  };
  // small-integer -> Fp cache for fixed values (0..3 and table values are converted on the fly)
  auto fixed_at = [&](const Expr& e, size_t row) {
    size_t r = (row + n + (size_t)((long)e.rotation)) % n;
    return Fp::from_u64(sh.fixed[e.column][r]);
  };
  auto advice_at = [&](const Expr& e, size_t row) {
    size_t r = (size_t)(((long)row + (long)e.rotation + (long)n) % (long)n);
    return Fp::from_u64(advice[e.column][r]);
  };
  // gates
  for (size_t row = 0; row < sh.usable_rows && f.ok; row++) {
    bool any = false;
    for (size_t c = 4; c < sh.fixed.size(); c++) any |= sh.fixed[c][row] != 0;  // compressed selector columns
    if (!any) continue;  // every polynomial carries a selector factor
    for (auto& g : sh.cs.gates)
      for (size_t pi = 0; pi < g.polys.size(); pi++) {
        Fp v = eval_fp(
            g.polys[pi], [&](const Expr& e) { return fixed_at(e, row); },
            [&](const Expr& e) { return advice_at(e, row); });
        if (!v.is_zero()) {
          std::ostringstream os;
          os << "ConstraintNotSatisfied gate='" << g.name << "' poly='" << g.poly_names[pi]
             << "' row=" << row;
          fail(os.str());
        }
      }
  }
  // lookups
  for (auto& l : sh.cs.lookups) {
    std::set<std::vector<uint64_t>> table;
    for (size_t row = 0; row < sh.usable_rows; row++) {
      std::vector<uint64_t> key;
      for (auto& te : l.table_expressions) {
        Fp v = eval_fp(
            te, [&](const Expr& e) { return fixed_at(e, row); },
            [&](const Expr& e) { return advice_at(e, row); });
        u64 raw[4];
        v.to_raw(raw);
        key.insert(key.end(), raw, raw + 4);
      }
      table.insert(key);
      if (row >= (1u << 16)) break;  // tail rows repeat row 0 of the table
    }
    for (size_t row = 0; row < sh.usable_rows && f.ok; row++) {
      std::vector<uint64_t> key;
      for (auto& ie : l.input_expressions) {
        Fp v = eval_fp(
            ie, [&](const Expr& e) { return fixed_at(e, row); },
            [&](const Expr& e) { return advice_at(e, row); });
        u64 raw[4];
        v.to_raw(raw);
        key.insert(key.end(), raw, raw + 4);
      }
      if (!table.count(key)) {
        std::ostringstream os;
        os << "Lookup input not in table, row=" << row;
        fail(os.str());
      }
    }
  }
  // permutation
  for (auto& c : sh.copies) {
    if (advice[c.lc][c.lr] != advice[c.rc][c.rr]) {
      std::ostringstream os;
      os << "Permutation: column " << c.lc << " row " << c.lr << " != column " << c.rc << " row "
         << c.rr;
      fail(os.str());
      break;
    }
  }
  return f;
}

}  // namespace zko
