// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Restatement of the parts of halo2_proofs 0.3.0 `plonk::circuit` (un-vendored dependency,
// /root/reference/Cargo.lock:842-857) that the reference chip drives from
// `Table16Chip::configure` (table16.rs:277-327), `SpreadTableChip::configure`
// (spread_table.rs:425-467) and `CompressionConfig::configure` (compression.rs:555-1074):
// `Expression`, `ConstraintSystem::{advice_column, lookup_table_column, selector,
// enable_equality, query_advice, create_gate, lookup, degree, blinding_factors,
// compress_selectors}`.  Parity unpinned.
#pragma once
#include <algorithm>
#include <functional>
#include <memory>
#include <string>
#include <vector>
#include "field.hpp"

namespace zko {

struct Expr;
typedef std::shared_ptr<const Expr> E;

struct Expr {
  enum Kind { Constant, Selector, Fixed, Advice, Negated, Sum, Product, Scaled } kind;
  Fp c;            // Constant / Scaled factor
  int index = 0;   // Selector index, or query index for Fixed/Advice
  int column = 0;  // column index for Fixed/Advice
  int rotation = 0;
  E a, b;

  int degree() const {
    switch (kind) {
      case Constant: return 0;
      case Selector: case Fixed: case Advice: return 1;
      case Negated: case Scaled: return a->degree();
      case Sum: return std::max(a->degree(), b->degree());
      case Product: return a->degree() + b->degree();
    }
    return 0;
  }
};

static inline E e_const(const Fp& c) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::Constant;
  e->c = c;
  return e;
}
static inline E e_u64(uint64_t v) { return e_const(Fp::from_u64(v)); }
static inline E operator+(const E& a, const E& b) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::Sum;
  e->a = a;
  e->b = b;
  return e;
}
static inline E operator-(const E& a) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::Negated;
  e->a = a;
  return e;
}
static inline E operator-(const E& a, const E& b) { return a + (-b); }
static inline E operator*(const E& a, const E& b) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::Product;
  e->a = a;
  e->b = b;
  return e;
}
static inline E operator*(const E& a, const Fp& f) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::Scaled;
  e->a = a;
  e->c = f;
  return e;
}

// Generic evaluation (halo2 `Expression::evaluate`).
template <class T>
T eval_expr(const E& e, const std::function<T(const Fp&)>& constant,
            const std::function<T(int)>& selector, const std::function<T(const Expr&)>& fixed,
            const std::function<T(const Expr&)>& advice, const std::function<T(T)>& neg,
            const std::function<T(T, T)>& sum, const std::function<T(T, T)>& prod,
            const std::function<T(T, const Fp&)>& scaled) {
  switch (e->kind) {
    case Expr::Constant: return constant(e->c);
    case Expr::Selector: return selector(e->index);
    case Expr::Fixed: return fixed(*e);
    case Expr::Advice: return advice(*e);
    case Expr::Negated:
      return neg(eval_expr<T>(e->a, constant, selector, fixed, advice, neg, sum, prod, scaled));
    case Expr::Sum: {
      T x = eval_expr<T>(e->a, constant, selector, fixed, advice, neg, sum, prod, scaled);
      T y = eval_expr<T>(e->b, constant, selector, fixed, advice, neg, sum, prod, scaled);
      return sum(x, y);
    }
    case Expr::Product: {
      T x = eval_expr<T>(e->a, constant, selector, fixed, advice, neg, sum, prod, scaled);
      T y = eval_expr<T>(e->b, constant, selector, fixed, advice, neg, sum, prod, scaled);
      return prod(x, y);
    }
    case Expr::Scaled:
      return scaled(eval_expr<T>(e->a, constant, selector, fixed, advice, neg, sum, prod, scaled),
                    e->c);
  }
  return T();
}

// Field evaluation given resolvers for the two leaf kinds (selectors must already be
// substituted).
static inline Fp eval_fp(const E& e, const std::function<Fp(const Expr&)>& fixed,
                         const std::function<Fp(const Expr&)>& advice) {
  switch (e->kind) {
    case Expr::Constant: return e->c;
    case Expr::Selector: return Fp::zero();  // unreachable after compress_selectors
    case Expr::Fixed: return fixed(*e);
    case Expr::Advice: return advice(*e);
    case Expr::Negated: return -eval_fp(e->a, fixed, advice);
    case Expr::Sum: return eval_fp(e->a, fixed, advice) + eval_fp(e->b, fixed, advice);
    case Expr::Product: {
      Fp x = eval_fp(e->a, fixed, advice);
      if (x.is_zero()) return x;
      return x * eval_fp(e->b, fixed, advice);
    }
    case Expr::Scaled: return eval_fp(e->a, fixed, advice) * e->c;
  }
  return Fp::zero();
}

struct Query {
  int column, rotation;
  bool operator==(const Query& o) const { return column == o.column && rotation == o.rotation; }
};

struct Gate {
  std::string name;
  std::vector<std::string> poly_names;
  std::vector<E> polys;
};

struct LookupArg {
  std::vector<E> input_expressions;
  std::vector<E> table_expressions;
  int required_degree() const {
    int in_deg = 1, t_deg = 1;
    for (auto& e : input_expressions) in_deg = std::max(in_deg, e->degree());
    for (auto& e : table_expressions) t_deg = std::max(t_deg, e->degree());
    return std::max(4, 2 + in_deg + t_deg);
  }
};

struct ConstraintSystem {
  int num_fixed_columns = 0, num_advice_columns = 0, num_instance_columns = 0, num_selectors = 0;
  std::vector<Gate> gates;
  std::vector<Query> advice_queries, fixed_queries;
  std::vector<int> num_advice_queries;  // per advice column
  std::vector<int> permutation_columns;  // advice column indices, in enable_equality order
  std::vector<LookupArg> lookups;
  int minimum_degree = -1;
  std::vector<int> selector_fixed_col;  // after compression: fixed column of each selector

  int advice_column() {
    num_advice_queries.push_back(0);
    return num_advice_columns++;
  }
  int fixed_column() { return num_fixed_columns++; }
  int lookup_table_column() { return fixed_column(); }
  int selector() { return num_selectors++; }
  void enable_equality(int advice_col) {
    if (std::find(permutation_columns.begin(), permutation_columns.end(), advice_col) ==
        permutation_columns.end())
      permutation_columns.push_back(advice_col);
  }
  E query_selector(int s) {
    auto e = std::make_shared<Expr>();
    e->kind = Expr::Selector;
    e->index = s;
    return e;
  }
  int query_advice_index(int col, int rot) {
    for (size_t i = 0; i < advice_queries.size(); i++)
      if (advice_queries[i] == Query{col, rot}) return (int)i;
    advice_queries.push_back(Query{col, rot});
    num_advice_queries[col]++;
    return (int)advice_queries.size() - 1;
  }
  int query_fixed_index(int col, int rot) {
    for (size_t i = 0; i < fixed_queries.size(); i++)
      if (fixed_queries[i] == Query{col, rot}) return (int)i;
    fixed_queries.push_back(Query{col, rot});
    return (int)fixed_queries.size() - 1;
  }
  E query_advice(int col, int rot) {
    auto e = std::make_shared<Expr>();
    e->kind = Expr::Advice;
    e->index = query_advice_index(col, rot);
    e->column = col;
    e->rotation = rot;
    return e;
  }
  E query_fixed(int col, int rot) {
    auto e = std::make_shared<Expr>();
    e->kind = Expr::Fixed;
    e->index = query_fixed_index(col, rot);
    e->column = col;
    e->rotation = rot;
    return e;
  }
  // meta.lookup(|meta| vec![(input, table_column), ...]): inputs are queried by the closure,
  // then each table column is queried at Rotation::cur().
  void lookup(const std::vector<std::pair<E, int>>& map) {
    LookupArg l;
    for (auto& p : map) {
      l.input_expressions.push_back(p.first);
      l.table_expressions.push_back(query_fixed(p.second, 0));
    }
    lookups.push_back(l);
  }
  void create_gate(const std::string& name, const std::vector<std::pair<std::string, E>>& polys) {
    Gate g;
    g.name = name;
    for (auto& p : polys) {
      g.poly_names.push_back(p.first);
      g.polys.push_back(p.second);
    }
    gates.push_back(g);
  }

  int permutation_required_degree() const { return 3; }
  int degree() const {
    int d = permutation_columns.empty() ? 1 : permutation_required_degree();
    for (auto& l : lookups) d = std::max(d, l.required_degree());
    for (auto& g : gates)
      for (auto& p : g.polys) d = std::max(d, p->degree());
    return std::max(d, minimum_degree);
  }
  int blinding_factors() const {
    int factors = 1;
    for (int q : num_advice_queries) factors = std::max(factors, q);
    factors = std::max(3, factors);
    return factors + 2;
  }
  int minimum_rows() const { return blinding_factors() + 1 + 1 + 1 + 1; }
};

// ---- halo2_proofs 0.3.0 plonk/circuit/compress_selectors.rs `process` ---------------------
struct SelectorAssignment {
  int selector;
  int combination_index;  // index into the returned fixed-column list
  int fixed_column;       // allocated fixed column
  int assigned_root;      // value the combined column takes where this selector is on
  int combination_len;
};

// Returns new fixed columns (values as small integers) and fills `assignments`.
static inline std::vector<std::vector<uint8_t>> compress_selectors(
    ConstraintSystem& cs, const std::vector<std::vector<uint8_t>>& activations,
    std::vector<SelectorAssignment>& assignments) {
  int ns = cs.num_selectors;
  std::vector<int> sel_degree(ns, 0);
  // max degree of any gate polynomial in which the selector appears
  std::function<void(const E&, std::vector<int>&)> collect = [&](const E& e, std::vector<int>& out) {
    if (!e) return;
    if (e->kind == Expr::Selector) out.push_back(e->index);
    collect(e->a, out);
    collect(e->b, out);
  };
  for (auto& g : cs.gates)
    for (auto& p : g.polys) {
      std::vector<int> sels;
      collect(p, sels);
      for (int s : sels) sel_degree[s] = std::max(sel_degree[s], p->degree());
    }
  int max_degree = cs.degree();
  size_t n = activations.empty() ? 0 : activations[0].size();
  std::vector<std::vector<uint8_t>> combos;
  assignments.clear();
  std::vector<int> simple;
  for (int s = 0; s < ns; s++) {
    if (sel_degree[s] == 0) {
      int col = cs.fixed_column();
      combos.push_back(activations[s]);
      assignments.push_back(SelectorAssignment{s, (int)combos.size() - 1, col, 1, 1});
    } else {
      simple.push_back(s);
    }
  }
  size_t m = simple.size();
  std::vector<std::vector<bool>> excl(m);
  for (size_t i = 0; i < m; i++) {
    excl[i].assign(i, false);
    for (size_t j = 0; j < i; j++) {
      const auto &x = activations[simple[i]], &y = activations[simple[j]];
      for (size_t r = 0; r < n; r++)
        if (x[r] & y[r]) {
          excl[i][j] = true;
          break;
        }
    }
  }
  std::vector<bool> added(m, false);
  for (size_t i = 0; i < m; i++) {
    if (added[i]) continue;
    added[i] = true;
    int d = sel_degree[simple[i]] - 1;
    std::vector<size_t> comb = {i};
    for (size_t j = i + 1; j < m; j++) {
      if (d + (int)comb.size() == max_degree) break;
      if (added[j]) continue;
      bool excluded = false;
      for (size_t c : comb)
        if (excl[j][c]) excluded = true;
      if (excluded) continue;
      int new_d = std::max(d, sel_degree[simple[j]] - 1);
      if (new_d + (int)comb.size() + 1 > max_degree) continue;
      d = new_d;
      comb.push_back(j);
      added[j] = true;
    }
    int col = cs.fixed_column();
    std::vector<uint8_t> values(n, 0);
    for (size_t c = 0; c < comb.size(); c++) {
      int s = simple[comb[c]];
      for (size_t r = 0; r < n; r++)
        if (activations[s][r]) values[r] = (uint8_t)(c + 1);
      assignments.push_back(
          SelectorAssignment{s, (int)combos.size(), col, (int)c + 1, (int)comb.size()});
    }
    combos.push_back(values);
  }
  // substitute: selector -> q * prod_{j != root} (j - q)
  std::vector<E> repl(ns);
  for (auto& a : assignments) {
    E q = cs.query_fixed(a.fixed_column, 0);
    E ex = q;
    for (int root = 1; root <= a.combination_len; root++)
      if (root != a.assigned_root) ex = ex * (e_u64(root) - q);
    repl[a.selector] = ex;
  }
  std::function<E(const E&)> subst = [&](const E& e) -> E {
    if (!e) return e;
    if (e->kind == Expr::Selector) return repl[e->index];
    if (e->kind == Expr::Constant || e->kind == Expr::Fixed || e->kind == Expr::Advice) return e;
    auto r = std::make_shared<Expr>(*e);
    r->a = subst(e->a);
    r->b = subst(e->b);
    return r;
  };
  for (auto& g : cs.gates)
    for (auto& p : g.polys) p = subst(p);
  for (auto& l : cs.lookups) {
    for (auto& e : l.input_expressions) e = subst(e);
    for (auto& e : l.table_expressions) e = subst(e);
  }
  cs.selector_fixed_col.assign(ns, -1);
  for (auto& a : assignments) cs.selector_fixed_col[a.selector] = a.fixed_column;
  return combos;
}

}  // namespace zko
