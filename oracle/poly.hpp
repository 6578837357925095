// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Restatement of halo2_proofs 0.3.0 `arithmetic.rs` (best_fft, eval_polynomial,
// kate_division, lagrange_interpolate, compute_inner_product) and `poly/domain.rs`
// (EvaluationDomain), which the reference reaches from `create_proof` / `verify_proof`
// (blake2f-circuit/benches/blake2f.rs:125,142).  Un-vendored dependency
// (Cargo.lock:842-857): the algorithms are the published ones; only results that reach the
// transcript are contractual (coefficients, evaluations), not intermediate buffers.
// Parity unpinned.
#pragma once
#include <functional>
#include <thread>
#include <vector>
#include "field.hpp"

namespace zko {

// Worker threads of the oracle: ZKO_THREADS or all hardware threads; settable at run time
// (set_oracle_threads(1) mirrors the pinned reference, whose rayon feature is off: Cargo.toml:15).
static inline int& oracle_threads_ref() {
  static int n = [] {
    const char* e = getenv("ZKO_THREADS");
    int v = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    return v < 1 ? 1 : v;
  }();
  return n;
}
static inline int oracle_threads() { return oracle_threads_ref(); }
static inline void set_oracle_threads(int t) {
  oracle_threads_ref() = t < 1 ? (int)std::max(1u, std::thread::hardware_concurrency()) : t;
}

// parallel_for over [0, n) in contiguous chunks: fn(begin, end)
static inline void parallel_for(size_t n, const std::function<void(size_t, size_t)>& fn,
                                size_t min_chunk = 1024) {
  int t = oracle_threads();
  if (t <= 1 || n < 2 * min_chunk) {
    fn(0, n);
    return;
  }
  size_t chunks = std::min<size_t>(t, (n + min_chunk - 1) / min_chunk);
  size_t per = (n + chunks - 1) / chunks;
  std::vector<std::thread> th;
  for (size_t c = 0; c < chunks; c++) {
    size_t b = c * per, e = std::min(n, b + per);
    if (b >= e) break;
    th.emplace_back([=, &fn] { fn(b, e); });
  }
  for (auto& x : th) x.join();
}

typedef std::vector<Fp> Poly;

// In-place radix-2 FFT over Fp: a[i] <- sum_j a[j] omega^(ij)   (halo2 `best_fft`)
static inline void fft(Fp* a, int log_n, const Fp& omega) {
  size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) {
    size_t r = 0;
    for (int b = 0; b < log_n; b++) r |= ((i >> b) & 1) << (log_n - 1 - b);
    if (i < r) std::swap(a[i], a[r]);
  }
  std::vector<Fp> tw(n / 2 ? n / 2 : 1);
  tw[0] = Fp::one();
  for (size_t i = 1; i < n / 2; i++) tw[i] = tw[i - 1] * omega;
  for (int s = 1; s <= log_n; s++) {
    size_t m = (size_t)1 << s, half = m / 2, stride = n / m;
    parallel_for(n / 2, [&](size_t b, size_t e) {
      for (size_t idx = b; idx < e; idx++) {
        size_t blk = idx / half, j = idx % half;
        Fp* lo = a + blk * m + j;
        Fp* hi = lo + half;
        Fp t = *hi * tw[j * stride];
        *hi = *lo - t;
        *lo = *lo + t;
      }
    }, 1 << 14);
  }
}

// eval_polynomial: Horner
static inline Fp eval_polynomial(const Fp* c, size_t n, const Fp& x) {
  // chunked Horner so it can run on several threads; same value as the serial fold
  int t = oracle_threads();
  if (t <= 1 || n < (1u << 14)) {
    Fp acc = Fp::zero();
    for (size_t i = n; i-- > 0;) acc = acc * x + c[i];
    return acc;
  }
  size_t chunks = t, per = (n + chunks - 1) / chunks;
  std::vector<Fp> part(chunks, Fp::zero());
  parallel_for(chunks, [&](size_t b, size_t e) {
    for (size_t cidx = b; cidx < e; cidx++) {
      size_t lo = cidx * per, hi = std::min(n, lo + per);
      Fp acc = Fp::zero();
      for (size_t i = hi; i-- > lo;) acc = acc * x + c[i];
      part[cidx] = acc;
    }
  }, 1);
  u64 e[4] = {per, 0, 0, 0};
  Fp xp = x.pow(e), acc = Fp::zero();
  for (size_t cidx = chunks; cidx-- > 0;) acc = acc * xp + part[cidx];
  return acc;
}
static inline Fp eval_polynomial(const Poly& p, const Fp& x) {
  return eval_polynomial(p.data(), p.size(), x);
}

// kate_division: quotient of a(X) / (X - b), remainder dropped; result has len - 1 entries
static inline Poly kate_division(const Poly& a, const Fp& b) {
  Poly q(a.size() - 1, Fp::zero());
  Fp tmp = Fp::zero();
  for (size_t i = a.size() - 1; i-- > 0;) {
    Fp lead = a[i + 1] + tmp;  // coefficient of X^i in the quotient
    q[i] = lead;
    tmp = lead * b;
  }
  return q;
}

// lagrange_interpolate: coefficients of the unique poly of degree < n through (points, evals)
static inline Poly lagrange_interpolate(const std::vector<Fp>& points, const std::vector<Fp>& evals) {
  size_t n = points.size();
  Poly res(n, Fp::zero());
  for (size_t j = 0; j < n; j++) {
    Poly num(1, Fp::one());
    Fp denom = Fp::one();
    for (size_t m = 0; m < n; m++) {
      if (m == j) continue;
      Poly next(num.size() + 1, Fp::zero());
      for (size_t i = 0; i < num.size(); i++) {
        next[i + 1] += num[i];
        next[i] -= num[i] * points[m];
      }
      num = next;
      denom = denom * (points[j] - points[m]);
    }
    Fp scale = evals[j] * denom.invert();
    for (size_t i = 0; i < num.size(); i++) res[i] += num[i] * scale;
  }
  return res;
}

static inline Fp inner_product(const Fp* a, const Fp* b, size_t n) {
  int t = oracle_threads();
  std::vector<Fp> part(t, Fp::zero());
  size_t per = (n + t - 1) / t;
  parallel_for(t, [&](size_t bb, size_t ee) {
    for (size_t c = bb; c < ee; c++) {
      Fp acc = Fp::zero();
      for (size_t i = c * per; i < std::min(n, (c + 1) * per); i++) acc += a[i] * b[i];
      part[c] = acc;
    }
  }, 1);
  Fp acc = Fp::zero();
  for (auto& p : part) acc += p;
  return acc;
}

struct Domain {  // EvaluationDomain
  int k, extended_k;
  size_t n, extended_n;
  int quotient_poly_degree;
  Fp omega, omega_inv, extended_omega, extended_omega_inv;
  Fp g_coset, g_coset_inv;  // zeta, zeta^2
  Fp ifft_divisor, extended_ifft_divisor, barycentric_weight;
  std::vector<Fp> t_evaluations_inv;  // 1 / (X^n - 1) on the extended coset (period 2^(ek-k))

  Domain() {}
  Domain(int cs_degree, int k_) {
    k = k_;
    n = (size_t)1 << k;
    quotient_poly_degree = cs_degree - 1;
    extended_k = k;
    while (((size_t)1 << extended_k) < n * (size_t)quotient_poly_degree) extended_k++;
    extended_n = (size_t)1 << extended_k;
    const auto& C = Fp::C();
    extended_omega = C.root_of_unity;
    for (int i = extended_k; i < 32; i++) extended_omega = extended_omega.square();
    omega = extended_omega;
    for (int i = k; i < extended_k; i++) omega = omega.square();
    omega_inv = omega.invert();
    extended_omega_inv = extended_omega.invert();
    g_coset = C.zeta;
    g_coset_inv = g_coset.square();
    ifft_divisor = Fp::from_u64(n).invert();
    extended_ifft_divisor = Fp::from_u64(extended_n).invert();
    barycentric_weight = Fp::from_u64(n).invert();
    Fp orig = g_coset.pow_u64(n), step = extended_omega.pow_u64(n), cur = orig;
    do {
      t_evaluations_inv.push_back((cur - Fp::one()).invert());
      cur = cur * step;
    } while (cur != orig);
  }
  Fp rotate_omega(const Fp& v, int rot) const {
    if (rot >= 0) return v * omega.pow_u64((u64)rot);
    return v * omega_inv.pow_u64((u64)(-rot));
  }
  Poly lagrange_to_coeff(Poly a) const {
    fft(a.data(), k, omega_inv);
    for (auto& x : a) x = x * ifft_divisor;
    return a;
  }
  Poly coeff_to_lagrange(Poly a) const {
    fft(a.data(), k, omega);
    return a;
  }
  // evaluations on the coset zeta * <extended_omega>
  Poly coeff_to_extended(const Poly& a) const {
    Poly e(extended_n, Fp::zero());
    Fp pw[3] = {Fp::one(), g_coset, g_coset_inv};
    for (size_t i = 0; i < a.size(); i++) e[i] = a[i] * pw[i % 3];
    fft(e.data(), extended_k, extended_omega);
    return e;
  }
  // inverse; truncated to n * quotient_poly_degree coefficients
  Poly extended_to_coeff(Poly e) const {
    fft(e.data(), extended_k, extended_omega_inv);
    Fp pw[3] = {Fp::one(), g_coset_inv, g_coset};
    for (size_t i = 0; i < e.size(); i++) e[i] = e[i] * extended_ifft_divisor * pw[i % 3];
    e.resize(n * (size_t)quotient_poly_degree);
    return e;
  }
  // l_i(x) for i in [from, to], i.e. omega^i (x^n - 1) / (n (x - omega^i))
  std::vector<Fp> l_i_range(const Fp& x, const Fp& xn, int from, int to) const {
    std::vector<Fp> r;
    for (int rot = from; rot <= to; rot++) r.push_back(x - rotate_omega(Fp::one(), rot));
    batch_invert(r.data(), r.size());
    Fp common = (xn - Fp::one()) * barycentric_weight;
    int idx = 0;
    for (int rot = from; rot <= to; rot++, idx++) r[idx] = rotate_omega(r[idx] * common, rot);
    return r;
  }
};

}  // namespace zko
