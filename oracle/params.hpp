// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// `best_multiexp` and `poly::commitment::Params<EqAffine>` of halo2_proofs 0.3.0 (un-vendored,
// Cargo.lock:842-857), reached from the reference at blake2f-circuit/benches/blake2f.rs:83-97
// (`Params::new` / `Params::read` / `Params::write`) and through every commitment inside
// `create_proof` (:125).
//
// Params::new hashes to the curve with pasta's simplified-SWU map, whose isogeny constants
// are not reproducible offline (SURVEY.md hard part H3).  `generate_substitute` therefore
// builds a *substitute* URS: g[i] = [s_i] G for scalars s_i drawn from the seeded
// XorShiftRng, g_lagrange = [ifft(s)_j] G (so that it is exactly `g_to_lagrange(g)`),
// w = [s_w] G, u = [s_u] G.  Discrete logs are known: benchmark / parity use only.  The
// on-disk format is halo2's, so a genuine params file can be dropped in (`read`).
// Parity unpinned.
#pragma once
#include <cstdio>
#include <stdexcept>
#include "curve.hpp"
#include "poly.hpp"
#include "xorshift.hpp"

namespace zko {

// sum_i [s_i] P_i by the bucket method; windows are independent and run on separate threads.
inline Jac msm(const Fp* scalars, const Affine* bases, size_t n) {
  if (n == 0) return Jac::identity();
  int c = n < 32 ? 3 : (n < (1u << 12) ? 8 : (n < (1u << 17) ? 12 : 14));
  int windows = (255 + c - 1) / c;
  std::vector<std::array<u64, 4>> raw(n);
  parallel_for(n, [&](size_t b, size_t e) {
    for (size_t i = b; i < e; i++) scalars[i].to_raw(raw[i].data());
  });
  std::vector<Jac> wsum(windows, Jac::identity());
  parallel_for(windows, [&](size_t wb, size_t we) {
    for (size_t w = wb; w < we; w++) {
      std::vector<Jac> buckets(((size_t)1 << c) - 1, Jac::identity());
      int bit = (int)w * c;
      for (size_t i = 0; i < n; i++) {
        int limb = bit / 64, off = bit % 64;
        u64 v = raw[i][limb] >> off;
        if (off + c > 64 && limb < 3) v |= raw[i][limb + 1] << (64 - off);
        v &= ((u64)1 << c) - 1;
        if (v) buckets[v - 1] = buckets[v - 1].add_affine(bases[i]);
      }
      Jac running = Jac::identity(), acc = Jac::identity();
      for (size_t b = buckets.size(); b-- > 0;) {
        running = running.add(buckets[b]);
        acc = acc.add(running);
      }
      wsum[w] = acc;
    }
  }, 1);
  Jac total = Jac::identity();
  for (int w = windows - 1; w >= 0; w--) {
    for (int i = 0; i < c; i++) total = total.dbl();
    total = total.add(wsum[w]);
  }
  return total;
}

struct Params {
  int k = 0;
  size_t n = 0;
  std::vector<Affine> g, g_lagrange;
  Affine w, u;

  // commit_lagrange(poly, r) = MSM(poly || r, g_lagrange || w); commit uses g.
  Jac commit_lagrange(const Poly& evals, const Fp& blind) const {
    Jac acc = msm(evals.data(), g_lagrange.data(), n);
    return acc.add(Jac::from_affine(w).mul(blind));
  }
  Jac commit(const Poly& coeffs, const Fp& blind) const {
    Jac acc = msm(coeffs.data(), g.data(), n);
    return acc.add(Jac::from_affine(w).mul(blind));
  }

  // [s] G for many scalars through an 8-bit fixed-base table
  static void fixed_base_mul(const std::vector<Fp>& scalars, std::vector<Affine>& out) {
    Affine G = vesta_generator();
    std::vector<Jac> tab_j(32 * 255);
    Jac base = Jac::from_affine(G);
    for (int w = 0; w < 32; w++) {
      Jac cur = base;
      for (int d = 1; d <= 255; d++) {
        tab_j[w * 255 + d - 1] = cur;
        cur = cur.add(base);
      }
      base = cur;  // 256 * previous base
    }
    std::vector<Affine> tab(tab_j.size());
    batch_normalize(tab_j.data(), tab.data(), tab.size());
    std::vector<Jac> res(scalars.size());
    parallel_for(scalars.size(), [&](size_t b, size_t e) {
      for (size_t i = b; i < e; i++) {
        uint8_t bytes[32];
        scalars[i].to_repr(bytes);
        Jac acc = Jac::identity();
        for (int w = 0; w < 32; w++)
          if (bytes[w]) acc = acc.add_affine(tab[w * 255 + bytes[w] - 1]);
        res[i] = acc;
      }
    }, 256);
    out.resize(scalars.size());
    size_t n = scalars.size(), chunk = 1 << 12;
    parallel_for((n + chunk - 1) / chunk, [&](size_t b, size_t e) {
      for (size_t c = b; c < e; c++) {
        size_t lo = c * chunk, hi = std::min(n, lo + chunk);
        batch_normalize(res.data() + lo, out.data() + lo, hi - lo);
      }
    }, 1);
  }

  static Params generate_substitute(int k, const uint8_t seed[16]) {
    Params p;
    p.k = k;
    p.n = (size_t)1 << k;
    XorShiftRng rng(seed);
    std::vector<Fp> s(p.n);
    for (auto& x : s) x = rng.random_field<Fp>();
    Fp sw = rng.random_field<Fp>(), su = rng.random_field<Fp>();
    fixed_base_mul(s, p.g);
    Domain d(2, k);
    Poly sl = d.lagrange_to_coeff(s);
    fixed_base_mul(sl, p.g_lagrange);
    std::vector<Affine> wu;
    fixed_base_mul({sw, su}, wu);
    p.w = wu[0];
    p.u = wu[1];
    return p;
  }

  // halo2 Params::write: k (u32 LE), g, g_lagrange (n compressed points each), w, u
  void write(std::vector<uint8_t>& out) const {
    out.resize(4 + (2 * n + 2) * 32);
    uint32_t kk = (uint32_t)k;
    memcpy(out.data(), &kk, 4);
    size_t off = 4;
    for (auto& pt : g) { pt.to_bytes(out.data() + off); off += 32; }
    for (auto& pt : g_lagrange) { pt.to_bytes(out.data() + off); off += 32; }
    w.to_bytes(out.data() + off); off += 32;
    u.to_bytes(out.data() + off);
  }
  static Params read(const uint8_t* data, size_t len) {
    if (len < 4) throw std::runtime_error("params: truncated");
    uint32_t kk;
    memcpy(&kk, data, 4);
    if (kk > 28) throw std::runtime_error("params: k too large");
    Params p;
    p.k = (int)kk;
    p.n = (size_t)1 << kk;
    if (len != 4 + (2 * p.n + 2) * 32) throw std::runtime_error("params: bad length");
    p.g.resize(p.n);
    p.g_lagrange.resize(p.n);
    std::vector<int> bad(1, 0);
    parallel_for(2 * p.n, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; i++) {
        Affine& dst = i < p.n ? p.g[i] : p.g_lagrange[i - p.n];
        if (!Affine::from_bytes(data + 4 + 32 * i, dst)) bad[0] = 1;
      }
    }, 256);
    if (!Affine::from_bytes(data + 4 + 64 * p.n, p.w) ||
        !Affine::from_bytes(data + 4 + 64 * p.n + 32, p.u) || bad[0])
      throw std::runtime_error("params: invalid point");
    return p;
  }
};

}  // namespace zko
