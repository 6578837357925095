// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Vesta (y^2 = x^3 + 5 over Fq; scalar field Fp) — the commitment curve `EqAffine` that the
// reference's prove/verify sequence uses (blake2f-circuit/benches/blake2f.rs:3,85
// `Params::<EqAffine>::new(k)`).  Arithmetic lives in pasta_curves 0.5.1 (Cargo.lock:1334-1347,
// not vendored); restated here from the published short-Weierstrass a=0 Jacobian formulas.
// Affine outputs are canonical, so any correct group law reproduces the same bytes.
// Parity unpinned (no golden commitment in the reference).
#pragma once
#include "field.hpp"

namespace zko {

struct Affine {
  Fq x, y;  // identity encoded as (0,0), as pasta_curves does
  bool is_identity() const { return x.is_zero() && y.is_zero(); }
  static Affine identity() { return Affine{Fq::zero(), Fq::zero()}; }
  bool operator==(const Affine& o) const { return x == o.x && y == o.y; }
  Affine neg() const { return is_identity() ? *this : Affine{x, -y}; }
  // group::GroupEncoding::to_bytes: x.to_repr() with the sign of y in bit 7 of byte 31
  void to_bytes(uint8_t out[32]) const {
    if (is_identity()) {
      memset(out, 0, 32);
      return;
    }
    x.to_repr(out);
    if (y.is_odd()) out[31] |= 0x80;
  }
  static bool from_bytes(const uint8_t in[32], Affine& out) {
    uint8_t tmp[32];
    memcpy(tmp, in, 32);
    bool ysign = tmp[31] >> 7;
    tmp[31] &= 0x7f;
    bool allz = true;
    for (int i = 0; i < 32; i++) allz &= (tmp[i] == 0);
    if (allz && !ysign) {
      out = identity();
      return true;
    }
    Fq x;
    if (!Fq::from_repr(tmp, x)) return false;
    Fq rhs = x.square() * x + Fq::from_u64(5), y;
    if (!rhs.sqrt(y)) return false;
    if (y.is_odd() != ysign) y = -y;
    out = Affine{x, y};
    return true;
  }
  bool on_curve() const {
    return is_identity() || y.square() == x.square() * x + Fq::from_u64(5);
  }
};

struct Jac {
  Fq x, y, z;  // (X/Z^2, Y/Z^3); identity has z == 0
  static Jac identity() { return Jac{Fq::zero(), Fq::one(), Fq::zero()}; }
  static Jac from_affine(const Affine& a) {
    return a.is_identity() ? identity() : Jac{a.x, a.y, Fq::one()};
  }
  bool is_identity() const { return z.is_zero(); }

  Jac dbl() const {  // dbl-2009-l (a = 0)
    if (is_identity()) return *this;
    Fq a = x.square(), b = y.square(), c = b.square();
    Fq d = ((x + b).square() - a - c).dbl();
    Fq e = a.dbl() + a;
    Fq f = e.square();
    Fq x3 = f - d.dbl();
    Fq y3 = e * (d - x3) - c.dbl().dbl().dbl();
    Fq z3 = (y * z).dbl();
    return Jac{x3, y3, z3};
  }
  Jac add(const Jac& o) const {  // add-2007-bl
    if (is_identity()) return o;
    if (o.is_identity()) return *this;
    Fq z1z1 = z.square(), z2z2 = o.z.square();
    Fq u1 = x * z2z2, u2 = o.x * z1z1;
    Fq s1 = y * o.z * z2z2, s2 = o.y * z * z1z1;
    if (u1 == u2) {
      if (s1 == s2) return dbl();
      return identity();
    }
    Fq h = u2 - u1;
    Fq i = h.dbl().square();
    Fq j = h * i;
    Fq r = (s2 - s1).dbl();
    Fq v = u1 * i;
    Fq x3 = r.square() - j - v.dbl();
    Fq y3 = r * (v - x3) - (s1 * j).dbl();
    Fq z3 = ((z + o.z).square() - z1z1 - z2z2) * h;
    return Jac{x3, y3, z3};
  }
  Jac add_affine(const Affine& o) const {  // madd-2007-bl
    if (o.is_identity()) return *this;
    if (is_identity()) return from_affine(o);
    Fq z1z1 = z.square();
    Fq u2 = o.x * z1z1;
    Fq s2 = o.y * z * z1z1;
    if (x == u2) {
      if (y == s2) return dbl();
      return identity();
    }
    Fq h = u2 - x;
    Fq hh = h.square();
    Fq i = hh.dbl().dbl();
    Fq j = h * i;
    Fq r = (s2 - y).dbl();
    Fq v = x * i;
    Fq x3 = r.square() - j - v.dbl();
    Fq y3 = r * (v - x3) - (y * j).dbl();
    Fq z3 = (z + h).square() - z1z1 - hh;
    return Jac{x3, y3, z3};
  }
  Jac neg() const { return Jac{x, -y, z}; }
  Affine to_affine() const {
    if (is_identity()) return Affine::identity();
    Fq zi = z.invert(), zi2 = zi.square();
    return Affine{x * zi2, y * zi2 * zi};
  }
  // scalar multiplication by a canonical 256-bit integer (double-and-add, MSB first)
  Jac mul_raw(const u64 e[4]) const {
    Jac acc = identity();
    for (int i = 255; i >= 0; i--) {
      acc = acc.dbl();
      if ((e[i / 64] >> (i % 64)) & 1) acc = acc.add(*this);
    }
    return acc;
  }
  Jac mul(const Fp& s) const {
    u64 e[4];
    s.to_raw(e);
    return mul_raw(e);
  }
};

// group::Curve::batch_normalize
static inline void batch_normalize(const Jac* in, Affine* out, size_t n) {
  std::vector<Fq> zs(n);
  for (size_t i = 0; i < n; i++) zs[i] = in[i].z;
  batch_invert(zs.data(), n);
  for (size_t i = 0; i < n; i++) {
    if (in[i].is_identity()) {
      out[i] = Affine::identity();
    } else {
      Fq zi2 = zs[i].square();
      out[i] = Affine{in[i].x * zi2, in[i].y * zi2 * zs[i]};
    }
  }
}

static inline Affine vesta_generator() {  // (-1, 2)
  return Affine{-Fq::one(), Fq::from_u64(2)};
}

// Multi-scalar multiplication sum_i [s_i] P_i — the function `best_multiexp` computes
// (halo2_proofs 0.3.0 arithmetic.rs, reached from Params::commit / commit_lagrange).
// Bucket method, window chosen like the serial halo2 path; the result is a group element so
// the choice of window does not affect any output byte.
Jac msm(const Fp* scalars, const Affine* bases, size_t n);

}  // namespace zko
