// Vesta group law (y^2 = x^3 + 5 over Fq) for host and device.  Product code.
//
// Replaces pasta_curves 0.5.1 `vesta::{Affine, Point}` as used by `Params<EqAffine>`
// (blake2f-circuit/benches/blake2f.rs:3,85).  Accumulators use extended Jacobian "XYZZ"
// coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; EFD shortw/xyzz, a = 0): mixed addition is
// 8M + 2S, the cheapest form for bucket accumulation.  Only affine results leave the library,
// so the coordinate system is not observable.
#pragma once
#include "field.cuh"

namespace zkodst {

struct Affine {  // 64 bytes; identity = (0, 0) as in pasta_curves
  Fq x, y;
  ZK_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
  ZK_HD static Affine identity() { return Affine{Fq::zero(), Fq::zero()}; }
  ZK_HD Affine neg() const { return is_identity() ? *this : Affine{x, y.neg()}; }
};

struct XYZZ {  // 128 bytes; identity has zz == 0
  Fq x, y, zz, zzz;
  ZK_HD static XYZZ identity() { return XYZZ{Fq::zero(), Fq::zero(), Fq::zero(), Fq::zero()}; }
  ZK_HD bool is_identity() const { return zz.is_zero(); }
  ZK_HD static XYZZ from_affine(const Affine& p) {
    if (p.is_identity()) return identity();
    return XYZZ{p.x, p.y, Fq::one(), Fq::one()};
  }
  ZK_HD XYZZ neg() const { return XYZZ{x, y.neg(), zz, zzz}; }

  // 2 * (affine point)                                             [mdbl-2008-s-1]
  ZK_HD static XYZZ dbl_affine(const Affine& p) {
    if (p.is_identity()) return identity();
    Fq u = p.y.dbl();
    Fq v = u.sqr();
    Fq w = u * v;
    Fq s = p.x * v;
    Fq xx = p.x.sqr();
    Fq m = xx.dbl() + xx;
    Fq x3 = m.sqr() - s.dbl();
    Fq y3 = m * (s - x3) - w * p.y;
    return XYZZ{x3, y3, v, w};
  }
  //                                                                  [dbl-2008-s-1]
  ZK_HD XYZZ dbl() const {
    if (is_identity()) return *this;
    Fq u = y.dbl();
    Fq v = u.sqr();
    Fq w = u * v;
    Fq s = x * v;
    Fq xx = x.sqr();
    Fq m = xx.dbl() + xx;
    Fq x3 = m.sqr() - s.dbl();
    Fq y3 = m * (s - x3) - w * y;
    return XYZZ{x3, y3, v * zz, w * zzz};
  }
  // same x: doubling or P + (-P); kept out of line on the device so the accumulation loops stay small
#if defined(__CUDACC__)
  __host__ __device__ __noinline__
#endif
  static XYZZ add_affine_same_x(const Affine& p, const Fq& r) {
    if (r.is_zero()) return dbl_affine(p);
    return identity();
  }
  // this + affine                                                    [madd-2008-s]
  ZK_HD XYZZ add_affine(const Affine& p) const {
    if (p.is_identity()) return *this;
    if (is_identity()) return from_affine(p);
    Fq u2 = p.x * zz;
    Fq s2 = p.y * zzz;
    Fq pp_ = u2 - x;
    Fq r = s2 - y;
    if (pp_.is_zero()) return add_affine_same_x(p, r);
    Fq pp = pp_.sqr();
    Fq ppp = pp_ * pp;
    Fq q = x * pp;
    Fq x3 = r.sqr() - ppp - q.dbl();
    Fq y3 = r * (q - x3) - y * ppp;
    return XYZZ{x3, y3, zz * pp, zzz * ppp};
  }
#if defined(__CUDACC__)
  __host__ __device__ __noinline__
#endif
  static XYZZ add_same_x(const XYZZ& a, const Fq& r) {
    if (r.is_zero()) return a.dbl();
    return identity();
  }
  // this + other                                                     [add-2008-s]
  ZK_HD XYZZ add(const XYZZ& o) const {
    if (o.is_identity()) return *this;
    if (is_identity()) return o;
    Fq u1 = x * o.zz;
    Fq u2 = o.x * zz;
    Fq s1 = y * o.zzz;
    Fq s2 = o.y * zzz;
    Fq pp_ = u2 - u1;
    Fq r = s2 - s1;
    if (pp_.is_zero()) return add_same_x(*this, r);
    Fq pp = pp_.sqr();
    Fq ppp = pp_ * pp;
    Fq q = u1 * pp;
    Fq x3 = r.sqr() - ppp - q.dbl();
    Fq y3 = r * (q - x3) - s1 * ppp;
    return XYZZ{x3, y3, zz * o.zz * pp, zzz * o.zzz * ppp};
  }
  // host-side normalisation (one inversion)
  ZK_HD Affine to_affine() const {
    if (is_identity()) return Affine::identity();
    Fq zi = zzz.inv();          // 1 / ZZZ
    Fq zz_inv = zi.sqr() * zz * zz;  // ZZ^2 / ZZZ^2 = ZZ^2 / ZZ^3 = 1 / ZZ
    return Affine{x * zz_inv, y * zi};
  }
};

// (t - 1) / 2 with q - 1 = 2^32 * t: the fixed exponent of Tonelli-Shanks in Fq
ZK_HD void fq_sqrt_exponent(uint64_t e[4]) {
  const uint64_t qm1[4] = {FqParams::MOD[0] - 1, FqParams::MOD[1], FqParams::MOD[2], FqParams::MOD[3]};
  uint64_t t[4];
  for (int i = 0; i < 4; i++) t[i] = (qm1[i] >> 32) | (i < 3 ? qm1[i + 1] << 32 : 0);
  t[0] -= 1;  // t is odd, no borrow
  for (int i = 0; i < 4; i++) e[i] = (t[i] >> 1) | (i < 3 ? t[i + 1] << 63 : 0);
}
// Tonelli-Shanks in Fq (2-adicity 32), exponent (t - 1) / 2 passed in
ZK_HD bool fq_sqrt(const Fq& a, const uint64_t tm1o2[4], Fq& out) {
  if (a.is_zero()) {
    out = a;
    return true;
  }
  Fq w = a.pow256(tm1o2);
  Fq v = a * w, b = v * w, z = Fq::root_of_unity(), x = v;
  int vexp = 32;
  while (b != Fq::one()) {
    int k = 0;
    Fq b2 = b;
    while (b2 != Fq::one()) {
      b2 = b2.sqr();
      k++;
      if (k == vexp) return false;
    }
    Fq ww = z;
    for (int i = 0; i < vexp - k - 1; i++) ww = ww.sqr();
    z = ww.sqr();
    b = b * z;
    x = x * ww;
    vexp = k;
  }
  out = x;
  return x.sqr() == a;
}
// group::GroupEncoding::from_bytes for vesta::Affine: x little-endian canonical, bit 255 = y is odd,
// 32 zero bytes = identity.  false: not a canonical encoding of a curve point.
ZK_HD bool decompress_point(const uint8_t* p, const uint64_t tm1o2[4], Affine& out) {
  uint64_t c[4];
  for (int l = 0; l < 4; l++) {
    uint64_t w = 0;
    for (int b = 0; b < 8; b++) w |= (uint64_t)p[8 * l + b] << (8 * b);
    c[l] = w;
  }
  const bool ysign = c[3] >> 63;
  c[3] &= 0x7fffffffffffffffULL;
  if (!(c[0] | c[1] | c[2] | c[3])) {
    out = Affine::identity();
    return !ysign;
  }
  if (Fq::geq_mod(c)) return false;
  Fq x = Fq::from_canonical(c);
  Fq rhs = x.sqr() * x + Fq::from_u64(5), y;
  if (!fq_sqrt(rhs, tm1o2, y)) return false;
  if (y.is_odd() != ysign) y = y.neg();
  out = Affine{x, y};
  return true;
}

}  // namespace zkodst
