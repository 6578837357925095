// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the Pasta field arithmetic that the reference reaches through
// `halo2_proofs 0.3.0` -> `pasta_curves 0.5.1` (un-vendored crates pinned by
// /root/reference/Cargo.lock:842-857 and :1334-1347; call sites
// blake2f-circuit/src/blake2f/table16.rs:23 `pasta::pallas`, :93-98 `pallas::Base::from`).
// Parity unpinned: the reference holds no golden field/commitment/proof vector
// (SURVEY.md §8c); the constants below are re-derived at start-up from the two moduli and
// cross-checked against SURVEY.md Appendix A.1 in tests/test_oracle_field.py.
//
// Representation mirrors pasta_curves: 4 x u64 little-endian limbs, Montgomery form
// (value * 2^256 mod p).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace zko {

typedef uint64_t u64;
typedef uint32_t u32;
typedef unsigned __int128 u128;

struct FpParams {  // pallas::Base == vesta::Scalar  (circuit field, MSM scalars)
  static constexpr u64 MOD[4] = {0x992d30ed00000001ULL, 0x224698fc094cf91bULL, 0x0ULL,
                                 0x4000000000000000ULL};
};
struct FqParams {  // vesta::Base == pallas::Scalar  (coordinate field of the commitment curve)
  static constexpr u64 MOD[4] = {0x8c46eb2100000001ULL, 0x224698fc0994a8ddULL, 0x0ULL,
                                 0x4000000000000000ULL};
};

// ---- raw 256-bit helpers -------------------------------------------------------------
static inline bool raw_geq(const u64 a[4], const u64 b[4]) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] != b[i]) return a[i] > b[i];
  }
  return true;
}
static inline u64 raw_add(u64 r[4], const u64 a[4], const u64 b[4]) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a[i] + b[i];
    r[i] = (u64)c;
    c >>= 64;
  }
  return (u64)c;
}
static inline u64 raw_sub(u64 r[4], const u64 a[4], const u64 b[4]) {
  u64 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a[i] - b[i] - borrow;
    r[i] = (u64)d;
    borrow = (u64)(d >> 64) & 1;
  }
  return borrow;
}

template <class P>
struct Fe {
  u64 l[4];

  struct Consts {
    u64 inv;        // -p^{-1} mod 2^64
    Fe R, R2, R3;   // 2^256, 2^512, 2^768 mod p  (as raw limbs == Montgomery 1, ...)
    Fe generator;   // 5
    Fe root_of_unity;      // 5^((p-1)/2^32), order 2^32
    Fe root_of_unity_inv;
    Fe delta;       // 5^(2^32)
    Fe zeta;        // primitive cube root of unity ((5^((p-1)/3))^2, pasta's ZETA)
    Fe two_inv;
    u64 t_minus1_over2[4];  // for Tonelli-Shanks
  };
  static const Consts& C() {
    static const Consts c = make_consts();
    return c;
  }

  static Fe zero() { return Fe{{0, 0, 0, 0}}; }
  static Fe one() { return C().R; }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  bool operator==(const Fe& o) const { return memcmp(l, o.l, 32) == 0; }
  bool operator!=(const Fe& o) const { return !(*this == o); }

  // Montgomery product (CIOS over 64-bit limbs).
  static Fe mont_mul(const Fe& a, const Fe& b, u64 inv) {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      u128 c = 0;
      for (int j = 0; j < 4; j++) {
        c += (u128)a.l[j] * b.l[i] + t[j];
        t[j] = (u64)c;
        c >>= 64;
      }
      c += t[4];
      t[4] = (u64)c;
      t[5] = (u64)(c >> 64);
      u64 m = t[0] * inv;
      c = (u128)m * P::MOD[0] + t[0];
      c >>= 64;
      for (int j = 1; j < 4; j++) {
        c += (u128)m * P::MOD[j] + t[j];
        t[j - 1] = (u64)c;
        c >>= 64;
      }
      c += t[4];
      t[3] = (u64)c;
      t[4] = t[5] + (u64)(c >> 64);
    }
    Fe r{{t[0], t[1], t[2], t[3]}};
    if (t[4] || raw_geq(r.l, P::MOD)) raw_sub(r.l, r.l, P::MOD);
    return r;
  }
  Fe operator*(const Fe& o) const { return mont_mul(*this, o, C().inv); }
  Fe square() const { return *this * *this; }
  Fe operator+(const Fe& o) const {
    Fe r;
    u64 c = raw_add(r.l, l, o.l);
    if (c || raw_geq(r.l, P::MOD)) raw_sub(r.l, r.l, P::MOD);
    return r;
  }
  Fe operator-(const Fe& o) const {
    Fe r;
    if (raw_sub(r.l, l, o.l)) raw_add(r.l, r.l, P::MOD);
    return r;
  }
  Fe operator-() const { return zero() - *this; }
  Fe& operator+=(const Fe& o) { return *this = *this + o; }
  Fe& operator-=(const Fe& o) { return *this = *this - o; }
  Fe& operator*=(const Fe& o) { return *this = *this * o; }
  Fe dbl() const { return *this + *this; }

  // integer -> field (Montgomery form)
  static Fe from_u64(u64 v) { return Fe{{v, 0, 0, 0}} * C().R2; }
  static Fe from_u128(u128 v) { return Fe{{(u64)v, (u64)(v >> 64), 0, 0}} * C().R2; }
  // canonical (non-Montgomery) limbs, must be < p
  static Fe from_raw(const u64 v[4]) { return Fe{{v[0], v[1], v[2], v[3]}} * C().R2; }
  void to_raw(u64 out[4]) const {
    Fe r = mont_mul(*this, Fe{{1, 0, 0, 0}}, C().inv);
    memcpy(out, r.l, 32);
  }
  // ff::PrimeField::to_repr: 32-byte little-endian canonical integer
  void to_repr(uint8_t out[32]) const {
    u64 r[4];
    to_raw(r);
    memcpy(out, r, 32);
  }
  // ff::PrimeField::from_repr; returns false if not canonical
  static bool from_repr(const uint8_t in[32], Fe& out) {
    u64 r[4];
    memcpy(r, in, 32);
    if (raw_geq(r, P::MOD)) return false;
    out = from_raw(r);
    return true;
  }
  // ff::FromUniformBytes<64>: 512-bit little-endian integer reduced mod p
  // (pasta_curves `from_u512`: lo*R2 + hi*R3 in Montgomery arithmetic)
  static Fe from_u512(const u64 v[8]) {
    Fe lo{{v[0], v[1], v[2], v[3]}}, hi{{v[4], v[5], v[6], v[7]}};
    return lo * C().R2 + hi * C().R3;
  }
  static Fe from_uniform_bytes(const uint8_t in[64]) {
    u64 v[8];
    memcpy(v, in, 64);
    return from_u512(v);
  }
  bool is_odd() const {
    u64 r[4];
    to_raw(r);
    return r[0] & 1;
  }
  // Ord on canonical integers (used by the lookup-argument sort)
  static int cmp(const Fe& a, const Fe& b) {
    u64 x[4], y[4];
    a.to_raw(x);
    b.to_raw(y);
    for (int i = 3; i >= 0; i--)
      if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return 0;
  }

  Fe pow(const u64 e[4]) const {
    Fe acc = one();
    for (int i = 255; i >= 0; i--) {
      acc = acc.square();
      if ((e[i / 64] >> (i % 64)) & 1) acc = acc * *this;
    }
    return acc;
  }
  Fe pow_u64(u64 e) const {
    u64 ee[4] = {e, 0, 0, 0};
    return pow(ee);
  }
  // Field::invert (Fermat); zero maps to zero (callers that need the Option check first)
  Fe invert() const {
    u64 e[4];
    u64 two[4] = {2, 0, 0, 0};
    raw_sub(e, P::MOD, two);
    return pow(e);
  }
  // Tonelli-Shanks, S = 32.  Returns false if not a square.
  bool sqrt(Fe& out) const {
    if (is_zero()) {
      out = zero();
      return true;
    }
    const Consts& c = C();
    Fe w = pow(c.t_minus1_over2);  // a^((t-1)/2)
    Fe v_ = *this * w;             // a^((t+1)/2)
    Fe b = v_ * w;                 // a^t
    Fe z = c.root_of_unity;
    Fe x = v_;
    int vexp = 32;
    while (b != one()) {
      int kk = 0;
      Fe b2 = b;
      while (b2 != one()) {
        b2 = b2.square();
        kk++;
        if (kk == vexp) return false;
      }
      Fe ww = z;
      for (int i = 0; i < vexp - kk - 1; i++) ww = ww.square();
      z = ww.square();
      b = b * z;
      x = x * ww;
      vexp = kk;
    }
    out = x;
    return (x.square() == *this);
  }

  std::string hex() const {  // "0x" + 64 lowercase hex digits, big-endian (ff Debug format)
    u64 r[4];
    to_raw(r);
    char buf[80];
    snprintf(buf, sizeof buf, "0x%016llx%016llx%016llx%016llx", (unsigned long long)r[3],
             (unsigned long long)r[2], (unsigned long long)r[1], (unsigned long long)r[0]);
    return buf;
  }

 private:
  static Consts make_consts() {
    Consts c;
    // inv = -p^{-1} mod 2^64 by Newton iteration
    u64 p0 = P::MOD[0], x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - p0 * x;
    c.inv = (u64)0 - x;
    // R = 2^256 mod p by repeated doubling of 1
    auto dbl_mod = [](Fe a) {
      Fe r;
      u64 cy = raw_add(r.l, a.l, a.l);
      if (cy || raw_geq(r.l, P::MOD)) raw_sub(r.l, r.l, P::MOD);
      return r;
    };
    Fe v{{1, 0, 0, 0}};
    for (int i = 0; i < 256; i++) v = dbl_mod(v);
    c.R = v;
    for (int i = 0; i < 256; i++) v = dbl_mod(v);
    c.R2 = v;
    for (int i = 0; i < 256; i++) v = dbl_mod(v);
    c.R3 = v;
    auto mul = [&](const Fe& a, const Fe& b) { return mont_mul(a, b, c.inv); };
    auto powr = [&](Fe base, const u64 e[4]) {
      Fe acc = c.R;
      for (int i = 255; i >= 0; i--) {
        acc = mul(acc, acc);
        if ((e[i / 64] >> (i % 64)) & 1) acc = mul(acc, base);
      }
      return acc;
    };
    c.generator = mul(Fe{{5, 0, 0, 0}}, c.R2);
    u64 pm1[4], one_[4] = {1, 0, 0, 0};
    raw_sub(pm1, P::MOD, one_);
    // t = (p-1) >> 32
    u64 t[4];
    for (int i = 0; i < 4; i++) t[i] = (pm1[i] >> 32) | (i < 3 ? (pm1[i + 1] << 32) : 0);
    c.root_of_unity = powr(c.generator, t);
    u64 pm2[4], two_[4] = {2, 0, 0, 0};
    raw_sub(pm2, P::MOD, two_);
    c.root_of_unity_inv = powr(c.root_of_unity, pm2);
    u64 e232[4] = {1ULL << 32, 0, 0, 0};
    c.delta = powr(c.generator, e232);
    // (p-1)/3
    u64 third[4];
    {
      u128 rem = 0;
      for (int i = 3; i >= 0; i--) {
        u128 cur = (rem << 64) | pm1[i];
        third[i] = (u64)(cur / 3);
        rem = cur % 3;
      }
    }
    Fe z3 = powr(c.generator, third);
    c.zeta = mul(z3, z3);
    c.two_inv = powr(mul(Fe{{2, 0, 0, 0}}, c.R2), pm2);
    // (t-1)/2
    u64 tm1[4];
    raw_sub(tm1, t, one_);
    for (int i = 0; i < 4; i++) c.t_minus1_over2[i] = (tm1[i] >> 1) | (i < 3 ? (tm1[i + 1] << 63) : 0);
    return c;
  }
};

typedef Fe<FpParams> Fp;
typedef Fe<FqParams> Fq;

// Montgomery batch inversion (zeros stay zero), as ff::BatchInvert.
template <class F>
static inline void batch_invert(F* v, size_t n) {
  std::vector<F> prefix(n);
  F acc = F::one();
  for (size_t i = 0; i < n; i++) {
    prefix[i] = acc;
    if (!v[i].is_zero()) acc = acc * v[i];
  }
  acc = acc.invert();
  for (size_t i = n; i-- > 0;) {
    if (v[i].is_zero()) continue;
    F t = acc * prefix[i];
    acc = acc * v[i];
    v[i] = t;
  }
}

}  // namespace zko
