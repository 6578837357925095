// Pasta field arithmetic for host and device (product code).
//
// Replaces pasta_curves 0.5.1 `Fp` / `Fq` (un-vendored dependency of the reference,
// Cargo.lock:1334-1347; used through `pallas::Base` at blake2f-circuit/src/blake2f/table16.rs:23):
// 4 x u64 little-endian limbs in Montgomery form — the same in-memory image, so buffers can be
// handed across the C ABI without conversion.
#pragma once
#include <cstdint>
#include <cstring>

#include "field_consts.h"
#if defined(__CUDACC__)
#include "mont_ptx.cuh"
#endif

#if defined(__CUDACC__)
#define ZK_HD __host__ __device__ __forceinline__
#else
#define ZK_HD inline
#endif

namespace zkodst {

typedef unsigned __int128 u128;

template <class P>
struct alignas(16) Fe {
  uint64_t l[4];

  ZK_HD static constexpr uint64_t M(int i) {
    return i == 0 ? P::MOD[0] : i == 1 ? P::MOD[1] : i == 2 ? P::MOD[2] : P::MOD[3];
  }
  ZK_HD static Fe zero() { return Fe{{0, 0, 0, 0}}; }
  ZK_HD static Fe one() { return Fe{{P::R[0], P::R[1], P::R[2], P::R[3]}}; }
  ZK_HD static Fe r2() { return Fe{{P::R2[0], P::R2[1], P::R2[2], P::R2[3]}}; }
  ZK_HD static Fe r3() { return Fe{{P::R3[0], P::R3[1], P::R3[2], P::R3[3]}}; }
  ZK_HD static Fe generator() { return Fe{{P::GENERATOR[0], P::GENERATOR[1], P::GENERATOR[2], P::GENERATOR[3]}}; }
  ZK_HD static Fe root_of_unity() { return Fe{{P::ROOT_OF_UNITY[0], P::ROOT_OF_UNITY[1], P::ROOT_OF_UNITY[2], P::ROOT_OF_UNITY[3]}}; }
  ZK_HD static Fe delta() { return Fe{{P::DELTA[0], P::DELTA[1], P::DELTA[2], P::DELTA[3]}}; }
  ZK_HD static Fe zeta() { return Fe{{P::ZETA[0], P::ZETA[1], P::ZETA[2], P::ZETA[3]}}; }

  ZK_HD bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  ZK_HD bool operator==(const Fe& o) const {
    return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3];
  }
  ZK_HD bool operator!=(const Fe& o) const { return !(*this == o); }

  // a >= MOD ?
  ZK_HD static bool geq_mod(const uint64_t a[4]) {
    if (a[3] != M(3)) return a[3] > M(3);
    if (a[2] != M(2)) return a[2] > M(2);
    if (a[1] != M(1)) return a[1] > M(1);
    return a[0] >= M(0);
  }
  ZK_HD static void sub_mod_raw(uint64_t a[4]) {  // a -= MOD
    u128 d = (u128)a[0] - M(0);
    a[0] = (uint64_t)d;
    d = (u128)a[1] - M(1) - ((uint64_t)(d >> 64) & 1);
    a[1] = (uint64_t)d;
    d = (u128)a[2] - M(2) - ((uint64_t)(d >> 64) & 1);
    a[2] = (uint64_t)d;
    d = (u128)a[3] - M(3) - ((uint64_t)(d >> 64) & 1);
    a[3] = (uint64_t)d;
  }

  ZK_HD Fe operator+(const Fe& o) const {
#if defined(__CUDA_ARCH__)
    return add_device(o);
#else
    Fe r;
    u128 c = (u128)l[0] + o.l[0];
    r.l[0] = (uint64_t)c;
    c = (u128)l[1] + o.l[1] + (uint64_t)(c >> 64);
    r.l[1] = (uint64_t)c;
    c = (u128)l[2] + o.l[2] + (uint64_t)(c >> 64);
    r.l[2] = (uint64_t)c;
    c = (u128)l[3] + o.l[3] + (uint64_t)(c >> 64);
    r.l[3] = (uint64_t)c;
    // MOD < 2^255 so the sum of two reduced elements never carries out of 256 bits
    if (geq_mod(r.l)) sub_mod_raw(r.l);
    return r;
#endif
  }
  ZK_HD Fe operator-(const Fe& o) const {
#if defined(__CUDA_ARCH__)
    return sub_device(o);
#else
    Fe r;
    u128 d = (u128)l[0] - o.l[0];
    r.l[0] = (uint64_t)d;
    d = (u128)l[1] - o.l[1] - ((uint64_t)(d >> 64) & 1);
    r.l[1] = (uint64_t)d;
    d = (u128)l[2] - o.l[2] - ((uint64_t)(d >> 64) & 1);
    r.l[2] = (uint64_t)d;
    d = (u128)l[3] - o.l[3] - ((uint64_t)(d >> 64) & 1);
    r.l[3] = (uint64_t)d;
    if ((uint64_t)(d >> 64) & 1) {
      u128 c = (u128)r.l[0] + M(0);
      r.l[0] = (uint64_t)c;
      c = (u128)r.l[1] + M(1) + (uint64_t)(c >> 64);
      r.l[1] = (uint64_t)c;
      c = (u128)r.l[2] + M(2) + (uint64_t)(c >> 64);
      r.l[2] = (uint64_t)c;
      c = (u128)r.l[3] + M(3) + (uint64_t)(c >> 64);
      r.l[3] = (uint64_t)c;
    }
    return r;
#endif
  }
  ZK_HD Fe neg() const { return zero() - *this; }
  ZK_HD Fe dbl() const { return *this + *this; }

  // Montgomery product.  Both operands must be < MOD (the device path drops no carries only
  // under that bound; the portable CIOS below also tolerates one operand < 2^256).
  ZK_HD Fe operator*(const Fe& o) const {
#if defined(__CUDA_ARCH__)
    return mul_device(o);
#else
    return mul_portable(o);
#endif
  }
#if defined(__CUDACC__)
  // Device product: PTX carry chains over 32-bit limbs (mont_ptx.cuh), one asm statement per row.
  __device__ __forceinline__ Fe mul_device(const Fe& o) const {
    uint32_t a[8], E[8], O[8];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      a[2 * i] = (uint32_t)l[i];
      a[2 * i + 1] = (uint32_t)(l[i] >> 32);
    }
    if constexpr (P::MOD[0] == FpParams::MOD[0]) {
      montptx::mont_first_row_fp(E, O, a, (uint32_t)o.l[0]);
      montptx::mont_row_fp(O, E, a, (uint32_t)(o.l[0] >> 32));
      montptx::mont_row_fp(E, O, a, (uint32_t)o.l[1]);
      montptx::mont_row_fp(O, E, a, (uint32_t)(o.l[1] >> 32));
      montptx::mont_row_fp(E, O, a, (uint32_t)o.l[2]);
      montptx::mont_row_fp(O, E, a, (uint32_t)(o.l[2] >> 32));
      montptx::mont_row_fp(E, O, a, (uint32_t)o.l[3]);
      montptx::mont_row_fp(O, E, a, (uint32_t)(o.l[3] >> 32));
    } else {
      montptx::mont_first_row_fq(E, O, a, (uint32_t)o.l[0]);
      montptx::mont_row_fq(O, E, a, (uint32_t)(o.l[0] >> 32));
      montptx::mont_row_fq(E, O, a, (uint32_t)o.l[1]);
      montptx::mont_row_fq(O, E, a, (uint32_t)(o.l[1] >> 32));
      montptx::mont_row_fq(E, O, a, (uint32_t)o.l[2]);
      montptx::mont_row_fq(O, E, a, (uint32_t)(o.l[2] >> 32));
      montptx::mont_row_fq(E, O, a, (uint32_t)o.l[3]);
      montptx::mont_row_fq(O, E, a, (uint32_t)(o.l[3] >> 32));
    }
    montptx::mont_merge(E, O);
    if constexpr (P::MOD[0] == FpParams::MOD[0]) montptx::final_sub_fp(E); else montptx::final_sub_fq(E);
    return pack(E);
  }
  __device__ __forceinline__ static Fe pack(const uint32_t (&w)[8]) {
    Fe r;
#pragma unroll
    for (int i = 0; i < 4; i++) r.l[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
    return r;
  }
  __device__ __forceinline__ void unpack(uint32_t (&w)[8]) const {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      w[2 * i] = (uint32_t)l[i];
      w[2 * i + 1] = (uint32_t)(l[i] >> 32);
    }
  }
  __device__ __forceinline__ Fe add_device(const Fe& o) const {
    uint32_t a[8], b[8], r[8];
    unpack(a);
    o.unpack(b);
    if constexpr (P::MOD[0] == FpParams::MOD[0]) montptx::add_fp(r, a, b); else montptx::add_fq(r, a, b);
    return pack(r);
  }
  __device__ __forceinline__ Fe sub_device(const Fe& o) const {
    uint32_t a[8], b[8], r[8];
    unpack(a);
    o.unpack(b);
    if constexpr (P::MOD[0] == FpParams::MOD[0]) montptx::sub_fp(r, a, b); else montptx::sub_fq(r, a, b);
    return pack(r);
  }
#endif
  ZK_HD Fe mul_portable(const Fe& o) const {
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint64_t b = o.l[i];
      u128 c = (u128)l[0] * b + t0;
      t0 = (uint64_t)c;
      c = (u128)l[1] * b + t1 + (uint64_t)(c >> 64);
      t1 = (uint64_t)c;
      c = (u128)l[2] * b + t2 + (uint64_t)(c >> 64);
      t2 = (uint64_t)c;
      c = (u128)l[3] * b + t3 + (uint64_t)(c >> 64);
      t3 = (uint64_t)c;
      uint64_t t5 = t4 + (uint64_t)(c >> 64);  // cannot overflow: operands < 2^255-ish
      const uint64_t m = t0 * P::INV;
      c = (u128)m * M(0) + t0;
      c = (u128)m * M(1) + t1 + (uint64_t)(c >> 64);
      t0 = (uint64_t)c;
      c = (u128)m * M(2) + t2 + (uint64_t)(c >> 64);
      t1 = (uint64_t)c;
      c = (u128)m * M(3) + t3 + (uint64_t)(c >> 64);
      t2 = (uint64_t)c;
      c = (u128)t5 + (uint64_t)(c >> 64);
      t3 = (uint64_t)c;
      t4 = (uint64_t)(c >> 64);
    }
    Fe r{{t0, t1, t2, t3}};
    if (t4 || geq_mod(r.l)) sub_mod_raw(r.l);
    return r;
  }
  ZK_HD Fe sqr() const { return *this * *this; }
  ZK_HD Fe& operator+=(const Fe& o) { return *this = *this + o; }
  ZK_HD Fe& operator-=(const Fe& o) { return *this = *this - o; }
  ZK_HD Fe& operator*=(const Fe& o) { return *this = *this * o; }

  ZK_HD static Fe from_u64(uint64_t v) { return Fe{{v, 0, 0, 0}} * r2(); }
  ZK_HD static Fe from_canonical(const uint64_t v[4]) { return Fe{{v[0], v[1], v[2], v[3]}} * r2(); }
  ZK_HD void to_canonical(uint64_t out[4]) const {
    Fe r = *this * Fe{{1, 0, 0, 0}};
    out[0] = r.l[0]; out[1] = r.l[1]; out[2] = r.l[2]; out[3] = r.l[3];
  }
  // 512-bit little-endian integer reduced mod MOD (ff::FromUniformBytes<64>, Field::random)
  ZK_HD static Fe from_u512(const uint64_t v[8]) {
    // the device product requires both operands < MOD; a raw 256-bit word is < 4 MOD
    Fe lo{{v[0], v[1], v[2], v[3]}}, hi{{v[4], v[5], v[6], v[7]}};
#pragma unroll
    for (int i = 0; i < 3; i++) {
      if (geq_mod(lo.l)) sub_mod_raw(lo.l);
      if (geq_mod(hi.l)) sub_mod_raw(hi.l);
    }
    return lo * r2() + hi * r3();
  }
  ZK_HD Fe pow_u64(uint64_t e) const {
    Fe acc = one(), base = *this;
    while (e) {
      if (e & 1) acc = acc * base;
      base = base.sqr();
      e >>= 1;
    }
    return acc;
  }
  ZK_HD Fe pow256(const uint64_t e[4]) const {
    Fe acc = one();
    for (int i = 255; i >= 0; i--) {
      acc = acc.sqr();
      if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * *this;
    }
    return acc;
  }
  // Fermat inversion; 0 -> 0
  ZK_HD Fe inv() const {
    uint64_t e[4] = {M(0) - 2, M(1), M(2), M(3)};
    return pow256(e);
  }
  ZK_HD bool is_odd() const {
    uint64_t c[4];
    to_canonical(c);
    return c[0] & 1;
  }
};

typedef Fe<FpParams> Fp;
typedef Fe<FqParams> Fq;

}  // namespace zkodst
