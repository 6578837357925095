// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// rand_xorshift 0.3.0 `XorShiftRng` (Cargo.lock:1571-1573), the seeded RNG the reference's
// timing harness hands to `create_proof`
// (benchmarking/src/blake2f_circuit_bench.rs:41-44).  Algorithm restated from the crate's
// published source (Marsaglia xorshift128, 32-bit words); pinned by tests/golden/xorshift.json
// (values computed by an independent pure-Python transcription, tests/golden/make_golden.py).
#pragma once
#include <cstdint>
#include <cstring>
#include "field.hpp"

namespace zko {

struct XorShiftRng {
  uint32_t x, y, z, w;
  static constexpr uint8_t REFERENCE_SEED[16] = {0x59, 0x62, 0xbe, 0x5d, 0x76, 0x3d, 0x31, 0x8d,
                                                 0x17, 0xdb, 0x37, 0x32, 0x54, 0x06, 0xbc, 0xe5};
  explicit XorShiftRng(const uint8_t seed[16]) {
    uint32_t s[4];
    memcpy(s, seed, 16);
    if ((s[0] | s[1] | s[2] | s[3]) == 0) {  // crate: all-zero seed replaced by a fixed state
      s[0] = 0x0BAD5EED;
      s[1] = 0x0BAD5EED;
      s[2] = 0x0BAD5EED;
      s[3] = 0x0BAD5EED;
    }
    x = s[0];
    y = s[1];
    z = s[2];
    w = s[3];
  }
  uint32_t next_u32() {
    uint32_t t = x ^ (x << 11);
    x = y;
    y = z;
    z = w;
    w = w ^ (w >> 19) ^ (t ^ (t >> 8));
    return w;
  }
  uint64_t next_u64() {  // rand_core::impls::next_u64_via_u32: low word first
    uint64_t lo = next_u32();
    uint64_t hi = next_u32();
    return lo | (hi << 32);
  }
  // ff::Field::random for pasta fields: 8 x next_u64 -> from_u512
  template <class F>
  F random_field() {
    uint64_t v[8];
    for (int i = 0; i < 8; i++) v[i] = next_u64();
    return F::from_u512(v);
  }
};

}  // namespace zko
