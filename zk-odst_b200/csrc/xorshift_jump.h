// Jump-ahead for the seeded XorShiftRng the reference's harness passes to `create_proof`
// (rand_xorshift 0.3.0, benchmarking/src/blake2f_circuit_bench.rs:41-44).  Product code.
//
// The xorshift128 step is linear over GF(2) on the 128-bit state, so skipping N outputs is a
// product with T^N.  The powers T^(2^j) are built once on the host by repeated squaring; the two
// n-element random polynomials of a proof (`vanishing::Argument::commit` and the IPA's `s`,
// halo2_proofs 0.3.0) are then generated ON THE DEVICE, each thread jumping to its own offset of
// the stream, while the host skips the same span in O(log n).  The byte stream is exactly the
// sequential one.
#pragma once
#include <cstdint>

#include "field.cuh"

namespace zkodst {

struct XsState {
  uint32_t s[4];
};
struct XsMatrix {  // column j = image of state bit j (bit j lives in word j / 32, bit j % 32)
  uint32_t col[128][4];
};

ZK_HD uint32_t xs_step(XsState& st) {
  uint32_t t = st.s[0] ^ (st.s[0] << 11);
  st.s[0] = st.s[1];
  st.s[1] = st.s[2];
  st.s[2] = st.s[3];
  st.s[3] = st.s[3] ^ (st.s[3] >> 19) ^ (t ^ (t >> 8));
  return st.s[3];
}

ZK_HD XsState xs_apply(const XsMatrix& m, const XsState& v) {
  XsState r{{0, 0, 0, 0}};
  for (int w = 0; w < 4; w++) {
    uint32_t bits = v.s[w];
    for (int b = 0; b < 32; b++) {
      if ((bits >> b) & 1) {
        const uint32_t* c = m.col[w * 32 + b];
        r.s[0] ^= c[0];
        r.s[1] ^= c[1];
        r.s[2] ^= c[2];
        r.s[3] ^= c[3];
      }
    }
  }
  return r;
}

constexpr int XS_JUMP_POWERS = 48;  // T^(2^j), j < 48

// host: table of T^(2^j), built on first use
inline const XsMatrix* xs_jump_table() {
  static XsMatrix* table = [] {
    XsMatrix* t = new XsMatrix[XS_JUMP_POWERS];
    for (int j = 0; j < 128; j++) {
      XsState e{{0, 0, 0, 0}};
      e.s[j >> 5] = 1u << (j & 31);
      xs_step(e);
      for (int w = 0; w < 4; w++) t[0].col[j][w] = e.s[w];
    }
    for (int p = 1; p < XS_JUMP_POWERS; p++)
      for (int j = 0; j < 128; j++) {
        XsState c{{t[p - 1].col[j][0], t[p - 1].col[j][1], t[p - 1].col[j][2], t[p - 1].col[j][3]}};
        XsState sq = xs_apply(t[p - 1], c);
        for (int w = 0; w < 4; w++) t[p].col[j][w] = sq.s[w];
      }
    return t;
  }();
  return table;
}

inline XsState xs_jump(XsState st, uint64_t steps) {
  const XsMatrix* t = xs_jump_table();
  for (int j = 0; j < XS_JUMP_POWERS && (steps >> j); j++)
    if ((steps >> j) & 1) st = xs_apply(t[j], st);
  return st;
}

}  // namespace zkodst
