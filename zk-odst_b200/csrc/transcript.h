// Host-side transcript, hash and RNG of the prover (product code).
//
// Replaces, for the proving path, halo2_proofs 0.3.0 `transcript::{Blake2bWrite, Blake2bRead,
// Challenge255}` over `blake2b_simd` (blake2f-circuit/benches/blake2f.rs:124 `Blake2bWrite::init`,
// :141 `Blake2bRead::init`) and the seeded `XorShiftRng` the reference's harness passes to
// `create_proof` (benchmarking/src/blake2f_circuit_bench.rs:41-44).  These are byte-serial host
// computations (a few KB per proof) and stay on the CPU by design.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "ec.cuh"

namespace zkodst {

class Blake2bState {  // RFC 7693 BLAKE2b-512, unkeyed, with personalisation
 public:
  explicit Blake2bState(const char* personal16) {
    static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                   0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                   0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    uint64_t param[8] = {0x01010040ULL, 0, 0, 0, 0, 0, 0, 0};  // digest 64, fanout 1, depth 1
    memcpy(&param[6], personal16, 16);
    for (int i = 0; i < 8; i++) h_[i] = iv[i] ^ param[i];
    t_ = 0;
    fill_ = 0;
  }
  void update(const void* data, size_t len) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    while (len) {
      if (fill_ == 128) {
        t_ += 128;
        compress(false);
        fill_ = 0;
      }
      size_t take = 128 - fill_ < len ? 128 - fill_ : len;
      memcpy(block_ + fill_, p, take);
      fill_ += take;
      p += take;
      len -= take;
    }
  }
  void finalize(uint8_t out[64]) const {  // does not disturb the running state
    Blake2bState c = *this;
    c.t_ += c.fill_;
    memset(c.block_ + c.fill_, 0, 128 - c.fill_);
    c.compress(true);
    memcpy(out, c.h_, 64);
  }

 private:
  static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
  void compress(bool last) {
    static const uint8_t sigma[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                   0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                   0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    uint64_t m[16], v[16];
    memcpy(m, block_, 128);
    for (int i = 0; i < 8; i++) {
      v[i] = h_[i];
      v[i + 8] = iv[i];
    }
    v[12] ^= t_;
    if (last) v[14] = ~v[14];
    for (int r = 0; r < 12; r++) {
      const uint8_t* s = sigma[r];
      auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
        v[a] += v[b] + x; v[d] = rotr(v[d] ^ v[a], 32);
        v[c] += v[d];     v[b] = rotr(v[b] ^ v[c], 24);
        v[a] += v[b] + y; v[d] = rotr(v[d] ^ v[a], 16);
        v[c] += v[d];     v[b] = rotr(v[b] ^ v[c], 63);
      };
      G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
      G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
      G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
      G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h_[i] ^= v[i] ^ v[i + 8];
  }
  uint64_t h_[8];
  uint64_t t_;  // proofs are far below 2^64 bytes: the high counter word stays zero
  uint8_t block_[128];
  size_t fill_;
};

inline void fe_to_repr(const Fp& v, uint8_t out[32]) {
  uint64_t c[4];
  v.to_canonical(c);
  memcpy(out, c, 32);
}
inline void fq_to_repr(const Fq& v, uint8_t out[32]) {
  uint64_t c[4];
  v.to_canonical(c);
  memcpy(out, c, 32);
}
// group::GroupEncoding::to_bytes for vesta::Affine
inline void point_to_bytes(const Affine& p, uint8_t out[32]) {
  if (p.is_identity()) {
    memset(out, 0, 32);
    return;
  }
  fq_to_repr(p.x, out);
  if (p.y.is_odd()) out[31] |= 0x80;
}

class TranscriptWriter {  // Blake2bWrite<Vec<u8>, EqAffine, Challenge255<EqAffine>>
 public:
  TranscriptWriter() : st_("Halo2-Transcript") {}
  void common_scalar(const Fp& s) {
    uint8_t b[33];
    b[0] = 2;  // BLAKE2B_PREFIX_SCALAR
    fe_to_repr(s, b + 1);
    st_.update(b, 33);
  }
  void common_point(const Affine& p) {
    if (p.is_identity()) throw std::runtime_error("cannot write points at infinity to the transcript");
    uint8_t b[65];
    b[0] = 1;  // BLAKE2B_PREFIX_POINT
    fq_to_repr(p.x, b + 1);
    fq_to_repr(p.y, b + 33);
    st_.update(b, 65);
  }
  void write_point(const Affine& p) {
    common_point(p);
    uint8_t b[32];
    point_to_bytes(p, b);
    proof_.insert(proof_.end(), b, b + 32);
  }
  void write_scalar(const Fp& s) {
    common_scalar(s);
    uint8_t b[32];
    fe_to_repr(s, b);
    proof_.insert(proof_.end(), b, b + 32);
  }
  Fp squeeze_challenge() {
    uint8_t z = 0;  // BLAKE2B_PREFIX_CHALLENGE
    st_.update(&z, 1);
    uint8_t out[64];
    st_.finalize(out);
    uint64_t w[8];
    memcpy(w, out, 64);
    return Fp::from_u512(w);
  }
  const std::vector<uint8_t>& proof() const { return proof_; }

 private:
  Blake2bState st_;
  std::vector<uint8_t> proof_;
};

class TranscriptReader {  // Blake2bRead<&[u8], EqAffine, Challenge255<EqAffine>>
 public:
  TranscriptReader(const uint8_t* proof, size_t len) : st_("Halo2-Transcript"), in_(proof), len_(len), pos_(0) {
    fq_sqrt_exponent(tm1o2_);
  }
  void common_scalar(const Fp& s) {
    uint8_t b[33];
    b[0] = 2;
    fe_to_repr(s, b + 1);
    st_.update(b, 33);
  }
  void common_point(const Affine& p) {
    if (p.is_identity()) throw std::runtime_error("point at infinity in the transcript");
    uint8_t b[65];
    b[0] = 1;
    fq_to_repr(p.x, b + 1);
    fq_to_repr(p.y, b + 33);
    st_.update(b, 65);
  }
  Affine read_point() {
    if (pos_ + 32 > len_) throw std::runtime_error("proof truncated");
    Affine p;
    if (!decompress_point(in_ + pos_, tm1o2_, p)) throw std::runtime_error("invalid point encoding in proof");
    pos_ += 32;
    common_point(p);
    return p;
  }
  Fp read_scalar() {
    if (pos_ + 32 > len_) throw std::runtime_error("proof truncated");
    uint64_t c[4];
    memcpy(c, in_ + pos_, 32);
    if (Fp::geq_mod(c)) throw std::runtime_error("invalid field element encoding in proof");
    pos_ += 32;
    Fp s = Fp::from_canonical(c);
    common_scalar(s);
    return s;
  }
  Fp squeeze_challenge() {
    uint8_t z = 0;
    st_.update(&z, 1);
    uint8_t out[64];
    st_.finalize(out);
    uint64_t w[8];
    memcpy(w, out, 64);
    return Fp::from_u512(w);
  }
  bool exhausted() const { return pos_ == len_; }

 private:
  Blake2bState st_;
  const uint8_t* in_;
  size_t len_, pos_;
  uint64_t tm1o2_[4];
};

class XorShift {  // rand_xorshift 0.3.0
 public:
  explicit XorShift(const uint8_t seed[16]) {
    memcpy(s_, seed, 16);
    if (!(s_[0] | s_[1] | s_[2] | s_[3])) s_[0] = s_[1] = s_[2] = s_[3] = 0x0BAD5EED;
  }
  uint32_t next_u32() {
    uint32_t t = s_[0] ^ (s_[0] << 11);
    s_[0] = s_[1];
    s_[1] = s_[2];
    s_[2] = s_[3];
    s_[3] = s_[3] ^ (s_[3] >> 19) ^ (t ^ (t >> 8));
    return s_[3];
  }
  uint64_t next_u64() {
    uint64_t lo = next_u32();
    return lo | ((uint64_t)next_u32() << 32);
  }
  // Field::random: eight next_u64 -> from_u512 (raw words, reduced later on host or device)
  void next_wide(uint64_t out[8]) {
    for (int i = 0; i < 8; i++) out[i] = next_u64();
  }
  Fp random_fp() {
    uint64_t w[8];
    next_wide(w);
    return Fp::from_u512(w);
  }

 private:
  uint32_t s_[4];
};

}  // namespace zkodst
