// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// (1) BLAKE2b `F` compression function with a variable round count, the function the
//     reference circuit arithmetises (blake2f-circuit/src/README.md "Function Compress" /
//     "Function Mix"; SIGMA = table16.rs:32-44 `ROUND_CONSTANTS`, IV = table16.rs:47-56) and
//     EIP-152 defines.  Pinned by EIP-152 vector 5 (blake2f-circuit/src/blake2f.rs:193-247) and
//     vectors 4/6/7 (tests/golden/eip152.json).
// (2) BLAKE2b-512 with personalisation (RFC 7693), the transcript / vk hash of halo2_proofs
//     0.3.0 (`blake2b_simd 1.0.1`, Cargo.lock:155-157).  Pinned by python hashlib in
//     tests/test_oracle_blake2b.py.
#pragma once
#include <cstdint>
#include <cstring>

namespace zko {

static const uint64_t BLAKE2B_IV[8] = {
    0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
    0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};

static const uint8_t BLAKE2B_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
    {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4},
    {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13},
    {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11},
    {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5},
    {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

static inline uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

// EIP-152 F: h (in/out), m[16], t[2], final flag, rounds.
static inline void blake2b_F(uint64_t h[8], const uint64_t m[16], const uint64_t t[2], bool f,
                             uint32_t rounds) {
  uint64_t v[16];
  for (int i = 0; i < 8; i++) {
    v[i] = h[i];
    v[i + 8] = BLAKE2B_IV[i];
  }
  v[12] ^= t[0];
  v[13] ^= t[1];
  if (f) v[14] = ~v[14];
  auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
    v[a] = v[a] + v[b] + x;
    v[d] = rotr64(v[d] ^ v[a], 32);
    v[c] = v[c] + v[d];
    v[b] = rotr64(v[b] ^ v[c], 24);
    v[a] = v[a] + v[b] + y;
    v[d] = rotr64(v[d] ^ v[a], 16);
    v[c] = v[c] + v[d];
    v[b] = rotr64(v[b] ^ v[c], 63);
  };
  for (uint32_t r = 0; r < rounds; r++) {
    const uint8_t* s = BLAKE2B_SIGMA[r % 10];
    G(0, 4, 8, 12, m[s[0]], m[s[1]]);
    G(1, 5, 9, 13, m[s[2]], m[s[3]]);
    G(2, 6, 10, 14, m[s[4]], m[s[5]]);
    G(3, 7, 11, 15, m[s[6]], m[s[7]]);
    G(0, 5, 10, 15, m[s[8]], m[s[9]]);
    G(1, 6, 11, 12, m[s[10]], m[s[11]]);
    G(2, 7, 8, 13, m[s[12]], m[s[13]]);
    G(3, 4, 9, 14, m[s[14]], m[s[15]]);
  }
  for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}

// Incremental BLAKE2b-512 (unkeyed, 16-byte personalisation).  Copyable, so a transcript can
// clone its running state to squeeze a challenge (halo2 Blake2bWrite::squeeze_challenge).
struct Blake2b {
  uint64_t h[8];
  uint64_t t[2];
  uint8_t buf[128];
  size_t buflen;
  size_t outlen;

  explicit Blake2b(const char personal[16] = nullptr, size_t outlen_ = 64) {
    uint8_t param[64];
    memset(param, 0, 64);
    param[0] = (uint8_t)outlen_;
    param[2] = 1;
    param[3] = 1;
    if (personal) memcpy(param + 48, personal, 16);
    for (int i = 0; i < 8; i++) {
      uint64_t w;
      memcpy(&w, param + 8 * i, 8);
      h[i] = BLAKE2B_IV[i] ^ w;
    }
    t[0] = t[1] = 0;
    buflen = 0;
    outlen = outlen_;
    memset(buf, 0, 128);
  }
  void compress(const uint8_t block[128], bool last) {
    uint64_t m[16];
    memcpy(m, block, 128);
    blake2b_F(h, m, t, last, 12);
  }
  void update(const void* data, size_t len) {
    const uint8_t* p = (const uint8_t*)data;
    while (len > 0) {
      if (buflen == 128) {  // buffer full and more input follows: it is not the last block
        t[0] += 128;
        if (t[0] < 128) t[1]++;
        compress(buf, false);
        buflen = 0;
      }
      size_t take = 128 - buflen;
      if (take > len) take = len;
      memcpy(buf + buflen, p, take);
      buflen += take;
      p += take;
      len -= take;
    }
  }
  void finalize(uint8_t* out) const {
    Blake2b c = *this;
    c.t[0] += c.buflen;
    if (c.t[0] < c.buflen) c.t[1]++;
    memset(c.buf + c.buflen, 0, 128 - c.buflen);
    c.compress(c.buf, true);
    memcpy(out, c.h, outlen);
  }
};

}  // namespace zko
