// Multi-block hashing driver and EIP-152 wire format (host code, product).
//
// Replaces the streaming gadget `Blake2f::{new, update, finalize, digest}`
// (blake2f-circuit/src/blake2f.rs:88-181): IV -> for every 128-byte block `initialization` +
// `compress` -> `digest`.  Here the driver turns a message into the chain of EIP-152 records that
// the batched circuit proves — record i carries the chaining value h_i, block i, the byte counter
// and the final flag.  What `zk_create_proof` over those records proves: every record is a correct
// evaluation of F (its h' cells are F of its h, m, t, f cells) and, when the keys were generated with
// record chaining (zk_blake2f_keygen_chained, docs/CIRCUIT.md "Chaining"), that record i + 1 starts from
// record i's output.  The circuit has no instance column (the reference passes `&[&[]]`,
// benches/blake2f.rs:125): message, counter, flag and digest are not exposed to the verifier, so the proof
// is a statement about *some* chain of compressions, bound to this message only for whoever knows the witness.
// The chaining values need F itself (README.md "Function Compress"); 12 rounds of it per block on
// the host are negligible next to the proof.
#include <cstring>

#include "../../include/zkodst.h"

namespace {

const uint8_t SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                        0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};

inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

// BLAKE2b F with a free round count (EIP-152; table16.rs:32-56 for SIGMA / IV)
void blake2b_f(uint32_t rounds, uint64_t h[8], const uint64_t m[16], const uint64_t t[2], bool last) {
  uint64_t v[16];
  for (int i = 0; i < 8; i++) {
    v[i] = h[i];
    v[i + 8] = IV[i];
  }
  v[12] ^= t[0];
  v[13] ^= t[1];
  if (last) v[14] = ~v[14];
  for (uint32_t r = 0; r < rounds; r++) {
    const uint8_t* s = SIGMA[r % 10];
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
  for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}

void put_record(uint8_t* out, uint32_t rounds, const uint64_t h[8], const uint64_t m[16], const uint64_t t[2],
                bool last) {
  out[0] = (uint8_t)(rounds >> 24);
  out[1] = (uint8_t)(rounds >> 16);
  out[2] = (uint8_t)(rounds >> 8);
  out[3] = (uint8_t)rounds;
  memcpy(out + 4, h, 64);     // little-endian host (as everywhere in the library)
  memcpy(out + 68, m, 128);
  memcpy(out + 196, t, 16);
  out[212] = last ? 1 : 0;
}

}  // namespace

extern "C" int32_t zk_eip152_validate(const uint8_t* input, uint64_t len, uint32_t* rounds) {
  if (!input) return ZK_E_INVALID;
  if (len != ZK_BLAKE2F_INPUT_BYTES) return ZK_E_INPUT;   // EIP-152: any other length is an error
  if (input[212] > 1) return ZK_E_INPUT;                  // f must be 0 or 1
  if (rounds) *rounds = ((uint32_t)input[0] << 24) | ((uint32_t)input[1] << 16) | ((uint32_t)input[2] << 8) | input[3];
  return ZK_OK;
}

extern "C" int32_t zk_blake2f_compress(const uint8_t input[ZK_BLAKE2F_INPUT_BYTES], uint8_t out[64]) {
  uint32_t rounds = 0;
  int32_t rc = zk_eip152_validate(input, ZK_BLAKE2F_INPUT_BYTES, &rounds);
  if (rc) return rc;
  if (!out) return ZK_E_INVALID;
  uint64_t h[8], m[16], t[2];
  memcpy(h, input + 4, 64);
  memcpy(m, input + 68, 128);
  memcpy(t, input + 196, 16);
  blake2b_f(rounds, h, m, t, input[212] == 1);
  memcpy(out, h, 64);
  return ZK_OK;
}

extern "C" int32_t zk_blake2b_records(const uint8_t* msg, uint64_t len, uint32_t rounds, uint8_t* records_out,
                                      uint64_t* n_records, uint8_t digest_out[64]) {
  if (!n_records || (len && !msg)) return ZK_E_INVALID;
  const uint64_t blocks = len == 0 ? 1 : (len + 127) / 128;
  const uint64_t capacity = *n_records;
  *n_records = blocks;
  if (!records_out && !digest_out) return ZK_OK;  // size query
  if (records_out && capacity < blocks) return ZK_E_BUFFER;
  // BLAKE2b-512, unkeyed, no salt / personalisation: h_0 = IV ^ 0x01010040
  uint64_t h[8];
  for (int i = 0; i < 8; i++) h[i] = IV[i];
  h[0] ^= 0x01010040ULL;
  for (uint64_t b = 0; b < blocks; b++) {
    uint64_t m[16] = {0};
    const uint64_t off = b * 128;
    const uint64_t take = len - off < 128 ? len - off : 128;
    if (take) memcpy(m, msg + off, take);
    const bool last = b + 1 == blocks;
    const uint64_t t[2] = {off + take, 0};
    if (records_out) put_record(records_out + b * ZK_BLAKE2F_INPUT_BYTES, rounds, h, m, t, last);
    blake2b_f(rounds, h, m, t, last);
  }
  if (digest_out) memcpy(digest_out, h, 64);
  return ZK_OK;
}
