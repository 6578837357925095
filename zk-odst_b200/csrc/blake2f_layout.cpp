// Region layout builder — see blake2f_layout.h and docs/CIRCUIT.md.
#include "blake2f_layout.h"

#include <stdexcept>

namespace zkodst {
namespace {

// a_0..a_9 of the reference's gate comments -> halo2 advice column index
// (table16.rs:281-310: idx0 = message_schedule = a_5, idx1..6 = extras = a_3,a_4,a_6,a_7,a_8,a_9,
//  idx7..9 = lookup inputs a_0,a_1,a_2).
const uint8_t COL_OF_A[10] = {7, 8, 9, 1, 2, 0, 3, 4, 5, 6};

const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                        0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};

const uint8_t SIGMA[10][16] = {
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

struct Val {      // an assigned cell: what it holds and where
  uint32_t desc;
  uint8_t a;      // a-number
  uint32_t row;
};
struct Word {
  Val dense[4], spread[4];
};

struct Builder {
  RegionLayout& L;
  uint32_t next_op = 0;
  explicit Builder(RegionLayout& l) : L(l) {}

  void set(uint8_t a, uint32_t row, uint32_t desc) {
    L.desc[(size_t)COL_OF_A[a] * L.rows + row] = desc;
  }
  Val assign(uint8_t a, uint32_t row, uint32_t desc) {
    set(a, row, desc);
    return Val{desc, a, row};
  }
  void copy(const Val& src, uint8_t a, uint32_t row) {
    set(a, row, src.desc);
    L.copies.push_back(CopyConstraint{COL_OF_A[src.a], src.row, COL_OF_A[a], row});
  }
  void enable(int sel, uint32_t row) { L.selectors[(size_t)sel * L.rows + row] = 1; }
  // one row of the (tag, dense, spread) lookup triple for piece rotr(word, rot)[0:len]
  void lookup(uint32_t row, uint32_t word, uint32_t rot, uint32_t len, Val& dense, Val& spread) {
    assign(0, row, cell_desc(CK_TAG, word, rot, len));
    dense = assign(1, row, cell_desc(CK_DENSE, word, rot, len));
    spread = assign(2, row, cell_desc(CK_SPREAD, word, rot, len));
  }
  Word limbs_at(uint32_t r, uint32_t word, uint32_t stride) {
    Word w;
    for (uint32_t i = 0; i < 4; i++) lookup(r + stride * i, word, 16 * i, 16, w.dense[i], w.spread[i]);
    return w;
  }
  uint32_t new_op() { return TR_OPS + 2 * next_op++; }

  Word input_word(uint32_t r, uint32_t word) {
    Word w = limbs_at(r, word, 1);
    assign(3, r + 1, cell_desc(CK_DENSE, word, 0, 64));
    copy(w.dense[3], 4, r + 1);
    enable(SEL_ABCD, r + 1);
    return w;
  }
  void window8(uint32_t r, const Val x[4], const Val y[4]) {
    for (int i = 0; i < 4; i++) copy(x[i], 3 + i, r);
    copy(y[0], 7, r);
    copy(y[1], 8, r);
    copy(y[2], 3, r + 1);
    copy(y[3], 4, r + 1);
  }
  Word add(uint32_t r, const Word& A, const Word& B, const Word* X, int sel) {
    uint32_t p = new_op();
    Word w = limbs_at(r, p, 1);
    window8(r, A.dense, B.dense);
    if (X)
      for (int i = 0; i < 4; i++) copy(X->dense[i], 5 + i, r + 1);
    copy(w.dense[3], 3, r + 2);
    assign(9, r + 1, cell_desc(CK_DENSE, p + 1, 0, 2));
    enable(sel, r + 1);
    return w;
  }
  Word xor_aligned(uint32_t r, const Word& X, const Word& Y, int sel, int rot_limbs) {
    uint32_t p = new_op();
    Word e;
    for (uint32_t i = 0; i < 4; i++) {
      Val od, os;
      lookup(r + 2 * i, p, 16 * i, 16, e.dense[i], e.spread[i]);
      lookup(r + 2 * i + 1, p + 1, 16 * i, 16, od, os);
      copy(X.spread[i], 3, r + 2 * i);
      copy(Y.spread[i], 4, r + 2 * i);
      enable(sel, r + 2 * i);
    }
    Word w;
    for (int i = 0; i < 4; i++) {
      w.dense[i] = e.dense[(i + rot_limbs) % 4];
      w.spread[i] = e.spread[(i + rot_limbs) % 4];
    }
    return w;
  }
  Word xor_rotr24(uint32_t r, const Word& X, const Word& Y) {
    uint32_t p = new_op();
    Val pd[5], ps[5], od[4], os[4];
    lookup(r + 0, p, 0, 8, pd[0], ps[0]);
    lookup(r + 1, p, 56, 8, pd[4], ps[4]);
    lookup(r + 2, p, 8, 16, pd[1], ps[1]);
    lookup(r + 3, p, 24, 16, pd[2], ps[2]);
    lookup(r + 4, p, 40, 16, pd[3], ps[3]);
    for (uint32_t i = 0; i < 4; i++) lookup(r + 5 + i, p + 1, 16 * i, 16, od[i], os[i]);
    Val r2d = assign(3, r, cell_desc(CK_DENSE, p, 56, 16));
    Val r2s = assign(4, r, cell_desc(CK_SPREAD, p, 56, 16));
    enable(SEL_EFGH, r);
    window8(r + 5, X.spread, Y.spread);
    for (int i = 0; i < 4; i++) copy(ps[i], 5 + i, r + 6);
    copy(ps[4], 3, r + 7);
    copy(os[3], 4, r + 7);
    enable(SEL_B1, r + 6);
    Word w;
    w.dense[0] = pd[2]; w.spread[0] = ps[2];
    w.dense[1] = pd[3]; w.spread[1] = ps[3];
    w.dense[2] = r2d;   w.spread[2] = r2s;
    w.dense[3] = pd[1]; w.spread[3] = ps[1];
    return w;
  }
  Word xor_rotr63(uint32_t r, const Word& X, const Word& Y) {
    uint32_t p = new_op();
    Val qd[4], qs[4], od[4], os[4];
    lookup(r + 0, p, 0, 15, qd[0], qs[0]);
    lookup(r + 1, p, 15, 16, qd[1], qs[1]);
    lookup(r + 2, p, 31, 16, qd[2], qs[2]);
    lookup(r + 3, p, 47, 16, qd[3], qs[3]);
    for (uint32_t i = 0; i < 4; i++) lookup(r + 4 + i, p + 1, 16 * i, 16, od[i], os[i]);
    Val r0d = assign(3, r, cell_desc(CK_DENSE, p, 63, 16));
    Val r0s = assign(4, r, cell_desc(CK_SPREAD, p, 63, 16));
    Val bit = assign(5, r, cell_desc(CK_DENSE, p, 63, 1));
    enable(SEL_IJKL, r);
    window8(r + 4, X.spread, Y.spread);
    for (int i = 0; i < 4; i++) copy(qs[i], 5 + i, r + 5);
    copy(bit, 3, r + 6);
    copy(os[3], 4, r + 6);
    enable(SEL_B2, r + 5);
    Word w;
    w.dense[0] = r0d;   w.spread[0] = r0s;
    w.dense[1] = qd[1]; w.spread[1] = qs[1];
    w.dense[2] = qd[2]; w.spread[2] = qs[2];
    w.dense[3] = qd[3]; w.spread[3] = qs[3];
    return w;
  }
  uint32_t xor_digest(uint32_t r, const Word& X, const Word& Y) {
    uint32_t p = new_op();
    Val ed[4], es[4], od[4], os[4];
    for (uint32_t i = 0; i < 4; i++) lookup(r + i, p, 16 * i, 16, ed[i], es[i]);
    for (uint32_t i = 0; i < 4; i++) lookup(r + 4 + i, p + 1, 16 * i, 16, od[i], os[i]);
    window8(r, X.spread, Y.spread);
    copy(es[3], 5, r + 1);
    copy(os[0], 6, r + 1);
    copy(os[1], 7, r + 1);
    copy(os[2], 8, r + 1);
    copy(os[3], 3, r + 2);
    copy(ed[3], 4, r + 2);
    assign(5, r + 2, cell_desc(CK_DENSE, p, 0, 64));
    enable(SEL_DIGEST, r + 1);
    return p;
  }
};

}  // namespace

void build_region_layout(uint32_t rounds, RegionLayout& L) {
  L.rounds = rounds;
  uint64_t rows = region_rows(rounds);
  if (rows > 0xffffffffull) throw std::runtime_error("rounds too large");
  L.rows = (uint32_t)rows;
  L.trace_words = TR_OPS + 2 * (19 + 64 * rounds);
  if (L.trace_words >= (1u << 18)) throw std::runtime_error("rounds too large for descriptors");
  L.desc.assign((size_t)NUM_USED_COLUMNS * L.rows, 0);
  L.selectors.assign((size_t)NUM_SELECTORS * L.rows, 0);
  L.constants.assign(L.rows, 0);
  L.copies.clear();
  Builder b(L);
  uint32_t r = 0;
  Word h[8], iv[8], m[16], v[16];
  for (int i = 0; i < 8; i++, r += 4) {
    h[i] = b.input_word(r, TR_H + i);
    L.h_word_row[i] = r + 1;
  }
  for (int i = 0; i < 8; i++, r += 4) {  // IV words: pinned to the constants column
    iv[i] = b.input_word(r, TR_IV + i);
    b.enable(SEL_CONST, r + 1);
    L.constants[r + 1] = IV[i];
  }
  for (int i = 0; i < 16; i++, r += 4) m[i] = b.input_word(r, TR_M + i);
  Word t0 = b.input_word(r, TR_T0); r += 4;
  Word t1 = b.input_word(r, TR_T1); r += 4;
  Word fm = b.input_word(r, TR_FMASK);  // final-flag mask = bit * (2^64 - 1), the bit in a_9
  b.assign(9, r + 1, cell_desc(CK_DENSE, TR_FMASK, 0, 1));
  b.enable(SEL_FMASK, r + 1);
  r += 4;
  for (int i = 0; i < 8; i++) {
    v[i] = h[i];
    v[i + 8] = iv[i];
  }
  v[12] = b.xor_aligned(r, iv[4], t0, SEL_D1, 0); r += 8;
  v[13] = b.xor_aligned(r, iv[5], t1, SEL_D1, 0); r += 8;
  v[14] = b.xor_aligned(r, iv[6], fm, SEL_D1, 0); r += 8;
  static const int GI[8][4] = {{0, 4, 8, 12}, {1, 5, 9, 13}, {2, 6, 10, 14}, {3, 7, 11, 15},
                               {0, 5, 10, 15}, {1, 6, 11, 12}, {2, 7, 8, 13}, {3, 4, 9, 14}};
  for (uint32_t round = 0; round < rounds; round++) {
    const uint8_t* s = SIGMA[round % 10];
    for (int g = 0; g < 8; g++) {
      int a = GI[g][0], bb = GI[g][1], c = GI[g][2], d = GI[g][3];
      v[a] = b.add(r, v[a], v[bb], &m[s[2 * g]], SEL_A1); r += 4;
      v[d] = b.xor_aligned(r, v[d], v[a], SEL_D1, 2); r += 8;
      v[c] = b.add(r, v[c], v[d], nullptr, SEL_C1); r += 4;
      v[bb] = b.xor_rotr24(r, v[bb], v[c]); r += 9;
      v[a] = b.add(r, v[a], v[bb], &m[s[2 * g + 1]], SEL_A2); r += 4;
      v[d] = b.xor_aligned(r, v[d], v[a], SEL_D2, 1); r += 8;
      v[c] = b.add(r, v[c], v[d], nullptr, SEL_C2); r += 4;
      v[bb] = b.xor_rotr63(r, v[bb], v[c]); r += 8;
    }
  }
  for (int i = 0; i < 8; i++) {
    Word t = b.xor_aligned(r, h[i], v[i], SEL_D1, 0); r += 8;
    L.digest_word[i] = b.xor_digest(r, t, v[i + 8]);
    L.out_word_row[i] = r + 2;
    r += 8;
  }
  if (r != L.rows) throw std::logic_error("region row count mismatch");
  if (TR_OPS + 2 * b.next_op != L.trace_words) throw std::logic_error("trace word count mismatch");
}

}  // namespace zkodst
