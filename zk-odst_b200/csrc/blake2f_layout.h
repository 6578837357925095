// Static layout of one BLAKE2f compression region (docs/CIRCUIT.md), as tables the device
// kernels and keygen consume.  Host-side, product code.
//
// The reference intends this to live in `CompressionConfig::compress` and the row-offset
// helpers of compression/compression_util.rs:32-43,112-205 (both un-compiled upstream); here
// the region is described once as data:
//   * a cell descriptor per (advice column, row): which trace word, which bit-piece of it,
//     and whether the cell holds the piece's dense value, its spread form or its range tag
//     (SpreadVar::with_lookup, spread_table.rs:257-285);
//   * the copy constraints the chip issues (`copy_advice`), in call order;
//   * which of the 12 selectors (compression.rs:561-577) is enabled on which row.
#pragma once
#include <cstdint>
#include <vector>

namespace zkodst {

// Cell descriptor, 32 bits:  [1:0] kind  [7:2] len-1  [13:8] rotate  [31:14] trace word index
//   piece = rotr64(trace[word], rotate) & ((1 << len) - 1)
enum CellKind : uint32_t { CK_ZERO = 0, CK_DENSE = 1, CK_SPREAD = 2, CK_TAG = 3 };
static inline uint32_t cell_desc(uint32_t kind, uint32_t word, uint32_t rot, uint32_t len) {
  return kind | ((len - 1) << 2) | (rot << 8) | (word << 14);
}

// Trace word indices: 0..7 h, 8..15 IV, 16..31 m, 32 t0, 33 t1, 34 final-flag mask, then two
// words per operation in region order: primary (sum / xor) and secondary (carry / and).
enum { TR_H = 0, TR_IV = 8, TR_M = 16, TR_T0 = 32, TR_T1 = 33, TR_FMASK = 34, TR_OPS = 35 };

// the reference's 12 selectors (compression.rs:561-577), then the two that pin the IV words to the constants
// column and the final-flag mask to {0, 2^64 - 1} (docs/CIRCUIT.md "Pinned inputs")
enum Selector {
  SEL_A1 = 0, SEL_B1, SEL_C1, SEL_D1, SEL_A2, SEL_B2, SEL_C2, SEL_D2,
  SEL_ABCD, SEL_EFGH, SEL_IJKL, SEL_DIGEST, SEL_CONST, SEL_FMASK, NUM_SELECTORS
};

static const int NUM_ADVICE_COLUMNS = 12;   // halo2 advice column indices 0..11
static const int NUM_USED_COLUMNS = 10;     // idx 10, 11 are allocated and never assigned

struct CopyConstraint {  // rows relative to the region start; columns are halo2 advice indices
  uint8_t left_col;
  uint32_t left_row;
  uint8_t right_col;
  uint32_t right_row;
};

struct RegionLayout {
  uint32_t rounds = 0;
  uint32_t rows = 0;          // R = 292 + 392 * rounds
  uint32_t trace_words = 0;   // 35 + 2 * (19 + 64 * rounds)
  std::vector<uint32_t> desc;             // [NUM_USED_COLUMNS][rows], by halo2 column index
  std::vector<CopyConstraint> copies;     // in copy_advice call order
  std::vector<uint8_t> selectors;         // [NUM_SELECTORS][rows]
  std::vector<uint64_t> constants;        // [rows]: the constants fixed column (IV words on SEL_CONST rows)
  uint32_t digest_word[8];                // trace word index of each output word h'_i
  // Word cells the chaining copies connect (CompressionConfig::initialize_with_state, compression.rs:1096-1111):
  // h_i enters in a_3 (advice column 1) at h_word_row[i]; h'_i leaves in a_5 (advice column 0) at out_word_row[i].
  uint32_t h_word_row[8], out_word_row[8];
};
static const uint8_t CHAIN_H_COLUMN = 1, CHAIN_OUT_COLUMN = 0;  // halo2 advice indices of a_3 and a_5

static inline uint64_t region_rows(uint32_t rounds) { return 292ull + 392ull * rounds; }

// Builds the layout tables for a given round count.
void build_region_layout(uint32_t rounds, RegionLayout& out);

}  // namespace zkodst
