// K1 — batched BLAKE2f witness generation.
//
// One thread block per compression region.  Phase 1: four lanes run the 12-round G mixing
// schedule (blake2f-circuit/src/README.md "Function Mix"; SIGMA = table16.rs:32-44, IV =
// table16.rs:47-56) and record every intermediate word — sums with their carries, XORs with
// their AND companions — in a shared-memory trace.  Phase 2: all threads walk the region's
// rows; each cell's descriptor (blake2f_layout.h) selects a bit-piece of a trace word and
// whether the cell carries its dense value, spread form (table16/util.rs:61-75 `spread_bits`)
// or range tag (spread_table.rs:213-222 `get_tag`); the value is converted to a
// Montgomery-form pallas::Base (table16.rs:93-98) and written with one 256-bit store per cell to the
// column-major advice buffer, consecutive threads writing consecutive rows of a column.
//
// Roofline: HBM store-bound.  Algorithmic bytes per compression = R * 12 * 32 written
// + 213 read; integer work per cell is ~30 ALU ops.
#include "zk_ctx.h"

namespace zkodst {

namespace {

__constant__ uint8_t c_sigma[10][16] = {
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

__constant__ uint64_t c_iv[8] = {
    0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
    0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};

// pallas::Base modulus p = 2^254 + t
constexpr uint64_t P0 = 0x992d30ed00000001ULL, P1 = 0x224698fc094cf91bULL, P3 = 0x4000000000000000ULL;
// 2^256 = 4p - 4t  =>  R = 2^256 mod p = p - 4t, and for 0 < v < 2^64:  v*R mod p = p - 4t*v
// (4t*v < 2^192 < p).  4t as two 64-bit words:
constexpr uint64_t C0 = P0 << 2, C1 = (P1 << 2) | (P0 >> 62);

__device__ __forceinline__ uint64_t rotr64(uint64_t x, uint32_t n) {
  return (x >> (n & 63)) | (x << ((64 - n) & 63));
}

// interleave_u16_with_zeros (spread_table.rs:729-736 in the commented test) == spread_bits
__device__ __forceinline__ uint32_t spread16(uint32_t x) {
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}

// one advice cell (4 x u64 limbs) with a single 256-bit store: a warp covers 1024 contiguous bytes of
// a column per instruction (STG.E.256, sm_100)
__device__ __forceinline__ void store_cell(uint64_t* p, uint64_t r0, uint64_t r1, uint64_t r2, uint64_t r3) {
  asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(r0), "l"(r1), "l"(r2), "l"(r3) : "memory");
}

// integer < 2^64 -> Montgomery-form Fp, stored at p:  v R mod p = p - 4t v  (0 < v < 2^64, see above).
// The 64 x 128-bit product runs as eight 32 x 32 -> 64 multiplies summed in two carry chains (even- and
// odd-aligned partial products), the subtraction from p as one borrow chain.
__device__ __forceinline__ void store_montgomery(uint64_t* p, uint64_t v) {
  if (v == 0) {
    store_cell(p, 0, 0, 0, 0);
    return;
  }
  constexpr uint32_t c0 = (uint32_t)C0, c1 = (uint32_t)(C0 >> 32), c2 = (uint32_t)C1, c3 = (uint32_t)(C1 >> 32);
  const uint32_t v0 = (uint32_t)v, v1 = (uint32_t)(v >> 32);
  uint64_t r0, r1, r2, r3;
  asm("{\n\t"
      ".reg .u64 a0, a1, b0, b1, d0, d1, e0, e1, x1, x2, y0, y1, y2, m1, m2, q0, q1, q2;\n\t"
      ".reg .u32 y0l, y0h, y1l, y1h, y2l, y2h, z;\n\t"
      "mul.wide.u32 a0, %4, %6;\n\t"   // v0 c0: limbs 0,1
      "mul.wide.u32 a1, %4, %8;\n\t"   // v0 c2: limbs 2,3
      "mul.wide.u32 b0, %4, %7;\n\t"   // v0 c1: limbs 1,2
      "mul.wide.u32 b1, %4, %9;\n\t"   // v0 c3: limbs 3,4
      "mul.wide.u32 d0, %5, %6;\n\t"   // v1 c0: limbs 1,2
      "mul.wide.u32 d1, %5, %8;\n\t"   // v1 c2: limbs 3,4
      "mul.wide.u32 e0, %5, %7;\n\t"   // v1 c1: limbs 2,3
      "mul.wide.u32 e1, %5, %9;\n\t"   // v1 c3: limbs 4,5
      "add.cc.u64 x1, a1, e0;\n\t"     // even-aligned words: a0 | a1 + e0 | e1 + carry
      "addc.u64 x2, e1, 0;\n\t"
      "add.cc.u64 y0, b0, d0;\n\t"     // odd-aligned words (limbs 1,2 | 3,4 | 5)
      "addc.cc.u64 y1, b1, d1;\n\t"
      "addc.u64 y2, 0, 0;\n\t"
      "mov.b64 {y0l, y0h}, y0;\n\t"
      "mov.b64 {y1l, y1h}, y1;\n\t"
      "mov.b64 {y2l, y2h}, y2;\n\t"
      "mov.u32 z, 0;\n\t"
      "mov.b64 q0, {z, y0l};\n\t"      // the odd chain shifted up by one limb
      "mov.b64 m1, {y0h, y1l};\n\t"
      "mov.b64 m2, {y1h, y2l};\n\t"
      "add.cc.u64 q0, q0, a0;\n\t"     // q = 4t v  (< 2^192)
      "addc.cc.u64 q1, m1, x1;\n\t"
      "addc.u64 q2, m2, x2;\n\t"
      "sub.cc.u64 %0, %10, q0;\n\t"    // p - q
      "subc.cc.u64 %1, %11, q1;\n\t"
      "subc.cc.u64 %2, 0, q2;\n\t"
      "subc.u64 %3, %12, 0;\n\t"
      "}"
      : "=l"(r0), "=l"(r1), "=l"(r2), "=l"(r3)
      : "r"(v0), "r"(v1), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(P0), "l"(P1), "l"(P3));
  store_cell(p, r0, r1, r2, r3);
}

// piece = rotr64(trace word, rot) & (2^len - 1), then dense value / spread form / range tag; straight-line
// (the three forms are a handful of ALU operations, cheaper than diverging on the kind)
__device__ __forceinline__ uint64_t eval_cell(uint32_t d, const uint64_t* __restrict__ trace) {
  const uint64_t w = trace[d >> 14];
  const uint32_t rot = (d >> 8) & 63, drop = 63 - ((d >> 2) & 63), kind = d & 3;  // drop = 64 - len
  uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
  if (rot & 32) {
    const uint32_t t = lo;
    lo = hi;
    hi = t;
  }
  const uint32_t xl = __funnelshift_r(lo, hi, rot), xh = __funnelshift_r(hi, lo, rot);  // shift amount mod 32
  const uint64_t x = (((uint64_t)xh << 32) | xl) & (~0ull >> drop);
  const uint32_t x32 = (uint32_t)x;
  const uint32_t tag = (x32 >= 256u) + (x32 >= 32768u);  // get_tag: pieces with a tag are <= 16 bits
  return kind == CK_DENSE ? x : (kind == CK_SPREAD ? (uint64_t)spread16(x32) : (uint64_t)tag);
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}

// One G (README.md "Function Mix"), recording the 8 primary/secondary trace word pairs.
__device__ __forceinline__ void mix_g(uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d,
                                      uint64_t x, uint64_t y, uint64_t* t, bool write) {
  uint64_t w[16];
  uint64_t s1 = a + b, c1 = s1 < a;
  uint64_t s2 = s1 + x;
  c1 += s2 < s1;
  w[0] = s2; w[1] = c1; a = s2;                         // a1 = a + b + x
  w[2] = d ^ a; w[3] = d & a; d = rotr64(d ^ a, 32);    // d1
  s1 = c + d;
  w[4] = s1; w[5] = s1 < c; c = s1;                     // c1
  w[6] = b ^ c; w[7] = b & c; b = rotr64(b ^ c, 24);    // b1
  s1 = a + b; c1 = s1 < a;
  s2 = s1 + y;
  c1 += s2 < s1;
  w[8] = s2; w[9] = c1; a = s2;                         // a2 = a1 + b1 + y
  w[10] = d ^ a; w[11] = d & a; d = rotr64(d ^ a, 16);  // d2
  s1 = c + d;
  w[12] = s1; w[13] = s1 < c; c = s1;                   // c2
  w[14] = b ^ c; w[15] = b & c; b = rotr64(b ^ c, 63);  // b2
  if (write) {
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = w[i];
  }
}

// Grid: n_compressions * slices blocks.  Block (comp, slice) rebuilds the compression's trace
// (cheap, and it keeps every block independent) and emits rows
// [slice * rows_per_slice, (slice + 1) * rows_per_slice) of that region, one thread per row: a cell
// is one 256-bit store, so every store instruction of a warp covers 1024 contiguous bytes of a column.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
blake2f_witness_kernel(const uint8_t* __restrict__ inputs, uint32_t rounds, uint32_t R,
                       const uint32_t* __restrict__ desc, uint64_t* __restrict__ advice, uint64_t n,
                       uint64_t* __restrict__ digests, uint64_t n_compressions, uint32_t slices,
                       uint32_t rows_per_slice, int* __restrict__ status) {
  extern __shared__ uint64_t trace[];  // trace_words
  __shared__ uint8_t rec[216];

  const uint64_t total = n_compressions * slices;
  for (uint64_t blk = blockIdx.x; blk < total; blk += gridDim.x) {
    const uint64_t comp = blk / slices;
    const uint32_t slice = (uint32_t)(blk % slices);
    // this item's rows; the first pass's cell descriptors are requested now so that their latency
    // overlaps the mixing schedule below
    const uint32_t row_begin = slice * rows_per_slice;
    const uint32_t row_end = min(R, row_begin + rows_per_slice);
    uint32_t dnext[NUM_USED_COLUMNS];
    if (row_begin + threadIdx.x < row_end) {
#pragma unroll
      for (int c = 0; c < NUM_USED_COLUMNS; c++) dnext[c] = __ldg(desc + (size_t)c * R + row_begin + threadIdx.x);
    }
    // ---- load and parse the 213-byte EIP-152 record ---------------------------------------
    for (int i = threadIdx.x; i < 213; i += THREADS) rec[i] = inputs[comp * 213 + i];
    __syncthreads();
    if (threadIdx.x < 26) {
      uint64_t w = 0;
      const uint8_t* p = rec + 4 + 8 * threadIdx.x;
#pragma unroll
      for (int b = 0; b < 8; b++) w |= (uint64_t)p[b] << (8 * b);
      // words 0..7 = h, 8..23 = m, 24,25 = t
      int dst = threadIdx.x < 8 ? TR_H + threadIdx.x
                                : (threadIdx.x < 24 ? TR_M + (threadIdx.x - 8)
                                                    : TR_T0 + (threadIdx.x - 24));
      trace[dst] = w;
    } else if (threadIdx.x < 34) {
      trace[TR_IV + (threadIdx.x - 26)] = c_iv[threadIdx.x - 26];
    } else if (threadIdx.x == 34) {
      uint32_t rr = ((uint32_t)rec[0] << 24) | ((uint32_t)rec[1] << 16) | ((uint32_t)rec[2] << 8) |
                    rec[3];
      if (rec[212] > 1 || rr != rounds) atomicOr(status, 1);
      trace[TR_FMASK] = rec[212] == 1 ? ~0ull : 0ull;
    }
    __syncthreads();

    // ---- phase 1: the mixing schedule.  Lane l of warp 0 owns column l of the 4x4 state
    // (a, b, c, d) = (v[l], v[4+l], v[8+l], v[12+l]); the diagonal step rotates b, c, d across
    // lanes with shuffles.  Four lanes = the four independent G of each half-round.
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x, l = lane & 3;
      const bool w = lane < 4;
      uint64_t a = trace[TR_H + l], b = trace[TR_H + 4 + l], c = c_iv[l], d = c_iv[4 + l];
      {  // v12 ^= t0, v13 ^= t1, v14 ^= fmask   (ops 0..2)
        uint64_t y = l < 3 ? trace[TR_T0 + l] : 0;
        if (lane < 3) {
          trace[TR_OPS + 2 * lane] = d ^ y;
          trace[TR_OPS + 2 * lane + 1] = d & y;
        }
        d ^= y;
      }
      for (uint32_t round = 0; round < rounds; round++) {
        const uint8_t* s = c_sigma[round % 10];
        uint64_t* t = trace + TR_OPS + 2 * (3 + (round * 8 + l) * 8);
        mix_g(a, b, c, d, trace[TR_M + s[2 * l]], trace[TR_M + s[2 * l + 1]], t, w);
        b = shfl64(b, (lane + 1) & 3);
        c = shfl64(c, (lane + 2) & 3);
        d = shfl64(d, (lane + 3) & 3);
        mix_g(a, b, c, d, trace[TR_M + s[8 + 2 * l]], trace[TR_M + s[8 + 2 * l + 1]], t + 64, w);
        b = shfl64(b, (lane + 3) & 3);
        c = shfl64(c, (lane + 2) & 3);
        d = shfl64(d, (lane + 1) & 3);
      }
      if (w) {  // h'_i = (h_i ^ v_i) ^ v_{i+8} for i = l and i = 4 + l
        uint64_t* t = trace + TR_OPS + 2 * (3 + rounds * 64);
        uint64_t h0 = trace[TR_H + l], h1 = trace[TR_H + 4 + l];
        uint64_t e0 = h0 ^ a, e1 = h1 ^ b;
        t[4 * l + 0] = e0;      t[4 * l + 1] = h0 & a;
        t[4 * l + 2] = e0 ^ c;  t[4 * l + 3] = e0 & c;
        t[4 * (4 + l) + 0] = e1;      t[4 * (4 + l) + 1] = h1 & b;
        t[4 * (4 + l) + 2] = e1 ^ d;  t[4 * (4 + l) + 3] = e1 & d;
        if (digests && slice == 0) {
          digests[comp * 8 + l] = e0 ^ c;
          digests[comp * 8 + 4 + l] = e1 ^ d;
        }
      }
    }
    __syncthreads();

    // ---- phase 2: emit this slice's cells ---------------------------------------------------
    const uint64_t base = comp * (uint64_t)R;
    for (uint32_t row = row_begin + threadIdx.x; row < row_end; row += THREADS) {
      uint32_t dcur[NUM_USED_COLUMNS];
#pragma unroll
      for (int c = 0; c < NUM_USED_COLUMNS; c++) dcur[c] = dnext[c];
      if (row + THREADS < row_end) {  // next pass's descriptors travel while this pass computes
#pragma unroll
        for (int c = 0; c < NUM_USED_COLUMNS; c++) dnext[c] = __ldg(desc + (size_t)c * R + row + THREADS);
      }
#pragma unroll
      for (int c = 0; c < NUM_USED_COLUMNS; c++)
        store_montgomery(advice + ((uint64_t)c * n + base + row) * 4, dcur[c] == 0 ? 0 : eval_cell(dcur[c], trace));
#pragma unroll
      for (int c = NUM_USED_COLUMNS; c < NUM_ADVICE_COLUMNS; c++)
        store_cell(advice + ((uint64_t)c * n + base + row) * 4, 0, 0, 0, 0);
    }
    __syncthreads();
  }
}

// rows [first_row, n) of every advice column := 0
__global__ void advice_tail_zero_kernel(uint4* __restrict__ advice, uint64_t n, uint64_t first_row) {
  uint64_t per_col = (n - first_row) * 2;  // uint4 units
  uint64_t total = per_col * NUM_ADVICE_COLUMNS;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t c = i / per_col, off = i % per_col;
    advice[(c * n + first_row) * 2 + off] = make_uint4(0, 0, 0, 0);
  }
}

}  // namespace

int32_t launch_witness(zk_ctx* ctx, int32_t k, uint32_t rounds, const uint8_t* d_inputs,
                       uint64_t n_compressions, void* d_advice, uint64_t* d_digests) {
  if (k < 17 || k > 28) return set_error(ctx, ZK_E_INVALID, "k out of range [17, 28]");
  if (!d_advice || (!d_inputs && n_compressions)) return set_error(ctx, ZK_E_INVALID, "null buffer");
  if ((uintptr_t)d_advice & 31) return set_error(ctx, ZK_E_INVALID, "advice buffer must be 32-byte aligned");
  DeviceRegionLayout* L = nullptr;
  int32_t rc = get_layout(ctx, rounds, &L);
  if (rc) return rc;
  const uint64_t n = 1ull << k, R = L->host.rows;
  const uint64_t usable = n - 6;  // blinding_factors 5 + 1 (docs/CIRCUIT.md)
  if (n_compressions > usable / R)  // no 64-bit wrap for absurd batch sizes
    return set_error(ctx, ZK_E_ROWS, "compressions do not fit in 2^k rows");
  const size_t smem = (size_t)L->host.trace_words * 8;
  constexpr int THREADS = 256;
  if (smem > 200 * 1024) return set_error(ctx, ZK_E_INVALID, "rounds too large for the trace buffer");
  if (smem > 48 * 1024)
    ZK_CUDA(ctx, cudaFuncSetAttribute(blake2f_witness_kernel<THREADS>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (n_compressions) {
    // One balanced wave: the work items are (compression, slice) pairs walked by a grid of at most
    // `cap` resident blocks.  Among the slice counts that keep a slice >= 256 rows, take the one that
    // fills the last round of the grid best (fewer slices on near-ties: every item recomputes the trace).
    int per_sm = 0;
    ZK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, blake2f_witness_kernel<THREADS>, THREADS, smem));
    const uint64_t cap = (uint64_t)ctx->sm_count * (uint64_t)(per_sm > 0 ? per_sm : 1);
    const uint32_t max_slices = (uint32_t)((R + 255) / 256);
    uint32_t slices = 1;
    double best = -1.0;
    for (uint32_t s = 1; s <= max_slices; s++) {
      const uint64_t items = n_compressions * s, rounds_of_grid = (items + cap - 1) / cap;
      const double fill = (double)items / (double)(rounds_of_grid * cap) - 0.004 * s;
      if (fill > best) {
        best = fill;
        slices = s;
      }
    }
    const uint32_t rows_per_slice = (uint32_t)((R + slices - 1) / slices);
    const uint64_t total = n_compressions * slices;
    const unsigned grid = (unsigned)(total < cap ? total : cap);
    KernelTimer timer(ctx, KC_WITNESS);
    blake2f_witness_kernel<THREADS><<<grid, THREADS, smem, ctx->stream>>>(
        d_inputs, rounds, (uint32_t)R, L->d_desc, (uint64_t*)d_advice, n, d_digests, n_compressions,
        slices, rows_per_slice, ctx->d_status);
    ctx->launches++;
  }
  ZK_CUDA(ctx, cudaGetLastError());
  {
    uint64_t first = n_compressions * R;
    unsigned grid = (unsigned)(ctx->sm_count * 8);
    advice_tail_zero_kernel<<<grid, 256, 0, ctx->stream>>>((uint4*)d_advice, n, first);
    ctx->launches++;
  }
  ZK_CUDA(ctx, cudaGetLastError());
  return ZK_OK;
}

}  // namespace zkodst
