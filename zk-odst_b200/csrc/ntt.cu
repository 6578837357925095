// K4/K5 — number-theoretic transforms over Fp.
//
// Replaces halo2_proofs 0.3.0 `best_fft` as reached through `EvaluationDomain::
// {lagrange_to_coeff, coeff_to_extended, extended_to_coeff}` inside `create_proof`
// (blake2f-circuit/benches/blake2f.rs:125).  Natural order in, natural order out.
//
// Decimation-in-time butterflies, executed as shared-memory passes of up to 10 stages over tiles of 1024
// elements: a tile is G independent groups x 2^S transform indices x W adjacent columns (W * 32 B
// contiguous per row), so a 2^19-point transform touches HBM twice instead of 19 times.  Inside a pass
// the stages run in rounds: a thread takes 2^Q elements into registers, runs Q stages (Q = 2: four
// butterflies, three twiddles) and puts them back, so shared memory and the barrier are visited once per
// Q stages.  The tile is padded by one element in eight, which spreads every round's stride over
// the banks.  The first pass fuses the bit-reversal gather, zero padding and an optional per-element
// input scaling (coset evaluation: c^i from a table); the last pass fuses the 1/N scaling.  Several
// transforms of one size run as one batch of launches (blockIdx.y, blockIdx.z), which also fills the
// grid: one 2^19-point transform is only 512 tiles.  Twiddles come from a per-domain
// table of N/2 powers; stage 0's twiddle is 1 and its multiplication is skipped.
//
// Roofline: algorithmic bytes 64*N per transform; work (N/2) log2 N Fp multiplications
// (~20 MAC/B at N = 2^19): integer-pipe bound; both fractions are reported.
#include "field.cuh"
#include "prover.h"
#include "zk_ctx.h"

namespace zkodst {
namespace {

__global__ void powers_kernel(Fp base, Fp* out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = base.pow_u64(i);
}

__device__ __forceinline__ uint32_t bitrev(uint32_t v, int bits) { return __brev(v) >> (32 - bits); }

constexpr int NTT_TILE_LOG = 10;     // elements per tile (log2)
constexpr int NTT_THREADS = 128;     // 8 elements per thread
constexpr int NTT_MAX_ROUNDS = 5;
// Stages per round and resident blocks per SM.  Measured on B200 inside the prover (k = 19, NTT ms per
// proof): 3 stages / 3 blocks (168 registers) 7.71, 3 / 4 (128 registers, spills) 7.22, 2 / 6 (80 registers)
// 6.87, 2 / 8 (64 registers, spills) 7.07 — the butterflies are bound by the half-rate IMAD.WIDE.X
// chains, so what the wider rounds save in shared-memory trips they lose in resident warps.
// Round 2 (profiles/r02_ntt_variants.json): 256-thread blocks — one radix-4 unit per thread and round — at 3 or 4
// blocks per SM are 1.6-1.9 % slower than 128 threads x 6 blocks; the six resident blocks are what the 221 KB of
// shared memory allow either way.
constexpr int NTT_ROUND_STAGES = 2;
constexpr int NTT_MIN_BLOCKS = 6;

struct NttPassArgs {
  const Fp* in;
  Fp* out;
  const Fp* tw;      // omega^i, i < N/2
  const Fp* scale_in;  // first pass: optional per-element multiplier
  size_t in_stride, out_stride, scale_stride;  // per transform of the batch (blockIdx.y)
  size_t in_stride2, out_stride2;              // second batch level (blockIdx.z)
  int log_n;         // L
  int s0;            // stages already done
  int S;             // stages in this pass
  int logW;          // log2 of adjacent columns per tile (0 for the first pass)
  int logG;          // log2 of independent high-index groups per tile
  int nrounds;       // the S stages as rounds of q[i] in {1, 2, 3} stages
  int q[NTT_MAX_ROUNDS];
  uint32_t n_in;     // first pass: input length (zero padded above)
  int first, last;
  int scale_out;     // last pass: multiply by `scale` (1/N of the inverse transform)
  Fp scale;
};

// slot of tile element e: one element of padding after every eight
__device__ __forceinline__ uint32_t ntt_slot(uint32_t e) { return e + (e >> 3); }

// One round: Q consecutive stages, starting at stage r of the pass, on 2^Q elements held in registers.
// Tile element e = (g << (S + logW)) | (t << logW) | l; a unit is the 2^Q elements whose t differ only in
// bits [r, r + Q).  The butterfly of stage r + q on the pair (m, m | 2^q) uses the twiddle of global index
// j = (((m mod 2^q) << r | t mod 2^r) << s0) | lo, scaled to the domain: omega^(j * N / 2^(s0 + r + q + 1)).
template <int Q>
__device__ __forceinline__ void ntt_round(Fp* __restrict__ sm, const NttPassArgs& a, int r, uint32_t lo0,
                                          uint32_t tile_elems) {
  const int logW = a.logW;
  const uint32_t W = 1u << logW;
  const uint32_t units = tile_elems >> Q;
  const uint32_t mstride = 1u << (r + logW);
  const bool unit_twiddle = a.s0 + r == 0;  // stage 0 of the transform: omega^0
  for (uint32_t u = threadIdx.x; u < units; u += NTT_THREADS) {
    const uint32_t l = u & (W - 1), ub = u >> logW;
    const uint32_t tlo = ub & ((1u << r) - 1), up = ub >> r;
    const uint32_t ebase = ((((up << (r + Q)) | tlo)) << logW) | l;
    const uint32_t lo = lo0 + l;
    Fp x[1 << Q];
#pragma unroll
    for (int m = 0; m < (1 << Q); m++) x[m] = sm[ntt_slot(ebase + m * mstride)];
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int shift = a.log_n - (a.s0 + r + q) - 1;
#pragma unroll
      for (int mlow = 0; mlow < (1 << q); mlow++) {
        const uint32_t j = ((((uint32_t)mlow << r) | tlo) << a.s0) | lo;
        const Fp w = a.tw[(size_t)j << shift];
#pragma unroll
        for (int mh = 0; mh < (1 << (Q - q - 1)); mh++) {
          const int i0 = (mh << (q + 1)) | mlow, i1 = i0 | (1 << q);
          // omega^0: stage 0 of the transform, and the mlow = 0 butterflies of stage 1 in the same first round
          // (r = 0, s0 = 0, no column offset: j = mlow)
          const Fp y = (unit_twiddle && (q == 0 || (q == 1 && mlow == 0 && logW == 0))) ? x[i1] : x[i1] * w;
          x[i1] = x[i0] - y;
          x[i0] = x[i0] + y;
        }
      }
    }
#pragma unroll
    for (int m = 0; m < (1 << Q); m++) sm[ntt_slot(ebase + m * mstride)] = x[m];
  }
}

template <int MAXQ, int MINB>
__global__ void __launch_bounds__(NTT_THREADS, MINB) ntt_pass_kernel(const __grid_constant__ NttPassArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Fp* sm = reinterpret_cast<Fp*>(smem_raw);
  const int L = a.log_n, S = a.S, logW = a.logW, logG = a.logG;
  const uint32_t W = 1u << logW;
  const uint32_t tile_elems = 1u << (S + logW + logG);
  // tile id -> (hi0, lo0): global index = (hi0 + g) << (s0 + S) | t << s0 | (lo0 + l)
  const uint32_t lo_tiles = (1u << a.s0) >> logW;
  const uint32_t tile = blockIdx.x;
  const uint32_t hi0 = (tile / lo_tiles) << logG, lo0 = (tile % lo_tiles) << logW;
  Fp* const out = a.out + blockIdx.y * a.out_stride + blockIdx.z * a.out_stride2;
  const Fp* const in = a.first ? a.in + blockIdx.y * a.in_stride + blockIdx.z * a.in_stride2 : out;
  auto global_index = [&](uint32_t e) {
    const uint32_t l = e & (W - 1), t = (e >> logW) & ((1u << S) - 1), g = e >> (logW + S);
    return ((hi0 + g) << (a.s0 + S)) | (t << a.s0) | (lo0 + l);
  };
  // ---- load
  for (uint32_t e = threadIdx.x; e < tile_elems; e += NTT_THREADS) {
    const uint32_t idx = global_index(e);
    Fp v;
    if (a.first) {
      uint32_t src = bitrev(idx, L);
      if (src < a.n_in) {
        v = in[src];
        if (a.scale_in) v = v * a.scale_in[blockIdx.y * a.scale_stride + src];
      } else {
        v = Fp::zero();
      }
    } else {
      v = in[idx];
    }
    sm[ntt_slot(e)] = v;
  }
  __syncthreads();
  // ---- stages
  int r = 0;
  for (int i = 0; i < a.nrounds; i++) {
    const int q = a.q[i];
    if (MAXQ >= 3 && q == 3) ntt_round<MAXQ >= 3 ? 3 : 1>(sm, a, r, lo0, tile_elems);
    else if (MAXQ >= 2 && q == 2) ntt_round<MAXQ >= 2 ? 2 : 1>(sm, a, r, lo0, tile_elems);
    else ntt_round<1>(sm, a, r, lo0, tile_elems);
    r += q;
    __syncthreads();
  }
  // ---- store
  for (uint32_t e = threadIdx.x; e < tile_elems; e += NTT_THREADS) {
    Fp v = sm[ntt_slot(e)];
    if (a.last && a.scale_out) v = v * a.scale;
    out[global_index(e)] = v;
  }
}

}  // namespace

int32_t ntt_tables(zk_ctx* ctx, int log_n, NttTables** out) {
  auto it = ctx->ntt_tables.find(log_n);
  if (it == ctx->ntt_tables.end()) {
    NttTables t;
    t.log_n = log_n;
    uint32_t half = log_n ? (1u << (log_n - 1)) : 1;
    Fp omega = Fp::root_of_unity();
    for (int i = log_n; i < 32; i++) omega = omega.sqr();
    t.omega = omega;
    t.omega_inv = omega.inv();
    t.n_inv = Fp::from_u64(1ull << log_n).inv();
    ZK_CUDA(ctx, cudaMalloc((void**)&t.tw_fwd, (size_t)half * sizeof(Fp)));
    ZK_CUDA(ctx, cudaMalloc((void**)&t.tw_inv, (size_t)half * sizeof(Fp)));
    powers_kernel<<<(half + 255) / 256, 256, 0, ctx->stream>>>(t.omega, t.tw_fwd, half);
    powers_kernel<<<(half + 255) / 256, 256, 0, ctx->stream>>>(t.omega_inv, t.tw_inv, half);
    ctx->launches += 2;
    ZK_CUDA(ctx, cudaGetLastError());
    it = ctx->ntt_tables.emplace(log_n, t).first;
  }
  *out = &it->second;
  return ZK_OK;
}

// out[i] = sum_j in'[j] w^(ij), in' = in zero-padded to 2^log_n with optional coset scaling.
// in == out is allowed only when no pass reads what another tile writes, i.e. never for the
// first (bit-reversing) pass: callers pass distinct buffers or accept the internal temp copy.
int32_t ntt_run(zk_ctx* ctx, const Fp* in, uint32_t n_in, Fp* out, int log_n, const NttOptions& opt) {
  NttTables* T = nullptr;
  int32_t rc = ntt_tables(ctx, log_n, &T);
  if (rc) return rc;
  const uint32_t N = 1u << log_n;
  const Fp* src = in;
  if (in == out && (opt.batch > 1 || opt.batch2 > 1)) return set_error(ctx, ZK_E_INVALID, "ntt: batched transforms need distinct buffers");
  if (in == out) {  // bit-reversal gather cannot run in place
    rc = ensure_buf(ctx, ctx->ntt_tmp, (size_t)n_in * sizeof(Fp));
    if (rc) return rc;
    ZK_CUDA(ctx, cudaMemcpyAsync(ctx->ntt_tmp.ptr, in, (size_t)n_in * sizeof(Fp), cudaMemcpyDeviceToDevice,
                                 ctx->stream));
    src = (const Fp*)ctx->ntt_tmp.ptr;
  }
  KernelTimer timer(ctx, KC_NTT);
  if (log_n == 0) {
    ZK_CUDA(ctx, cudaMemcpyAsync(out, src, sizeof(Fp), cudaMemcpyDeviceToDevice, ctx->stream));
    return ZK_OK;
  }
  const int passes = (log_n + NTT_TILE_LOG - 1) / NTT_TILE_LOG;
  int s0 = 0;
  for (int pass = 0; pass < passes; pass++) {
    NttPassArgs a;
    memset(&a, 0, sizeof a);
    // stages balanced over the passes, the longer ones first
    const int S = log_n / passes + (pass < log_n % passes ? 1 : 0);
    const int tile_log = log_n < NTT_TILE_LOG ? log_n : NTT_TILE_LOG;
    const int logW = s0 < tile_log - S ? s0 : tile_log - S;
    const int logG = tile_log - S - logW;  // <= log_n - s0 - S: the first pass has all the high bits, later ones logG = 0
    a.in = s0 == 0 ? src : out;
    a.out = out;
    a.tw = opt.inverse ? T->tw_inv : T->tw_fwd;
    a.log_n = log_n;
    a.s0 = s0;
    a.S = S;
    a.logW = logW;
    a.logG = logG;
    for (int left = S; left > 0;) {  // rounds of NTT_ROUND_STAGES stages
      const int q = left >= NTT_ROUND_STAGES ? NTT_ROUND_STAGES : left;
      a.q[a.nrounds++] = q;
      left -= q;
    }
    a.n_in = n_in;
    a.first = s0 == 0;
    a.last = s0 + S == log_n;
    a.scale_in = opt.scale_in;
    a.in_stride = opt.in_stride;
    a.out_stride = opt.out_stride;
    a.scale_stride = opt.scale_stride;
    a.in_stride2 = opt.in_stride2;
    a.out_stride2 = opt.out_stride2;
    a.scale_out = opt.inverse;
    a.scale = T->n_inv;
    const uint32_t tile_elems = 1u << tile_log;
    const uint32_t tiles = N / tile_elems;
    const size_t smem = (size_t)(tile_elems + (tile_elems >> 3) + 1) * sizeof(Fp);
    const dim3 grid(tiles, (unsigned)opt.batch, (unsigned)opt.batch2);
    ntt_pass_kernel<NTT_ROUND_STAGES, NTT_MIN_BLOCKS><<<grid, NTT_THREADS, smem, ctx->stream>>>(a);
    ctx->launches++;
    s0 += S;
  }
  ZK_CUDA(ctx, cudaGetLastError());
  return ZK_OK;
}

}  // namespace zkodst

using namespace zkodst;

// data: 2^log_n Montgomery-form Fp, transformed in place (host or device memory).
// inverse = 0: coefficients -> evaluations on <omega>; 1: evaluations -> coefficients.
extern "C" int32_t zk_ntt_fp(zk_ctx* ctx, void* data, int32_t log_n, int32_t inverse, int32_t on_device) {
  if (!ctx || !data || log_n < 0 || log_n > 28) return ZK_E_INVALID;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t bytes = sizeof(Fp) << log_n;
  Fp* d = (Fp*)data;
  if (!on_device) {
    int32_t rc = ensure_buf(ctx, ctx->scratch_a, bytes);
    if (rc) return rc;
    ZK_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_a.ptr, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
    d = (Fp*)ctx->scratch_a.ptr;
  }
  NttOptions opt;
  opt.inverse = inverse != 0;
  int32_t rc = ntt_run(ctx, d, 1u << log_n, d, log_n, opt);
  if (rc) return rc;
  if (!on_device) {
    ZK_CUDA(ctx, cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return zk_ctx_synchronize(ctx);
  }
  return ZK_OK;
}
