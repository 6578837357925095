// K4/K5 — number-theoretic transforms over Fp.
//
// Replaces halo2_proofs 0.3.0 `best_fft` as reached through `EvaluationDomain::
// {lagrange_to_coeff, coeff_to_extended, extended_to_coeff}` inside `create_proof`
// (blake2f-circuit/benches/blake2f.rs:125).  Natural order in, natural order out.
//
// Decimation-in-time radix-2 butterflies, executed as shared-memory passes of up to 8 stages:
// a pass loads a tile of 2^S elements x W adjacent columns (W * 32 B contiguous per row), runs
// S stages out of shared memory and stores the tile back, so a 2^21-point transform touches
// HBM 3 times instead of 21.  The first pass fuses the bit-reversal gather, zero padding and
// an optional per-element input scaling (coset evaluation: c^i from a table); the last pass fuses the
// 1/N scaling.  Several transforms of one size run as one batch of launches (blockIdx.y).  Twiddles come from a per-domain table of N/2 powers.
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

struct NttPassArgs {
  const Fp* in;
  Fp* out;
  const Fp* tw;      // omega^i, i < N/2
  const Fp* scale_in;  // first pass: optional per-element multiplier
  size_t in_stride, out_stride, scale_stride;  // per transform of the batch (blockIdx.y)
  int log_n;         // L
  int s0;            // stages already done
  int S;             // stages in this pass
  int logW;          // log2 of adjacent columns per tile (0 for the first pass)
  uint32_t n_in;     // first pass: input length (zero padded above)
  int first, last;
  int scale_out;     // last pass: multiply by `scale` (1/N of the inverse transform)
  Fp scale;
};

__global__ void __launch_bounds__(256, 5) ntt_pass_kernel(NttPassArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Fp* sm = reinterpret_cast<Fp*>(smem_raw);
  const int L = a.log_n, S = a.S, logW = a.logW, W = 1 << logW;
  const uint32_t tile_elems = 1u << (S + logW);
  // tile id -> (hi, lo0): global index = hi << (s0 + S) | t << s0 | lo,  lo in [lo0, lo0 + W)
  const uint32_t lo_groups = (1u << a.s0) >> logW;  // number of W-wide lo groups (>= 1)
  const uint32_t tile = blockIdx.x;
  const uint32_t hi = tile / lo_groups, lo0 = (tile % lo_groups) << logW;
  const Fp* const in = a.in + (a.first ? blockIdx.y * a.in_stride : blockIdx.y * a.out_stride);
  Fp* const out = a.out + blockIdx.y * a.out_stride;
  // ---- load
  for (uint32_t e = threadIdx.x; e < tile_elems; e += blockDim.x) {
    uint32_t t = e >> logW, l = e & (W - 1);
    uint32_t idx = (hi << (a.s0 + S)) | (t << a.s0) | (lo0 + l);
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
    sm[e] = v;
  }
  __syncthreads();
  // ---- stages: one butterfly per thread (blockDim = tile_elems / 2); the twiddle of the next stage is
  // requested before the barrier so that its latency overlaps the wait
  {
    const uint32_t b = threadIdx.x;
    const uint32_t l = b & (W - 1), tb = b >> logW;  // tb in [0, 2^(S-1))
    auto twiddle_index = [&](int r) {
      const uint32_t tlo = tb & ((1u << r) - 1);
      const uint32_t j = (tlo << a.s0) | (lo0 + l);  // index within the half
      return (size_t)j << (L - (a.s0 + r + 1));
    };
    Fp w = a.tw[twiddle_index(0)];
    for (int r = 0; r < S; r++) {
      const uint32_t tlo = tb & ((1u << r) - 1), thi = tb >> r;
      const uint32_t t0 = (thi << (r + 1)) | tlo, t1 = t0 | (1u << r);
      const Fp x = sm[(t0 << logW) | l], y = sm[(t1 << logW) | l] * w;
      sm[(t0 << logW) | l] = x + y;
      sm[(t1 << logW) | l] = x - y;
      if (r + 1 < S) w = a.tw[twiddle_index(r + 1)];
      __syncthreads();
    }
  }
  // ---- store
  for (uint32_t e = threadIdx.x; e < tile_elems; e += blockDim.x) {
    uint32_t t = e >> logW, l = e & (W - 1);
    uint32_t idx = (hi << (a.s0 + S)) | (t << a.s0) | (lo0 + l);
    Fp v = sm[e];
    if (a.last && a.scale_out) v = v * a.scale;
    out[idx] = v;
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
  if (in == out && opt.batch > 1) return set_error(ctx, ZK_E_INVALID, "ntt: batched transforms need distinct buffers");
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
  const int MAXS = 8;
  int s0 = 0;
  while (s0 < log_n) {
    NttPassArgs a;
    memset(&a, 0, sizeof a);
    int S = log_n - s0 < MAXS ? log_n - s0 : MAXS;
    // balance the remaining passes so the last one is not tiny
    int remaining = log_n - s0, passes = (remaining + MAXS - 1) / MAXS;
    S = (remaining + passes - 1) / passes;
    int logW = s0 == 0 ? 0 : (s0 < 2 ? s0 : 2);
    if (S + logW > 9) logW = 9 - S;  // one butterfly per thread, at most 256 threads
    a.in = s0 == 0 ? src : out;
    a.out = out;
    a.tw = opt.inverse ? T->tw_inv : T->tw_fwd;
    a.log_n = log_n;
    a.s0 = s0;
    a.S = S;
    a.logW = logW;
    a.n_in = n_in;
    a.first = s0 == 0;
    a.last = s0 + S == log_n;
    a.scale_in = opt.scale_in;
    a.in_stride = opt.in_stride;
    a.out_stride = opt.out_stride;
    a.scale_stride = opt.scale_stride;
    a.scale_out = opt.inverse;
    a.scale = T->n_inv;
    uint32_t tile_elems = 1u << (S + logW);
    uint32_t tiles = N / tile_elems;
    size_t smem = (size_t)tile_elems * sizeof(Fp);
    ntt_pass_kernel<<<dim3(tiles, (unsigned)opt.batch), tile_elems / 2, smem, ctx->stream>>>(a);
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
