// Integer-pipe micro-benchmark: measured peak of the 32x32-bit multiply-add forms that
// Montgomery arithmetic is built from.  Its result is the compute roofline for the MSM / NTT
// kernels ("the builder must add an IMAD micro-benchmark").
//   mode 0: mad.lo.u32                    (IMAD,      1 result word per instruction)
//   mode 1: mad.wide.u32                  (IMAD.WIDE, full 64-bit product + 64-bit add)
//   mode 2: mad.lo.cc.u32 + madc.hi.cc.u32 (carry-chained pair = one 32x32->64 MAC)
// Reports instructions/s; a "MAC" (32x32->64 multiply-accumulate) is 1 instruction in mode 1
// and 2 instructions in mode 2.
#include "ec.cuh"
#include "zk_ctx.h"

namespace zkodst {
namespace {

template <int MODE>
__global__ void __launch_bounds__(256) imad_kernel(uint32_t* out, uint32_t iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
  uint32_t r0 = 1, r1 = 2, r2 = 3, r3 = 4, r4 = 5, r5 = 6, r6 = 7, r7 = 8;
  uint64_t w0 = 1, w1 = 2, w2 = 3, w3 = 4, w4 = 5, w5 = 6, w6 = 7, w7 = 8;
  for (uint32_t i = 0; i < iters; i++) {
    if (MODE == 0) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        asm volatile("mad.lo.u32 %0, %8, %9, %0;\n\tmad.lo.u32 %1, %8, %9, %1;\n\t"
                     "mad.lo.u32 %2, %8, %9, %2;\n\tmad.lo.u32 %3, %8, %9, %3;\n\t"
                     "mad.lo.u32 %4, %8, %9, %4;\n\tmad.lo.u32 %5, %8, %9, %5;\n\t"
                     "mad.lo.u32 %6, %8, %9, %6;\n\tmad.lo.u32 %7, %8, %9, %7;"
                     : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7)
                     : "r"(a), "r"(b));
      }
    } else if (MODE == 1) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        asm volatile("mad.wide.u32 %0, %8, %9, %0;\n\tmad.wide.u32 %1, %8, %9, %1;\n\t"
                     "mad.wide.u32 %2, %8, %9, %2;\n\tmad.wide.u32 %3, %8, %9, %3;\n\t"
                     "mad.wide.u32 %4, %8, %9, %4;\n\tmad.wide.u32 %5, %8, %9, %5;\n\t"
                     "mad.wide.u32 %6, %8, %9, %6;\n\tmad.wide.u32 %7, %8, %9, %7;"
                     : "+l"(w0), "+l"(w1), "+l"(w2), "+l"(w3), "+l"(w4), "+l"(w5), "+l"(w6), "+l"(w7)
                     : "r"(a), "r"(b));
      }
    } else {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                     "madc.lo.cc.u32 %2, %8, %9, %2;\n\tmadc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                     "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\t"
                     "madc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %9, %7;"
                     : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7)
                     : "r"(a), "r"(b));
      }
    }
  }
  uint32_t acc = r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7;
  uint64_t wacc = w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ (uint32_t)wacc ^ (uint32_t)(wacc >> 32);
}

// mode 9: fma.rn.f64 (DFMA), 8 independent chains; mode 10: the same DFMA stream interleaved one to one with
// mad.wide.u32 — do the FP64 and the integer multiply pipes issue side by side?  (Counts DFMA + IMAD.)
template <int MIX>
__global__ void __launch_bounds__(256) dfma_kernel(uint32_t* out, uint32_t iters, uint32_t seed) {
  double a = 1.0 + 1e-9 * (seed + threadIdx.x), b = 1.0 - 1e-9 * blockIdx.x;
  double d0 = 1, d1 = 2, d2 = 3, d3 = 4, d4 = 5, d5 = 6, d6 = 7, d7 = 8;
  uint32_t ia = seed + threadIdx.x, ib = seed * 3 + blockIdx.x;
  uint64_t w0 = 1, w1 = 2, w2 = 3, w3 = 4, w4 = 5, w5 = 6, w6 = 7, w7 = 8;
  for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (MIX) {
        asm volatile("fma.rn.f64 %0, %16, %17, %0;\n\tmad.wide.u32 %8, %18, %19, %8;\n\t"
                     "fma.rn.f64 %1, %16, %17, %1;\n\tmad.wide.u32 %9, %18, %19, %9;\n\t"
                     "fma.rn.f64 %2, %16, %17, %2;\n\tmad.wide.u32 %10, %18, %19, %10;\n\t"
                     "fma.rn.f64 %3, %16, %17, %3;\n\tmad.wide.u32 %11, %18, %19, %11;\n\t"
                     "fma.rn.f64 %4, %16, %17, %4;\n\tmad.wide.u32 %12, %18, %19, %12;\n\t"
                     "fma.rn.f64 %5, %16, %17, %5;\n\tmad.wide.u32 %13, %18, %19, %13;\n\t"
                     "fma.rn.f64 %6, %16, %17, %6;\n\tmad.wide.u32 %14, %18, %19, %14;\n\t"
                     "fma.rn.f64 %7, %16, %17, %7;\n\tmad.wide.u32 %15, %18, %19, %15;"
                     : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5), "+d"(d6), "+d"(d7), "+l"(w0),
                       "+l"(w1), "+l"(w2), "+l"(w3), "+l"(w4), "+l"(w5), "+l"(w6), "+l"(w7)
                     : "d"(a), "d"(b), "r"(ia), "r"(ib));
      } else {
        asm volatile("fma.rn.f64 %0, %8, %9, %0;\n\tfma.rn.f64 %1, %8, %9, %1;\n\t"
                     "fma.rn.f64 %2, %8, %9, %2;\n\tfma.rn.f64 %3, %8, %9, %3;\n\t"
                     "fma.rn.f64 %4, %8, %9, %4;\n\tfma.rn.f64 %5, %8, %9, %5;\n\t"
                     "fma.rn.f64 %6, %8, %9, %6;\n\tfma.rn.f64 %7, %8, %9, %7;"
                     : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5), "+d"(d6), "+d"(d7)
                     : "d"(a), "d"(b));
      }
    }
  }
  const double acc = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
  const uint64_t wacc = w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7;
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)__double_as_longlong(acc) ^ (uint32_t)wacc ^ (uint32_t)(wacc >> 32);
}

// mode 3: Fq Montgomery multiplications (4 independent chains per thread);
// mode 4: XYZZ mixed additions (the MSM inner loop)
__global__ void __launch_bounds__(128) fieldmul_kernel(uint64_t* out, uint32_t iters, uint64_t seed) {
  Fq a{{seed + threadIdx.x, 2, 3, 4}}, b{{5, seed ^ blockIdx.x, 7, 8}}, c{{9, 10, seed, 12}}, d{{13, 14, 15, 1}};
  Fq m{{seed | 1, 77, 99, 1234}};
  for (uint32_t i = 0; i < iters; i++) {
    a = a * m;
    b = b * m;
    c = c * m;
    d = d * m;
  }
  Fq r = a + b + c + d;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r.l[0] ^ r.l[3];
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) madd_kernel(uint64_t* out, uint32_t iters, uint64_t seed) {
  Affine p{Fq{{seed + threadIdx.x, 2, 3, 4}}, Fq{{5, seed ^ blockIdx.x, 7, 8}}};
  XYZZ acc = XYZZ::from_affine(Affine{Fq{{11, 12, 13, 14}}, Fq{{1, 2, 3, 5}}});
  for (uint32_t i = 0; i < iters; i++) {
    acc = acc.add_affine(p);
    p.x.l[0] += i;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x.l[0] ^ acc.zzz.l[3];
}

// mode 5: the same mixed addition with the field product as a real call (small loop body: the
// instruction-cache footprint drops from ~40 KB to a few KB at the price of call overhead)
__device__ __noinline__ Fq fq_mul_call(Fq a, Fq b) { return a * b; }
__device__ __forceinline__ XYZZ add_affine_calls(const XYZZ& a, const Affine& p) {
  Fq u2 = fq_mul_call(p.x, a.zz);
  Fq s2 = fq_mul_call(p.y, a.zzz);
  Fq pp_ = u2 - a.x;
  Fq r = s2 - a.y;
  if (pp_.is_zero()) return XYZZ::add_affine_same_x(p, r);
  Fq pp = fq_mul_call(pp_, pp_);
  Fq ppp = fq_mul_call(pp_, pp);
  Fq q = fq_mul_call(a.x, pp);
  Fq x3 = fq_mul_call(r, r) - ppp - q.dbl();
  Fq y3 = fq_mul_call(r, q - x3) - fq_mul_call(a.y, ppp);
  return XYZZ{x3, y3, fq_mul_call(a.zz, pp), fq_mul_call(a.zzz, ppp)};
}
__global__ void __launch_bounds__(128) madd_calls_kernel(uint64_t* out, uint32_t iters, uint64_t seed) {
  Affine p{Fq{{seed + threadIdx.x, 2, 3, 4}}, Fq{{5, seed ^ blockIdx.x, 7, 8}}};
  XYZZ acc = XYZZ::from_affine(Affine{Fq{{11, 12, 13, 14}}, Fq{{1, 2, 3, 5}}});
  for (uint32_t i = 0; i < iters; i++) {
    acc = add_affine_calls(acc, p);
    p.x.l[0] += i;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x.l[0] ^ acc.zzz.l[3];
}

}  // namespace
}  // namespace zkodst

using namespace zkodst;

extern "C" int32_t zk_bench_int_pipe(zk_ctx* ctx, int32_t mode, uint32_t iters,
                                     double* instr_per_sec) {
  if (!ctx || !instr_per_sec || mode < 0 || mode > 10) return ZK_E_INVALID;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool field_mode = mode >= 3 && mode <= 8;
  const int blocks = ctx->sm_count * (field_mode ? 120 : 8), threads = field_mode ? 128 : 256;
  int32_t rc = ensure_buf(ctx, ctx->scratch_digests, (size_t)blocks * threads * 8);
  if (rc) return rc;
  uint32_t* out = (uint32_t*)ctx->scratch_digests.ptr;
  cudaEvent_t e0, e1;
  ZK_CUDA(ctx, cudaEventCreate(&e0));
  ZK_CUDA(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    ZK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (mode == 0) imad_kernel<0><<<blocks, threads, 0, ctx->stream>>>(out, iters, 12345u + rep);
    if (mode == 1) imad_kernel<1><<<blocks, threads, 0, ctx->stream>>>(out, iters, 12345u + rep);
    if (mode == 2) imad_kernel<2><<<blocks, threads, 0, ctx->stream>>>(out, iters, 12345u + rep);
    if (mode == 3) fieldmul_kernel<<<blocks, threads, 0, ctx->stream>>>((uint64_t*)out, iters, 12345u + rep);
    if (mode == 4) madd_kernel<4><<<blocks, threads, 0, ctx->stream>>>((uint64_t*)out, iters, 12345u + rep);
    if (mode == 6) madd_kernel<5><<<blocks, threads, 0, ctx->stream>>>((uint64_t*)out, iters, 12345u + rep);
    if (mode == 7) madd_kernel<6><<<blocks, threads, 0, ctx->stream>>>((uint64_t*)out, iters, 12345u + rep);
    if (mode == 8) madd_kernel<8><<<blocks, threads, 0, ctx->stream>>>((uint64_t*)out, iters, 12345u + rep);
    if (mode == 9) dfma_kernel<0><<<blocks, threads, 0, ctx->stream>>>(out, iters, 12345u + rep);
    if (mode == 10) dfma_kernel<1><<<blocks, threads, 0, ctx->stream>>>(out, iters, 12345u + rep);
    if (mode == 5) madd_calls_kernel<<<blocks, threads, 0, ctx->stream>>>((uint64_t*)out, iters, 12345u + rep);
    ctx->launches++;
    ZK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    ZK_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    ZK_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  // field mults / madds / instructions per loop iteration
  double per_iter = mode == 3 ? 4.0 : (field_mode ? 1.0 : (mode == 10 ? 128.0 : 64.0));
  double instr = (double)blocks * threads * (double)iters * per_iter;
  *instr_per_sec = instr / (best * 1e-3);
  return ZK_OK;
}
