// Internal interfaces between the kernels' host drivers (product code).
#pragma once
#include "ec.cuh"
#include "zk_ctx.h"

namespace zkodst {

struct NttOptions {
  bool inverse = false;         // use omega^-1 and scale the output by 1/N
  // Batched transforms: `batch` independent transforms in one sequence of launches (blockIdx.y);
  // transform b reads in + b * in_stride (0: all read the same vector), writes out + b * out_stride.
  int batch = 1;
  size_t in_stride = 0, out_stride = 0;
  // Second batch level (blockIdx.z), e.g. columns x cosets: transform (b, c) reads
  // in + b * in_stride + c * in_stride2 and writes out + b * out_stride + c * out_stride2.
  int batch2 = 1;
  size_t in_stride2 = 0, out_stride2 = 0;
  // Optional per-element input scaling: input element i of transform b is multiplied by
  // scale_in[b * scale_stride + i] while it is loaded (coset evaluation: scale_in[i] = c^i).
  const Fp* scale_in = nullptr;
  size_t scale_stride = 0;
};

// ntt.cu
int32_t ntt_tables(zk_ctx* ctx, int log_n, NttTables** out);
int32_t ntt_run(zk_ctx* ctx, const Fp* in, uint32_t n_in, Fp* out, int log_n, const NttOptions& opt);

// msm.cu
int msm_window_bits(uint64_t n);
int32_t msm_run(zk_ctx* ctx, const Fp* d_scalars, const Affine* d_bases, uint64_t n, XYZZ* result,
                const Fp* d_extra = nullptr);

}  // namespace zkodst
