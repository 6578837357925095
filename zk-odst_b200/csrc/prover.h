// Internal interfaces between the kernels' host drivers (product code).
#pragma once
#include "ec.cuh"
#include "zk_ctx.h"

namespace zkodst {

struct NttOptions {
  bool inverse = false;        // use omega^-1 and scale the output by 1/N
  int coset_in = 0;            // multiply input i by coset_in_pow[(i % 3) - 1] before the transform
  int coset_out = 0;           // multiply output i by coset_out_pow[(i % 3) - 1] after it
  Fp coset_in_pow[2] = {Fp::zero(), Fp::zero()};
  Fp coset_out_pow[2] = {Fp::zero(), Fp::zero()};
};

// ntt.cu
int32_t ntt_tables(zk_ctx* ctx, int log_n, NttTables** out);
int32_t ntt_run(zk_ctx* ctx, const Fp* in, uint32_t n_in, Fp* out, int log_n, const NttOptions& opt);

// msm.cu
int msm_window_bits(uint64_t n);
int32_t msm_run(zk_ctx* ctx, const Fp* d_scalars, const Affine* d_bases, uint64_t n, XYZZ* result,
                const Fp* d_extra = nullptr);

}  // namespace zkodst
