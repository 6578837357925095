// Fixed-base MSM with precomputed window tables (msm_fixed.cu).  Product code.
#pragma once
#include "ec.cuh"
#include "zk_ctx.h"

namespace zkodst {

struct FixedBase {
  int c = 0, nwin = 0;
  uint64_t npoints = 0;
  Affine* table = nullptr;  // [nwin][npoints]: table[w][i] = 2^(c w) * base_i
};

int fixed_window_bits(uint64_t npoints);
int32_t fixed_base_build(zk_ctx* ctx, const Affine* d_bases, uint64_t npoints, FixedBase* out);
void fixed_base_free(FixedBase& fb);

// result = sum_{t < count} scalars[t] * base_t + sum_e extra[e] * base_{extra_index[e]}.
// side_bit_mask != 0 restricts the first sum to indices t with ((t & mask) != 0) == side_select.
int32_t msm_fixed(zk_ctx* ctx, const FixedBase& fb, const Fp* d_scalars, uint64_t count, const Fp* extra_host,
                  const uint32_t* extra_index_host, int n_extra, XYZZ* result, uint32_t side_bit_mask = 0,
                  int side_select = 0);

}  // namespace zkodst
