// Fixed-base MSM with precomputed window tables (msm_fixed.cu).  Product code.
#pragma once
#include "ec.cuh"
#include "zk_ctx.h"

namespace zkodst {

struct FixedBase {
  int c = 0, nwin = 0;
  uint64_t npoints = 0;     // local table width = nmain + nextra
  uint64_t lo = 0;          // global index of the first main point held here
  uint64_t nmain = 0;       // main points held here: global indices [lo, lo + nmain)
  uint64_t total_main = 0;  // main points of the whole base; extras have global index >= total_main
  uint64_t nextra = 0;      // extra points (W, U) held here: all of them on rank 0, none elsewhere
  bool split = false;       // built over this rank's range of a multi-GPU group: results are partial sums
  Affine* table = nullptr;  // [nwin][npoints]: table[w][i] = 2^(c w) * local point i
};

int fixed_window_bits(uint64_t npoints);
// d_bases: total_main main points followed by n_extra extra points (global indexing).  With the
// context in a multi-GPU group (dist.cu) only this rank's contiguous range of the main points is
// tabulated (and the extras on rank 0).
int32_t fixed_base_build(zk_ctx* ctx, const Affine* d_bases, uint64_t total_main, uint64_t n_extra,
                         FixedBase* out, int force_c = 0);
// ipa_fold.cu: the IPA generators after r folds (needs a c = 8 table over all points), and a window
// table built in caller-owned storage in a single launch
size_t ipa_fold_workspace_bytes(uint64_t len, int world);
// fb8 may cover only this rank's range of the points (MSM split): the partial sums are exchanged with
// one all-gather and added, so every rank ends with all folded generators
int32_t ipa_fold_generators(zk_ctx* ctx, const FixedBase& fb8, const Fp* u, int r, void* workspace, Affine* out);
int32_t fixed_base_build_inplace(zk_ctx* ctx, uint64_t total_main, uint64_t n_extra, int c, Affine* storage,
                                 void* tmp, FixedBase* out);
void fixed_base_free(FixedBase& fb);

constexpr int MSM_MAX_BATCH = 16;
struct MsmJob {  // one MSM of a batch over the same base
  const Fp* scalars = nullptr;  // device, `count` entries (jobs may share the vector)
  Fp extra[4];                  // host scalars of up to 4 extra terms (blinds on W, IPA terms on U)
  uint32_t extra_index[4] = {};
  int n_extra = 0;
  uint32_t side_mask = 0;       // != 0: keep index t only if ((t & mask) != 0) == side_select
  int side_select = 0;
};
// results[m] = sum_{t < count} jobs[m].scalars[t] * base_t + extras, for all jobs in one pipeline
int32_t msm_fixed_batch(zk_ctx* ctx, const FixedBase& fb, const MsmJob* jobs, int nb, uint64_t count,
                        XYZZ* results);

// result = sum_{t < count} scalars[t] * base_t + sum_e extra[e] * base_{extra_index[e]}.
// side_bit_mask != 0 restricts the first sum to indices t with ((t & mask) != 0) == side_select.
int32_t msm_fixed(zk_ctx* ctx, const FixedBase& fb, const Fp* d_scalars, uint64_t count, const Fp* extra_host,
                  const uint32_t* extra_index_host, int n_extra, XYZZ* result, uint32_t side_bit_mask = 0,
                  int side_select = 0);

}  // namespace zkodst
