// verify_proof for the BLAKE2f Table16 circuit (Pasta / IPA, SingleVerifier strategy).
//
// Replaces halo2_proofs 0.3.0 `plonk::verify_proof(&params, pk.get_vk(), SingleVerifier::new(&params),
// &[&[]], &mut Blake2bRead<_, _, Challenge255<_>>)` as the reference calls it
// (blake2f-circuit/benches/blake2f.rs:138-144).  The transcript, the evaluation of the constraint
// system at the challenge x (gates.cuh) and the multiopen bookkeeping are byte-serial host work on a
// few hundred field elements; the one large object, the final multi-scalar multiplication
//     [q', S, L_j, R_j, commitments] + [-v + s] . g + u_scalar U + w_scalar W  ==  identity,
// runs on the device: the size-n part over the fixed-base window tables (msm_fixed.cu, split across
// GPUs when the context joined a group), the ~80 proof points through the variable-base MSM (msm.cu).
#include <algorithm>
#include <set>

#include "gates.cuh"
#include "polyops.cuh"
#include "prover_state.h"
#include "transcript.h"
#include "xorshift_jump.h"

namespace zkodst {

int32_t msm_run(zk_ctx* ctx, const Fp* d_scalars, const Affine* d_bases, uint64_t n, XYZZ* result,
                const Fp* d_extra);

namespace {

struct Term {
  Fp scalar;
  Affine point;
};
struct CommitmentAcc {  // poly/commitment/msm.rs `MSM` restricted to explicit terms
  std::vector<Term> terms;
  void scale(const Fp& f) {
    for (auto& t : terms) t.scalar = t.scalar * f;
  }
};

struct Horner {
  Fp h, y;
  void fold(const Fp& v) { h = h * y + v; }
};

// s_m = neg_c * prod_b us[k - 1 - b]^(bit b of m)   (commitment::Guard::compute_s), plus a constant
// term on s_0
// (batch verification: accumulate != 0 adds to s instead of overwriting; neg_c and constant then carry the proof's
// random weight)
__global__ void compute_s_kernel(Fp* __restrict__ s, uint64_t n, int k, const Fp* __restrict__ us, Fp neg_c,
                                 Fp constant, int accumulate) {
  uint64_t m = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (m >= n) return;
  Fp acc = neg_c;
  for (int b = 0; b < k; b++)
    if ((m >> b) & 1) acc = acc * us[k - 1 - b];
  if (m == 0) acc = acc + constant;
  s[m] = accumulate ? s[m] + acc : acc;
}

// What the transcript phase of one proof leaves for the final check:
//   sum_t terms[t].scalar * terms[t].point + sum_m s_m g_m + u_scalar U + w_scalar W == identity,
//   s_m = neg_c * prod_b us[k-1-b]^(bit b of m) + [m = 0] s0_constant
struct Guard {
  std::vector<Term> terms;
  std::vector<Fp> us;
  Fp neg_c, s0_constant, u_scalar, w_scalar;
};

// host part: transcript, constraint evaluation at x, multiopen bookkeeping (throws on malformed proofs)
int32_t verify_transcript(zk_ctx* ctx, const uint8_t* proof, size_t proof_len, Guard& out) {
  ProverState* S = prover_state(ctx);
  if (!S->has_params || !S->has_keys) return set_error(ctx, ZK_E_STATE, "verify_proof before params/keygen");
  const DeviceKeys& K = S->keys;
  const uint64_t n = K.n;
  const int k = K.k;
  NttTables* TN = nullptr;
  int32_t rc = ntt_tables(ctx, k, &TN);
  if (rc) return rc;

  TranscriptReader tr(proof, proof_len);
  tr.common_scalar(K.transcript_repr);
  Affine advice_c[NUM_ADVICE_COLUMNS];
  for (auto& c : advice_c) c = tr.read_point();
  const Fp theta = tr.squeeze_challenge();
  const Affine pin_c = tr.read_point(), ptab_c = tr.read_point();
  const Fp beta = tr.squeeze_challenge(), gamma = tr.squeeze_challenge();
  Affine z_c[NUM_SETS];
  for (auto& c : z_c) c = tr.read_point();
  const Affine zl_c = tr.read_point();
  const Affine random_c = tr.read_point();
  const Fp y = tr.squeeze_challenge();
  Affine h_c[3];
  for (auto& c : h_c) c = tr.read_point();
  const Fp x = tr.squeeze_challenge();
  Fp advice_e[24], fixed_e[NUM_FIXED], sigma_e[NUM_PERM];
  for (auto& e : advice_e) e = tr.read_scalar();
  for (auto& e : fixed_e) e = tr.read_scalar();
  const Fp random_e = tr.read_scalar();
  for (auto& e : sigma_e) e = tr.read_scalar();
  Fp z_e[NUM_SETS], z_next_e[NUM_SETS], z_last_e[NUM_SETS];
  for (int s = 0; s < NUM_SETS; s++) {
    z_e[s] = tr.read_scalar();
    z_next_e[s] = tr.read_scalar();
    if (s + 1 != NUM_SETS) z_last_e[s] = tr.read_scalar();
  }
  const Fp zl_e = tr.read_scalar(), zl_next_e = tr.read_scalar(), pin_e = tr.read_scalar(),
           pin_prev_e = tr.read_scalar(), ptab_e = tr.read_scalar();

  // ---- expected h(x) ------------------------------------------------------------------------------------
  const Fp one = Fp::one();
  const Fp xn = x.pow_u64(n);
  auto rotate = [&](const Fp& v, int rot) {
    return rot >= 0 ? v * TN->omega.pow_u64((uint64_t)rot) : v * TN->omega_inv.pow_u64((uint64_t)(-rot));
  };
  // l_i(x) = (omega^i / n) (x^n - 1) / (x - omega^i) for i = -(BLINDING + 1) .. 0
  Fp l_evals[BLINDING + 2];
  {
    const Fp common = (xn - one) * TN->n_inv;
    for (int j = 0; j < BLINDING + 2; j++) {
      const Fp w = rotate(one, j - (BLINDING + 1));
      l_evals[j] = common * w * (x - w).inv();
    }
  }
  const Fp l_last = l_evals[0], l_0 = l_evals[BLINDING + 1];
  Fp l_blind = Fp::zero();
  for (int j = 1; j <= BLINDING; j++) l_blind = l_blind + l_evals[j];
  const Fp active = one - (l_last + l_blind);
  auto advice_at = [&](int col, int rot) -> const Fp& {
    for (int i = 0; i < 24; i++)
      if (ADVICE_QUERIES[i][0] == col && ADVICE_QUERIES[i][1] == rot) return advice_e[i];
    throw std::runtime_error("advice query missing");
  };
  GateConsts kc;
  for (int i = 0; i < 4; i++) kc.small[i] = Fp::from_u64(i);
  kc.pow2[0] = one;
  for (int e = 1; e < 127; e++) kc.pow2[e] = kc.pow2[e - 1].dbl();
  GateCells v;
  {
    const int* A = A_NUMBER_COLUMN;
    v.a0c = advice_at(A[0], 0), v.a0n = advice_at(A[0], 1);
    v.a1p = advice_at(A[1], -1), v.a1c = advice_at(A[1], 0), v.a1n = advice_at(A[1], 1);
    v.a2p = advice_at(A[2], -1), v.a2c = advice_at(A[2], 0), v.a2n = advice_at(A[2], 1);
    v.a3p = advice_at(A[3], -1), v.a3c = advice_at(A[3], 0), v.a3n = advice_at(A[3], 1);
    v.a4p = advice_at(A[4], -1), v.a4c = advice_at(A[4], 0), v.a4n = advice_at(A[4], 1);
    v.a5p = advice_at(A[5], -1), v.a5c = advice_at(A[5], 0), v.a5n = advice_at(A[5], 1);
    v.a6p = advice_at(A[6], -1), v.a6c = advice_at(A[6], 0);
    v.a7p = advice_at(A[7], -1), v.a7c = advice_at(A[7], 0);
    v.a8p = advice_at(A[8], -1), v.a8c = advice_at(A[8], 0);
    v.a9c = advice_at(A[9], 0);
  }
  Fp sel[NUM_SELECTORS];
  for (int s = 0; s < NUM_SELECTORS; s++)
    sel[s] = selector_expr(fixed_e[K.selectors[s].fixed_col], K.selectors[s].root, K.selectors[s].len, kc.small);
  Horner H{Fp::zero(), y};
  fold_gates(H, v, sel, kc, fixed_e[FIXED_CONSTANTS]);
  // permutation argument (columns in enable_equality order: a1,a2 | a3,a4 | a5,a6 | a7,a8, all cur)
  H.fold(l_0 * (one - z_e[0]));
  H.fold(l_last * (z_e[NUM_SETS - 1] * z_e[NUM_SETS - 1] - z_e[NUM_SETS - 1]));
  for (int s = 1; s < NUM_SETS; s++) H.fold((z_e[s] - z_last_e[s - 1]) * l_0);
  {
    Fp cur = beta * x;
    for (int s = 0; s < NUM_SETS; s++) {
      Fp left = z_next_e[s], right = z_e[s];
      for (int j = 0; j < 2; j++) {
        const int ci = 2 * s + j;
        const Fp& val = advice_at(PERM_COLUMNS[ci], 0);
        left = left * (val + beta * sigma_e[ci] + gamma);
        right = right * (val + cur + gamma);
        cur = cur * Fp::delta();
      }
      H.fold((left - right) * active);
    }
  }
  // lookup argument
  {
    const Fp cin = (v.a0c * theta + v.a1c) * theta + v.a2c;
    const Fp ctab = (fixed_e[0] * theta + fixed_e[1]) * theta + fixed_e[2];
    H.fold(l_0 * (one - zl_e));
    H.fold(l_last * (zl_e * zl_e - zl_e));
    H.fold((zl_next_e * (pin_e + beta) * (ptab_e + gamma) - zl_e * (cin + beta) * (ctab + gamma)) * active);
    H.fold(l_0 * (pin_e - ptab_e));
    H.fold((pin_e - ptab_e) * (pin_e - pin_prev_e) * active);
  }
  const Fp expected_h = H.h * (xn - one).inv();

  // ---- queries (multiopen) --------------------------------------------------------------------------------
  struct Query {
    int commitment;
    Fp point, eval;
  };
  std::vector<CommitmentAcc> commitments;
  std::vector<Query> queries;
  auto single = [&](const Affine& p) {
    CommitmentAcc c;
    c.terms.push_back(Term{one, p});
    commitments.push_back(c);
    return (int)commitments.size() - 1;
  };
  const Fp x_next = rotate(x, 1), x_prev = rotate(x, -1), x_last = rotate(x, -(BLINDING + 1));
  auto point_of = [&](int rot) { return rot == 0 ? x : (rot == 1 ? x_next : x_prev); };
  int advice_id[NUM_ADVICE_COLUMNS];
  for (int c = 0; c < NUM_ADVICE_COLUMNS; c++) advice_id[c] = -1;
  for (int i = 0; i < 24; i++) {
    const int col = ADVICE_QUERIES[i][0];
    if (advice_id[col] < 0) advice_id[col] = single(advice_c[col]);
    queries.push_back(Query{advice_id[col], point_of(ADVICE_QUERIES[i][1]), advice_e[i]});
  }
  int z_id[NUM_SETS];
  for (int s = 0; s < NUM_SETS; s++) {
    z_id[s] = single(z_c[s]);
    queries.push_back(Query{z_id[s], x, z_e[s]});
    queries.push_back(Query{z_id[s], x_next, z_next_e[s]});
  }
  for (int s = NUM_SETS - 2; s >= 0; s--) queries.push_back(Query{z_id[s], x_last, z_last_e[s]});
  {
    const int idz = single(zl_c), idi = single(pin_c), idt = single(ptab_c);
    queries.push_back(Query{idz, x, zl_e});
    queries.push_back(Query{idi, x, pin_e});
    queries.push_back(Query{idt, x, ptab_e});
    queries.push_back(Query{idi, x_prev, pin_prev_e});
    queries.push_back(Query{idz, x_next, zl_next_e});
  }
  for (int c = 0; c < NUM_FIXED; c++) queries.push_back(Query{single(K.fixed_commitments[c]), x, fixed_e[c]});
  for (int c = 0; c < NUM_PERM; c++) queries.push_back(Query{single(K.sigma_commitments[c]), x, sigma_e[c]});
  {
    CommitmentAcc hm;  // h = h_0 + x^n h_1 + x^2n h_2
    for (int p = 2; p >= 0; p--) {
      hm.scale(xn);
      hm.terms.push_back(Term{one, h_c[p]});
    }
    commitments.push_back(hm);
    queries.push_back(Query{(int)commitments.size() - 1, x, expected_h});
    queries.push_back(Query{single(random_c), x, random_e});
  }
  // point indices by first appearance; per commitment the ordered set of its point indices;
  // point sets numbered by first appearance over the commitments
  std::vector<Fp> points;
  auto point_index = [&](const Fp& p) {
    for (size_t i = 0; i < points.size(); i++)
      if (points[i] == p) return (int)i;
    points.push_back(p);
    return (int)points.size() - 1;
  };
  struct Opened {
    int commitment;
    std::vector<int> pts;     // ascending point indices
    std::vector<Fp> evals;    // same order
    int set = -1;
  };
  std::vector<Opened> opened;
  for (auto& q : queries) {
    const int pi = point_index(q.point);
    Opened* o = nullptr;
    for (auto& c : opened)
      if (c.commitment == q.commitment) o = &c;
    if (!o) {
      opened.push_back(Opened{q.commitment, {}, {}, -1});
      o = &opened.back();
    }
    size_t at = std::lower_bound(o->pts.begin(), o->pts.end(), pi) - o->pts.begin();
    if (at < o->pts.size() && o->pts[at] == pi) {
      o->evals[at] = q.eval;
    } else {
      o->pts.insert(o->pts.begin() + at, pi);
      o->evals.insert(o->evals.begin() + at, q.eval);
    }
  }
  std::vector<std::vector<int>> sets;
  for (auto& o : opened) {
    int found = -1;
    for (size_t i = 0; i < sets.size(); i++)
      if (sets[i] == o.pts) found = (int)i;
    if (found < 0) {
      sets.push_back(o.pts);
      found = (int)sets.size() - 1;
    }
    o.set = found;
  }
  const Fp x1 = tr.squeeze_challenge();
  const Fp x2 = tr.squeeze_challenge();
  const size_t nsets = sets.size();
  std::vector<CommitmentAcc> q_commitments(nsets);
  std::vector<Fp> x1_power(nsets, one);
  std::vector<std::vector<Fp>> q_evals(nsets);
  for (size_t s = 0; s < nsets; s++) q_evals[s].assign(sets[s].size(), Fp::zero());
  for (size_t ci = opened.size(); ci-- > 0;) {
    const Opened& o = opened[ci];
    CommitmentAcc m = commitments[o.commitment];
    m.scale(x1_power[o.set]);
    auto& dst = q_commitments[o.set].terms;
    dst.insert(dst.end(), m.terms.begin(), m.terms.end());
    for (size_t i = 0; i < o.evals.size(); i++) q_evals[o.set][i] = q_evals[o.set][i] + o.evals[i] * x1_power[o.set];
    x1_power[o.set] = x1_power[o.set] * x1;
  }
  const Affine q_prime_c = tr.read_point();
  const Fp x3 = tr.squeeze_challenge();
  std::vector<Fp> u(nsets);
  for (auto& e : u) e = tr.read_scalar();
  Fp msm_eval = Fp::zero();
  for (size_t s = 0; s < nsets; s++) {
    // r(x3) for the interpolant r of (point_i, q_evals_i), then (u_s - r(x3)) / prod (x3 - point_i)
    const size_t m = sets[s].size();
    Fp r_eval = Fp::zero();
    for (size_t i = 0; i < m; i++) {
      Fp num = q_evals[s][i], den = one;
      for (size_t j = 0; j < m; j++) {
        if (j == i) continue;
        num = num * (x3 - points[sets[s][j]]);
        den = den * (points[sets[s][i]] - points[sets[s][j]]);
      }
      r_eval = r_eval + num * den.inv();
    }
    Fp e = u[s] - r_eval;
    for (size_t i = 0; i < m; i++) e = e * (x3 - points[sets[s][i]]).inv();
    msm_eval = msm_eval * x2 + e;
  }
  const Fp x4 = tr.squeeze_challenge();
  CommitmentAcc acc;
  acc.terms.push_back(Term{one, q_prime_c});
  Fp vv = msm_eval;
  for (size_t s = 0; s < nsets; s++) {
    acc.scale(x4);
    acc.terms.insert(acc.terms.end(), q_commitments[s].terms.begin(), q_commitments[s].terms.end());
    vv = vv * x4 + u[s];
  }
  // ---- inner product argument -----------------------------------------------------------------------------
  const Affine s_c = tr.read_point();
  const Fp xi = tr.squeeze_challenge();
  acc.terms.push_back(Term{xi, s_c});
  const Fp zc = tr.squeeze_challenge();
  std::vector<Fp> us(k);
  for (int j = 0; j < k; j++) {
    const Affine lj = tr.read_point(), rj = tr.read_point();
    us[j] = tr.squeeze_challenge();
    acc.terms.push_back(Term{us[j].inv(), lj});
    acc.terms.push_back(Term{us[j], rj});
  }
  const Fp c = tr.read_scalar();
  const Fp f = tr.read_scalar();
  if (!tr.exhausted()) throw std::runtime_error("trailing bytes in proof");
  Fp b = one, cur = x3;
  for (int j = k - 1; j >= 0; j--) {
    b = b * (one + us[j] * cur);
    cur = cur.sqr();
  }
  out.terms = std::move(acc.terms);
  out.us = us;
  out.neg_c = c.neg();
  out.s0_constant = vv.neg();
  out.u_scalar = out.neg_c * b * zc;
  out.w_scalar = f.neg();
  return ZK_OK;
}

// device part: the final MSM of one or several proofs.  With several, proof i is weighted by weights[i]: the sum
// of the proofs' MSMs is the identity for random weights only if every one of them is (halo2's BatchVerifier) —
// one size-n fixed-base MSM and one variable-base MSM for the whole batch.
int32_t verify_final_msm(zk_ctx* ctx, const std::vector<Guard>& guards, const std::vector<Fp>& weights) {
  ProverState* S = prover_state(ctx);
  const DeviceKeys& K = S->keys;
  const DeviceParams& P = S->params;
  const uint64_t n = K.n;
  const int k = K.k;
  cudaStream_t st = ctx->stream;
  int32_t rc;
  std::vector<Fp> hs;
  std::vector<Affine> hp;
  Fp u_scalar = Fp::zero(), w_scalar = Fp::zero();
  for (size_t i = 0; i < guards.size(); i++) {
    for (auto& t : guards[i].terms) {
      hs.push_back(t.scalar * weights[i]);
      hp.push_back(t.point);
    }
    u_scalar = u_scalar + guards[i].u_scalar * weights[i];
    w_scalar = w_scalar + guards[i].w_scalar * weights[i];
  }
  const size_t nt = hs.size();
  if ((rc = ensure_buf(ctx, ctx->scratch_a, (size_t)(n + k + 8) * sizeof(Fp) + nt * sizeof(Fp)))) return rc;
  if ((rc = ensure_buf(ctx, ctx->scratch_b, nt * sizeof(Affine)))) return rc;
  Fp* d_s = (Fp*)ctx->scratch_a.ptr;
  Fp* d_us = d_s + n;
  Fp* d_ts = d_us + k + 8;
  Affine* d_tp = (Affine*)ctx->scratch_b.ptr;
  ZK_CUDA(ctx, cudaMemcpyAsync(d_ts, hs.data(), nt * sizeof(Fp), cudaMemcpyHostToDevice, st));
  ZK_CUDA(ctx, cudaMemcpyAsync(d_tp, hp.data(), nt * sizeof(Affine), cudaMemcpyHostToDevice, st));
  for (size_t i = 0; i < guards.size(); i++) {
    ZK_CUDA(ctx, cudaMemcpyAsync(d_us, guards[i].us.data(), (size_t)k * sizeof(Fp), cudaMemcpyHostToDevice, st));
    compute_s_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_s, n, k, d_us, guards[i].neg_c * weights[i],
                                                                  guards[i].s0_constant * weights[i], i ? 1 : 0);
    ctx->launches++;
  }
  ZK_CUDA(ctx, cudaGetLastError());
  const Fp extra[2] = {u_scalar, w_scalar};
  const uint32_t extra_idx[2] = {(uint32_t)n + 1, (uint32_t)n};
  XYZZ g_part, t_part;
  if ((rc = msm_fixed(ctx, P.fb_g, d_s, n, extra, extra_idx, 2, &g_part))) return rc;
  if ((rc = msm_run(ctx, d_ts, d_tp, nt, &t_part, nullptr))) return rc;
  if (!g_part.add(t_part).is_identity()) return set_error(ctx, ZK_E_VERIFY, "final MSM is not the identity");
  return ZK_OK;
}

}  // namespace
}  // namespace zkodst

using namespace zkodst;

extern "C" int32_t zk_verify_proof(zk_ctx* ctx, const uint8_t* proof, uint64_t proof_len) {
  if (!ctx || !proof) return ZK_E_INVALID;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  try {
    std::vector<Guard> g(1);
    int32_t rc = verify_transcript(ctx, proof, proof_len, g[0]);
    if (rc) return rc;
    return verify_final_msm(ctx, g, {Fp::one()});
  } catch (std::exception& e) {
    return set_error(ctx, ZK_E_VERIFY, e.what());
  }
}

// Batch verification: the transcript phase of every proof on the host, then ONE final MSM for all of them,
// proof i weighted by a field element drawn from XorShiftRng(seed) (the first weight is 1).
extern "C" int32_t zk_verify_proofs_batch(zk_ctx* ctx, const uint8_t* proofs, const uint64_t* proof_lens,
                                          uint64_t count, const uint8_t seed[16]) {
  if (!ctx || !proofs || !proof_lens || !seed || count == 0) return ZK_E_INVALID;
  ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  try {
    std::vector<Guard> guards(count);
    std::vector<Fp> weights(count);
    XsState st;
    memcpy(st.s, seed, 16);
    if (!(st.s[0] | st.s[1] | st.s[2] | st.s[3])) st.s[0] = st.s[1] = st.s[2] = st.s[3] = 0x0BAD5EED;
    uint64_t off = 0;
    for (uint64_t i = 0; i < count; i++) {
      int32_t rc = verify_transcript(ctx, proofs + off, proof_lens[i], guards[i]);
      if (rc) return rc;
      off += proof_lens[i];
      if (i == 0) {
        weights[i] = Fp::one();
      } else {
        uint64_t w[8];
        for (int j = 0; j < 8; j++) {
          const uint64_t lo = xs_step(st);
          w[j] = lo | ((uint64_t)xs_step(st) << 32);
        }
        weights[i] = Fp::from_u512(w);
      }
    }
    return verify_final_msm(ctx, guards, weights);
  } catch (std::exception& e) {
    return set_error(ctx, ZK_E_VERIFY, e.what());
  }
}
