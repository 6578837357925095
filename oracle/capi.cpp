// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C entry points over the oracle so that tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs can drive it through ctypes.  Nothing under
// zk-odst_b200/ may load this library.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include "blake2f_circuit.hpp"
#include "curve.hpp"
#include "mock_prover.hpp"
#include "plonk.hpp"
#include "xorshift.hpp"

using namespace zko;

static void set_msg(char* msg, size_t len, const std::string& s) {
  if (msg && len) {
    snprintf(msg, len, "%s", s.c_str());
  }
}

extern "C" {

// ---- BLAKE2b -----------------------------------------------------------------------------
int zko_blake2f_F(const uint8_t in[213], uint8_t out[64]) {
  Blake2fInput x;
  if (parse_eip152(in, x)) return -1;
  uint64_t h[8];
  memcpy(h, x.h, 64);
  blake2b_F(h, x.m, x.t, x.f, x.rounds);
  memcpy(out, h, 64);
  return 0;
}
void zko_blake2b(const char* personal16, const uint8_t* data, size_t len, uint8_t out[64]) {
  Blake2b st(personal16);
  st.update(data, len);
  st.finalize(out);
}
void zko_xorshift_u64(const uint8_t seed[16], size_t n, uint64_t* out) {
  XorShiftRng rng(seed);
  for (size_t i = 0; i < n; i++) out[i] = rng.next_u64();
}

// ---- fields (which: 0 = Fp, 1 = Fq); all elements are 4 x u64 Montgomery limbs ------------
// consts layout: MOD, R, R2, R3, {INV,0,0,0}, GENERATOR, ROOT_OF_UNITY, DELTA, ZETA  (raw
// canonical integers for GENERATOR.. so tests can compare against SURVEY Appendix A.1)
void zko_field_consts(int which, uint64_t out[9 * 4]) {
  auto fill = [&](auto tag) {
    typedef decltype(tag) F;
    const auto& c = F::C();
    memset(out, 0, 9 * 32);
    memcpy(out + 4, c.R.l, 32);
    memcpy(out + 8, c.R2.l, 32);
    memcpy(out + 12, c.R3.l, 32);
    out[16] = c.inv;
    c.generator.to_raw(out + 20);
    c.root_of_unity.to_raw(out + 24);
    c.delta.to_raw(out + 28);
    c.zeta.to_raw(out + 32);
  };
  if (which == 0) {
    memcpy(out, FpParams::MOD, 32);
    uint64_t m[4];
    memcpy(m, FpParams::MOD, 32);
    fill(Fp());
    memcpy(out, m, 32);
  } else {
    uint64_t m[4];
    memcpy(m, FqParams::MOD, 32);
    fill(Fq());
    memcpy(out, m, 32);
  }
}
// op: 0 mul, 1 add, 2 sub, 3 invert(a), 4 from_raw(a) (canonical -> Montgomery),
//     5 to_raw(a), 6 sqrt(a) (returns 0/1 in rc)
int zko_field_op(int which, int op, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
  auto run = [&](auto tag) -> int {
    typedef decltype(tag) F;
    F x, y, r;
    memcpy(x.l, a, 32);
    if (b) memcpy(y.l, b, 32);
    int rc = 0;
    switch (op) {
      case 0: r = x * y; break;
      case 1: r = x + y; break;
      case 2: r = x - y; break;
      case 3: r = x.invert(); break;
      case 4: r = F::from_raw(a); break;
      case 5: x.to_raw(r.l); break;
      case 6: rc = x.sqrt(r) ? 1 : 0; break;
      default: return -1;
    }
    memcpy(out, r.l, 32);
    return rc;
  };
  return which == 0 ? run(Fp()) : run(Fq());
}
void zko_field_from_u512(int which, const uint64_t v[8], uint64_t out[4]) {
  if (which == 0) {
    Fp r = Fp::from_u512(v);
    memcpy(out, r.l, 32);
  } else {
    Fq r = Fq::from_u512(v);
    memcpy(out, r.l, 32);
  }
}

// ---- circuit ------------------------------------------------------------------------------
uint64_t zko_rows_per_compression(uint32_t rounds) { return rows_per_compression(rounds); }

uint32_t zko_spread16(uint32_t x) { return spread16(x); }
uint32_t zko_get_tag(uint32_t x) { return get_tag(x); }
uint32_t zko_even_bits32(uint32_t x) { return even_bits32(x); }
uint32_t zko_odd_bits32(uint32_t x) { return odd_bits32(x); }

static int parse_inputs(const uint8_t* inputs213, size_t n, std::vector<Blake2fInput>& v,
                        uint32_t rounds) {
  v.resize(n);
  for (size_t i = 0; i < n; i++) {
    if (parse_eip152(inputs213 + 213 * i, v[i])) return -1;
    if (v[i].rounds != rounds) return -2;
  }
  return 0;
}

// Witness generation (Circuit::synthesize).  advice_mont: 12 * n * 4 u64 (column-major by
// halo2 column index, 32-byte Montgomery Fp per cell) or NULL; advice_raw: 12 * n u64 or
// NULL; digests: n_compressions * 8 u64 or NULL.
int zko_blake2f_witness(int k, uint32_t rounds, const uint8_t* inputs213, size_t n_compressions,
                        uint64_t* advice_mont, uint64_t* advice_raw, uint64_t* digests) {
  try {
    std::vector<Blake2fInput> in;
    int rc = parse_inputs(inputs213, n_compressions, in, rounds);
    if (rc) return rc;
    ConstraintSystem cs;
    blake2f_configure(cs);
    Blake2fAssignment as;
    as.want_shape = false;
    blake2f_synthesize(as, k, rounds, in.data(), n_compressions, cs.blinding_factors());
    size_t n = as.n;
    for (int c = 0; c < NUM_ADVICE; c++)
      for (size_t r = 0; r < n; r++) {
        if (advice_raw) advice_raw[c * n + r] = as.advice[c][r];
        if (advice_mont) {
          Fp v = Fp::from_u64(as.advice[c][r]);
          memcpy(advice_mont + (c * n + r) * 4, v.l, 32);
        }
      }
    if (digests)
      for (size_t j = 0; j < n_compressions; j++) memcpy(digests + 8 * j, as.outputs[j].data(), 64);
    return 0;
  } catch (std::exception& e) {
    fprintf(stderr, "zko_blake2f_witness: %s\n", e.what());
    return -3;
  }
}

// MockProver-equivalent over an advice matrix given as integers (12 * n u64, column-major).
static int mock_verify_raw_impl(int k, uint32_t rounds, size_t n_compressions, const uint8_t* chain,
                                const uint64_t* advice_raw, char* msg, size_t msg_len) {
  try {
    CircuitShape sh;
    build_shape(sh, k, rounds, n_compressions, chain);
    std::vector<std::vector<uint64_t>> adv(NUM_ADVICE, std::vector<uint64_t>(sh.n));
    for (int c = 0; c < NUM_ADVICE; c++) memcpy(adv[c].data(), advice_raw + c * sh.n, 8 * sh.n);
    MockFailure f = mock_verify(sh, adv);
    set_msg(msg, msg_len, f.what);
    return f.ok ? 0 : 1;
  } catch (std::exception& e) {
    set_msg(msg, msg_len, e.what());
    return -3;
  }
}
int zko_mock_verify_raw(int k, uint32_t rounds, size_t n_compressions, const uint64_t* advice_raw,
                        char* msg, size_t msg_len) {
  return mock_verify_raw_impl(k, rounds, n_compressions, nullptr, advice_raw, msg, msg_len);
}
// Same, over Montgomery-form cells (what the CUDA path emits).  Any cell >= 2^64 fails.
// chain: n_compressions flags or NULL (compression j continues compression j - 1).
int zko_mock_verify_mont_chained(int k, uint32_t rounds, size_t n_compressions, const uint8_t* chain,
                                 const uint64_t* advice_mont, char* msg, size_t msg_len);
int zko_mock_verify_mont(int k, uint32_t rounds, size_t n_compressions,
                         const uint64_t* advice_mont, char* msg, size_t msg_len) {
  return zko_mock_verify_mont_chained(k, rounds, n_compressions, nullptr, advice_mont, msg, msg_len);
}
int zko_mock_verify_mont_chained(int k, uint32_t rounds, size_t n_compressions, const uint8_t* chain,
                                 const uint64_t* advice_mont, char* msg, size_t msg_len) {
  size_t n = (size_t)1 << k;
  std::vector<uint64_t> raw(NUM_ADVICE * n);
  for (size_t i = 0; i < NUM_ADVICE * n; i++) {
    Fp v;
    memcpy(v.l, advice_mont + 4 * i, 32);
    if (!raw_geq(FpParams::MOD, v.l) || raw_geq(v.l, FpParams::MOD)) {
      set_msg(msg, msg_len, "non-canonical Montgomery limb");
      return 2;
    }
    u64 r[4];
    v.to_raw(r);
    if (r[1] | r[2] | r[3]) {
      set_msg(msg, msg_len, "cell value >= 2^64 at index " + std::to_string(i));
      return 2;
    }
    raw[i] = r[0];
  }
  return mock_verify_raw_impl(k, rounds, n_compressions, chain, raw.data(), msg, msg_len);
}

static uint64_t fnv1a(uint64_t h, const void* data, size_t len) {
  const uint8_t* p = (const uint8_t*)data;
  for (size_t i = 0; i < len; i++) {
    h ^= p[i];
    h *= 0x100000001b3ULL;
  }
  return h;
}
// FNV-1a digests of one region's copy constraints and selector activations (same definition
// as zk_blake2f_layout_hash in include/zkodst.h), from the oracle's own synthesize.
int zko_layout_hash(uint32_t rounds, uint64_t* copies_hash, uint64_t* selectors_hash,
                    uint64_t* n_copies) {
  try {
    Blake2fAssignment as;
    as.want_witness = false;
    int k = 17;
    while (rows_per_compression(rounds) + 6 > ((size_t)1 << k)) k++;
    blake2f_synthesize(as, k, rounds, nullptr, 1, 5);
    size_t R = rows_per_compression(rounds);
    uint64_t h = 0xcbf29ce484222325ULL;
    for (auto& c : as.copies) {
      uint32_t rec[4] = {(uint32_t)c.lc, (uint32_t)c.lr, (uint32_t)c.rc, (uint32_t)c.rr};
      h = fnv1a(h, rec, sizeof rec);
    }
    *copies_hash = h;
    uint64_t hs = 0xcbf29ce484222325ULL;
    for (int s = 0; s < NUM_SEL; s++) hs = fnv1a(hs, as.selectors[s].data(), R);
    *selectors_hash = hs;
    *n_copies = as.copies.size();
    return 0;
  } catch (std::exception& e) {
    return -3;
  }
}

// Circuit description for cross-checking the product's hard-wired tables.
// Writes "col,rot;" lists etc. as text.
int zko_describe_circuit(int k, uint32_t rounds, size_t n_compressions, char* out, size_t out_len) {
  try {
    CircuitShape sh;
    build_shape(sh, k, rounds, n_compressions);
    std::ostringstream os;
    os << "degree=" << sh.cs.degree() << "\nblinding_factors=" << sh.blinding_factors
       << "\nnum_fixed=" << sh.cs.num_fixed_columns << "\nnum_advice=" << sh.cs.num_advice_columns
       << "\nadvice_queries=";
    for (auto& q : sh.cs.advice_queries) os << q.column << "," << q.rotation << ";";
    os << "\nfixed_queries=";
    for (auto& q : sh.cs.fixed_queries) os << q.column << "," << q.rotation << ";";
    os << "\npermutation_columns=";
    for (int c : sh.cs.permutation_columns) os << c << ";";
    os << "\nselectors=";
    for (auto& a : sh.selector_assignments)
      os << a.selector << ":" << a.fixed_column << ":" << a.assigned_root << ":" << a.combination_len
         << ";";
    os << "\nn_copies=" << sh.copies.size() << "\nn_polys=";
    size_t np = 0;
    for (auto& g : sh.cs.gates) np += g.polys.size();
    os << np << "\n";
    set_msg(out, out_len, os.str());
    return 0;
  } catch (std::exception& e) {
    set_msg(out, out_len, e.what());
    return -3;
  }
}


// ---- params / keygen / prove / verify ----------------------------------------------------------
struct OracleProver {
  Params params;
  ProvingKey pk;
  bool has_pk = false, has_vk = false;
  ProverTrace trace;
  double synth_ms = 0, prove_ms = 0;  // last zko_create_proof: witness synthesis, the rest of create_proof
};

// worker threads: t >= 1, or 0 for all hardware threads; returns the count now in use
int zko_set_threads(int t) {
  set_oracle_threads(t);
  return oracle_threads();
}
// milliseconds of the last zko_create_proof: out[0] = witness synthesis (Circuit::synthesize), out[1] = the rest
void zko_last_proof_ms(void* h, double out[2]) {
  auto* p = (OracleProver*)h;
  out[0] = p->synth_ms;
  out[1] = p->prove_ms;
}

void* zko_prover_new_substitute(int k, const uint8_t seed[16]) {
  try {
    auto* p = new OracleProver();
    p->params = Params::generate_substitute(k, seed);
    return p;
  } catch (std::exception& e) {
    fprintf(stderr, "zko_prover_new_substitute: %s\n", e.what());
    return nullptr;
  }
}
void* zko_prover_new_from_params(const uint8_t* bytes, size_t len) {
  try {
    auto* p = new OracleProver();
    p->params = Params::read(bytes, len);
    return p;
  } catch (std::exception& e) {
    fprintf(stderr, "zko_prover_new_from_params: %s\n", e.what());
    return nullptr;
  }
}
void zko_prover_free(void* h) { delete (OracleProver*)h; }
// halo2 Params::write format; returns the size, writes if out != NULL
size_t zko_params_write(void* h, uint8_t* out, size_t cap) {
  auto* p = (OracleProver*)h;
  size_t need = 4 + (2 * p->params.n + 2) * 32;
  if (out && cap >= need) {
    std::vector<uint8_t> v;
    p->params.write(v);
    memcpy(out, v.data(), need);
  }
  return need;
}
// raw affine points (Montgomery x, y; 64 bytes each): which 0 = g, 1 = g_lagrange, 2 = {w, u}
void zko_params_points(void* h, int which, uint64_t* out) {
  auto* p = (OracleProver*)h;
  const std::vector<Affine>* v = which == 0 ? &p->params.g : &p->params.g_lagrange;
  if (which == 2) {
    memcpy(out, &p->params.w, 64);
    memcpy(out + 8, &p->params.u, 64);
    return;
  }
  memcpy(out, v->data(), v->size() * 64);
}
int zko_keygen(void* h, uint32_t rounds, size_t n_compressions) {
  auto* p = (OracleProver*)h;
  try {
    keygen(p->params, rounds, n_compressions, p->pk);
    p->has_pk = p->has_vk = true;
    return 0;
  } catch (std::exception& e) {
    fprintf(stderr, "zko_keygen: %s\n", e.what());
    return -3;
  }
}
// keygen with record chaining (chain: n_compressions flags; compression j continues j - 1); vk_only as below
int zko_keygen_chained(void* h, uint32_t rounds, size_t n_compressions, const uint8_t* chain, int vk_only) {
  auto* p = (OracleProver*)h;
  try {
    p->has_pk = false;
    keygen(p->params, rounds, n_compressions, p->pk, vk_only != 0, chain);
    p->has_vk = true;
    p->has_pk = !vk_only;
    return 0;
  } catch (std::exception& e) {
    fprintf(stderr, "zko_keygen_chained: %s\n", e.what());
    return -3;
  }
}
// keygen_vk only: enough for zko_vk_bytes and zko_verify_proof (no proving key)
int zko_keygen_vk(void* h, uint32_t rounds, size_t n_compressions) {
  auto* p = (OracleProver*)h;
  try {
    p->has_pk = false;
    keygen(p->params, rounds, n_compressions, p->pk, true);
    p->has_vk = true;
    return 0;
  } catch (std::exception& e) {
    fprintf(stderr, "zko_keygen_vk: %s\n", e.what());
    return -3;
  }
}
// vk pieces for comparison: fixed commitments then permutation commitments (compressed, 32 B each),
// then transcript_repr (32 B repr).  Returns the number of bytes.
size_t zko_vk_bytes(void* h, uint8_t* out, size_t cap) {
  auto* p = (OracleProver*)h;
  const VerifyingKey& vk = p->pk.vk;
  size_t need = (vk.fixed_commitments.size() + vk.permutation_commitments.size() + 1) * 32;
  if (!out || cap < need) return need;
  size_t off = 0;
  for (auto& c : vk.fixed_commitments) { c.to_bytes(out + off); off += 32; }
  for (auto& c : vk.permutation_commitments) { c.to_bytes(out + off); off += 32; }
  vk.transcript_repr.to_repr(out + off);
  return need;
}
// vk commitments as raw affine points (Montgomery x, y; 64 bytes each): fixed then permutation; returns the count
size_t zko_vk_points(void* h, uint64_t* out) {
  auto* p = (OracleProver*)h;
  const VerifyingKey& vk = p->pk.vk;
  size_t i = 0;
  for (auto& c : vk.fixed_commitments) memcpy(out + 8 * i++, &c, 64);
  for (auto& c : vk.permutation_commitments) memcpy(out + 8 * i++, &c, 64);
  return i;
}
// the `{:?}` rendering of vk.pinned() that transcript_repr hashes; returns the length, writes if it fits
size_t zko_vk_pinned_debug(void* h, char* out, size_t cap) {
  auto* p = (OracleProver*)h;
  if (!p->has_vk) return 0;
  const std::string s = vk_pinned_debug(p->pk.vk);
  if (out && cap >= s.size()) memcpy(out, s.data(), s.size());
  return s.size();
}
int zko_create_proof(void* h, const uint8_t* inputs213, size_t n_compressions, const uint8_t seed[16],
                     uint8_t* proof_out, size_t* proof_len) {
  auto* p = (OracleProver*)h;
  try {
    if (!p->has_pk) return -6;
    uint32_t rounds = p->pk.vk.shape.rounds;
    if (n_compressions != p->pk.vk.shape.n_compressions) return -1;
    std::vector<Blake2fInput> in;
    int rc = parse_inputs(inputs213, n_compressions, in, rounds);
    if (rc) return rc;
    Blake2fAssignment as;
    as.want_shape = false;
    const auto t0 = std::chrono::steady_clock::now();
    blake2f_synthesize(as, p->params.k, rounds, in.data(), n_compressions,
                       p->pk.vk.shape.blinding_factors);
    const auto t1 = std::chrono::steady_clock::now();
    XorShiftRng rng(seed);
    std::vector<uint8_t> proof = create_proof(p->params, p->pk, as.advice, rng, &p->trace);
    const auto t2 = std::chrono::steady_clock::now();
    p->synth_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    p->prove_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
    if (*proof_len < proof.size()) {
      *proof_len = proof.size();
      return -8;
    }
    memcpy(proof_out, proof.data(), proof.size());
    *proof_len = proof.size();
    return 0;
  } catch (std::exception& e) {
    fprintf(stderr, "zko_create_proof: %s\n", e.what());
    return -3;
  }
}
int zko_verify_proof(void* h, const uint8_t* proof, size_t len, char* msg, size_t msg_len) {
  auto* p = (OracleProver*)h;
  if (!p->has_vk) return -6;
  std::string why;
  bool ok = verify_proof(p->params, p->pk.vk, proof, len, &why);
  set_msg(msg, msg_len, why);
  return ok ? 0 : 1;
}
// MSM / NTT primitives for kernel-level parity tests
void zko_msm(const uint64_t* scalars_mont, const uint64_t* bases_affine_mont, size_t n,
             uint64_t out_affine[8]) {
  Jac r = msm((const Fp*)scalars_mont, (const Affine*)bases_affine_mont, n);
  Affine a = r.to_affine();
  memcpy(out_affine, &a, 64);
}
// in-place: inverse = 0 -> coeff->lagrange (omega), 1 -> lagrange->coeff (omega^-1, scaled)
void zko_ntt(uint64_t* data_mont, int log_n, int inverse) {
  Domain d(2, log_n);
  Poly a((Fp*)data_mont, (Fp*)data_mont + ((size_t)1 << log_n));
  a = inverse ? d.lagrange_to_coeff(a) : d.coeff_to_lagrange(a);
  memcpy(data_mont, a.data(), a.size() * 32);
}
// coeffs (n) -> extended coset evaluations (4n for this circuit's degree)
void zko_coeff_to_extended(const uint64_t* coeffs_mont, int k, int cs_degree, uint64_t* out) {
  Domain d(cs_degree, k);
  Poly a((const Fp*)coeffs_mont, (const Fp*)coeffs_mont + d.n);
  Poly e = d.coeff_to_extended(a);
  memcpy(out, e.data(), e.size() * 32);
}
void zko_random_fields(const uint8_t seed[16], size_t n, uint64_t* out_mont) {
  XorShiftRng rng(seed);
  for (size_t i = 0; i < n; i++) {
    Fp v = rng.random_field<Fp>();
    memcpy(out_mont + 4 * i, v.l, 32);
  }
}

}  // extern "C"
