// zkodst.hpp — the reference's plugin surface for the BLAKE2f proving path, in C++, over the C ABI.
//
// The reference is a Rust crate and there is no Rust toolchain in the build image, so the host side
// above `zkodst.h` that a maintainer can compile and run here is this header (C++17, header-only; it
// includes nothing but the C ABI and the standard library).  It keeps the reference's names, argument
// meaning and error behaviour (paths relative to the reference repository):
//
//   Blake2fInstructions { initialization_vector, initialization, compress, digest }
//                                                         blake2f-circuit/src/blake2f.rs:40-72
//   Blake2f<Chip> { new_, update, finalize, digest }, Blake2fDigest      src/blake2f.rs:75-181
//   BlockWord, Table16Config, Table16Chip { configure, construct, load } src/blake2f/table16.rs:59-61, 250-336
//   Circuit { without_witnesses, configure, synthesize }                 src/blake2f.rs:257-277 (commented test
//                                                         circuit), table16/spread_table.rs:630-655
//   Params { new_, read, write }, keygen_vk, keygen_pk, create_proof, verify_proof, SingleVerifier,
//   Blake2bWrite / Blake2bRead, MockProver::run(..).verify()
//                                                         blake2f-circuit/benches/blake2f.rs:83-142,
//                                                         table16/spread_table.rs:759-763
//
// What changes underneath: `synthesize` does not assign cells one by one.  The layouter handed to it
// records every compression region the circuit lays out as one EIP-152 record (rounds, h, m, t, f); the
// cells of all regions are then produced on the GPU in one call (`zk_blake2f_witness_batch` inside
// `zk_create_proof` / `zk_mock_verify`).  `new` is a C++ keyword: Rust's `T::new` is spelled `new_`.
//
// The Rust shim with the same shape is rust/zkodst-backend (source only); tests/cpp/facade_test.cpp is
// the reference's benches/blake2f.rs and its commented test module rewritten against this header.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "zkodst.h"

namespace zkodst {

// ---- halo2_proofs::plonk::Error ------------------------------------------------------------------------
// The variants the path can raise, plus `Backend` for what has no halo2 counterpart (CUDA, allocation,
// call order).  `code` is the ABI status the variant was made from.
class Error : public std::runtime_error {
 public:
  enum Kind { Synthesis, ConstraintSystemFailure, NotEnoughRowsAvailable, Opening, Transcript, Backend };
  Error(Kind k, int32_t c, const std::string& what) : std::runtime_error(what), kind(k), code(c) {}
  Kind kind;
  int32_t code;
};

// halo2_proofs::circuit::Value<T>: a witness that is unknown during key generation
template <class T>
using Value = std::optional<T>;
template <class T>
inline Value<T> known(T v) { return Value<T>(v); }

constexpr size_t BLOCK_SIZE = 16;   // words of a message block (src/blake2f.rs:34)
constexpr size_t DIGEST_SIZE = 8;   // words of a digest (src/blake2f.rs:36)
constexpr uint32_t BLAKE2B_ROUNDS = 12;

// table16.rs:59-61 (64-bit words for BLAKE2b).  Default = the zero word, what `finalize` pads with.
struct BlockWord {
  Value<uint64_t> v = uint64_t(0);
  BlockWord() = default;
  BlockWord(uint64_t w) : v(w) {}
  static BlockWord unknown() {
    BlockWord b;
    b.v.reset();
    return b;
  }
};

// ---- the device ------------------------------------------------------------------------------------------
// One `zk_ctx` (one host thread, one GPU).  The reference has no counterpart; the Rust shim keeps the same
// handle inside its chip backend state.
class Device {
 public:
  explicit Device(int32_t device_id = 0) : device_id_(device_id) {
    zk_ctx* c = nullptr;
    const int32_t rc = zk_ctx_create(device_id, &c);
    if (rc != ZK_OK) throw Error(Error::Backend, rc, "zk_ctx_create failed (no CUDA device? there is no CPU fallback)");
    ctx_.reset(c, zk_ctx_destroy);
  }
  zk_ctx* ctx() const { return ctx_.get(); }
  int32_t device_id() const { return device_id_; }
  // ABI status -> plonk::Error, as the call site `where` would see it from halo2
  void check(int32_t rc, const char* where, Error::Kind on_verify = Error::ConstraintSystemFailure) const {
    if (rc == ZK_OK) return;
    const char* msg = zk_last_error(ctx_.get());
    const std::string what = std::string(where) + ": " + (msg ? msg : "");
    switch (rc) {
      case ZK_E_ROWS: throw Error(Error::NotEnoughRowsAvailable, rc, what);
      case ZK_E_INPUT: throw Error(Error::Synthesis, rc, what);
      case ZK_E_VERIFY: throw Error(on_verify, rc, what);
      default: throw Error(Error::Backend, rc, what);
    }
  }

 private:
  int32_t device_id_;
  std::shared_ptr<zk_ctx> ctx_;
};

// ---- what a circuit lays out -----------------------------------------------------------------------------
// The witness of one compression region: the EIP-152 precompile input (the `Blake2fWitness` of the
// reference's commented test circuit, src/blake2f.rs:201-240).
struct Blake2fWitness {
  uint32_t rounds = BLAKE2B_ROUNDS;
  std::array<uint64_t, 8> h{};
  std::array<uint64_t, 16> m{};
  std::array<uint64_t, 2> t{};
  bool f = false;
  std::array<uint8_t, ZK_BLAKE2F_INPUT_BYTES> to_eip152() const {
    std::array<uint8_t, ZK_BLAKE2F_INPUT_BYTES> out{};
    out[0] = uint8_t(rounds >> 24);
    out[1] = uint8_t(rounds >> 16);
    out[2] = uint8_t(rounds >> 8);
    out[3] = uint8_t(rounds);
    auto le = [&](size_t off, uint64_t w) {
      for (int b = 0; b < 8; b++) out[off + b] = uint8_t(w >> (8 * b));
    };
    for (size_t i = 0; i < 8; i++) le(4 + 8 * i, h[i]);
    for (size_t i = 0; i < 16; i++) le(68 + 8 * i, m[i]);
    for (size_t i = 0; i < 2; i++) le(196 + 8 * i, t[i]);
    out[212] = f ? 1 : 0;
    return out;
  }
};

class ConstraintSystem {};  // `configure` only declares the column / gate / lookup plan, which is fixed
                            // (docs/CIRCUIT.md) and lives in the library

// halo2_proofs::circuit::Layouter, as this backend needs it: it records the regions.
class Layouter {
 public:
  explicit Layouter(bool witnesses_known) : known_(witnesses_known) {}
  Layouter& namespace_(const char*) { return *this; }
  bool witnesses_known() const { return known_; }
  bool table_loaded = false;
  uint32_t rounds = 0;               // every region of one circuit has the same round count
  std::vector<uint8_t> records;      // 213 bytes per compression region, in layout order
  uint64_t regions = 0;
  void push_region(const Blake2fWitness* w, uint32_t r) {
    if (regions && r != rounds) throw Error(Error::Synthesis, ZK_E_INPUT, "regions of one circuit must share the round count");
    rounds = r;
    regions++;
    if (known_ && w) {
      const auto rec = w->to_eip152();
      records.insert(records.end(), rec.begin(), rec.end());
    }
  }

 private:
  bool known_;
};

// table16/compression.rs:286-525 `State`: the chaining value (as values, the cells live on the device),
// and what BLAKE2b's F needs besides it: the byte counter, the final flag and the round count.
struct State {
  std::array<Value<uint64_t>, 8> h;
  uint64_t t = 0;                 // BLAKE2b byte counter of the block about to be compressed (it counts that block)
  bool last = false;
  uint32_t rounds = BLAKE2B_ROUNDS;
};

// ---- src/blake2f.rs:40-72 --------------------------------------------------------------------------------
class Blake2fInstructions {
 public:
  virtual ~Blake2fInstructions() = default;
  // Places the IV in the circuit, returning the initial state variable (BLAKE2b-512, unkeyed: IV ^ 0x01010040).
  virtual State initialization_vector(Layouter& layouter) const = 0;
  // Creates an initial state from the output state of a previous block.
  virtual State initialization(Layouter& layouter, const State& init_state) const = 0;
  // Starting from the given initialized state, processes a block of input and returns the final state.
  virtual State compress(Layouter& layouter, const State& initialized_state,
                         const std::array<BlockWord, BLOCK_SIZE>& input) const = 0;
  // Converts the given state into a message digest.
  virtual std::array<BlockWord, DIGEST_SIZE> digest(Layouter& layouter, const State& state) const = 0;
};

struct Table16Config {
  bool configured = false;
};

// src/blake2f/table16.rs:250-384
class Table16Chip : public Blake2fInstructions {
 public:
  static Table16Config configure(ConstraintSystem&) { return Table16Config{true}; }
  static Table16Chip construct(Table16Config config) { return Table16Chip(config); }
  // SpreadTableChip::load (table16/spread_table.rs:470-508): the 2^16-row table is a fixed column set the
  // library generates at keygen
  static void load(const Table16Config& config, Layouter& layouter) {
    if (!config.configured) throw Error(Error::Synthesis, ZK_E_STATE, "Table16Chip::load before configure");
    layouter.table_loaded = true;
  }
  const Table16Config& config() const { return config_; }

  State initialization_vector(Layouter&) const override {
    static const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                   0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                   0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};  // table16.rs:47-56
    State s;
    for (int i = 0; i < 8; i++) s.h[i] = IV[i];
    s.h[0] = *s.h[0] ^ 0x01010040ULL;
    return s;
  }
  State initialization(Layouter&, const State& init_state) const override { return init_state; }
  // One compression region.  The chaining value of the next block is F's output, which the host needs
  // only as a value (zk_blake2f_compress); the region's cells are assigned on the device.
  State compress(Layouter& layouter, const State& st, const std::array<BlockWord, BLOCK_SIZE>& input) const override {
    if (!layouter.table_loaded) throw Error(Error::Synthesis, ZK_E_STATE, "Table16Chip::compress before load");
    State out = st;
    bool all_known = layouter.witnesses_known();
    Blake2fWitness w;
    w.rounds = st.rounds;
    for (size_t i = 0; i < 8 && all_known; i++) {
      if (!st.h[i]) all_known = false; else w.h[i] = *st.h[i];
    }
    for (size_t i = 0; i < BLOCK_SIZE && all_known; i++) {
      if (!input[i].v) all_known = false; else w.m[i] = *input[i].v;
    }
    w.t = {st.t, 0};
    w.f = st.last;
    if (layouter.witnesses_known() && !all_known)
      throw Error(Error::Synthesis, ZK_E_INPUT, "unknown witness value in a proving run");
    layouter.push_region(all_known ? &w : nullptr, st.rounds);
    if (all_known) {
      const auto rec = w.to_eip152();
      uint8_t hout[64];
      const int32_t rc = zk_blake2f_compress(rec.data(), hout);
      if (rc != ZK_OK) throw Error(Error::Synthesis, rc, "zk_blake2f_compress");
      for (int i = 0; i < 8; i++) {
        uint64_t v = 0;
        for (int b = 0; b < 8; b++) v |= uint64_t(hout[8 * i + b]) << (8 * b);
        out.h[i] = v;
      }
    } else {
      for (auto& v : out.h) v.reset();
    }
    return out;
  }
  std::array<BlockWord, DIGEST_SIZE> digest(Layouter&, const State& state) const override {
    std::array<BlockWord, DIGEST_SIZE> out;
    for (size_t i = 0; i < DIGEST_SIZE; i++) out[i].v = state.h[i];
    return out;
  }

 private:
  explicit Table16Chip(Table16Config c) : config_(c) {}
  Table16Config config_;
};

// src/blake2f.rs:75-76
struct Blake2fDigest {
  std::array<BlockWord, DIGEST_SIZE> words;
  // the 64 digest bytes, if the witness is known
  Value<std::array<uint8_t, 64>> bytes() const {
    std::array<uint8_t, 64> out{};
    for (size_t i = 0; i < DIGEST_SIZE; i++) {
      if (!words[i].v) return std::nullopt;
      for (int b = 0; b < 8; b++) out[8 * i + b] = uint8_t(*words[i].v >> (8 * b));
    }
    return out;
  }
};

// src/blake2f.rs:80-181: the streaming gadget, at a granularity of one 64-bit word.  BLAKE2b flags its
// last block, so a full block is compressed only once more data (or `finalize`) arrives.
template <class Chip>
class Blake2f {
 public:
  // Create a new hasher instance.
  static Blake2f new_(Chip chip, Layouter& layouter, uint32_t rounds = BLAKE2B_ROUNDS) {
    Blake2f h(std::move(chip));
    h.state_ = h.chip_.initialization_vector(layouter);
    h.state_.rounds = rounds;
    return h;
  }
  // Digest data, updating the internal state.
  void update(Layouter& layouter, const BlockWord* data, size_t len) {
    for (size_t i = 0; i < len; i++) {
      if (cur_block_.size() == BLOCK_SIZE) flush(layouter, false);
      cur_block_.push_back(data[i]);
    }
    length_ += len * 64;
  }
  void update(Layouter& layouter, const std::vector<BlockWord>& data) { update(layouter, data.data(), data.size()); }
  // Retrieve result and consume hasher instance.  `unused_bytes_of_last_word` (< 8) trims the byte counter when
  // the message does not end on a word boundary (the caller zero-pads its last word).
  Blake2fDigest finalize(Layouter& layouter, size_t unused_bytes_of_last_word = 0) {
    const size_t words = cur_block_.size();
    cur_block_.resize(BLOCK_SIZE);  // pad with BlockWord::default()
    const uint64_t bytes = words * 8 - (words ? unused_bytes_of_last_word : 0);
    flush(layouter, true, bytes);
    return Blake2fDigest{chip_.digest(layouter, state_)};
  }
  // Convenience function to compute hash of the data.
  static Blake2fDigest digest(Chip chip, Layouter& layouter, const std::vector<BlockWord>& data,
                              size_t unused_bytes_of_last_word = 0, uint32_t rounds = BLAKE2B_ROUNDS) {
    Blake2f hasher = new_(std::move(chip), layouter.namespace_("init"), rounds);
    hasher.update(layouter.namespace_("update"), data);
    return hasher.finalize(layouter.namespace_("finalize"), unused_bytes_of_last_word);
  }

 private:
  explicit Blake2f(Chip chip) : chip_(std::move(chip)) {}
  void flush(Layouter& layouter, bool last, uint64_t block_bytes = 128) {
    State st = chip_.initialization(layouter, state_);
    st.t = bytes_done_ + block_bytes;  // BLAKE2b's counter includes the block being compressed
    st.last = last;
    std::array<BlockWord, BLOCK_SIZE> block;
    for (size_t i = 0; i < BLOCK_SIZE; i++) block[i] = cur_block_[i];
    state_ = chip_.compress(layouter, st, block);
    bytes_done_ += block_bytes;
    cur_block_.clear();
  }
  Chip chip_;
  State state_;
  std::vector<BlockWord> cur_block_;
  uint64_t bytes_done_ = 0;
  size_t length_ = 0;
};

// ---- halo2_proofs::plonk::Circuit -------------------------------------------------------------------------
class Circuit {
 public:
  using Config = Table16Config;
  virtual ~Circuit() = default;
  virtual std::unique_ptr<Circuit> without_witnesses() const = 0;
  virtual Config configure(ConstraintSystem& meta) const { return Table16Chip::configure(meta); }
  virtual void synthesize(const Config& config, Layouter& layouter) const = 0;
};

// The reference's (commented) test circuit, src/blake2f.rs:184-279, over a batch: one compression region
// per EIP-152 input.
class Blake2fCircuit : public Circuit {
 public:
  explicit Blake2fCircuit(std::vector<Value<Blake2fWitness>> inputs, uint32_t rounds = BLAKE2B_ROUNDS)
      : inputs_(std::move(inputs)), rounds_(rounds) {}
  std::unique_ptr<Circuit> without_witnesses() const override {
    return std::make_unique<Blake2fCircuit>(std::vector<Value<Blake2fWitness>>(inputs_.size()), rounds_);
  }
  void synthesize(const Config& config, Layouter& layouter) const override {
    Table16Chip::load(config, layouter);
    const Table16Chip chip = Table16Chip::construct(config);
    for (const auto& in : inputs_) {
      State st;
      st.rounds = in ? in->rounds : rounds_;
      std::array<BlockWord, BLOCK_SIZE> block;
      if (in) {
        for (size_t i = 0; i < 8; i++) st.h[i] = in->h[i];
        for (size_t i = 0; i < BLOCK_SIZE; i++) block[i] = BlockWord(in->m[i]);
        st.t = in->t[0];
        if (in->t[1]) throw Error(Error::Synthesis, ZK_E_INPUT, "byte counters above 2^64 are not laid out");
        st.last = in->f;
      } else {
        for (auto& b : block) b = BlockWord::unknown();
      }
      st = chip.initialization(layouter, st);
      const State out = chip.compress(layouter, st, block);
      (void)chip.digest(layouter, out);
    }
  }

 private:
  std::vector<Value<Blake2fWitness>> inputs_;
  uint32_t rounds_;
};

// ---- halo2_proofs::poly::commitment::Params<EqAffine> ---------------------------------------------------------
// The reference seeds its RNG with this array (benchmarking/src/blake2f_circuit_bench.rs:41-44).
using Seed = std::array<uint8_t, 16>;
constexpr Seed REFERENCE_SEED = {0x59, 0x62, 0xbe, 0x5d, 0x76, 0x3d, 0x31, 0x8d, 0x17, 0xdb, 0x37, 0x32, 0x54, 0x06, 0xbc, 0xe5};

class Params {
 public:
  // `Params::new(k)` (benches/blake2f.rs:85).  halo2 derives the generators by hashing to the curve, which
  // cannot be reproduced offline: this is the substitute URS (known discrete logs — benchmark / parity
  // only); load a genuine params file with `read`.
  static Params new_(const Device& dev, uint32_t k, const Seed& seed = REFERENCE_SEED) {
    dev.check(zk_params_generate_substitute(dev.ctx(), int32_t(k), seed.data()), "Params::new");
    return Params(dev, k);
  }
  // `Params::read` (benches/blake2f.rs:92-97): halo2's params file format
  static Params read(const Device& dev, const std::vector<uint8_t>& bytes) {
    dev.check(zk_params_load(dev.ctx(), bytes.data(), bytes.size()), "Params::read");
    if (bytes.size() < 4) throw Error(Error::Backend, ZK_E_INVALID, "Params::read: truncated");
    const uint32_t k = uint32_t(bytes[0]) | uint32_t(bytes[1]) << 8 | uint32_t(bytes[2]) << 16 | uint32_t(bytes[3]) << 24;
    return Params(dev, k);
  }
  std::vector<uint8_t> write() const {
    uint64_t len = 0;
    zk_params_write(dev_.ctx(), nullptr, &len);
    std::vector<uint8_t> out(len);
    dev_.check(zk_params_write(dev_.ctx(), out.data(), &len), "Params::write");
    out.resize(len);
    return out;
  }
  uint32_t k() const { return k_; }
  const Device& device() const { return dev_; }

 private:
  Params(const Device& d, uint32_t k) : dev_(d), k_(k) {}
  Device dev_;
  uint32_t k_;
};

// what `synthesize` laid out, collected by a recording layouter
inline Layouter lay_out(const Circuit& circuit, bool witnesses_known) {
  ConstraintSystem meta;
  const Circuit::Config config = circuit.configure(meta);
  Layouter layouter(witnesses_known);
  circuit.synthesize(config, layouter);
  if (layouter.regions == 0) layouter.rounds = BLAKE2B_ROUNDS;
  return layouter;
}

class VerifyingKey {
 public:
  // fixed + permutation commitments followed by transcript_repr (zk_vk_bytes)
  std::vector<uint8_t> bytes;
  uint32_t rounds = 0;
  uint64_t regions = 0;
};
class ProvingKey {
 public:
  const VerifyingKey& get_vk() const { return vk_; }
  VerifyingKey vk_;
};

// `keygen_vk(&params, &empty_circuit)` (benches/blake2f.rs:102): the keys live in the device context
inline VerifyingKey keygen_vk(const Params& params, const Circuit& circuit) {
  const Layouter shape = lay_out(*circuit.without_witnesses(), false);
  const Device& dev = params.device();
  dev.check(zk_blake2f_keygen(dev.ctx(), shape.rounds, shape.regions), "keygen_vk");
  VerifyingKey vk;
  uint64_t len = 0;
  zk_vk_bytes(dev.ctx(), nullptr, &len);
  vk.bytes.resize(len);
  dev.check(zk_vk_bytes(dev.ctx(), vk.bytes.data(), &len), "keygen_vk");
  vk.bytes.resize(len);
  vk.rounds = shape.rounds;
  vk.regions = shape.regions;
  return vk;
}
// `keygen_pk(&params, vk, &empty_circuit)` (benches/blake2f.rs:103)
inline ProvingKey keygen_pk(const Params& params, VerifyingKey vk, const Circuit& circuit) {
  const Layouter shape = lay_out(*circuit.without_witnesses(), false);
  if (shape.rounds != vk.rounds || shape.regions != vk.regions)
    throw Error(Error::Synthesis, ZK_E_STATE, "keygen_pk: circuit differs from the verifying key's");
  (void)params;
  return ProvingKey{std::move(vk)};
}

// ---- transcripts (Blake2bWrite / Blake2bRead with Challenge255, benches/blake2f.rs:121-123, 139-140) ------------
// The BLAKE2b transcript itself runs inside the library; these carry the proof bytes in and out.
class Blake2bWrite {
 public:
  static Blake2bWrite init(std::vector<uint8_t> sink = {}) { return Blake2bWrite{std::move(sink)}; }
  std::vector<uint8_t> finalize() { return std::move(buf); }
  std::vector<uint8_t> buf;
};
class Blake2bRead {
 public:
  static Blake2bRead init(const std::vector<uint8_t>& proof) { return Blake2bRead{proof}; }
  std::vector<uint8_t> proof;
};
// rand_xorshift's XorShiftRng::from_seed (benchmarking/src/blake2f_circuit_bench.rs:41-44)
struct XorShiftRng {
  Seed seed = REFERENCE_SEED;
  static XorShiftRng from_seed(const Seed& s) { return XorShiftRng{s}; }
};

// `create_proof(&params, &pk, &[circuit], &[&[]], rng, &mut transcript)` (benches/blake2f.rs:124-127).
// One circuit, no instance columns (the BLAKE2f circuit has none).
inline void create_proof(const Params& params, const ProvingKey& pk, const std::vector<const Circuit*>& circuits,
                         const std::vector<std::vector<std::vector<uint8_t>>>& instances, XorShiftRng rng,
                         Blake2bWrite& transcript) {
  if (circuits.size() != 1) throw Error(Error::Backend, ZK_E_INVALID, "create_proof: one circuit per proof");
  for (const auto& per_circuit : instances)
    if (!per_circuit.empty()) throw Error(Error::Backend, ZK_E_INVALID, "create_proof: the circuit has no instance columns");
  const Layouter laid = lay_out(*circuits[0], true);
  if (laid.rounds != pk.get_vk().rounds || laid.regions != pk.get_vk().regions)
    throw Error(Error::Synthesis, ZK_E_STATE, "create_proof: circuit differs from the proving key's");
  const Device& dev = params.device();
  std::vector<uint8_t> proof(8192);  // a proof of this circuit is (23 + 2k) points and ~70 scalars: 4-5 KB
  uint64_t len = proof.size();
  int32_t rc = zk_create_proof(dev.ctx(), laid.records.data(), laid.regions, rng.seed.data(), proof.data(), &len);
  if (rc == ZK_E_BUFFER) {  // the required size was written back
    proof.resize(len);
    rc = zk_create_proof(dev.ctx(), laid.records.data(), laid.regions, rng.seed.data(), proof.data(), &len);
  }
  dev.check(rc, "create_proof");
  proof.resize(len);
  transcript.buf.insert(transcript.buf.end(), proof.begin(), proof.end());
}

// `SingleVerifier::new(&params)` / `verify_proof(&params, pk.get_vk(), strategy, &[&[]], &mut transcript)`
// (benches/blake2f.rs:138-144).  Throws Error::Opening (or ConstraintSystemFailure) on rejection, like the
// `Err` the reference unwraps.
struct SingleVerifier {
  static SingleVerifier new_(const Params&) { return SingleVerifier{}; }
};
inline void verify_proof(const Params& params, const VerifyingKey& vk, SingleVerifier,
                         const std::vector<std::vector<std::vector<uint8_t>>>& instances, Blake2bRead& transcript) {
  for (const auto& per_circuit : instances)
    if (!per_circuit.empty()) throw Error(Error::Backend, ZK_E_INVALID, "verify_proof: the circuit has no instance columns");
  const Device& dev = params.device();
  uint64_t len = 0;
  zk_vk_bytes(dev.ctx(), nullptr, &len);
  std::vector<uint8_t> cur(len);
  dev.check(zk_vk_bytes(dev.ctx(), cur.data(), &len), "verify_proof");
  cur.resize(len);
  if (cur != vk.bytes) throw Error(Error::Backend, ZK_E_STATE, "verify_proof: the context holds another verifying key");
  dev.check(zk_verify_proof(dev.ctx(), transcript.proof.data(), transcript.proof.size()), "verify_proof", Error::Opening);
}

// ---- halo2_proofs::dev::MockProver (table16/spread_table.rs:759-763) --------------------------------------------
struct VerifyFailure {
  enum Kind { ConstraintNotSatisfied = 1, Lookup = 2, Permutation = 3 } kind;
  uint64_t row;
  uint64_t index;  // gate / copy-constraint index
  std::string description;
};
class MockProver {
 public:
  // `MockProver::run(k, &circuit, vec![])`: lays the circuit out; NotEnoughRowsAvailable if it does not fit 2^k
  // rows.  The checker keeps its own context on the device of `on`, so a prover's params and keys stay.
  static MockProver run(const Device& on, uint32_t k, const Circuit& circuit,
                        const std::vector<std::vector<uint8_t>>& instance = {}) {
    if (!instance.empty()) throw Error(Error::Backend, ZK_E_INVALID, "MockProver::run: the circuit has no instance columns");
    const Device dev(on.device_id());
    MockProver p(dev, k, lay_out(circuit, true));
    int32_t need = 0;
    dev.check(zk_blake2f_min_k(p.laid_.rounds, p.laid_.regions, &need), "MockProver::run");
    if (uint32_t(need) > k) throw Error(Error::NotEnoughRowsAvailable, ZK_E_ROWS, "MockProver::run: circuit does not fit 2^k rows");
    return p;
  }
  // `prover.verify()`: Ok(()) = empty vector; otherwise the first failure
  std::vector<VerifyFailure> verify() const {
    dev_.check(zk_params_generate_substitute(dev_.ctx(), int32_t(k_), REFERENCE_SEED.data()), "MockProver::verify");
    dev_.check(zk_blake2f_keygen(dev_.ctx(), laid_.rounds, laid_.regions), "MockProver::verify");
    uint64_t failure[3] = {0, 0, 0};
    const int32_t rc = zk_mock_verify(dev_.ctx(), laid_.records.data(), laid_.regions, nullptr, failure);
    if (rc == ZK_OK) return {};
    if (rc != ZK_E_VERIFY) dev_.check(rc, "MockProver::verify");
    const char* msg = zk_last_error(dev_.ctx());
    return {VerifyFailure{VerifyFailure::Kind(failure[0]), failure[1], failure[2], msg ? msg : ""}};
  }

 private:
  MockProver(const Device& d, uint32_t k, Layouter l) : dev_(d), k_(k), laid_(std::move(l)) {}
  Device dev_;
  uint32_t k_;
  Layouter laid_;
};

}  // namespace zkodst
