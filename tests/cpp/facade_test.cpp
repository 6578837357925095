// The reference's bench (blake2f-circuit/benches/blake2f.rs:83-142) and its commented test module
// (blake2f-circuit/src/blake2f.rs:184-279; MockProver call shape of table16/spread_table.rs:759-763)
// rewritten against include/zkodst.hpp.  Built and run by tests/test_facade.py (needs a GPU to run).
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "zkodst.hpp"

using namespace zkodst;

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      fprintf(stderr, "CHECK failed: %s (line %d)\n", #cond, __LINE__);    \
      return 1;                                                            \
    }                                                                      \
  } while (0)

static std::string hex(const uint8_t* p, size_t n) {
  static const char* d = "0123456789abcdef";
  std::string s;
  for (size_t i = 0; i < n; i++) {
    s += d[p[i] >> 4];
    s += d[p[i] & 15];
  }
  return s;
}

// EIP-152 test vector 5, the literal of src/blake2f.rs:193-247: h = BLAKE2b-512 IV ^ param block, m = "abc"
static Blake2fWitness vector5() {
  Blake2fWitness w;
  w.rounds = 12;
  w.h = {0x6a09e667f2bdc948ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
         0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
  w.m = {};
  w.m[0] = 0x0000000000636261ULL;
  w.t = {3, 0};
  w.f = true;
  return w;
}
static const char* VECTOR5_OUT =
    "ba80a53f981c4d0d6a2797b69f12f6e94c212f14685ac4b74b12bb6fdbffa2d1"
    "7d87c5392aab792dc252d5de4533cc9518d38aa8dbf1925ab92386edd4009923";

int main(int argc, char** argv) {
  const bool host_only = argc > 1 && std::string(argv[1]) == "--host-only";
  try {
    const uint32_t k = 17;

    // ---- the streaming gadget: Blake2f::digest over "abc" lays out one region and returns BLAKE2b-512("abc")
    {
      ConstraintSystem meta;
      const Table16Config config = Table16Chip::configure(meta);
      Layouter layouter(true);
      Table16Chip::load(config, layouter);
      const Blake2fDigest d =
          Blake2f<Table16Chip>::digest(Table16Chip::construct(config), layouter, {BlockWord(0x636261ULL)}, 5);
      CHECK(layouter.regions == 1);
      const auto bytes = d.bytes();
      CHECK(bytes.has_value());
      CHECK(hex(bytes->data(), 64) == VECTOR5_OUT);
      // the record the gadget laid out is the EIP-152 vector itself
      CHECK(layouter.records.size() == ZK_BLAKE2F_INPUT_BYTES);
      CHECK(hex(layouter.records.data(), 213) == hex(vector5().to_eip152().data(), 213));
      // 129 bytes -> two chained regions, the second flagged final
      std::vector<BlockWord> long_msg(17, BlockWord(0x0101010101010101ULL));
      Layouter l2(true);
      Table16Chip::load(config, l2);
      (void)Blake2f<Table16Chip>::digest(Table16Chip::construct(config), l2, long_msg, 7);
      CHECK(l2.regions == 2);
      CHECK(l2.records[212] == 0 && l2.records[213 + 212] == 1);
      CHECK(l2.records[196] == 128 && l2.records[213 + 196] == 129);
    }

    // ---- what a circuit lays out, with and without witnesses (no device involved)
    const Blake2fCircuit circuit({known(vector5()), known(vector5())});
    {
      const Layouter proving = lay_out(circuit, true);
      CHECK(proving.regions == 2 && proving.rounds == 12 && proving.records.size() == 2 * ZK_BLAKE2F_INPUT_BYTES);
      const Layouter keygen = lay_out(*circuit.without_witnesses(), false);
      CHECK(keygen.regions == 2 && keygen.rounds == 12 && keygen.records.empty());
      bool threw = false;
      try {  // an unknown witness cannot be proved: Error::Synthesis, as halo2 reports a missing Value
        (void)lay_out(*circuit.without_witnesses(), true);
      } catch (const Error& e) {
        threw = e.kind == Error::Synthesis;
      }
      CHECK(threw);
    }
    if (host_only) {
      printf("facade host part ok\n");
      return 0;
    }

    // ---- MockProver::run(k, &circuit, vec![]).verify() == Ok(())
    const Device dev(0);
    {
      const MockProver prover = MockProver::run(dev, k, circuit);
      const auto failures = prover.verify();
      CHECK(failures.empty());
      // too many regions for 2^17 rows -> NotEnoughRowsAvailable, as halo2 reports it
      bool threw = false;
      try {
        MockProver::run(dev, k, Blake2fCircuit(std::vector<Value<Blake2fWitness>>(27, known(vector5()))));
      } catch (const Error& e) {
        threw = e.kind == Error::NotEnoughRowsAvailable;
      }
      CHECK(threw);
    }

    // ---- benches/blake2f.rs: params, keys, proof, verification
    const Params params = Params::new_(dev, k);
    const auto empty_circuit = circuit.without_witnesses();
    const VerifyingKey vk = keygen_vk(params, *empty_circuit);
    const ProvingKey pk = keygen_pk(params, vk, *empty_circuit);

    Blake2bWrite transcript = Blake2bWrite::init();
    create_proof(params, pk, {&circuit}, {{}}, XorShiftRng::from_seed(REFERENCE_SEED), transcript);
    const std::vector<uint8_t> proof = transcript.finalize();
    CHECK(proof.size() > 3000 && proof.size() < 8192);

    {
      const SingleVerifier strategy = SingleVerifier::new_(params);
      Blake2bRead rd = Blake2bRead::init(proof);
      verify_proof(params, pk.get_vk(), strategy, {{}}, rd);  // throws on rejection
    }
    {  // a flipped bit is rejected with the error the reference's `unwrap()` would hit
      std::vector<uint8_t> bad = proof;
      bad[40] ^= 1;
      Blake2bRead rd = Blake2bRead::init(bad);
      bool rejected = false;
      try {
        verify_proof(params, pk.get_vk(), SingleVerifier::new_(params), {{}}, rd);
      } catch (const Error& e) {
        rejected = e.kind == Error::Opening && e.code == ZK_E_VERIFY;
      }
      CHECK(rejected);
    }
    {  // same seed, same records -> same bytes; the params file round-trips
      Blake2bWrite again = Blake2bWrite::init();
      create_proof(params, pk, {&circuit}, {{}}, XorShiftRng::from_seed(REFERENCE_SEED), again);
      CHECK(again.finalize() == proof);
      const std::vector<uint8_t> file = params.write();
      CHECK(file.size() == 4 + (2 * (size_t(1) << k) + 2) * 32);
    }
    {  // benches/blake2f.rs:79-97,119-136: params and proof cached on disk, read back, the cached proof verified
      const std::string dir = std::string(getenv("TMPDIR") ? getenv("TMPDIR") : "/tmp") + "/blake2f_assets";
      (void)!system(("mkdir -p " + dir).c_str());
      const std::string params_path = dir + "/blake2f_params", proof_path = dir + "/blake2f_proof";
      auto write_file = [](const std::string& path, const std::vector<uint8_t>& bytes) -> void {
        FILE* f = fopen(path.c_str(), "wb");
        if (!f || fwrite(bytes.data(), 1, bytes.size(), f) != bytes.size()) throw std::runtime_error("cannot write " + path);
        fclose(f);
      };
      auto read_file = [](const std::string& path) -> std::vector<uint8_t> {
        std::vector<uint8_t> bytes;
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("cannot read " + path);
        uint8_t buf[1 << 16];
        for (size_t got; (got = fread(buf, 1, sizeof buf, f)) > 0;) bytes.insert(bytes.end(), buf, buf + got);
        fclose(f);
        return bytes;
      };
      write_file(params_path, params.write());
      write_file(proof_path, proof);
      const Params cached = Params::read(dev, read_file(params_path));     // Params::read (benches/blake2f.rs:95-97)
      const VerifyingKey vk2 = keygen_vk(cached, *empty_circuit);
      const ProvingKey pk2 = keygen_pk(cached, vk2, *empty_circuit);
      const std::vector<uint8_t> cached_proof = read_file(proof_path);
      CHECK(cached_proof == proof);
      Blake2bRead rd = Blake2bRead::init(cached_proof);
      verify_proof(cached, pk2.get_vk(), SingleVerifier::new_(cached), {{}}, rd);
      remove(params_path.c_str());
      remove(proof_path.c_str());
    }
    {  // a circuit of another shape (11 rounds) is refused before anything reaches the device
      Blake2fWitness w = vector5();
      w.rounds = 11;
      const Blake2fCircuit other({known(w), known(w)}, 11);
      bool threw = false;
      try {
        Blake2bWrite t2 = Blake2bWrite::init();
        create_proof(params, pk, {&other}, {{}}, XorShiftRng{}, t2);
      } catch (const Error& e) {
        threw = e.kind == Error::Synthesis;
      }
      CHECK(threw);
    }
    printf("facade ok: %zu-byte proof\n", proof.size());
    return 0;
  } catch (const Error& e) {
    fprintf(stderr, "zkodst::Error kind %d code %d: %s\n", int(e.kind), e.code, e.what());
    return 2;
  }
}
