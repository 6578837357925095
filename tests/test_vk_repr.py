"""vk.transcript_repr as halo2 derives it: the Rust `{:?}` rendering of vk.pinned().
Two independent renderers must agree character by character — the library's string builder, which restates
`configure` (zk-odst_b200/csrc/vk_repr.cpp, through the host-only zk_blake2f_pinned_debug), and the oracle's
generic walk over its constraint-system expression trees (oracle/plonk.hpp) — and the transcript_repr both
publish must be the BLAKE2b hash halo2's `VerifyingKey::from_parts` takes of that string.  (Neither has been
compared with halo2 itself here: rust/xcheck does that where cargo exists.)"""
import hashlib

import numpy as np
import pytest

import oracle_lib


@pytest.fixture(scope="module")
def keys17(oracle, zk):
    op = oracle_lib.OracleProver(oracle, k=17, seed=zk.REFERENCE_SEED)
    op.keygen_vk(12, 2)
    yield op
    op.close()


def test_pinned_debug_strings_agree(keys17, zk):
    mine = zk.pinned_debug(17, 12, keys17.vk_points())
    ref = keys17.vk_pinned_debug()
    if mine != ref:
        i = next(i for i in range(min(len(mine), len(ref))) if mine[i] != ref[i])
        raise AssertionError("first difference at %d: %r vs %r" % (i, mine[i - 60:i + 60], ref[i - 60:i + 60]))
    assert mine.startswith('PinnedVerificationKey { base_modulus: "0x40000000000000000000000000000000224698fc0994a8dd'
                           '8c46eb2100000001", scalar_modulus: "0x40000000000000000000000000000000224698fc094cf91b'
                           '992d30ed00000001", domain: PinnedEvaluationDomain { k: 17, extended_k: 19, omega: 0x')
    assert ("cs: PinnedConstraintSystem { num_fixed_columns: 12, num_advice_columns: 12, num_instance_columns: 0, "
            "num_selectors: 14, gates: [Product(") in mine
    assert mine.count("Argument { input_expressions: [Advice { query_index: 0, column_index: 7,") == 1
    assert "constants: [], minimum_degree: None }, fixed_commitments: [(0x" in mine
    assert mine.endswith(")] } }")
    # other round counts change rows, not the constraint system: only the commitments may differ
    assert zk.pinned_debug(17, 1, keys17.vk_points()) == mine


def test_transcript_repr_is_the_hash_of_the_pinned_string(keys17, oracle):
    s = keys17.vk_pinned_debug().encode()
    digest = hashlib.blake2b(len(s).to_bytes(8, "little") + s, digest_size=64, person=b"Halo2-Verify-Key").digest()
    vk = keys17.vk_bytes()   # ... | transcript_repr (32 B canonical little-endian = from_uniform_bytes(digest))
    repr_le = int.from_bytes(vk[-32:], "little")
    p = oracle.consts(0)["MOD"]
    assert repr_le == int.from_bytes(digest, "little") % p


def test_layout_tables_export(zk, oracle):
    for rounds in (0, 1, 12):
        copies, sel, const, chain = zk.layout_tables(rounds)
        R = zk.rows_per_compression(rounds)
        assert sel.shape == (14, R) and const.shape == (R,) and copies.shape[1] == 4
        assert copies.shape[0] == zk.layout_hash(rounds)[2]
        assert list(chain[:8]) == [4 * i + 1 for i in range(8)]
        assert list(chain[8:]) == [R - 128 + 16 * i + 10 for i in range(8)]
        # s_const rows carry the IV words, every other row of the constants column is zero
        iv = [0x6a09e667f3bcc908, 0xbb67ae8584caa73b, 0x3c6ef372fe94f82b, 0xa54ff53a5f1d36f1,
              0x510e527fade682d1, 0x9b05688c2b3e6c1f, 0x1f83d9abfb41bd6b, 0x5be0cd19137e2179]
        rows = np.flatnonzero(sel[12])
        assert list(rows) == [32 + 4 * i + 1 for i in range(8)] and [int(const[r]) for r in rows] == iv
        assert np.count_nonzero(const) == 8
        assert list(np.flatnonzero(sel[13])) == [137]
