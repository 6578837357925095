"""The completed BLAKE2f circuit in the oracle: EIP-152 outputs appear in the digest cells and
the MockProver-equivalent check (the reference's only live assertion,
spread_table.rs:759-763) passes; tampering is rejected."""
import json
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(__file__), "golden")
VECS = json.load(open(os.path.join(G, "eip152.json")))


@pytest.mark.parametrize("idx", [0, 1, 2, 3, 5])
def test_vector_mock_prover(oracle, idx):
    v = VECS[idx]
    rec = bytes.fromhex(v["input"])
    rounds = int.from_bytes(rec[:4], "big")
    _, raw, dig = oracle.witness(17, rounds, rec, 1, mont=False, raw=True)
    assert dig.tobytes().hex() == v["output"]
    rc, msg = oracle.mock_verify_raw(17, rounds, 1, raw)
    assert rc == 0, msg


def test_rows_and_capacity(oracle, zk):
    assert oracle.rows_per_compression(12) == 4996 == zk.rows_per_compression(12)
    assert oracle.rows_per_compression(0) == 292
    assert zk.min_k(12, 1) == 17 and zk.min_k(12, 26) == 17 and zk.min_k(12, 27) == 18
    assert zk.min_k(12, 64) == 19 and zk.min_k(12, 256) == 21 and zk.min_k(12, 1024) == 23


def test_batch_mock_prover_and_tamper(oracle, zk):
    n = 3
    inputs = zk.synthetic_inputs(n)
    _, raw, dig = oracle.witness(17, 12, inputs, n, mont=False, raw=True)
    for i in range(n):
        rc, out = oracle.blake2f(inputs[213 * i:213 * (i + 1)])
        assert rc == 0 and out == dig[i].tobytes()
    assert oracle.mock_verify_raw(17, 12, n, raw)[0] == 0
    R = 4996
    # a lookup cell, a carry cell, a copied cell, a digest word: each must be caught
    for col, row in [(8, R + 1000), (6, 165), (1, 2 * R + 164), (0, 3 * R - 6)]:
        bad = raw.copy()
        bad[col, row] ^= 1
        rc, msg = oracle.mock_verify_raw(17, 12, n, bad)
        assert rc == 1, (col, row, msg)


def test_layout_matches_product_tables(oracle, zk):
    for rounds in (0, 1, 2, 12):
        assert oracle.layout_hash(rounds) == zk.layout_hash(rounds)


def test_circuit_description(oracle):
    d = oracle.describe(17, 12, 1)
    assert d["degree"] == "4" and d["blinding_factors"] == "5"
    # 3 table columns, the constants column, 8 columns from compress_selectors
    assert d["num_advice"] == "12" and d["num_fixed"] == "12"
    assert d["permutation_columns"] == "8;9;1;2;0;3;4;5;"
    assert d["n_polys"] == "26"
    assert d["fixed_queries"] == "".join("%d,0;" % c for c in range(12))
    # selector:fixed column:root:combination length — {a1} {b1,c1} {d1,b2,d2} {a2} {c2,abcd} {efgh,ijkl}
    # {digest,const} {fmask} (docs/CIRCUIT.md "Selectors")
    assert d["selectors"] == ("0:4:1:1;1:5:1:2;2:5:2:2;3:6:1:3;5:6:2:3;7:6:3:3;4:7:1:1;6:8:1:2;8:8:2:2;"
                              "9:9:1:2;10:9:2:2;11:10:1:2;12:10:2:2;13:11:1:1;")


def test_pinned_inputs_are_constrained(oracle, zk):
    """The IV words and the final-flag mask are no longer free witnesses: changing an IV word cell (and its
    limbs, consistently) or giving the mask a value other than 0 / 2^64 - 1 must fail a gate."""
    rec = zk.synthetic_inputs(1)
    _, raw, _ = oracle.witness(17, 12, rec, 1, mont=False, raw=True)
    assert oracle.mock_verify_raw(17, 12, 1, raw)[0] == 0
    # IV_0 lives in the S_ABCD slot at rows 32..35: word in a_3 (column 1) row 33, limbs in a_1 (column 8)
    bad = raw.copy()
    bad[1, 33] ^= 1          # word
    bad[8, 32] ^= 1          # limb 0 (dense), keeps `decompose ABCD` satisfied
    bad[9, 32] ^= 1          # its spread form (bit 0 spreads to bit 0)
    rc, msg = oracle.mock_verify_raw(17, 12, 1, bad)
    assert rc == 1 and "pin constant" in msg, msg
    # final-flag mask slot: rows 35 * 4 - 4 = 136..139 (h 8, IV 8, m 16, t0, t1 precede it); bit in a_9 (column 6)
    bad = raw.copy()
    bad[6, 137] ^= 1
    rc, msg = oracle.mock_verify_raw(17, 12, 1, bad)
    assert rc == 1 and "final flag" in msg, msg


def test_chained_records(oracle, zk):
    """Record chaining (CompressionConfig::initialize_with_state, compression.rs:1096-1111): the records of a
    multi-block hash satisfy the cross-region copies, unrelated records do not."""
    import hashlib
    msg = bytes(range(256)) + b"tail"            # 3 blocks
    records, digest = zk.blake2b_records(msg)
    assert digest == hashlib.blake2b(msg).digest()
    n = len(records) // 213
    assert n == 3
    adv, _, _ = oracle.witness(17, 12, records, n)
    chain = bytes([0, 1, 1])
    assert oracle.mock_verify_mont(17, 12, n, adv, chain)[0] == 0
    other = zk.synthetic_inputs(n)              # independent records: the chain copies cannot hold
    adv2, _, _ = oracle.witness(17, 12, other, n)
    assert oracle.mock_verify_mont(17, 12, n, adv2)[0] == 0
    rc, why = oracle.mock_verify_mont(17, 12, n, adv2, chain)
    assert rc == 1 and "Permutation" in why, why
    rc, why = oracle.mock_verify_mont(17, 12, n, adv, bytes([1, 0, 0]))   # the first record has no predecessor
    assert rc == -3 and "first compression" in why, why
