"""The prover half of the oracle (poly.hpp, curve.hpp, params.hpp, plonk.hpp) against arithmetic written
here in plain Python integers — a naive DFT, double-and-add on Vesta, the Lagrange / monomial basis
relation of the URS — and BASELINE configs[0] end to end on the CPU: one 12-round compression (EIP-152
vector 5) proved and verified at the reference's default k = 17."""
import json
import os
import random

import numpy as np
import pytest

import oracle_lib

G = os.path.join(os.path.dirname(__file__), "golden")
SEED = bytes([0x59, 0x62, 0xbe, 0x5d, 0x76, 0x3d, 0x31, 0x8d, 0x17, 0xdb, 0x37, 0x32, 0x54, 0x06, 0xbc, 0xe5])


def _ints(a):
    return [oracle_lib.Oracle._int(row) for row in a]


def _omega(oracle, log_n):
    c = oracle.consts(0)
    p, R = c["MOD"], c["R"]
    root = c["ROOT_OF_UNITY"]
    if pow(root, 1 << 32, p) != 1 or pow(root, 1 << 31, p) == 1:  # stored in Montgomery form
        root = root * pow(R, -1, p) % p
    assert pow(root, 1 << 32, p) == 1 and pow(root, 1 << 31, p) != 1
    return pow(root, 1 << (32 - log_n), p)


@pytest.mark.parametrize("log_n", [1, 3, 6])
def test_ntt_is_the_dft_over_omega(oracle, log_n):
    """best_fft's convention: natural order in and out, out[i] = sum_j in[j] omega^(ij);
    the inverse uses omega^-1 and divides by n."""
    c = oracle.consts(0)
    p, R = c["MOD"], c["R"]
    rinv = pow(R, -1, p)
    n = 1 << log_n
    w = _omega(oracle, log_n)
    data = oracle_lib.random_fields(oracle, SEED, n)
    x = [v * rinv % p for v in _ints(data)]
    fwd = [v * rinv % p for v in _ints(oracle_lib.ntt(oracle, data, log_n, False))]
    assert fwd == [sum(x[j] * pow(w, i * j, p) for j in range(n)) % p for i in range(n)]
    inv = [v * rinv % p for v in _ints(oracle_lib.ntt(oracle, data, log_n, True))]
    ninv, winv = pow(n, -1, p), pow(w, -1, p)
    assert inv == [sum(x[j] * pow(winv, i * j, p) for j in range(n)) * ninv % p for i in range(n)]


class Vesta:
    """y^2 = x^3 + 5 over Fq, in affine Python integers (None = identity)."""

    def __init__(self, q):
        self.q = q

    def add(self, a, b):
        q = self.q
        if a is None:
            return b
        if b is None:
            return a
        if a[0] == b[0]:
            if (a[1] + b[1]) % q == 0:
                return None
            lam = 3 * a[0] * a[0] * pow(2 * a[1], -1, q) % q
        else:
            lam = (b[1] - a[1]) * pow(b[0] - a[0], -1, q) % q
        x = (lam * lam - a[0] - b[0]) % q
        return (x, (lam * (a[0] - x) - a[1]) % q)

    def mul(self, s, pt):
        acc = None
        while s:
            if s & 1:
                acc = self.add(acc, pt)
            pt = self.add(pt, pt)
            s >>= 1
        return acc


def _points(oracle, arr):
    c = oracle.consts(1)
    q, rinv = c["MOD"], pow(c["R"], -1, c["MOD"])
    out = []
    for row in arr:
        x, y = oracle_lib.Oracle._int(row[:4]) * rinv % q, oracle_lib.Oracle._int(row[4:]) * rinv % q
        out.append(None if x == 0 and y == 0 else (x, y))
    return out


def test_msm_matches_double_and_add(oracle):
    """best_multiexp (curve.hpp) against schoolbook scalar multiplication; the URS points are on the curve."""
    p, q = oracle.consts(0)["MOD"], oracle.consts(1)["MOD"]
    rinv = pow(oracle.consts(0)["R"], -1, p)
    E = Vesta(q)
    prover = oracle_lib.OracleProver(oracle, k=4, seed=SEED)
    bases = prover.points(0, 16)
    prover.close()
    pts = _points(oracle, bases)
    for pt in pts:
        assert (pt[1] * pt[1] - pt[0] ** 3 - 5) % q == 0
    scalars = oracle_lib.random_fields(oracle, SEED, 16)
    scalars[3] = 0
    scalars[5] = oracle_lib.Oracle._limbs(oracle.field_op(0, 4, 1)[1])          # 1
    scalars[6] = oracle_lib.Oracle._limbs(oracle.field_op(0, 4, p - 1)[1])      # -1
    s = [v * rinv % p for v in _ints(scalars)]
    want = None
    for si, pt in zip(s, pts):
        want = E.add(want, E.mul(si, pt))
    got = _points(oracle, oracle_lib.msm(oracle, scalars, np.ascontiguousarray(bases)).reshape(1, 8))[0]
    assert got == want


def test_lagrange_basis_is_the_inverse_dft_of_the_monomial_basis(oracle):
    """Params: commit_lagrange(values) must equal commit(coefficients), i.e.
    g_lagrange[i] = sum_j (omega^(-ij) / n) g[j]  (halo2 `Params::new`: g_lagrange = ifft of g in the exponent)."""
    k, n = 3, 8
    p, q = oracle.consts(0)["MOD"], oracle.consts(1)["MOD"]
    E = Vesta(q)
    prover = oracle_lib.OracleProver(oracle, k=k, seed=SEED)
    g, gl = _points(oracle, prover.points(0, n)), _points(oracle, prover.points(1, n))
    prover.close()
    winv, ninv = pow(_omega(oracle, k), -1, p), pow(n, -1, p)
    for i in range(n):
        acc = None
        for j in range(n):
            acc = E.add(acc, E.mul(pow(winv, i * j, p) * ninv % p, g[j]))
        assert acc == gl[i], i


def test_config0_single_compression_prove_and_verify(oracle):
    """BASELINE configs[0]: EIP-152 vector 5 (the reference's own literal, src/blake2f.rs:193-247) through
    keygen, create_proof and verify_proof of the CPU oracle at the reference's default k = 17; a flipped bit
    anywhere in the proof is rejected, and the proof is a deterministic function of (records, seed)."""
    vecs = json.load(open(os.path.join(G, "eip152.json")))
    rec = next(bytes.fromhex(v["input"]) for v in vecs if v["input"].startswith("0000000c48c9bdf2") and v["input"].endswith("01"))
    prover = oracle_lib.OracleProver(oracle, k=17, seed=SEED)
    prover.keygen(12, 1)
    proof = prover.create_proof(rec, 1, SEED)
    assert 3000 < len(proof) < 8192
    rc, msg = prover.verify(proof)
    assert rc == 0, msg
    rnd = random.Random(7)
    for pos in (0, 40, len(proof) // 2, len(proof) - 1, rnd.randrange(len(proof))):
        bad = bytearray(proof)
        bad[pos] ^= 1
        assert prover.verify(bytes(bad))[0] != 0, pos
    assert prover.verify(proof[:-32])[0] != 0 and prover.verify(proof + b"\0" * 32)[0] != 0
    assert prover.create_proof(rec, 1, SEED) == proof
    prover.close()
