"""ctypes wrapper of the CPU oracle (oracle/libzkoracle.so).  Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
_L = None


def build():
    subprocess.run(["make", "-C", ODIR, "-s"], check=True)
    return os.path.join(ODIR, "libzkoracle.so")


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        c = ctypes
        vp, u64, u32, sz = c.c_void_p, c.c_uint64, c.c_uint32, c.c_size_t
        lib.zko_blake2f_F.argtypes = [c.c_char_p, c.c_char_p]
        lib.zko_blake2b.argtypes = [c.c_char_p, c.c_char_p, sz, c.c_char_p]
        lib.zko_xorshift_u64.argtypes = [c.c_char_p, sz, vp]
        lib.zko_field_consts.argtypes = [c.c_int, vp]
        lib.zko_field_op.argtypes = [c.c_int, c.c_int, vp, vp, vp]
        lib.zko_field_from_u512.argtypes = [c.c_int, vp, vp]
        lib.zko_rows_per_compression.restype = u64
        lib.zko_rows_per_compression.argtypes = [u32]
        for f in ("zko_spread16", "zko_get_tag", "zko_even_bits32", "zko_odd_bits32"):
            getattr(lib, f).restype = u32
            getattr(lib, f).argtypes = [u32]
        lib.zko_blake2f_witness.argtypes = [c.c_int, u32, c.c_char_p, sz, vp, vp, vp]
        lib.zko_mock_verify_raw.argtypes = [c.c_int, u32, sz, vp, c.c_char_p, sz]
        lib.zko_mock_verify_mont.argtypes = [c.c_int, u32, sz, vp, c.c_char_p, sz]
        lib.zko_mock_verify_mont_chained.argtypes = [c.c_int, u32, sz, c.c_char_p, vp, c.c_char_p, sz]
        lib.zko_layout_hash.argtypes = [u32, c.POINTER(u64), c.POINTER(u64), c.POINTER(u64)]
        lib.zko_describe_circuit.argtypes = [c.c_int, u32, sz, c.c_char_p, sz]

    # -- BLAKE2 ---------------------------------------------------------------------------
    def blake2f(self, record):
        out = ctypes.create_string_buffer(64)
        rc = self.lib.zko_blake2f_F(record, out)
        return rc, out.raw

    def blake2b(self, personal, data):
        out = ctypes.create_string_buffer(64)
        self.lib.zko_blake2b(personal, data, len(data), out)
        return out.raw

    def xorshift(self, seed, n):
        out = np.zeros(n, dtype=np.uint64)
        self.lib.zko_xorshift_u64(seed, n, out.ctypes.data)
        return out

    # -- fields ---------------------------------------------------------------------------
    @staticmethod
    def _limbs(v):
        return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)

    @staticmethod
    def _int(a):
        return sum(int(a[i]) << (64 * i) for i in range(len(a)))

    def consts(self, which):
        out = np.zeros(36, dtype=np.uint64)
        self.lib.zko_field_consts(which, out.ctypes.data)
        names = ["MOD", "R", "R2", "R3", "INV", "GENERATOR", "ROOT_OF_UNITY", "DELTA", "ZETA"]
        return {n: self._int(out[4 * i:4 * i + 4]) for i, n in enumerate(names)}

    def field_op(self, which, op, a, b=0):
        out = np.zeros(4, dtype=np.uint64)
        la, lb = self._limbs(a), self._limbs(b)
        rc = self.lib.zko_field_op(which, op, la.ctypes.data, lb.ctypes.data, out.ctypes.data)
        return rc, self._int(out)

    def from_u512(self, which, v):
        out = np.zeros(4, dtype=np.uint64)
        lv = np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(8)], dtype=np.uint64)
        self.lib.zko_field_from_u512(which, lv.ctypes.data, out.ctypes.data)
        return self._int(out)

    # -- circuit --------------------------------------------------------------------------
    def rows_per_compression(self, rounds):
        return self.lib.zko_rows_per_compression(rounds)

    def witness(self, k, rounds, inputs, n, mont=True, raw=False):
        nrows = 1 << k
        a_m = np.zeros((12, nrows, 4), dtype=np.uint64) if mont else None
        a_r = np.zeros((12, nrows), dtype=np.uint64) if raw else None
        dig = np.zeros((n, 8), dtype=np.uint64)
        rc = self.lib.zko_blake2f_witness(
            k, rounds, inputs, n, a_m.ctypes.data if mont else None,
            a_r.ctypes.data if raw else None, dig.ctypes.data)
        if rc:
            raise RuntimeError(f"zko_blake2f_witness rc={rc}")
        return a_m, a_r, dig

    def mock_verify_raw(self, k, rounds, n, advice_raw):
        msg = ctypes.create_string_buffer(512)
        rc = self.lib.zko_mock_verify_raw(k, rounds, n, advice_raw.ctypes.data, msg, 512)
        return rc, msg.value.decode()

    def mock_verify_mont(self, k, rounds, n, advice_mont, chain=None):
        msg = ctypes.create_string_buffer(512)
        rc = self.lib.zko_mock_verify_mont_chained(k, rounds, n, bytes(chain) if chain is not None else None,
                                                   advice_mont.ctypes.data, msg, 512)
        return rc, msg.value.decode()

    def layout_hash(self, rounds):
        a, b, n = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        rc = self.lib.zko_layout_hash(rounds, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n))
        assert rc == 0
        return a.value, b.value, n.value

    def describe(self, k, rounds, n):
        buf = ctypes.create_string_buffer(1 << 16)
        rc = self.lib.zko_describe_circuit(k, rounds, n, buf, 1 << 16)
        assert rc == 0, buf.value
        return dict(line.split("=", 1) for line in buf.value.decode().strip().split("\n"))


def load():
    global _L
    if _L is None:
        path = os.path.join(ODIR, "libzkoracle.so")
        if not os.path.exists(path) or any(
            os.path.getmtime(os.path.join(ODIR, f)) > os.path.getmtime(path)
            for f in os.listdir(ODIR) if f.endswith((".hpp", ".cpp"))
        ):
            try:
                build()
            except Exception:
                if not os.path.exists(path):
                    raise
        _L = Oracle(ctypes.CDLL(path))
    return _L


# ---- prover-level helpers (appended) ----------------------------------------------------------
def _setup_prover_api(lib):
    c = ctypes
    vp, sz, u32 = c.c_void_p, c.c_size_t, c.c_uint32
    lib.zko_prover_new_substitute.restype = vp
    lib.zko_prover_new_substitute.argtypes = [c.c_int, c.c_char_p]
    lib.zko_prover_new_from_params.restype = vp
    lib.zko_prover_new_from_params.argtypes = [c.c_char_p, sz]
    lib.zko_prover_free.argtypes = [vp]
    lib.zko_params_write.restype = sz
    lib.zko_params_write.argtypes = [vp, vp, sz]
    lib.zko_params_points.argtypes = [vp, c.c_int, vp]
    lib.zko_keygen.argtypes = [vp, u32, sz]
    lib.zko_keygen_vk.argtypes = [vp, u32, sz]
    lib.zko_vk_points.restype = sz
    lib.zko_vk_points.argtypes = [vp, vp]
    lib.zko_vk_pinned_debug.restype = sz
    lib.zko_vk_pinned_debug.argtypes = [vp, c.c_char_p, sz]
    lib.zko_keygen_chained.argtypes = [vp, u32, sz, c.c_char_p, c.c_int]
    lib.zko_vk_bytes.restype = sz
    lib.zko_vk_bytes.argtypes = [vp, vp, sz]
    lib.zko_create_proof.argtypes = [vp, c.c_char_p, sz, c.c_char_p, vp, c.POINTER(sz)]
    lib.zko_verify_proof.argtypes = [vp, c.c_char_p, sz, c.c_char_p, sz]
    lib.zko_msm.argtypes = [vp, vp, sz, vp]
    lib.zko_ntt.argtypes = [vp, c.c_int, c.c_int]
    lib.zko_coeff_to_extended.argtypes = [vp, c.c_int, c.c_int, vp]
    lib.zko_random_fields.argtypes = [c.c_char_p, sz, vp]
    lib.zko_set_threads.argtypes = [c.c_int]
    lib.zko_last_proof_ms.argtypes = [vp, c.POINTER(c.c_double)]


class OracleProver:
    """Params + keygen + create_proof + verify_proof of the CPU oracle."""

    def __init__(self, oracle, k=None, seed=None, params_bytes=None):
        self.o = oracle
        _setup_prover_api(oracle.lib)
        if params_bytes is not None:
            self.h = oracle.lib.zko_prover_new_from_params(params_bytes, len(params_bytes))
        else:
            self.h = oracle.lib.zko_prover_new_substitute(k, seed)
        assert self.h
        self.k = k

    def close(self):
        if self.h:
            self.o.lib.zko_prover_free(self.h)
            self.h = None

    def params_bytes(self):
        need = self.o.lib.zko_params_write(self.h, None, 0)
        buf = np.zeros(need, dtype=np.uint8)
        self.o.lib.zko_params_write(self.h, buf.ctypes.data, need)
        return buf.tobytes()

    def points(self, which, n):
        out = np.zeros((2 if which == 2 else n, 8), dtype=np.uint64)
        self.o.lib.zko_params_points(self.h, which, out.ctypes.data)
        return out

    def keygen(self, rounds, n_compressions):
        rc = self.o.lib.zko_keygen(self.h, rounds, n_compressions)
        assert rc == 0, rc

    def keygen_chained(self, rounds, n_compressions, chain, vk_only=False):
        """keygen with chain[j] != 0 meaning compression j continues compression j - 1."""
        rc = self.o.lib.zko_keygen_chained(self.h, rounds, n_compressions, bytes(chain), 1 if vk_only else 0)
        assert rc == 0, rc

    def keygen_vk(self, rounds, n_compressions):
        """keygen_vk only (commitments + transcript_repr): enough for vk_bytes() and verify()."""
        rc = self.o.lib.zko_keygen_vk(self.h, rounds, n_compressions)
        assert rc == 0, rc

    def vk_points(self):
        """The fixed then the permutation commitments as (count, 8) u64: affine Montgomery x, y."""
        out = np.zeros((64, 8), dtype=np.uint64)
        cnt = self.o.lib.zko_vk_points(self.h, out.ctypes.data)
        return np.ascontiguousarray(out[:cnt])

    def vk_pinned_debug(self):
        """The Rust `{:?}` rendering of vk.pinned() as the oracle derives it from its constraint system."""
        need = self.o.lib.zko_vk_pinned_debug(self.h, None, 0)
        buf = ctypes.create_string_buffer(need)
        self.o.lib.zko_vk_pinned_debug(self.h, buf, need)
        return buf.raw[:need].decode()

    def vk_bytes(self):
        need = self.o.lib.zko_vk_bytes(self.h, None, 0)
        buf = np.zeros(need, dtype=np.uint8)
        self.o.lib.zko_vk_bytes(self.h, buf.ctypes.data, need)
        return buf.tobytes()

    def create_proof(self, inputs, n, seed):
        buf = np.zeros(1 << 16, dtype=np.uint8)
        ln = ctypes.c_size_t(buf.size)
        rc = self.o.lib.zko_create_proof(self.h, inputs, n, seed, buf.ctypes.data, ctypes.byref(ln))
        assert rc == 0, rc
        return buf[:ln.value].tobytes()

    def last_proof_ms(self):
        """(witness synthesis ms, rest of create_proof ms) of the last create_proof."""
        out = (ctypes.c_double * 2)()
        self.o.lib.zko_last_proof_ms(self.h, out)
        return out[0], out[1]

    def verify(self, proof):
        msg = ctypes.create_string_buffer(256)
        rc = self.o.lib.zko_verify_proof(self.h, proof, len(proof), msg, 256)
        return rc, msg.value.decode()


def random_fields(oracle, seed, n):
    _setup_prover_api(oracle.lib)
    out = np.zeros((n, 4), dtype=np.uint64)
    oracle.lib.zko_random_fields(seed, n, out.ctypes.data)
    return out


def msm(oracle, scalars, bases):
    _setup_prover_api(oracle.lib)
    out = np.zeros(8, dtype=np.uint64)
    oracle.lib.zko_msm(scalars.ctypes.data, bases.ctypes.data, len(scalars), out.ctypes.data)
    return out


def ntt(oracle, data, log_n, inverse):
    _setup_prover_api(oracle.lib)
    d = np.ascontiguousarray(data).copy()
    oracle.lib.zko_ntt(d.ctypes.data, log_n, 1 if inverse else 0)
    return d


def set_threads(oracle, t):
    """Worker threads of the oracle (0 = all hardware threads); returns the count in use."""
    _setup_prover_api(oracle.lib)
    return oracle.lib.zko_set_threads(t)


def coeff_to_extended(oracle, coeffs, k, cs_degree=4):
    """halo2 `coeff_to_extended`: n coefficients -> 4n evaluations on zeta * <omega_4n>."""
    _setup_prover_api(oracle.lib)
    n = 1 << k
    out = np.zeros((n << 2, 4), dtype=np.uint64)
    c = np.ascontiguousarray(coeffs)
    oracle.lib.zko_coeff_to_extended(c.ctypes.data, k, cs_degree, out.ctypes.data)
    return out
