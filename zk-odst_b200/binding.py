"""ctypes binding of include/zkodst.h."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ERRORS = {
    -1: "ZK_E_INVALID", -2: "ZK_E_CUDA", -3: "ZK_E_NOMEM", -4: "ZK_E_ROWS", -5: "ZK_E_INPUT",
    -6: "ZK_E_STATE", -7: "ZK_E_VERIFY", -8: "ZK_E_BUFFER",
}


class ZkError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


def library_path():
    return os.path.join(_HERE, "libzkodst.so")


def load_library():
    """Loads libzkodst.so; fails loudly if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise ZkError(-6, f"{path} is missing: run `python zk-odst_b200/build.py` "
                              "(or __graft_entry__.build()); there is no CPU fallback")
        lib = ctypes.CDLL(path)
        c = ctypes
        vp, u64, i32, u32 = c.c_void_p, c.c_uint64, c.c_int32, c.c_uint32
        sigs = {
            "zk_ctx_create": (i32, [i32, c.POINTER(vp)]),
            "zk_ctx_destroy": (None, [vp]),
            "zk_last_error": (c.c_char_p, [vp]),
            "zk_ctx_set_stream": (i32, [vp, vp]),
            "zk_ctx_synchronize": (i32, [vp]),
            "zk_ctx_set_blocking_sync": (i32, [vp, i32]),
            "zk_ctx_launch_count": (u64, [vp]),
            "zk_ctx_last_kernel_ms": (i32, [vp, i32, c.POINTER(c.c_float)]),
            "zk_ctx_enable_timing": (i32, [vp, i32]),
            "zk_ctx_timing_report": (i32, [vp, c.POINTER(c.c_float), c.POINTER(u32)]),
            "zk_bench_int_pipe": (i32, [vp, i32, u32, c.POINTER(c.c_double)]),
            "zk_blake2f_rows_per_compression": (i32, [u32, c.POINTER(u64)]),
            "zk_blake2f_min_k": (i32, [u32, u64, c.POINTER(i32)]),
            "zk_blake2f_layout_hash": (i32, [u32, c.POINTER(u64), c.POINTER(u64), c.POINTER(u64)]),
            "zk_blake2f_layout_tables": (i32, [u32, vp, c.POINTER(u64), vp, vp, vp]),
            "zk_params_generate_substitute": (i32, [vp, i32, c.c_char_p]),
            "zk_params_load": (i32, [vp, vp, u64]),
            "zk_params_write": (i32, [vp, vp, c.POINTER(u64)]),
            "zk_blake2f_keygen": (i32, [vp, u32, u64]),
            "zk_blake2f_keygen_chained": (i32, [vp, u32, u64, c.c_char_p]),
            "zk_vk_bytes": (i32, [vp, vp, c.POINTER(u64)]),
            "zk_vk_repr_override": (i32, [vp, c.c_char_p]),
            "zk_vk_pinned_debug": (i32, [vp, vp, c.POINTER(u64)]),
            "zk_blake2f_pinned_debug": (i32, [i32, u32, vp, vp, c.POINTER(u64)]),
            "zk_create_proof": (i32, [vp, vp, u64, c.c_char_p, vp, c.POINTER(u64)]),
            "zk_create_proof_device_inputs": (i32, [vp, vp, u64, c.c_char_p, vp, c.POINTER(u64)]),
            "zk_msm_vesta": (i32, [vp, vp, vp, u64, i32, vp]),
            "zk_ntt_fp": (i32, [vp, vp, i32, i32, i32]),
            "zk_commit_batch": (i32, [vp, i32, vp, u32, vp, u32, i32, i32, vp]),
            "zk_ntt_fp_batch": (i32, [vp, vp, vp, i32, u32, i32, i32]),
            "zk_coeff_to_cosets": (i32, [vp, vp, u32, i32, vp]),
            "zk_blake2f_witness_batch": (i32, [vp, i32, u32, vp, u64, vp, vp]),
            "zk_blake2f_witness_batch_device": (i32, [vp, i32, u32, vp, u64, vp, vp]),
            "zk_eip152_validate": (i32, [c.c_char_p, u64, c.POINTER(u32)]),
            "zk_blake2f_compress": (i32, [c.c_char_p, c.c_char_p]),
            "zk_blake2b_records": (i32, [c.c_char_p, u64, u32, c.c_char_p, c.POINTER(u64), c.c_char_p]),
            "zk_verify_proof": (i32, [vp, c.c_char_p, u64]),
            "zk_verify_proofs_batch": (i32, [vp, c.c_char_p, c.POINTER(u64), u64, c.c_char_p]),
            "zk_mock_verify": (i32, [vp, vp, u64, vp, c.POINTER(u64)]),
            "zk_dist_unique_id": (i32, [c.c_char_p]),
            "zk_dist_init": (i32, [vp, c.c_char_p, i32, i32]),
            "zk_dist_info": (i32, [vp, c.POINTER(i32), c.POINTER(i32)]),
            "zk_dist_range": (i32, [u64, i32, i32, c.POINTER(u64), c.POINTER(u64)]),
            "zk_dist_quotient_rows": (i32, [u64, i32, i32, c.POINTER(u64), c.POINTER(u64), c.POINTER(u64),
                                            c.POINTER(c.c_uint32)]),
            "zk_dist_column_block": (i32, [i32, i32, c.POINTER(c.c_uint32), c.POINTER(c.c_uint32),
                                           c.POINTER(c.c_uint32)]),
        }
        for name, (res, args) in sigs.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def eip152_validate(record):
    """Round count of a well-formed EIP-152 input; raises ZkError(ZK_E_INPUT) otherwise."""
    r = ctypes.c_uint32()
    rc = load_library().zk_eip152_validate(bytes(record), len(record), ctypes.byref(r))
    if rc:
        raise ZkError(rc, "malformed EIP-152 input")
    return r.value


def blake2f_compress(record):
    out = ctypes.create_string_buffer(64)
    rc = load_library().zk_blake2f_compress(bytes(record), out)
    if rc:
        raise ZkError(rc, "malformed EIP-152 input")
    return out.raw


def blake2b_records(msg, rounds=12):
    """(records, digest): the EIP-152 record chain whose proof is a proof of BLAKE2b-512(msg)."""
    lib = load_library()
    msg = bytes(msg)
    n = ctypes.c_uint64(0)
    rc = lib.zk_blake2b_records(msg, len(msg), rounds, None, ctypes.byref(n), None)
    if rc:
        raise ZkError(rc)
    buf = ctypes.create_string_buffer(n.value * 213)
    dig = ctypes.create_string_buffer(64)
    rc = lib.zk_blake2b_records(msg, len(msg), rounds, buf, ctypes.byref(n), dig)
    if rc:
        raise ZkError(rc)
    return buf.raw, dig.raw


def dist_unique_id():
    buf = ctypes.create_string_buffer(128)
    rc = load_library().zk_dist_unique_id(buf)
    if rc:
        raise ZkError(rc, "zk_dist_unique_id failed (libnccl.so.2 missing?)")
    return buf.raw


def dist_range(n_points, rank, world):
    lo, hi = ctypes.c_uint64(), ctypes.c_uint64()
    rc = load_library().zk_dist_range(n_points, rank, world, ctypes.byref(lo), ctypes.byref(hi))
    if rc:
        raise ZkError(rc)
    return lo.value, hi.value


def dist_column_block(rank, world):
    """(lo, hi, per_rank): the witness-column slots `rank` of `world` transforms (zk_dist_column_block)."""
    lib = load_library()
    lo, hi, per = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    rc = lib.zk_dist_column_block(rank, world, ctypes.byref(lo), ctypes.byref(hi), ctypes.byref(per))
    if rc:
        raise ZkError(rc)
    return lo.value, hi.value, per.value


def dist_quotient_rows(n, rank, world):
    """((row_lo, row_hi), [(start, length), ...]): the quotient rows `rank` evaluates and the row segments
    of every coset column it reads (zk_dist_quotient_rows)."""
    lib = load_library()
    lo, hi, cnt = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32(16)
    segs = (ctypes.c_uint64 * 32)()
    rc = lib.zk_dist_quotient_rows(n, rank, world, ctypes.byref(lo), ctypes.byref(hi), segs, ctypes.byref(cnt))
    if rc:
        raise ZkError(rc)
    return (lo.value, hi.value), [(segs[2 * i], segs[2 * i + 1]) for i in range(cnt.value)]


def pinned_debug(k, rounds, commitments):
    """Host-only: the `{:?}` rendering of vk.pinned() for 20 commitments (12 fixed, 8 permutation) given as an
    array of 64-byte affine Montgomery points."""
    lib = load_library()
    ln = ctypes.c_uint64(0)
    lib.zk_blake2f_pinned_debug(k, rounds, _ptr(commitments), None, ctypes.byref(ln))
    buf = ctypes.create_string_buffer(ln.value)
    rc = lib.zk_blake2f_pinned_debug(k, rounds, _ptr(commitments), ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(ln))
    if rc:
        raise ZkError(rc)
    return buf.raw[:ln.value].decode()


def rows_per_compression(rounds):
    out = ctypes.c_uint64()
    rc = load_library().zk_blake2f_rows_per_compression(rounds, ctypes.byref(out))
    if rc:
        raise ZkError(rc)
    return out.value


def min_k(rounds, n_compressions):
    out = ctypes.c_int32()
    rc = load_library().zk_blake2f_min_k(rounds, n_compressions, ctypes.byref(out))
    if rc:
        raise ZkError(rc)
    return out.value


def layout_hash(rounds):
    a, b, n = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    rc = load_library().zk_blake2f_layout_hash(rounds, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n))
    if rc:
        raise ZkError(rc)
    return a.value, b.value, n.value


def layout_tables(rounds):
    """(copies (n, 4) u32, selectors (14, R) u8, constants (R,) u64, chain_rows (16,) u32) of one region."""
    import numpy as np
    lib = load_library()
    n = ctypes.c_uint64(0)
    rc = lib.zk_blake2f_layout_tables(rounds, None, ctypes.byref(n), None, None, None)
    if rc:
        raise ZkError(rc)
    R = rows_per_compression(rounds)
    copies = np.zeros((n.value, 4), dtype=np.uint32)
    sel = np.zeros((14, R), dtype=np.uint8)
    const = np.zeros(R, dtype=np.uint64)
    chain = np.zeros(16, dtype=np.uint32)
    rc = lib.zk_blake2f_layout_tables(rounds, copies.ctypes.data, ctypes.byref(n), sel.ctypes.data, const.ctypes.data,
                                      chain.ctypes.data)
    if rc:
        raise ZkError(rc)
    return copies, sel, const, chain


def _ptr(x):
    """Address of a bytes / numpy array / torch tensor / int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, (bytes, bytearray)):
        return ctypes.cast(ctypes.c_char_p(bytes(x)), ctypes.c_void_p).value
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(type(x))


class Context:
    """`zk_ctx`: one per (host thread, device)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.zk_ctx_create(device, ctypes.byref(h))
        if rc:
            raise ZkError(rc, "zk_ctx_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.lib.zk_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise ZkError(rc, self.lib.zk_last_error(self.h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.zk_ctx_set_stream(self.h, cuda_stream_ptr))

    def set_blocking_sync(self, on=True):
        self._check(self.lib.zk_ctx_set_blocking_sync(self.h, 1 if on else 0))

    def synchronize(self):
        self._check(self.lib.zk_ctx_synchronize(self.h))

    def launch_count(self):
        return self.lib.zk_ctx_launch_count(self.h)

    def enable_timing(self, on=True):
        self._check(self.lib.zk_ctx_enable_timing(self.h, 1 if on else 0))

    def timing_report(self):
        ms = (ctypes.c_float * 8)()
        cnt = (ctypes.c_uint32 * 8)()
        self._check(self.lib.zk_ctx_timing_report(self.h, ms, cnt))
        names = ["witness", "msm", "ntt", "quotient", "collapse", "msm_accumulate", "6", "7"]
        return {names[i]: (ms[i], cnt[i]) for i in range(8)}

    def last_kernel_ms(self, which=0):
        ms = ctypes.c_float()
        self._check(self.lib.zk_ctx_last_kernel_ms(self.h, which, ctypes.byref(ms)))
        return ms.value

    def bench_int_pipe(self, mode, iters=20000):
        out = ctypes.c_double()
        self._check(self.lib.zk_bench_int_pipe(self.h, mode, iters, ctypes.byref(out)))
        return out.value

    # ---- one MSM split across GPUs (zk_dist_*) -------------------------------------------
    def dist_init(self, unique_id, rank, world):
        """Joins the NCCL group described by `unique_id` (128 bytes from dist_unique_id() on
        rank 0, shared by the caller); must precede params_* on this context."""
        self._check(self.lib.zk_dist_init(self.h, bytes(unique_id), rank, world))

    def dist_info(self):
        r, w = ctypes.c_int32(), ctypes.c_int32()
        self._check(self.lib.zk_dist_info(self.h, ctypes.byref(r), ctypes.byref(w)))
        return r.value, w.value

    # ---- params / keygen / prove ---------------------------------------------------------
    def params_generate_substitute(self, k, seed):
        self._check(self.lib.zk_params_generate_substitute(self.h, k, bytes(seed)))

    def params_load(self, data):
        keep = bytes(data)
        self._check(self.lib.zk_params_load(self.h, _ptr(keep), len(keep)))

    def params_write(self):
        ln = ctypes.c_uint64(0)
        self.lib.zk_params_write(self.h, None, ctypes.byref(ln))
        buf = ctypes.create_string_buffer(ln.value)
        self._check(self.lib.zk_params_write(self.h, ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(ln)))
        return buf.raw[:ln.value]

    def keygen(self, rounds, n_compressions, chain=None):
        """chain: n_compressions flags; chain[j] != 0 makes compression j continue compression j - 1."""
        if chain is None:
            self._check(self.lib.zk_blake2f_keygen(self.h, rounds, n_compressions))
        else:
            assert len(chain) == n_compressions
            self._check(self.lib.zk_blake2f_keygen_chained(self.h, rounds, n_compressions, bytes(chain)))

    def vk_bytes(self):
        ln = ctypes.c_uint64(0)
        self.lib.zk_vk_bytes(self.h, None, ctypes.byref(ln))
        buf = ctypes.create_string_buffer(ln.value)
        self._check(self.lib.zk_vk_bytes(self.h, ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(ln)))
        return buf.raw[:ln.value]

    def vk_pinned_debug(self):
        """The Rust `{:?}` rendering of vk.pinned() that transcript_repr hashes."""
        ln = ctypes.c_uint64(0)
        self.lib.zk_vk_pinned_debug(self.h, None, ctypes.byref(ln))
        buf = ctypes.create_string_buffer(ln.value)
        self._check(self.lib.zk_vk_pinned_debug(self.h, ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(ln)))
        return buf.raw[:ln.value].decode()

    def create_proof(self, inputs, n_compressions, seed, on_device=False):
        keep = bytes(inputs) if isinstance(inputs, (bytes, bytearray)) else inputs
        buf = ctypes.create_string_buffer(1 << 16)
        ln = ctypes.c_uint64(len(buf))
        fn = self.lib.zk_create_proof_device_inputs if on_device else self.lib.zk_create_proof
        self._check(fn(self.h, _ptr(keep), n_compressions, bytes(seed),
                       ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(ln)))
        return buf.raw[:ln.value]

    def verify_proof(self, proof):
        """True if accepted; False (reason in .last_error()) if rejected."""
        rc = self.lib.zk_verify_proof(self.h, bytes(proof), len(proof))
        if rc == 0:
            return True
        if rc == -7:
            return False
        self._check(rc)

    def verify_proofs_batch(self, proofs, seed):
        """True if every proof of the list is accepted (one combined final MSM); False if any is rejected."""
        lens = (ctypes.c_uint64 * len(proofs))(*[len(p) for p in proofs])
        rc = self.lib.zk_verify_proofs_batch(self.h, b"".join(bytes(p) for p in proofs), lens, len(proofs), bytes(seed))
        if rc == 0:
            return True
        if rc == -7:
            return False
        self._check(rc)

    def last_error(self):
        return self.lib.zk_last_error(self.h).decode()

    def mock_verify(self, inputs, n_compressions, advice_override=None):
        """None if every constraint holds, else (kind, row, index) of the first failure."""
        keep = bytes(inputs) if isinstance(inputs, (bytes, bytearray)) else inputs
        fail = (ctypes.c_uint64 * 3)()
        rc = self.lib.zk_mock_verify(self.h, _ptr(keep), n_compressions, _ptr(advice_override), fail)
        if rc == 0:
            return None
        if rc == -7:
            return (fail[0], fail[1], fail[2])
        self._check(rc)

    # ---- K2/K3, K4/K5 ------------------------------------------------------------------
    def msm(self, scalars, bases, n, out_affine, on_device=False):
        self._check(self.lib.zk_msm_vesta(self.h, _ptr(scalars), _ptr(bases), n,
                                          1 if on_device else 0, _ptr(out_affine)))

    def ntt(self, data, log_n, inverse=False, on_device=False):
        self._check(self.lib.zk_ntt_fp(self.h, _ptr(data), log_n, 1 if inverse else 0,
                                       1 if on_device else 0))

    def commit_batch(self, basis, scalars, ncols, blinds, out_affine, index_mask=0, index_select=0,
                     on_device=False):
        """Params::commit (basis 0) / commit_lagrange (basis 1) of ncols columns in one batched pipeline."""
        self._check(self.lib.zk_commit_batch(self.h, basis, _ptr(scalars), ncols, _ptr(blinds), index_mask,
                                             index_select, 1 if on_device else 0, _ptr(out_affine)))

    def ntt_batch(self, data_in, data_out, log_n, batch, inverse=False, on_device=False):
        self._check(self.lib.zk_ntt_fp_batch(self.h, _ptr(data_in), _ptr(data_out), log_n, batch,
                                             1 if inverse else 0, 1 if on_device else 0))

    def coeff_to_cosets(self, coeffs, ncols, out, on_device=False):
        self._check(self.lib.zk_coeff_to_cosets(self.h, _ptr(coeffs), ncols, 1 if on_device else 0, _ptr(out)))

    # ---- K1 ------------------------------------------------------------------------------
    def witness_batch(self, k, rounds, inputs, n_compressions, advice_out, digests_out=None):
        """Host buffers in, host buffers out (numpy arrays / pinned torch tensors)."""
        keep = bytes(inputs) if isinstance(inputs, (bytes, bytearray)) else inputs
        self._check(self.lib.zk_blake2f_witness_batch(
            self.h, k, rounds, _ptr(keep), n_compressions, _ptr(advice_out), _ptr(digests_out)))

    def witness_batch_device(self, k, rounds, d_inputs, n_compressions, d_advice, d_digests=None):
        """Device buffers (torch CUDA tensors or raw device addresses); asynchronous."""
        self._check(self.lib.zk_blake2f_witness_batch_device(
            self.h, k, rounds, _ptr(d_inputs), n_compressions, _ptr(d_advice), _ptr(d_digests)))
