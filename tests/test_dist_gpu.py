"""BASELINE configs[3] shape on 2 GPUs: every MSM split by point range across the ranks, partial
points exchanged with one NCCL all-gather.  The keys and the proof bytes must equal the single-GPU
ones (and therefore the oracle's, tests/test_prover_gpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_msm_split_two_gpus_same_proof_bytes(zk, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    k, ncomp, world = 17, 3, 2
    uid = zk.dist_unique_id()
    out = str(tmp_path / "proof")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), str(r), str(world),
                               uid.hex(), str(k), str(ncomp), out]) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    single = zk.Context(0)
    single.params_generate_substitute(k, zk.REFERENCE_SEED)
    single.keygen(12, ncomp)
    want = single.vk_bytes() + single.create_proof(zk.synthetic_inputs(ncomp), ncomp, zk.REFERENCE_SEED)
    single.close()
    for r in range(world):
        assert open(out + ".%d" % r, "rb").read() == want, "rank %d" % r
