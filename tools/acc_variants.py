#!/usr/bin/env python
"""Accumulation-kernel experiment: times the bucket accumulation of one batched
commitment of 12 full-width columns at n = 2^19 (the 8.39 M-entry-per-job launch of a proof) for every
ZK_ACC_VARIANT, one process each, and checks that all variants return the same commitments.

  python tools/acc_variants.py > gpurun_out/acc_variants.json
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {0: "register prefetch, 4 blocks/SM (product)", 1: "cp.async staging, 5 blocks/SM",
            2: "cp.async.bulk + mbarrier staging, 5 blocks/SM", 11: "cp.async staging, 4 blocks/SM",
            12: "cp.async.bulk + mbarrier staging, 4 blocks/SM"}
K, NCOLS, REPS = 19, 12, 5


def worker():
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import zk_odst_b200 as zk
    n = 1 << K
    rnd = np.random.RandomState(1)
    sc = rnd.randint(0, 1 << 62, size=(NCOLS, n, 4), dtype=np.uint64)   # any 256-bit word < p is a field element
    d = torch.from_numpy(sc.view(np.int64)).cuda()
    blinds = np.zeros((NCOLS, 4), dtype=np.uint64)
    out = np.zeros((NCOLS, 8), dtype=np.uint64)
    ctx = zk.Context(0)
    ctx.params_generate_substitute(K, zk.REFERENCE_SEED)
    ctx.commit_batch(1, d, NCOLS, blinds, out, on_device=True)   # warm-up
    ctx.enable_timing(True)
    ctx.timing_report()
    for _ in range(REPS):
        ctx.commit_batch(1, d, NCOLS, blinds, out, on_device=True)
    rep = ctx.timing_report()
    print(json.dumps({"accumulate_ms": rep["msm_accumulate"][0] / REPS, "msm_ms": rep["msm"][0] / REPS,
                      "digest": out.tobytes().hex()[:64], "all": __import__("hashlib").sha256(out.tobytes()).hexdigest()}))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        return worker()
    res = {}
    for v, name in VARIANTS.items():
        env = dict(os.environ, ZK_ACC_VARIANT=str(v))
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "worker"], env=env, capture_output=True,
                           text=True, timeout=600)
        try:
            res[name] = json.loads(p.stdout.strip().split("\n")[-1])
        except Exception:
            res[name] = {"error": (p.stdout + p.stderr)[-800:]}
    base = res[VARIANTS[0]].get("all")
    for name in res:
        res[name]["same_commitments_as_product"] = res[name].get("all") == base
    print(json.dumps({"k": K, "columns": NCOLS, "entries_per_launch": NCOLS * 16 * (1 << K), "variants": res}, indent=1))


if __name__ == "__main__":
    main()
