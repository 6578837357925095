#!/usr/bin/env python
"""Times one proof stream of configs[2] (k = 19, 64 compressions) per kernel class under each value of an
experiment switch (one process per value; the proof bytes of all values must agree).

  python tools/variants_bench.py ZK_REDUCE_VARIANT 0 1 2 3 > gpurun_out/reduce_variants.json
"""
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker():
    import torch
    sys.path.insert(0, ROOT)
    import zk_odst_b200 as zk
    n, rounds, reps = 64, 12, 6
    k = zk.min_k(rounds, n)
    ctx = zk.Context(0)
    ctx.params_generate_substitute(k, zk.REFERENCE_SEED)
    ctx.keygen(rounds, n)
    d_in = torch.frombuffer(bytearray(zk.synthetic_inputs(n)), dtype=torch.uint8).cuda()
    for _ in range(3):
        proof = ctx.create_proof(d_in, n, zk.REFERENCE_SEED, on_device=True)
    ctx.enable_timing(True)
    ctx.timing_report()
    t0 = time.perf_counter()
    for _ in range(reps):
        proof = ctx.create_proof(d_in, n, zk.REFERENCE_SEED, on_device=True)
    ms = (time.perf_counter() - t0) / reps * 1e3
    rep = ctx.timing_report()
    out = {"ms_per_proof_single_stream": ms, "proof_sha256": hashlib.sha256(proof).hexdigest()}
    out.update({name + "_ms": v[0] / reps for name, v in rep.items() if v[1]})
    print(json.dumps(out))


def main():
    if sys.argv[1] == "worker":
        return worker()
    var, values = sys.argv[1], sys.argv[2:]
    res = {}
    for v in values:
        env = dict(os.environ)
        env[var] = v
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "worker"], env=env, capture_output=True, text=True,
                           timeout=900)
        try:
            res[v] = json.loads(p.stdout.strip().split("\n")[-1])
        except Exception:
            res[v] = {"error": (p.stdout + p.stderr)[-600:]}
    shas = {r.get("proof_sha256") for r in res.values()}
    print(json.dumps({"switch": var, "same_proof_bytes": len(shas) == 1, "values": res}, indent=1))


if __name__ == "__main__":
    main()
