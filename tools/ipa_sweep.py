"""Stage-2 shapes of the inner-product argument (rounds on the folded generators: small MSMs bound by the
length of their chains of dependent point additions).  One process per configuration (the knobs are read once):
window bits of the table over the folded generators, minimum accumulation chunk, segment sums per tree thread.
Prints, per configuration, the `ZK_PHASE_TRACE` steps of the argument and the proof's wall time (median of 5)."""
import json
import os
import re
import subprocess
import sys

CHILD = r"""
import sys, time
sys.path.insert(0, ".")
import zk_odst_b200 as zk
k, n = int(sys.argv[1]), int(sys.argv[2])
ctx = zk.Context(0)
ctx.params_generate_substitute(k, zk.REFERENCE_SEED)
ctx.keygen(12, n)
inputs = zk.synthetic_inputs(n)
import hashlib
for _ in range(8):
    t = time.perf_counter()
    proof = ctx.create_proof(inputs, n, zk.REFERENCE_SEED)
    print("PROOF_MS %.3f %s" % ((time.perf_counter() - t) * 1e3, hashlib.sha256(proof).hexdigest()[:16]), file=sys.stderr)
"""


def run(env, k, n, trace):
    e = dict(os.environ)
    e.update(env)
    if trace:
        e["ZK_PHASE_TRACE"] = "1"
    out = subprocess.run([sys.executable, "-c", CHILD, str(k), str(n)], env=e, capture_output=True, text=True)
    if out.returncode:
        return {"error": out.stderr[-400:]}
    ms = sorted(float(m.group(1)) for m in re.finditer(r"PROOF_MS ([0-9.]+)", out.stderr))[: 5]
    sha = set(m.group(1) for m in re.finditer(r"PROOF_MS [0-9.]+ (\w+)", out.stderr))
    steps = {}
    for m in re.finditer(r"\[phase\]\s+\. (ipa: .*?)\s+([0-9.]+) ms", out.stderr):
        steps[m.group(1)] = float(m.group(2))   # last proof wins
    return {"proof_ms_median5": ms[len(ms) // 2], "sha": sorted(sha), **steps}


if __name__ == "__main__":
    k, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (19, 64)
    configs = [{}]
    for c in ("12", "13"):
        for mc in ("16", "8"):
            for per in ("8", "2", "1"):
                configs.append({"ZK_IPA_STAGE2_C": c, "ZK_SMALL_MIN_CHUNK": mc, "ZK_SMALL_TREE_PER": per})
    for cfg in configs:
        traced = run(cfg, k, n, True)
        plain = run(cfg, k, n, False)
        print(json.dumps({"config": cfg, "k": k, "untraced_proof_ms": plain.get("proof_ms_median5"), **traced}), flush=True)
