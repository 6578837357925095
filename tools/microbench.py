#!/usr/bin/env python
"""Runs every mode of zk_bench_int_pipe on cuda:0 and prints one JSON object (profiles/rNN_microbench.json).

  python tools/microbench.py > gpurun_out/microbench.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_odst_b200 as zk  # noqa: E402

NAMES = {0: "mad.lo.u32 instr/s", 1: "mad.wide.u32 instr/s", 2: "mad.lo.cc + madc.hi.cc chain instr/s",
         3: "Fq Montgomery products/s", 4: "XYZZ mixed additions/s (4 blocks/SM)",
         5: "XYZZ mixed additions/s, product as a call", 6: "XYZZ mixed additions/s (5 blocks/SM)",
         7: "XYZZ mixed additions/s (6 blocks/SM)", 8: "XYZZ mixed additions/s (8 blocks/SM)",
         9: "fma.rn.f64 instr/s", 10: "fma.rn.f64 + mad.wide.u32 interleaved, instr/s (both counted)"}


def main():
    ctx = zk.Context(0)
    out = {}
    for mode, name in NAMES.items():
        iters = 20000 if mode < 3 or mode > 8 else (2000 if mode == 3 else 300)
        out[name] = ctx.bench_int_pipe(mode, iters)
    ctx.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
