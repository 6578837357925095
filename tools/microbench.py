"""Integer-pipe and field/curve micro-benchmarks on cuda:0 (zk_bench_int_pipe modes)."""
import json
import sys

sys.path.insert(0, ".")
import zk_odst_b200 as zk

ctx = zk.Context(0)
names = {0: "mad.lo.u32 instr/s", 1: "mad.wide.u32 instr/s", 2: "mad.lo.cc+madc.hi.cc instr/s",
         3: "Fq mont mul/s", 4: "XYZZ mixed add/s (inline)", 5: "XYZZ mixed add/s (mul as call)",
         6: "XYZZ mixed add/s (inline, 5 blocks/SM)", 7: "XYZZ mixed add/s (inline, 6 blocks/SM)",
         8: "XYZZ mixed add/s (inline, 8 blocks/SM)"}
out = {}
for mode in range(9):
    iters = 2000 if mode < 3 else (400 if mode == 3 else 40)
    out[names[mode]] = ctx.bench_int_pipe(mode, iters)
print(json.dumps(out, indent=1))
