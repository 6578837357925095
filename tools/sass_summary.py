#!/usr/bin/env python
"""Per-kernel SASS mnemonic histogram of libzkodst.so (cuobjdump -sass), for the kernels DESIGN.md quotes:
how many IMAD.WIDE.U32(.X) a kernel carries, whether its stores are 256-bit, whether the staged accumulation
uses LDGSTS (cp.async) or UBLKCP (cp.async.bulk) + SYNCS (mbarrier).  usage: sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = ["fixed_accumulate_kernel", "fixed_accumulate_staged_kernelILi1ELi5", "fixed_accumulate_staged_kernelILi2ELi5",
           "ntt_pass_kernel", "blake2f_witness_kernel", "quotient_gates_kernel", "fixed_digits_kernel",
           "fixed_scatter_kernel", "fold_accumulate_kernel"]
WATCH = re.compile(r"\b(IMAD\.WIDE\.U32\.X|IMAD\.WIDE\.U32|IMAD\.WIDE|IMAD|IADD3|LDGSTS[\w.]*|UBLKCP[\w.]*|SYNCS[\w.]*|"
                   r"STG\.E[\w.]*|LDG\.E[\w.]*|STS[\w.]*|LDS[\w.]*|ATOMG[\w.]*|RED[\w.]*|MATCH[\w.]*|SHFL[\w.]*|BAR[\w.]*|"
                   r"STL[\w.]*|LDL[\w.]*|CALL[\w.]*)\b")
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "zk-odst_b200", "libzkodst.so")], capture_output=True,
                      text=True).stdout
cur, hist, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next((k for k in KERNELS if k in m.group(1)), None)
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        total[cur] += 1
        w = WATCH.search(line)
        if w:
            hist[cur][w.group(1)] += 1
print("cuobjdump -sass zk-odst_b200/libzkodst.so (sm_100a): instructions per kernel and selected mnemonics")
for k in KERNELS:
    if not total[k]:
        continue
    print("\n== %s: %d instructions" % (k, total[k]))
    for name, cnt in hist[k].most_common(18):
        print("   %6d  %s" % (cnt, name))
