"""Row plan of one BLAKE2f compression region, as JSON (and optionally an SVG strip chart).

The reference's layout renderer is commented out (`print_blake2f_circuit`, feature `test-dev-graph`,
`CircuitLayout::render(17, ...)`, table16.rs:462-528; SURVEY.md section 5 lists "dump per-region row ranges as JSON"
as the replacement).  This tool reads the frozen layout through the C ABI (`zk_blake2f_layout_tables`, host-only: no
GPU needed) and reports, per selector, the row ranges it is enabled on, the rows that pin constants, the rows the
chaining copies connect, and the copy constraints per column pair.

usage: python tools/layout_dump.py [--rounds 12] [--svg layout.svg]
"""
import argparse
import collections
import json
import sys

sys.path.insert(0, ".")
import zk_odst_b200 as zk

SELECTORS = ["s_spread_a1", "s_spread_b1", "s_spread_c1", "s_spread_d1", "s_spread_a2", "s_spread_b2", "s_spread_c2",
             "s_spread_d2", "s_decompose_abcd", "s_decompose_efgh", "s_decompose_ijkl", "s_digest", "s_const",
             "s_fmask"]   # compression.rs:561-577, then the two pinning selectors of docs/CIRCUIT.md
COLOURS = ["#1f77b4", "#ff7f0e", "#2ca02c", "#d62728", "#9467bd", "#8c564b", "#e377c2", "#7f7f7f", "#bcbd22",
           "#17becf", "#aec7e8", "#000000", "#ffbb78", "#98df8a"]


def ranges(rows):
    """sorted row numbers -> [[first, last], ...] of maximal runs"""
    out = []
    for r in rows:
        if out and r == out[-1][1] + 1:
            out[-1][1] = r
        else:
            out.append([r, r])
    return out


def describe(rounds):
    copies, sel, const, chain = zk.layout_tables(rounds)
    R = sel.shape[1]
    per_sel = {}
    for s, name in enumerate(SELECTORS):
        rows = [int(r) for r in sel[s].nonzero()[0]]
        per_sel[name] = {"rows_enabled": len(rows), "ranges": ranges(rows)}
    pairs = collections.Counter((int(c[0]), int(c[2])) for c in copies)
    return {
        "rounds": rounds,
        "rows": R,
        "rows_formula": "292 + 392 * rounds",
        "min_k_for_one_compression": zk.min_k(rounds, 1),
        "selectors": per_sel,
        "rows_without_selector": int((sel.sum(axis=0) == 0).sum()),
        "constant_rows": {str(int(r)): hex(int(const[r])) for r in const.nonzero()[0]},
        "copy_constraints": len(copies),
        "copies_by_column_pair": {"%d->%d" % k: v for k, v in sorted(pairs.items())},
        "chain_rows": {"h_in (advice column 1)": [int(x) for x in chain[:8]],
                       "h_out (advice column 0)": [int(x) for x in chain[8:]]},
    }, sel


def svg(sel, path, px_per_row=None):
    R = sel.shape[1]
    width = 1200
    scale = width / R
    height = 16 * len(SELECTORS) + 30
    parts = ['<svg xmlns="http://www.w3.org/2000/svg" width="%d" height="%d" font-family="monospace" font-size="10">'
             % (width + 160, height)]
    for s, name in enumerate(SELECTORS):
        y = 10 + 16 * s
        parts.append('<text x="0" y="%d">%s</text>' % (y + 10, name))
        for a, b in ranges([int(r) for r in sel[s].nonzero()[0]]):
            parts.append('<rect x="%.2f" y="%d" width="%.2f" height="12" fill="%s"/>'
                         % (150 + a * scale, y, max((b - a + 1) * scale, 0.6), COLOURS[s]))
    parts.append('<text x="150" y="%d">row 0 .. %d (one compression region)</text>' % (height - 4, R - 1))
    parts.append("</svg>")
    open(path, "w").write("\n".join(parts))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=12)
    ap.add_argument("--svg")
    a = ap.parse_args()
    d, sel = describe(a.rounds)
    if a.svg:
        svg(sel, a.svg)
    print(json.dumps(d))
