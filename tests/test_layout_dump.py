"""tools/layout_dump.py: the row plan of a compression region as JSON (the replacement SURVEY.md section 5 names for
the reference's commented-out `CircuitLayout::render`, table16.rs:462-528).  Host-only: runs without a GPU."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tool():
    spec = importlib.util.spec_from_file_location("layout_dump", os.path.join(ROOT, "tools", "layout_dump.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("rounds", [0, 1, 12])
def test_row_plan(tool, rounds, tmp_path):
    d, sel = tool.describe(rounds)
    assert d["rows"] == 292 + 392 * rounds == sel.shape[1]
    s = d["selectors"]
    # per G step: one row of each a / b / c selector and of the efgh / ijkl decompositions; 8 G steps per round
    for name in ("s_spread_a1", "s_spread_b1", "s_spread_c1", "s_spread_a2", "s_spread_b2", "s_spread_c2",
                 "s_decompose_efgh", "s_decompose_ijkl"):
        assert s[name]["rows_enabled"] == 8 * rounds, name
    assert s["s_digest"]["rows_enabled"] == 8 and s["s_const"]["rows_enabled"] == 8
    assert s["s_fmask"]["rows_enabled"] == 1
    # the pinned constants are the eight IV words, on the s_const rows
    iv = ["0x6a09e667f3bcc908", "0xbb67ae8584caa73b", "0x3c6ef372fe94f82b", "0xa54ff53a5f1d36f1",
          "0x510e527fade682d1", "0x9b05688c2b3e6c1f", "0x1f83d9abfb41bd6b", "0x5be0cd19137e2179"]
    assert list(d["constant_rows"].values()) == iv
    assert [[int(r), int(r)] for r in d["constant_rows"]] == s["s_const"]["ranges"]
    # every row of a range list is inside the region; ranges are disjoint and ordered
    for name, v in s.items():
        last = -1
        for a, b in v["ranges"]:
            assert last < a <= b < d["rows"], name
            last = b
        assert sum(b - a + 1 for a, b in v["ranges"]) == v["rows_enabled"]
    assert len(d["chain_rows"]["h_in (advice column 1)"]) == 8
    assert sum(d["copies_by_column_pair"].values()) == d["copy_constraints"]
    out = tmp_path / "layout.svg"
    tool.svg(sel, str(out))
    assert out.read_text().startswith("<svg") and out.read_text().rstrip().endswith("</svg>")
