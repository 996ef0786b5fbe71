"""tools/compare_rust_dump.py is what will pin the oracle's placement against the real crate the day someone has cargo + network
(rust/tools/dump_witness).  Until then this keeps the comparer itself known-good: a synthetic dump written from the oracle in the
dump tool's exact file format compares clean, and every kind of placement / value error is reported where it was injected."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import compare_rust_dump as CRD  # noqa: E402

MSGS, PRE = [b"abc", b""], [0, 0]


@pytest.fixture(scope="module")
def synthetic(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("dump") / "synthetic.bin")
    rc = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "compare_rust_dump.py"), "--make-synthetic", path, MSGS[0].hex(), "0", "", "0"],
                        capture_output=True, text=True)
    assert rc.returncode == 0, rc.stdout + rc.stderr
    return path


def test_synthetic_dump_has_the_dump_tools_format_and_compares_clean(synthetic):
    raw = open(synthetic, "rb").read()
    assert raw[:8] == b"H2SHADMP" and len(raw) == 16 + 8 * (1 << 17) * 32       # 3 gate + 1 lookup + 2 dense + 2 spread columns, k = 17
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "compare_rust_dump.py"), synthetic, MSGS[0].hex(), "0", "", "0"], capture_output=True, text=True)
    assert out.returncode == 0 and "0 cells differ" in out.stdout, out.stdout + out.stderr
    # canonical, not Montgomery: the one-time zero cell aside, the first gate cell is the input length 3
    cells = CRD.load_dump(synthetic)
    assert cells[0, 0].tolist() == [3, 0, 0, 0]


def test_value_and_placement_errors_are_reported_where_they_are(synthetic, tmp_path):
    cells = CRD.load_dump(synthetic).copy()
    # (1) one wrong value in gate column 1
    c1 = cells.copy(); c1[1, 777, 0] ^= 1
    p = str(tmp_path / "v.bin"); CRD.write_dump(p, c1)
    r = CRD.compare(p, MSGS, PRE)
    assert r["differing"] == 1 and r["per_column"] == {"gate_1": (1, 777)}
    # (2) a placement error: everything in gate column 0 from row 5000 on shifted down by one cell (an op one cell longer)
    c2 = cells.copy(); c2[0, 5001:] = cells[0, 5000:-1]; c2[0, 5000] = 0
    p = str(tmp_path / "s.bin"); CRD.write_dump(p, c2)
    r = CRD.compare(p, MSGS, PRE)
    assert r["per_column"]["gate_0"][1] == 5000 and r["per_column"]["gate_0"][0] > 10000 and r["first_gate_stream_index"] == 5000
    # (3) lookup column in another push order; (4) a spread row swapped between the two column pairs
    c3 = cells.copy(); c3[3, [10, 11]] = cells[3, [11, 10]]
    r20 = int(np.nonzero((cells[4] != cells[5]).any(axis=-1))[0][3])     # a row whose two dense limbs differ
    c3[4, r20], c3[5, r20] = cells[5, r20].copy(), cells[4, r20].copy()
    p = str(tmp_path / "l.bin"); CRD.write_dump(p, c3)
    r = CRD.compare(p, MSGS, PRE)
    assert set(r["per_column"]) == {"lookup", "dense_0", "dense_1"} and r["per_column"]["lookup"] == (2, 10) and r["per_column"]["dense_0"] == (1, r20)
    # (5) a dump with the wrong number of columns / a truncated file is refused, not mis-compared
    p = str(tmp_path / "c.bin"); CRD.write_dump(p, cells[:7])
    with pytest.raises(ValueError):
        CRD.compare(p, MSGS, PRE)
    open(str(tmp_path / "t.bin"), "wb").write(open(synthetic, "rb").read()[:-32])
    with pytest.raises(ValueError):
        CRD.load_dump(str(tmp_path / "t.bin"))
