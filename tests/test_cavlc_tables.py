"""CAVLC code tables of the host front end (m2dec_b200/csrc/h264_cavlc_tables.h; shared with the
test-only writer, so a wrong entry is invisible to writer-vs-parser tests).  Checked here against
the reference's own table dumps (extras/h264vld.txt, h264zeros*.txt, h264run*.txt) where the
reference tree exists, and against the entries the reference decodes without a table
(h264.cpp:1783-1926: total_zeros of 15 coefficients, chroma DC total_zeros, run_before 1 and 3)."""
import os
import re
from fractions import Fraction

import pytest

import util

HEADER = os.path.join(util.ROOT, "m2dec_b200", "csrc", "h264_cavlc_tables.h")
EXTRAS = "/root/reference/extras"


def _tables():
    src = open(HEADER).read()
    tabs = {}
    for m in re.finditer(r"static const VlcCode (\w+)\[\] = \{(.*?)\};", src, re.S):
        tabs[m.group(1)] = {int(v): c.replace(" ", "") for c, v in re.findall(r'\{"([01 ]+)", (\d+)\}', m.group(2))}
    return tabs


def test_tables_are_prefix_free():
    for name, t in _tables().items():
        codes = sorted(t.values())
        assert len(set(codes)) == len(codes), name
        for a in codes:
            for b in codes:
                assert a == b or not b.startswith(a), (name, a, b)
        assert sum(Fraction(1, 2 ** len(c)) for c in codes) <= 1, name


def test_entries_the_reference_decodes_without_a_table():
    t = _tables()
    # h264.cpp:1856-1858: total_zeros of a 15-coefficient block is the bit itself
    assert t["total_zeros_15"] == {0: "0", 1: "1"}
    # h264.cpp:1866-1883 (chroma DC)
    assert t["total_zeros_cdc_1"] == {0: "1", 1: "01", 2: "001", 3: "000"}
    assert t["total_zeros_cdc_2"] == {0: "1", 1: "01", 2: "00"}
    assert t["total_zeros_cdc_3"] == {0: "1", 1: "0"}
    # h264.cpp:1893-1903: run_before with 1 and 3 zeros left
    assert t["run_before_1"] == {0: "1", 1: "0"}
    assert t["run_before_3"] == {0: "11", 1: "10", 2: "01", 3: "00"}


@pytest.mark.skipif(not os.path.isdir(EXTRAS), reason="reference tree not present")
def test_tables_match_the_reference_dumps():
    t = _tables()
    names = ["coeff_token_nc0", "coeff_token_nc2", "coeff_token_nc4", "coeff_token_nc8", "coeff_token_chroma_dc"]
    seen = 0
    for line in open(os.path.join(EXTRAS, "h264vld.txt")):
        f = [x.strip().replace(" ", "") for x in line.strip().split(",")]
        if len(f) < 7:
            continue
        t1, tc = int(f[0]), int(f[1])
        for k, n in enumerate(names):
            if n == "coeff_token_chroma_dc" and tc > 4:
                continue
            assert t[n].get(tc * 4 + t1) == f[2 + k], (n, t1, tc)
            seen += 1
    assert seen == 4 * 62 + 14
    for i in range(1, 15):
        rows = [l.strip().split(",") for l in open(os.path.join(EXTRAS, "h264zeros%d.txt" % i)) if "," in l]
        assert {int(v): c.replace(" ", "") for v, c in rows} == t["total_zeros_%d" % i], i
    for i in (2, 4, 5, 6, 7):
        rows = [l.strip().split(",") for l in open(os.path.join(EXTRAS, "h264run%d.txt" % i)) if "," in l]
        assert {int(v): c.replace(" ", "") for v, c in rows} == t["run_before_%d" % i], i
