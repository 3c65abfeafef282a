"""CPU tier: the oracle (oracle/lq_oracle.c) against the committed golden tables, which were produced by the
unmodified reference binary (oracle/make_golden.py).  This is what pins the oracle on a box without
/root/reference."""
import json
import os

import pytest

import cases
import liblq

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAN = json.load(open(os.path.join(GOLD, "manifest.json")))


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_matches_reference_table(name):
    T, Q = cases.make_case(name)
    assert cases.inputs_md5(T, Q) == MAN[name]["inputs_md5"], "generator drift: regenerate the goldens in the build container"
    got, mid, parts = liblq.oracle_table(T, Q, liblq.oracle_opt(**cases.opts(name)))
    want = open(os.path.join(GOLD, name + ".tsv"), "rb").read()
    assert got == want
    assert MAN[name]["mid_occ_line"] in (None, "mid_occ = %d" % mid)


@pytest.mark.parametrize("name", ["plain_pb", "tandem", "ambiguous_fasta", "junk_adapters"])
def test_oracle_matches_reference_sdust(name):
    _, Q = cases.make_case(name)
    assert liblq.oracle_sdust_table(Q) == open(os.path.join(GOLD, name + ".sdust.tsv"), "rb").read()
