"""BASELINE.json configs[0] (the reference's own CPU-runnable case) and the consumer of the table.

CPU tier: the oracle port reproduces the reference's C1 tables (committed goldens made by oracle/make_consumer_golden.py); where
/root/reference exists the UNMODIFIED lq_coverage.LqCoverage is run on them (tests/consumer_harness.py) and its fields must be
the committed ones -- "bit-identical lq_coverage.py output" follows from a byte-identical table, this pins the numbers it means.
GPU tier: the CUDA path reproduces the same tables."""
import json
import os

import pytest

import c1_cases
import consumer_harness
import liblq

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _golden(name):
    return open(os.path.join(GOLD, name + ".tsv"), "rb").read()


@pytest.mark.parametrize("name", c1_cases.NAMES)
def test_oracle_reproduces_c1(name):
    T, Q = c1_cases.make(name)
    _, oopt = liblq.opt_pair(**c1_cases.OPTS)
    got, mid_occ, parts = liblq.oracle_table(T, Q, oopt)
    assert parts == 1
    assert got == _golden(name)


@pytest.mark.skipif(not consumer_harness.available(), reason="needs /root/reference (lq_coverage.py)")
@pytest.mark.parametrize("name", c1_cases.NAMES)
def test_consumer_fields(name):
    want = json.load(open(os.path.join(GOLD, "consumer_c1.json")))[name]
    got = consumer_harness.consumer_fields(os.path.join(GOLD, name + ".tsv"))
    for k in ("unmapped_frac_trimmed", "unmapped_frac_untrimmed", "unmapped_frac_med", "high_div_frac", "low_coverage"):
        assert got[k] == want[k], k                       # deterministic functions of the table
    for k in ("mean", "sd"):                              # seeded GMM: same library, same seed
        assert got[k] == pytest.approx(want[k], rel=1e-6), k
    # the non-sense-read flag of north_star == column 4 is '0'
    rows = [ln.split(b"\t") for ln in _golden(name).splitlines()]
    assert got["unmapped_frac_med"] == pytest.approx(sum(1 for r in rows if r[4] == b"0") / len(rows))


def test_oracle_reproduces_many_targets():
    """> 131 072 target reads in one part (three values of rid>>16): the oracle's sort restatement against the reference table"""
    T, Q = c1_cases.make("many_targets")
    _, oopt = liblq.opt_pair(**c1_cases.OPTS)
    got, _, parts = liblq.oracle_table(T, Q, oopt)
    assert parts == 1 and got == _golden("many_targets")


@pytest.mark.gpu
def test_gpu_reproduces_many_targets():
    """the few-region walk of the s48 sort level (tied queries with > 2048 seeds per strand against 140 000 targets)"""
    import longqc_b200 as L
    T, Q = c1_cases.make("many_targets")
    opt, _ = liblq.opt_pair(**c1_cases.OPTS)
    assert L.coverage_table(T, Q, opt) == _golden("many_targets")


@pytest.mark.gpu
@pytest.mark.parametrize("name", c1_cases.NAMES)
def test_gpu_reproduces_c1(name):
    import longqc_b200 as L
    T, Q = c1_cases.make(name)
    opt, _ = liblq.opt_pair(**c1_cases.OPTS)
    assert L.coverage_table(T, Q, opt) == _golden(name)


# ---- the consumer's deterministic core restated (longqc_b200/consumer.py) against the unmodified class's committed fields ----
def _consumer_gold():
    g = json.load(open(os.path.join(GOLD, "consumer_c1.json")))
    g.update(json.load(open(os.path.join(GOLD, "consumer_c5.json"))))
    return g


def _check_consumer(table, want):
    from longqc_b200 import consumer
    got = consumer.zero_fractions(table)
    for k in ("unmapped_frac_trimmed", "unmapped_frac_untrimmed", "unmapped_frac_med", "high_div_frac"):
        assert got[k] == want[k], k
    h, e = consumer.coverage_histogram(table, want["mean"], want["cov_main"])
    assert [float(x) for x in e] == want["hist_edges"] and [float(x) for x in h] == want["hist_density"]


@pytest.mark.parametrize("name", list(c1_cases.NAMES) + ["c5_small"])
def test_consumer_restatement_matches_lq_coverage(name):
    """zero-coverage fractions (non-sense reads) and histogram bins of lq_coverage.py:211-241, from the reference's table"""
    _check_consumer(_golden(name), _consumer_gold()[name])


@pytest.mark.gpu
def test_gpu_c5_small_consumer_fields():
    """BASELINE configs[4] reduced (mixed GC, junk + adapter reads, 5 index parts): every row identical AND the consumer's
    unmapped_frac_med / histogram bins computed from OUR table equal what the unmodified lq_coverage.py got from the reference's"""
    import cases
    import longqc_b200 as L
    T, Q = cases.make_case("c5_small")
    got = L.coverage_table(T, Q, L.Opt(**cases.opts("c5_small")))
    assert got == _golden("c5_small")
    _check_consumer(got, _consumer_gold()["c5_small"])
