"""GPU tier: the CUDA path against the committed golden tables of the unmodified reference binary
(tests/golden, oracle/make_golden.py) -- through the C ABI and through the drop-in executables the way
LongQC spawns them (lq_exec.py / lq_mask.py)."""
import json
import os
import subprocess

import pytest

import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAN = json.load(open(os.path.join(GOLD, "manifest.json")))


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_abi_table_equals_reference_golden(name):
    import longqc_b200 as L
    T, Q = cases.make_case(name)
    assert cases.inputs_md5(T, Q) == MAN[name]["inputs_md5"]
    got = L.coverage_table(T, Q, L.Opt(**cases.opts(name)))
    assert got == open(os.path.join(GOLD, name + ".tsv"), "rb").read()


@pytest.mark.parametrize("name", ["plain_pb", "tandem", "ambiguous_fasta", "junk_adapters", "spike_hpc_filter"])
def test_abi_sdust_equals_reference_golden(name):
    import longqc_b200 as L
    _, Q = cases.make_case(name)
    assert L.sdust_table(Q) == open(os.path.join(GOLD, name + ".sdust.tsv"), "rb").read()


@pytest.mark.parametrize("name,gz", [("plain_pb", False), ("tandem_parts", True), ("spike_hpc_filter", False), ("ambiguous_fasta", True), ("ava_X", False), ("hifi_fast_k19", False),
                                     ("wide_k17_parts", False)])
def test_dropin_executables_like_longqc(name, gz, tmp_path):
    """argv + stdout bytes + exit status: exactly what longQC.py:438-445 / lq_mask.py:17-23 rely on"""
    import longqc_b200 as L
    T, Q = cases.make_case(name)
    tf = str(tmp_path / ("t" + (".fa" if T.qual is None else ".fq") + (".gz" if gz else "")))
    qf = str(tmp_path / ("q" + (".fa" if Q.qual is None else ".fq")))
    T.write_fastx(tf)
    Q.write_fastx(qf, line_width=70 if Q.qual is None else 0)
    out, err = str(tmp_path / "coverage_out.txt"), str(tmp_path / "coverage_err.txt")
    le = L.LqExec(L.bin_path("minimap2-coverage"))
    le.exec(*(cases.CASES[name][2].split() + ["-t", "4", tf, qf]), out=out, err=err)
    assert le.wait() == 0
    assert open(out, "rb").read() == open(os.path.join(GOLD, name + ".tsv"), "rb").read()
    assert b"Real time" in open(err, "rb").read()
    sd = subprocess.run([L.bin_path("sdust"), qf], capture_output=True, check=True)   # lq_mask.py:19
    assert sd.stdout == open(os.path.join(GOLD, name + ".sdust.tsv"), "rb").read()


def test_reset_and_determinism():
    import longqc_b200 as L
    T, Q = cases.make_case("tandem")
    opt = L.Opt(**cases.opts("tandem"))
    want = open(os.path.join(GOLD, "tandem.tsv"), "rb").read()
    with L.Coverage(opt) as c:
        for _ in range(3):
            L.load().lqcov_reset(c._h)
            c.set_queries(Q)
            c.add_targets(T)
            assert c.table() == want


def test_small_seed_budget_batches():
    """many query batches (seed budget far below one job) give the same table"""
    import longqc_b200 as L
    T, Q = cases.make_case("plain_pb")
    o = cases.opts("plain_pb")
    o["seed_budget"] = 20000
    with L.Coverage(L.Opt(**o)) as c:
        c.set_queries(Q)
        c.add_targets(T)
        assert c.table() == open(os.path.join(GOLD, "plain_pb.tsv"), "rb").read()
        assert c.stats()["batches"] > 5
