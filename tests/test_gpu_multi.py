"""GPU tier, needs >= 2 GPUs (skipped otherwise): several GPUs on one index part.
  * the drop-in executable driving 2 GPUs from one process (NCCL communicator inside liblqcov.so, lq_comm.cu) must print the table
    the unmodified reference printed -- cases with several index parts, tie-order sensitive sorts and > 131 072 target reads;
  * one process per GPU under torchrun (tools/dist_check.py: the path bench.py --gpus N takes)."""
import os
import subprocess
import sys

import pytest

import c1_cases
import cases

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")


def _cli(T, Q, flags, tmp_path, gpus):
    import longqc_b200 as L
    tf, qf = str(tmp_path / "t.fq"), str(tmp_path / "q.fq")
    T.write_fastx(tf)
    Q.write_fastx(qf)
    env = dict(os.environ, LQCOV_GPUS=str(gpus))
    p = subprocess.run([L.bin_path("minimap2-coverage")] + flags.split() + ["-t", "8", tf, qf], capture_output=True, env=env)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return p.stdout, p.stderr


@needs2
@pytest.mark.parametrize("name", ["plain_pb", "tandem", "tandem_parts", "parts", "covt_gate", "c5_small", "ava_X"])
def test_executable_on_two_gpus_equals_reference_golden(name, tmp_path):
    T, Q = cases.make_case(name)
    out, err = _cli(T, Q, cases.CASES[name][2], tmp_path, 2)
    assert b"2 GPUs" in err
    assert out == open(os.path.join(GOLD, name + ".tsv"), "rb").read()


@needs2
def test_executable_on_two_gpus_many_targets(tmp_path):
    """> 131 072 target reads in one part: the few-region walk of the s48 sort level on every GPU"""
    T, Q = c1_cases.make("many_targets")
    out, _ = _cli(T, Q, c1_cases.FLAGS.replace("-t 4", ""), tmp_path, 2)
    assert out == open(os.path.join(GOLD, "many_targets.tsv"), "rb").read()


@pytest.mark.skipif(_n_gpus() < 4, reason="needs at least 4 GPUs")
def test_executable_on_four_gpus(tmp_path):
    T, Q = cases.make_case("tandem_parts")
    out, _ = _cli(T, Q, cases.CASES["tandem_parts"][2], tmp_path, 4)
    assert out == open(os.path.join(GOLD, "tandem_parts.tsv"), "rb").read()


@needs2
def test_one_process_per_gpu_under_torchrun():
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py"), "parts", "tandem_parts", "plain_pb", "c5_small"],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    assert p.stdout.count("IDENTICAL to the reference golden") == 4
