#!/usr/bin/env python
"""Generate tests/golden/*.tsv by running the UNMODIFIED reference binaries (oracle/_ref, built from
/root/reference by oracle/Makefile) on the seeded cases of tests/cases.py.  Run in the build container
(the reference sources are not available on the GPU box); the outputs are committed.

    python oracle/make_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "minimap2-coverage")
SDUST = os.path.join(ROOT, "oracle", "_ref", "sdust")
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    assert os.path.exists(REF), "build oracle/_ref first: make -C oracle ref"
    os.makedirs(GOLD, exist_ok=True)
    manifest = {}
    for name in sorted(cases.CASES):
        T, Q = cases.make_case(name)
        flags = cases.CASES[name][2]
        with tempfile.TemporaryDirectory() as d:
            ext_t = ".fa" if T.qual is None else ".fq"
            ext_q = ".fa" if Q.qual is None else ".fq"
            tf, qf = os.path.join(d, "t" + ext_t), os.path.join(d, "q" + ext_q)
            T.write_fastx(tf)
            Q.write_fastx(qf, line_width=70 if Q.qual is None else 0)
            out = subprocess.run([REF] + flags.split() + ["-t", "4", tf, qf], capture_output=True, check=True)
            open(os.path.join(GOLD, name + ".tsv"), "wb").write(out.stdout)
            sd = subprocess.run([SDUST, qf], capture_output=True, check=True)
            open(os.path.join(GOLD, name + ".sdust.tsv"), "wb").write(sd.stdout)
            mid = [ln for ln in out.stderr.decode().split("\n") if "mid_occ" in ln]
        manifest[name] = {"flags": flags, "inputs_md5": cases.inputs_md5(T, Q), "targets": T.n, "queries": Q.n,
                          "rows": out.stdout.count(b"\n"), "mid_occ_line": mid[0].split("] ")[-1] if mid else None}
        print(name, manifest[name])
    json.dump(manifest, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
