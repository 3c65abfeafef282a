#!/usr/bin/env python
"""Generate tests/golden/*.tsv by running the UNMODIFIED reference binaries (oracle/_ref, built from
/root/reference by oracle/Makefile) on the seeded cases of tests/cases.py.  Run in the build container
(the reference sources are not available on the GPU box); the outputs are committed.

    python oracle/make_golden.py [case ...]
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
    only = set(sys.argv[1:])   # case names: regenerate just those (the manifest keeps the others)
    mf = os.path.join(GOLD, "manifest.json")
    manifest = json.load(open(mf)) if only and os.path.exists(mf) else {}
    for name in sorted(cases.CASES):
        if only and name not in only:
            continue
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
    if not only:
        make_dump_goldens()


def make_dump_goldens():
    """the reference's `-d` image of a tiny set (3 parts) and the tables it prints when it maps against its own dumps
    (tests/test_index_dump.py): the longQC.py --db flow, with the command line's k and with --fast's k = 15 index"""
    import gzip
    import hashlib
    import test_index_dump as tid
    T, Q = tid.tiny_set()
    with tempfile.TemporaryDirectory() as d:
        tf, qf = os.path.join(d, "t.fq"), os.path.join(d, "q.fq")
        T.write_fastx(tf); Q.write_fastx(qf)
        db, db15 = os.path.join(d, "db.mmi"), os.path.join(d, "db15.mmi")
        # the reference only survives -d when a query file is given too (without one it reads an uninitialised vector and dies before
        # the first part is indexed: minimap2-coverage.c:406-444 -- so longQC.py --db never got its file from the reference)
        subprocess.run([REF, "-Y", "-k", "12", "-w", "5", "-I", "20K", "-d", db, tf, qf], check=True, capture_output=True)
        subprocess.run([REF, "-Y", "-k", "12", "-w", "10", "-d", db15, tf, qf], check=True, capture_output=True)
        with gzip.GzipFile(os.path.join(GOLD, "dump_tiny.mmi.gz"), "wb", mtime=0) as g:
            g.write(open(db, "rb").read())
        # an index built with OTHER parameters than the mapping command line's defaults (k = 12, w = 5): the reference sizes its per-minimizer
        # counters with the command line's sketch and indexes them with the index's, so it only survives when the former is the denser one
        # (w = 10 index: yes; the k = 15 index of longQC.py --fast: heap corruption, abort)
        open(os.path.join(GOLD, "dump_tiny_w10.md5"), "w").write(hashlib.md5(open(db15, "rb").read()).hexdigest() + "\n")
        for name, f in (("dump_tiny.map.tsv", db), ("dump_tiny_w10.map.tsv", db15)):
            out = subprocess.run([REF] + "-Y -l 0 -q 160 -p 80 -t 4".split() + [f, qf], capture_output=True)
            print(name, "rc", out.returncode, out.stdout.count(b"\n"), "rows")
            open(os.path.join(GOLD, name), "wb").write(out.stdout)


if __name__ == "__main__":
    main()
