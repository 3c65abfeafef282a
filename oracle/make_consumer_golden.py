#!/usr/bin/env python
"""Makes tests/golden/c1_*.tsv and tests/golden/consumer_c1.json: BASELINE.json configs[0] (1 000 synthetic 10 kb reads, all of
them queries, `-Y -l 0 -q 160 -k 12 -w 5 -I 4G -p 80 -t 4`) in a clean and a junk/adapter flavour, run through the UNMODIFIED
reference binary (oracle/_ref) and the unmodified lq_coverage.LqCoverage (tests/consumer_harness.py).  Run in the container that
has /root/reference; the outputs are committed so that the pin travels."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import consumer_harness  # noqa: E402
import c1_cases  # noqa: E402


def main():
    out = {}
    for name in c1_cases.NAMES + c1_cases.EXTRA:
        T, Q = c1_cases.make(name)
        d = tempfile.mkdtemp(prefix="lqc1_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        tf, qf = os.path.join(d, "t.fq"), os.path.join(d, "q.fq")
        T.write_fastx(tf); Q.write_fastx(qf)
        table = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "minimap2-coverage")] + c1_cases.FLAGS.split() + [tf, qf],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        path = os.path.join(ROOT, "tests", "golden", name + ".tsv")
        open(path, "wb").write(table)
        if name in c1_cases.NAMES:
            out[name] = consumer_harness.consumer_fields(path)
        print(name, table.count(b"\n"), "rows", out.get(name))
        for f in (tf, qf):
            os.unlink(f)
        os.rmdir(d)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "consumer_c1.json"), "w"), indent=1, sort_keys=True)
    # the reduced configs[4] case (tests/cases.py "c5_small": its table golden is made by make_golden.py)
    c5 = consumer_harness.consumer_fields(os.path.join(ROOT, "tests", "golden", "c5_small.tsv"))
    json.dump({"c5_small": c5}, open(os.path.join(ROOT, "tests", "golden", "consumer_c5.json"), "w"), indent=1, sort_keys=True)
    print("c5_small", {k: v for k, v in c5.items() if not k.startswith("hist_")})


if __name__ == "__main__":
    main()
