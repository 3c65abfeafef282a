"""Seeded parity cases shared by the golden generator (oracle/make_golden.py), the CPU tests (oracle vs
golden) and the GPU tests (CUDA path vs golden / oracle).  A case = generator spec + option overrides
(field names of lqcov_opt_t) + the argv the reference binary was run with."""
import hashlib

import numpy as np

CASES = {
    # name: (generator spec, options, reference argv flags)
    "plain_pb": (("std", 300, 6000, 0.13, 101, 40), dict(min_score_med=80, min_score_good=160), "-Y -l 0 -q 160 -k 12 -w 5 -I 4G -p 80"),
    "plain_ont": (("std", 250, 8000, 0.15, 102, 30), dict(min_score_med=160, min_score_good=160), "-Y -l 0 -q 160 -k 12 -w 5 -I 4G -p 160"),
    "tandem": (("tandem", 103, 250, 5000, 30), dict(min_score_med=80, min_score_good=160), "-Y -l 0 -q 160 -k 12 -w 5 -I 4G -p 80"),
    "tandem_parts": (("tandem", 104, 250, 5000, 30), dict(min_score_med=80, min_score_good=160, batch_size=300000), "-Y -l 0 -q 160 -k 12 -w 5 -I 300K -p 80"),
    "parts": (("std", 300, 6000, 0.13, 105, 40), dict(min_score_med=80, min_score_good=160, batch_size=400000), "-Y -l 0 -q 160 -k 12 -w 5 -I 400K -p 80"),
    "fast_k15": (("std", 250, 6000, 0.10, 106, 30), dict(k=15, min_score_med=160, min_score_good=160), "-Y -l 0 -q 160 -k 15 -w 5 -I 4G -p 160"),
    "ava_X": (("std", 250, 5000, 0.13, 107, 30), dict(ava=1, min_score_med=80, min_score_good=160), "-X -l 0 -q 160 -k 12 -w 5 -p 80"),
    "spike_hpc_filter": (("spike", 108, 200, 5000), dict(is_hpc=1, k=15, w=10, min_coverage=1, filter=1), "-Y -Hk15 -w 10 -c 1 -l 0 --filter"),
    "ambiguous_fasta": (("nfasta", 109, 250, 5000, 30), dict(min_score_med=80, min_score_good=160), "-Y -l 0 -q 160 -k 12 -w 5 -p 80"),
    "junk_adapters": (("junk", 110, 250, 5000, 40), dict(min_score_med=80, min_score_good=160), "-Y -l 0 -q 160 -k 12 -w 5 -p 80"),
    # ~300x of a 6 kb genome cut into >= 6 index parts: the COVT gate (esterr.c:87-91, lambda/len > 150) closes for the later parts
    "covt_gate": (("deep", 600, 3000, 0.10, 32, 300.0, 25), dict(min_score_med=80, min_score_good=160, batch_size=150000), "-Y -l 0 -q 160 -k 12 -w 5 -I 150K -p 80"),
    # BASELINE configs[3] flavour: ONT ultra-long 50 kb reads, -x ont-rapid (-p 160): long chains, thousands of anchors per target
    "ultralong": (("std", 60, 50000, 0.15, 31, 12), dict(min_score_med=160, min_score_good=160), "-Y -l 0 -q 160 -k 12 -w 5 -I 4G -p 160"),
    # BASELINE configs[4] reduced: mixed-GC genome, 10 % iid junk + 10 % adapter/low-complexity reads, several index parts, -x pb-sequel
    # k > 15: the open-address index (lq_widx.cu).  longQC.py:222-226: `-x pb-hifi --fast` runs -k 19 -w 10; then other widths / several parts
    "hifi_fast_k19": (("std", 250, 8000, 0.01, 112, 30), dict(k=19, w=10, min_score_med=80, min_score_good=160), "-Y -l 0 -q 160 -k 19 -w 10 -I 4G -p 80"),
    "wide_k17_parts": (("std", 300, 6000, 0.04, 113, 40), dict(k=17, min_score_med=80, min_score_good=160, batch_size=400000), "-Y -l 0 -q 160 -k 17 -w 5 -I 400K -p 80"),
    "wide_k28_tandem": (("tandem", 114, 200, 5000, 30), dict(k=28, w=10, min_score_med=80, min_score_good=160), "-Y -l 0 -q 160 -k 28 -w 10 -I 4G -p 80"),
    "wide_k16_nfasta": (("nfasta", 115, 200, 5000, 30), dict(k=16, min_score_med=80, min_score_good=160), "-Y -l 0 -q 160 -k 16 -w 5 -p 80"),
    "c5_small": (("junk", 111, 1500, 6000, 400), dict(min_score_med=80, min_score_good=160, batch_size=2000000), "-Y -l 0 -q 160 -k 12 -w 5 -I 2M -p 80"),
}

ADP = b"ATCTCTCTCAACAACAACAACGGAGGAGGAGGAAAAGAGAGAGAT"  # longQC.py:184 (pb-sequel preset)


def make_case(name):
    """-> (targets ReadSet, queries ReadSet)"""
    from longqc_b200 import synth
    spec = CASES[name][0]
    kind = spec[0]
    if kind == "std":
        _, n, L, err, seed, nq = spec
        return synth.standard_set(n, L, err, seed=seed, n_query=nq)
    if kind == "deep":
        _, n, L, err, seed, cov, nq = spec
        return synth.standard_set(n, L, err, seed=seed, coverage=cov, n_query=nq)
    if kind == "tandem":
        _, seed, n, L, nq = spec
        rng = np.random.default_rng(seed)
        g = synth.add_tandem_repeats(synth.make_genome(n * L // 30, rng), rng, 40, unit_len=(2, 40), copies=(10, 150))
        T = synth.simulate_reads(g, n, L, 0.10, rng)
        return T, T.subset(np.sort(rng.choice(n, nq, replace=False)))
    if kind == "spike":  # a 4 kb control sequence as the only target; some queries carry it (longQC.py:553-557)
        _, seed, nq, L = spec
        rng = np.random.default_rng(seed)
        ctrl = synth.make_genome(4100, rng)
        g = synth.make_genome(nq * L // 20, rng)
        Q = synth.simulate_reads(g, nq - 30, L, 0.13, rng)
        C_ = synth.simulate_reads(ctrl, 30, 3000, 0.13, rng, name_prefix=b"c")
        Q = synth.ReadSet.concat([Q, C_]).shuffled(rng)
        T = synth.simulate_reads(ctrl, 1, 4100, 0.0, rng, name_prefix=b"control")
        T = synth.ReadSet(T.seq, T.seq_off, None, T.names)
        return T, Q
    if kind == "nfasta":
        _, seed, n, L, nq = spec
        rng = np.random.default_rng(seed)
        T, Q = synth.standard_set(n, L, 0.10, seed=seed, n_query=nq)
        T = synth.sprinkle_n(T, 0.01, rng)
        Q = synth.sprinkle_n(Q, 0.005, rng)
        return T, synth.ReadSet(Q.seq, Q.seq_off, None, Q.names)
    if kind == "junk":  # config-5 flavour: 10 % iid junk reads + 10 % reads with adapters / low-complexity inserts
        _, seed, n, L, nq = spec
        rng = np.random.default_rng(seed)
        g = synth.make_genome(n * L // 30, rng, gc_blocks=True, block=5000)
        good = synth.simulate_reads(g, int(n * 0.8), L, 0.13, rng)
        junk = synth.random_reads(n // 10, L, rng)
        adp = synth.with_adapters(synth.simulate_reads(g, n - int(n * 0.8) - n // 10, L, 0.13, rng), ADP, ADP, rng)
        T = synth.ReadSet.concat([good, junk, adp]).shuffled(rng).renamed()
        return T, T.subset(np.sort(rng.choice(T.n, nq, replace=False)))
    raise KeyError(kind)


def inputs_md5(T, Q):
    h = hashlib.md5()
    for rs in (T, Q):
        h.update(rs.seq.tobytes()); h.update(rs.seq_off.tobytes())
        if rs.qual is not None:
            h.update(rs.qual.tobytes())
        h.update(b"\n".join(rs.names))
    return h.hexdigest()


def opts(name):
    return dict(CASES[name][1])
