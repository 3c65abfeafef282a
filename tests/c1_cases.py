"""BASELINE.json configs[0] ("sampleqc -x pb-sequel on 1k synthetic 10 kb fastq reads, CPU minimap2-coverage, 4 host threads"):
the argv longQC.py issues for it (longQC.py:177-231, 438-445) on seeded inputs, clean and with junk/adapter reads."""
import numpy as np

FLAGS = "-Y -l 0 -q 160 -k 12 -w 5 -I 4G -p 80 -t 4"
NAMES = ("c1_sampleqc", "c1_junk")
# more than 131 072 target reads in one index part: the rid>>16 digit of the seed key takes three values, so the s48 level of the exact
# seed sort is a walk over few, large regions (the regime of the multi-GPU runs, where the replicated index holds N x 100 000 reads)
EXTRA = ("many_targets",)
OPTS = dict(min_score_med=80, min_score_good=160)


def make(name):
    from longqc_b200 import synth
    import cases
    if name == "c1_sampleqc":
        return synth.standard_set(1000, 10000, 0.13, seed=7, n_query=1000)     # every read is a query (<= 5 000 reads: longQC.py:414-418)
    if name == "c1_junk":
        rng = np.random.default_rng(8)
        n, L = 1000, 10000
        g = synth.make_genome(n * L // 30, rng, gc_blocks=True, block=5000)
        good = synth.simulate_reads(g, 800, L, 0.13, rng)
        junk = synth.random_reads(100, L, rng)
        adp = synth.with_adapters(synth.simulate_reads(g, 100, L, 0.13, rng), cases.ADP, cases.ADP, rng)
        T = synth.ReadSet.concat([good, junk, adp]).shuffled(rng).renamed()
        return T, T
    if name == "many_targets":
        rng = np.random.default_rng(77)
        n, L = 140000, 1500
        g = synth.add_tandem_repeats(synth.make_genome(n * L // 25, rng), rng, 300, unit_len=(2, 40), copies=(10, 150))
        T = synth.simulate_reads(g, n, L, 0.08, rng)
        return T, T.subset(np.sort(rng.choice(n, 60, replace=False)))
    raise KeyError(name)
