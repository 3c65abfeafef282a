"""Seeded synthetic long-read generator.

Not part of the hot path: it produces the inputs the hot path is measured and
parity-tested on (bench.py, tests/, oracle/make_golden.py).  Everything is
driven by ``numpy.random.default_rng(seed)`` so a (seed, parameters) pair names
a read set exactly; the reference has no generator of its own (it has no tests).

Read model: genome of G iid bases (optionally 50 kb blocks with
GC in [0.30, 0.70]); reads start uniformly, 50 % are reverse-complemented,
errors are split evenly between substitutions, insertions and deletions, names
are ``r<i>``, qualities are uniform Q5..Q19 (``sequel=True``: all '!' as
lq_utils.py:249 writes for BAM input).
"""
from __future__ import annotations

import dataclasses
import gzip
import io
from typing import Optional, Sequence

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclasses.dataclass
class ReadSet:
    """Reads in the layout the C-ABI takes: one ASCII blob + offsets (n+1)."""

    seq: np.ndarray          # uint8, concatenated ASCII bases
    seq_off: np.ndarray      # int64[n+1]
    qual: Optional[np.ndarray]  # uint8, same length/offsets as seq (None = FASTA)
    names: list              # list[bytes]

    @property
    def n(self) -> int:
        return len(self.seq_off) - 1

    @property
    def n_bases(self) -> int:
        return int(self.seq_off[-1])

    def lengths(self) -> np.ndarray:
        return np.diff(self.seq_off)

    def subset(self, idx: Sequence[int]) -> "ReadSet":
        idx = np.asarray(idx, dtype=np.int64)
        lens = self.lengths()[idx]
        off = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        # gather indices for all bases of the selected reads
        starts = self.seq_off[idx]
        flat = np.repeat(starts - off[:-1], lens) + np.arange(off[-1], dtype=np.int64)
        return ReadSet(self.seq[flat], off, None if self.qual is None else self.qual[flat],
                       [self.names[i] for i in idx])

    @staticmethod
    def concat(sets: Sequence["ReadSet"]) -> "ReadSet":
        seq = np.concatenate([s.seq for s in sets])
        lens = np.concatenate([s.lengths() for s in sets])
        off = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        if any(s.qual is None for s in sets):
            qual = None
        else:
            qual = np.concatenate([s.qual for s in sets])
        names = [n for s in sets for n in s.names]
        return ReadSet(seq, off, qual, names)

    def renamed(self, prefix: bytes = b"r") -> "ReadSet":
        return ReadSet(self.seq, self.seq_off, self.qual, [prefix + str(i).encode() for i in range(self.n)])

    def shuffled(self, rng: np.random.Generator) -> "ReadSet":
        return self.subset(rng.permutation(self.n))

    # ---- text formats (what longQC.py hands the binaries) ----
    def write_fastx(self, path: str, line_width: int = 0) -> None:
        """FASTQ if qualities are present, FASTA otherwise; ``.gz`` suffix => gzip."""
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "wb") as raw:
            fh = io.BufferedWriter(raw, buffer_size=1 << 22) if not str(path).endswith(".gz") else raw
            seqb = self.seq.tobytes()
            qualb = None if self.qual is None else self.qual.tobytes()
            off = self.seq_off
            for i, name in enumerate(self.names):
                s, e = int(off[i]), int(off[i + 1])
                if qualb is None:
                    fh.write(b">" + name + b"\n")
                    if line_width > 0:
                        for p in range(s, e, line_width):
                            fh.write(seqb[p:min(e, p + line_width)] + b"\n")
                    else:
                        fh.write(seqb[s:e] + b"\n")
                else:
                    fh.write(b"@" + name + b"\n" + seqb[s:e] + b"\n+\n" + qualb[s:e] + b"\n")
            if fh is not raw:
                fh.flush()


def make_genome(G: int, rng: np.random.Generator, gc_blocks: bool = False, block: int = 50_000) -> np.ndarray:
    """uint8 codes 0..3 (A,C,G,T)."""
    if not gc_blocks:
        return rng.integers(0, 4, size=G, dtype=np.uint8)
    out = np.empty(G, dtype=np.uint8)
    for s in range(0, G, block):
        e = min(G, s + block)
        gc = rng.uniform(0.30, 0.70)
        p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
        out[s:e] = rng.choice(4, size=e - s, p=p).astype(np.uint8)
    return out


def add_tandem_repeats(genome: np.ndarray, rng: np.random.Generator, n_loci: int, unit_len=(2, 60),
                       copies=(10, 200)) -> np.ndarray:
    """Overwrite ``n_loci`` random loci with tandem repeats (tie-order stress)."""
    g = genome.copy()
    G = len(g)
    for _ in range(n_loci):
        ul = int(rng.integers(unit_len[0], unit_len[1] + 1))
        nc = int(rng.integers(copies[0], copies[1] + 1))
        unit = rng.integers(0, 4, size=ul, dtype=np.uint8)
        rep = np.tile(unit, nc)
        if len(rep) >= G:
            rep = rep[: G // 2]
        st = int(rng.integers(0, G - len(rep)))
        g[st:st + len(rep)] = rep
    return g


def _mutate(codes: np.ndarray, lens: np.ndarray, err: float, rng: np.random.Generator):
    """Apply sub/ins/del (err split evenly) to a flat code array made of reads of ``lens``."""
    n = len(codes)
    if err <= 0 or n == 0:
        return codes, lens
    r = rng.random(n, dtype=np.float32)
    e3 = np.float32(err / 3.0)
    sub = r < e3
    dele = (r >= e3) & (r < 2 * e3)
    ins = (r >= 2 * e3) & (r < 3 * e3)
    codes = codes.copy()
    nsub = int(sub.sum())
    if nsub:
        codes[sub] = (codes[sub] + rng.integers(1, 4, size=nsub, dtype=np.uint8)) & 3
    cnt = (~dele).astype(np.int64) + ins.astype(np.int64)
    # never delete a read down to nothing: keep the first base of every read
    starts = np.zeros(len(lens), dtype=np.int64)
    np.cumsum(lens[:-1], out=starts[1:])
    nz = lens > 0
    cnt[starts[nz]] = np.maximum(cnt[starts[nz]], 1)
    out = np.repeat(codes, cnt)
    # positions of inserted bases: the second copy of a kept+ins base, or the only copy of a del+ins(impossible)
    cs = np.cumsum(cnt)
    ins_idx = np.nonzero(ins & (cnt == 2))[0]
    if len(ins_idx):
        out[cs[ins_idx] - 1] = rng.integers(0, 4, size=len(ins_idx), dtype=np.uint8)
    read_id = np.repeat(np.arange(len(lens)), lens)
    new_lens = np.bincount(read_id, weights=cnt, minlength=len(lens)).astype(np.int64)
    return out, new_lens


def simulate_reads(genome: np.ndarray, n: int, L: int, err: float, rng: np.random.Generator,
                   len_cv: float = 0.0, sequel: bool = False, chunk: int = 2000,
                   name_prefix: bytes = b"r", name_start: int = 0) -> ReadSet:
    """n reads of nominal length L (gamma-distributed with CV ``len_cv`` if > 0) sampled from ``genome``."""
    G = len(genome)
    seqs, lens_all = [], []
    comp = np.array([3, 2, 1, 0], dtype=np.uint8)
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        if len_cv > 0:
            shape = 1.0 / (len_cv * len_cv)
            lens = np.maximum(200, rng.gamma(shape, L / shape, size=m)).astype(np.int64)
        else:
            lens = np.full(m, L, dtype=np.int64)
        lens = np.minimum(lens, G)
        starts = (rng.random(m) * (G - lens + 1)).astype(np.int64)
        rev = rng.random(m) < 0.5
        off = np.zeros(m + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        within = np.arange(off[-1], dtype=np.int64) - np.repeat(off[:-1], lens)
        rl = np.repeat(lens, lens)
        rr = np.repeat(rev, lens)
        gpos = np.repeat(starts, lens) + np.where(rr, rl - 1 - within, within)
        codes = genome[gpos]
        codes = np.where(rr, comp[codes], codes).astype(np.uint8)
        codes, lens = _mutate(codes, lens, err, rng)
        seqs.append(codes)
        lens_all.append(lens)
    codes = np.concatenate(seqs) if seqs else np.zeros(0, dtype=np.uint8)
    lens = np.concatenate(lens_all) if lens_all else np.zeros(0, dtype=np.int64)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    seq = _ACGT[codes]
    if sequel:
        qual = np.full(len(seq), 33, dtype=np.uint8)
    else:
        qual = (rng.integers(5, 20, size=len(seq), dtype=np.uint8) + 33).astype(np.uint8)
    names = [name_prefix + str(name_start + i).encode() for i in range(len(lens))]
    return ReadSet(seq, off, qual, names)


def random_reads(n: int, L: int, rng: np.random.Generator, **kw) -> ReadSet:
    """iid junk reads (no overlap partners): the 'non-sense read' population of config 5."""
    g = rng.integers(0, 4, size=max(n * L, 1), dtype=np.uint8)
    rs = simulate_reads(g, n, L, 0.0, rng, **kw)
    return rs


def with_adapters(rs: ReadSet, adp5: bytes, adp3: bytes, rng: np.random.Generator, low_complexity: float = 0.3) -> ReadSet:
    """Attach preset adapters (longQC.py:184-185) to both ends; a share also gets a low-complexity insert."""
    seqb = rs.seq.tobytes()
    out_seq, out_len = [], []
    for i in range(rs.n):
        s = seqb[int(rs.seq_off[i]):int(rs.seq_off[i + 1])]
        if rng.random() < low_complexity and len(s) > 200:
            unit = bytes(_ACGT[rng.integers(0, 4, size=int(rng.integers(1, 7)))])
            ins = unit * int(rng.integers(20, 120))
            p = int(rng.integers(0, len(s)))
            s = s[:p] + ins + s[p:]
        s = adp5 + s + adp3
        out_seq.append(s)
        out_len.append(len(s))
    seq = np.frombuffer(b"".join(out_seq), dtype=np.uint8).copy()
    off = np.zeros(rs.n + 1, dtype=np.int64)
    np.cumsum(np.asarray(out_len, dtype=np.int64), out=off[1:])
    qual = None
    if rs.qual is not None:
        qual = (rng.integers(5, 20, size=len(seq), dtype=np.uint8) + 33).astype(np.uint8)
    return ReadSet(seq, off, qual, list(rs.names))


def sprinkle_n(rs: ReadSet, frac: float, rng: np.random.Generator) -> ReadSet:
    """Replace a fraction of bases with 'N' (ambiguous-base edge cases)."""
    seq = rs.seq.copy()
    m = rng.random(len(seq)) < frac
    seq[m] = ord("N")
    return ReadSet(seq, rs.seq_off, rs.qual, rs.names)


def standard_set(n_reads: int, read_len: int, err: float, seed: int, coverage: float = 30.0,
                 n_query: int = 5000, **kw):
    """(targets, queries) for a BASELINE.json-style config: G = N*L/coverage, queries = a random subsample
    (what longQC.py:414-418 writes as subsample.fastq), in target-file order."""
    rng = np.random.default_rng(seed)
    G = max(int(n_reads * read_len / coverage), 2 * read_len)
    genome = make_genome(G, rng, gc_blocks=kw.pop("gc_blocks", False))
    targets = simulate_reads(genome, n_reads, read_len, err, rng, **kw)
    nq = min(n_query, n_reads)
    qidx = np.sort(rng.choice(n_reads, size=nq, replace=False))
    return targets, targets.subset(qidx)
