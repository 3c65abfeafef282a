"""The deterministic part of the table's consumer, lq_coverage.LqCoverage.__est_coverage (reference lq_coverage.py:211-241),
restated on numpy: the three zero-coverage fractions (`unmapped_frac_med` is north_star's non-sense-read fraction), the
high-divergence fraction, and the coverage histogram whose bins follow the fitted main component (mean, variance).

The Gaussian-mixture fit itself (lq_coverage.py:570-621, sklearn, unseeded in the reference) is NOT restated: `coverage_histogram`
takes its two outputs as arguments.  Pinned against the unmodified class in tests/test_zz_c1_consumer.py (build container) through
the committed tests/golden/consumer_*.json."""
from __future__ import annotations

import numpy as np

# column map of lq_coverage.py:77-85
QLENGTH, N_MBASE, MED_READ_COV_CORS, T1_COVERAGE, QV, DIV, COVERAGE = 1, 2, 4, 5, 6, 7, 8
DIV_SCORE_THRESHOLD, COV_THRESHOLD_FOR_DIV_SC = 0.25, 25


def parse_table(table: bytes):
    """rows of the minimap2-coverage table -> columns, typed the way pd.read_table(dtype={3: str, 4: str}) types them"""
    rows = [ln.split(b"\t") for ln in table.split(b"\n") if ln]
    col = lambda i, f: np.array([f(r[i]) for r in rows])
    return {"n": len(rows), QLENGTH: col(QLENGTH, int), N_MBASE: col(N_MBASE, int), MED_READ_COV_CORS: [r[MED_READ_COV_CORS] for r in rows],
            T1_COVERAGE: col(T1_COVERAGE, float), QV: col(QV, float), DIV: col(DIV, float), COVERAGE: col(COVERAGE, float)}


def zero_fractions(table: bytes):
    """lq_coverage.py:212-224"""
    t = parse_table(table)
    n = t["n"]
    med0 = np.array([c == b"0" for c in t[MED_READ_COV_CORS]])
    return {
        "unmapped_frac_trimmed": float(np.count_nonzero(t[T1_COVERAGE] == 0.0) / n),
        "unmapped_frac_untrimmed": float(np.count_nonzero(t[N_MBASE] == 0) / n),
        "unmapped_frac_med": float(np.count_nonzero(med0) / n),
        "high_div_frac": float(np.count_nonzero((t[DIV] >= DIV_SCORE_THRESHOLD) & (t[T1_COVERAGE] >= COV_THRESHOLD_FOR_DIV_SC) & ~med0) / n),
    }


def coverage_histogram(table: bytes, mean_main: float, cov_main: float):
    """lq_coverage.py:234-241: density histogram of lambda / read length, bins of mean/10 up to mean + 10 sd + mean/10"""
    t = parse_table(table)
    x = t[N_MBASE] / t[QLENGTH]
    bins = np.arange(0, mean_main + 10 * np.sqrt(cov_main) + mean_main / 10, mean_main / 10)
    h, e = np.histogram(x, bins=bins, density=True)
    return h, e
