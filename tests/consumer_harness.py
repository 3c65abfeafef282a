"""Consumer-side parity harness: the UNMODIFIED lq_coverage.LqCoverage of the reference, run on a coverage table.

LongQC's own entry point cannot start in this container (no pysam / edlib / matplotlib / h5py), but lq_coverage.py itself only
needs numpy, pandas, scipy, sklearn and the vendored mixEM; matplotlib is stubbed (plt.hist == np.histogram), pandas' Arrow-backed
strings are switched off (lq_coverage.py:216 compares column 4 with '0'), and NumPy's global RNG is seeded because sklearn's
GaussianMixture falls back to it (lq_coverage.py:588 passes no random_state).  Test infrastructure only."""
import os
import sys
import types

import numpy as np

REF_DIR = "/root/reference"


def available():
    return os.path.exists(os.path.join(REF_DIR, "lq_coverage.py"))


def _stub_matplotlib():
    if "matplotlib" in sys.modules:   # the real one, or our stub from an earlier call (lq_coverage keeps its reference to it)
        return
    mpl = types.ModuleType("matplotlib")
    mpl._lq_stub = True
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")

    def hist(x, bins=10, density=False, **k):
        h, e = np.histogram(x, bins=bins, density=density)
        plt._lq_last_hist = (np.asarray(h), np.asarray(e))   # lq_coverage.py:234-241: the coverage histogram behind is_low_coverage()
        return h, e, None
    plt.hist = hist
    for name in ("close", "figure", "grid", "axvline", "axhline", "xlabel", "ylabel", "legend", "savefig", "plot", "xlim", "ylim",
                 "subplots", "title", "bar", "boxplot", "scatter", "fill_between"):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def consumer_fields(table_path, seed=12345):
    """the fields longQC.py reports from the table (lq_coverage.py:211-285): deterministic ones + the seeded GMM's mean / sd"""
    import warnings
    import pandas as pd
    _stub_matplotlib()
    try:
        pd.set_option("future.infer_string", False)
    except Exception:
        pass
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import lq_coverage
        np.random.seed(seed)
        c = lq_coverage.LqCoverage(table_path)
    return {
        "unmapped_frac_trimmed": float(c.unmapped_frac_trimmed),      # share of rows with column 5 == 0.0   (lq_coverage.py:212)
        "unmapped_frac_untrimmed": float(c.unmapped_frac_untrimmed),  # share of rows with column 2 == 0     (:214)
        "unmapped_frac_med": float(c.unmapped_frac_med),              # non-sense reads: column 4 == '0'     (:216)
        "high_div_frac": float(c.high_div_frac),                      # (:220-224)
        "mean": float(c.get_mean()), "sd": float(c.get_sd()), "cov_main": float(c.cov_main),
        "hist_density": [float(x) for x in sys.modules["matplotlib.pyplot"]._lq_last_hist[0]],
        "hist_edges": [float(x) for x in sys.modules["matplotlib.pyplot"]._lq_last_hist[1]],
        "low_coverage": bool(c.is_low_coverage()) if c.is_low_coverage() is not None else None,
    }
