"""longqc_b200 -- B200-native drop-in for LongQC's overlap/coverage pass.

The product is the C-ABI library ``liblqcov.so`` (CUDA sm_100a kernels + C host, see include/lqcov.h)
and the two executables ``bin/minimap2-coverage`` and ``bin/sdust`` that LongQC spawns
(lq_exec.py:13-38, lq_mask.py:17-23 of the reference).  This Python package is plumbing only:
ctypes bindings used by the tests and bench.py, the synthetic read generator, the multi-GPU
launcher glue (torch.distributed) and an ``LqExec`` mirror of the reference's process wrapper.
"""
from ._lib import (LqcovError, Opt, Coverage, lib_path, load, reads_struct, sketch, sdust_table,  # noqa: F401
                   coverage_table, bin_path)
from .exec import LqExec  # noqa: F401

__all__ = ["LqcovError", "Opt", "Coverage", "lib_path", "load", "sketch", "sdust_table", "coverage_table", "bin_path", "LqExec"]
