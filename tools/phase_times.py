"""Host-clock time of every C-ABI call of one coverage job (diagnostic; run on the GPU box)."""
import ctypes as C
import sys
import time

sys.path.insert(0, ".")
import longqc_b200 as L
from longqc_b200 import _lib, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
T, Q = synth.standard_set(n, 8000, 0.15, seed=1, n_query=min(5000, n // 20))
lib = L.load()
opt = L.Opt(min_score_med=160, min_score_good=160)
cov = L.Coverage(opt)
tk, qk = _lib.reads_struct(T), _lib.reads_struct(Q)
for it in range(3):
    lib.lqcov_reset(cov._h)
    t = [time.perf_counter()]
    lib.lqcov_set_queries(cov._h, C.byref(qk.st)); t.append(time.perf_counter())
    lib.lqcov_part_sketch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.lqcov_part_sketch(cov._h, C.byref(tk.st), 0); t.append(time.perf_counter())
    lib.lqcov_part_finish.argtypes = [C.c_void_p, C.c_void_p]
    lib.lqcov_part_finish(cov._h, C.byref(tk.st)); t.append(time.perf_counter())
    lib.lqcov_map_part.argtypes = [C.c_void_p]
    lib.lqcov_map_part(cov._h); t.append(time.perf_counter())
    tab = cov.table(); t.append(time.perf_counter())
    names = ["set_queries", "part_sketch", "part_finish", "map_part", "table"]
    print("iter", it, " ".join("%s=%.1fms" % (nm, 1e3 * (b - a)) for nm, a, b in zip(names, t, t[1:])), "total=%.1fms" % (1e3 * (t[-1] - t[0])))
    print("   stats", {k: round(v, 1) for k, v in cov.stats().items() if k.startswith("t_")})
