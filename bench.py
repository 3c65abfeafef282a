#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: read Gbases/s through minimap2-coverage on B200 vs host CPU.

One "step" = one complete all-vs-subsample coverage job on the synthetic workload (sketch every
target read, build the minimizer index, map the sampled queries, emit the table):
    Gbases/s = (target bases indexed + query bases mapped) / step time.

  value  whole-job throughput with the reads' ASCII bases already resident in HBM
  e2e    the same job through the C ABI with HOST (pinned) buffers: H2D of every base and D2H of the
         per-query results are inside the timed region
  roofline      the dominant kernel of the step, timed with CUDA events on the library's stream
  cpu_baseline  the reference CPU binary (oracle/_ref) on a bounded sample of the same workload
  --impl reference   the reference's own CPU implementation as the measured arm (bounded sample per step)

N > 1: launched under torchrun; see longqc_b200/dist.py for the sharding (targets sharded for sketching,
minimizer counts all-reduced, index replicated, queries sharded).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLAGS = "-Y -l 0 -q 160 -k 12 -w 5 -I 4G -p 160"   # longQC.py:171-231 for -x ont-ligation
METRIC = "read Gbases/s through minimap2-coverage (target bases indexed + query bases mapped per second)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100_000)       # BASELINE.json configs[1]
    ap.add_argument("--read-len", type=int, default=8_000)
    ap.add_argument("--err", type=float, default=0.15)
    ap.add_argument("--queries", type=int, default=5_000)
    ap.add_argument("--seed", type=int, default=20260925)
    ap.add_argument("--cpu-sample-reads", type=int, default=12_000)
    ap.add_argument("--cpu-sample-queries", type=int, default=600)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", action="store_true", help="time the CPU reference on the bounded sample instead of the whole workload")
    ap.add_argument("--ref-full", action="store_true", help="--impl reference at N > 1: time the whole N-GPU workload (N x 25 s per step)")
    ap.add_argument("--no-cli", action="store_true", help="skip the drop-in executable's end-to-end run (cli_e2e)")
    ap.add_argument("--no-sdust", action="store_true", help="skip the sdust line")
    ap.add_argument("--preset", default="ont-ligation", choices=["ont-ligation", "ont-rapid", "pb-sequel"],
                    help="LongQC preset whose native flags are used (longQC.py:171-231); the default is BASELINE configs[1]'s")
    a = ap.parse_args()
    global FLAGS
    if a.preset == "pb-sequel":      # minimap2_med_score_threshold = 80
        FLAGS = "-Y -l 0 -q 160 -k 12 -w 5 -I 4G -p 80"
    return a


def workload_name(a, world=1):
    s = "%dk synthetic %s %d kb reads (%.0f%% error), -x %s (%s), %d sampled queries" % (
        a.reads // 1000, "Sequel" if a.preset == "pb-sequel" else "ONT", a.read_len // 1000, a.err * 100, a.preset, FLAGS, a.queries)
    if world > 1:   # weak scaling: every rank brings its own reads of one shared genome, the queries are split
        s += "; x%d GPUs = %dk target reads in one replicated index, %d queries per GPU" % (world, a.reads * world // 1000, a.queries // world)
    return s


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _nvml_handle(self):
        """NVML handle of the sampled device (fast path: a sample every 25 ms instead of one nvidia-smi process per 200 ms); None
        when NVML is not usable, then nvidia-smi is polled as before."""
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            return pynvml, h
        except Exception:
            return None

    def _run_nvml(self, nv):
        pynvml, h = nv
        HW, SWT, HWT, PWR = 0x8, 0x20, 0x40, 0x4       # nvmlClocksThrottleReason{HwSlowdown, SwThermalSlowdown, HwThermalSlowdown, SwPowerCap}
        act = lambda m, b: "Active" if m & b else "Not Active"
        try:
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception:
            mx = 0
        while not self._stop.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    m = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                except Exception:
                    m = 0
                self.rows.append([str(int(sm)), str(int(mx)), "0", act(m, HW), act(m, HWT), act(m, SWT), act(m, PWR)])
            except Exception:
                break
            self._stop.wait(0.025)

    def _run(self):
        nv = self._nvml_handle()
        if nv is not None:
            self._run_nvml(nv)
            if self.rows:
                return
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nm in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- data
def make_data(a, n_reads, n_query, rank=0):
    from longqc_b200 import synth
    return synth.standard_set(n_reads, a.read_len, a.err, seed=a.seed + 1000 * rank, n_query=n_query)


def global_workload(a, world):
    """(targets, queries) of the whole job exactly as our arm's ranks generate them (longqc_b200/dist.py rank_inputs), rank-major"""
    from longqc_b200 import dist as lqdist, synth
    parts = [lqdist.rank_inputs(a, r, world) for r in range(world)]
    if world == 1:
        return parts[0]
    return synth.ReadSet.concat([p[0] for p in parts]), synth.ReadSet.concat([p[1] for p in parts])


class FastqFiles:
    """the workload as the FASTQ files LongQC hands the binaries, in /dev/shm (written once, outside every timed region)"""

    def __init__(self, targets, queries):
        self.d = tempfile.mkdtemp(prefix="lqbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        self.tf, self.qf, self.out = os.path.join(self.d, "t.fq"), os.path.join(self.d, "q.fq"), os.path.join(self.d, "out.tsv")
        targets.write_fastx(self.tf)
        queries.write_fastx(self.qf)
        self.n_bases = targets.n_bases + queries.n_bases
        self.bytes = os.path.getsize(self.tf) + os.path.getsize(self.qf)

    def run(self, cmd, env=None):
        """wall seconds of one invocation `cmd <targets> <queries> > out`, and the table it printed"""
        t0 = time.perf_counter()
        with open(self.out, "wb") as out:
            subprocess.run(cmd + [self.tf, self.qf], stdout=out, stderr=subprocess.DEVNULL, check=True, env=env)
        dt = time.perf_counter() - t0
        return dt, open(self.out, "rb").read()

    def close(self):
        for f in (self.tf, self.qf, self.out):
            if os.path.exists(f):
                os.unlink(f)
        os.rmdir(self.d)


def reference_cmd(threads):
    """the unmodified reference binary (oracle/_ref) -- or the oracle port when it was not built"""
    ref = os.path.join(ROOT, "oracle", "_ref", "minimap2-coverage")
    port = os.path.join(ROOT, "oracle", "lq_oracle_cli")
    if os.path.exists(ref):
        return [ref] + FLAGS.split() + ["-t", str(threads)], "reference", threads
    return [port, "cov"] + FLAGS.split(), "port", 1


def cpu_reference_run(a, targets, queries, threads, files=None):
    """Time the reference CPU implementation on (targets, queries).  Returns (Gbases/s, seconds, kind, cores, table bytes)."""
    own = files is None
    if own:
        files = FastqFiles(targets, queries)
    cmd, kind, cores = reference_cmd(threads)
    dt, table = files.run(cmd)
    if own:
        files.close()
    return files.n_bases / dt / 1e9, dt, kind, cores, table


def host_threads():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, min(n, 64))


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference_arm(a):
    """The unmodified reference binary (oracle/_ref/minimap2-coverage, all host threads) as the measured arm.
    N=1: every timed step is the WHOLE benchmarked workload (same inputs as our arm: same generator, same seed), ~25 s per step;
    the untimed warm-up steps run on a bounded sample (a CPU process has nothing to warm but the page cache).
    N>1: our arm's workload is N x as large (one replicated index over all ranks' reads) and would take N x as long per step on the
    CPU, so each step is a bounded sample of it unless --ref-full is given; `config.sample` says which."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    thr = host_threads()
    full = a.gpus == 1 or a.ref_full
    st, sq = make_data(a, a.cpu_sample_reads, a.cpu_sample_queries)
    for _ in range(a.warmup):
        cpu_reference_run(a, st, sq, thr)
    if full:
        targets, queries = global_workload(a, a.gpus)
        sample = "the whole workload (%d target reads, %d queries): same inputs as the GPU arm; warm-up steps on %d reads + %d queries" % (
            targets.n, queries.n, st.n, sq.n)
    else:
        targets, queries = st, sq
        sample = "%d target reads x %d b + %d queries of the same generator (bounded sample of the %d-GPU workload)" % (st.n, a.read_len, sq.n, a.gpus)
    vals, secs, kind, cores = [], [], None, None
    files = FastqFiles(targets, queries)
    for _ in range(a.steps):
        v, dt, kind, cores, _ = cpu_reference_run(a, targets, queries, thr, files)
        vals.append(v)
        secs.append(dt)
    files.close()
    v = (targets.n_bases + queries.n_bases) * len(secs) / sum(secs) / 1e9
    line = {"impl": "reference", "metric": METRIC,
            "value": v, "unit": "Gbases/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(a, a.gpus), "target_bases": int(targets.n_bases), "query_bases": int(queries.n_bases),
                       "sample": sample, "same_workload_as_gpu_arm": bool(full)},
            "cpu_baseline": {"value": v, "unit": "Gbases/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- our arm
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(scope):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one profiling scope per job, from the committed ncu launch list of
    this same command (profiles/*_traffic.json, written by tools/summarize_launches.py); None when no capture names the scope."""
    import glob
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            d = json.load(open(f))
            if scope in d:
                return d[scope]["dram_bytes_per_job"], os.path.relpath(f, ROOT)
        except Exception:
            pass
    return None, None


def profile_json(L):
    lib = L.load()
    lib.lqcov_profile_json.restype = C.c_size_t
    lib.lqcov_profile_json.argtypes = [C.c_char_p, C.c_size_t]
    n = lib.lqcov_profile_json(None, 0)
    buf = C.create_string_buffer(n + 16)
    lib.lqcov_profile_json(buf, n + 16)
    return json.loads(buf.value.decode())


def sdust_line(a, L, reads):
    """the second executable of the path: `sdust` over EVERY read of the input (lq_mask.py:17-23), through the C ABI with host buffers"""
    import torch
    t0 = time.perf_counter()
    tab = L.sdust_table(reads)
    torch.cuda.synchronize()
    lib = L.load()
    lib.lqcov_profile_reset(); lib.lqcov_profile_enable(1)
    t1 = time.perf_counter()
    tab = L.sdust_table(reads)
    dt = time.perf_counter() - t1
    prof = profile_json(L)
    lib.lqcov_profile_enable(0)
    kms = sum(k["ms"] for k in prof["kernels"] if k["name"] == "sdust")
    peak, _ = load_peaks()
    out = {"value": reads.n_bases / dt / 1e9, "unit": "Gbases/s", "seconds": dt, "first_call_seconds": t1 - t0, "reads": reads.n, "rows": tab.count(b"\n"),
           "kernel_ms": kms, "kernel_gbs": (2 * reads.n_bases / (kms * 1e-3) / 1e9) if kms > 0 else None, "kernel_frac_of_hbm_peak": (2 * reads.n_bases / (kms * 1e-3) / 1e9 / peak) if kms > 0 else None,
           "what": "lqcov_sdust_table on all target reads (pageable host buffers in, table out; chunks of 48 MB: copy of chunk c+1 beside the kernel of chunk c); "
                   "kernel_*: the sdust kernels alone (segment scan, redo of segments with long interval lists, per-read fold, quality sums; CUDA events), 2 bytes per base (base + quality)"}
    ref = os.path.join(ROOT, "oracle", "_ref", "sdust")
    if os.path.exists(ref):   # the reference's sdust on a bounded sample (one thread, as lq_mask.py runs it per chunk)
        n = min(reads.n, 4000)
        sub = reads.subset(range(n))
        d = tempfile.mkdtemp(prefix="lqsd_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        fq = os.path.join(d, "s.fq")
        sub.write_fastx(fq)
        t2 = time.perf_counter()
        want = subprocess.run([ref, fq], capture_output=True, check=True).stdout
        dtc = time.perf_counter() - t2
        os.unlink(fq); os.rmdir(d)
        out["cpu_reference"] = {"value": sub.n_bases / dtc / 1e9, "unit": "Gbases/s", "cores": 1, "sample": "%d reads" % n}
        out["parity"] = "identical on the sample" if L.sdust_table(sub) == want else "DIFFERS"
    return out


def run_ours(a):
    import torch
    import torch.distributed as dist
    import longqc_b200 as L
    from longqc_b200 import _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == a.gpus, "launch with torchrun --nproc-per-node %d for --gpus %d" % (a.gpus, a.gpus)
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from longqc_b200 import dist as lqdist

    opt = L.Opt(min_score_med=80 if a.preset == "pb-sequel" else 160, min_score_good=160, device=local)
    runner = lqdist.Runner(a, opt, rank, world, local)
    targets, queries = runner.make_inputs()        # this rank's shard of the workload (weak scaling)
    n_bases_job = runner.job_bases()                # all ranks together

    lib = L.load()
    lib.lqcov_profile_enable(0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ----
    for _ in range(a.warmup):
        runner.step(resident=True)
    lib.lqcov_profile_enable(1)
    lib.lqcov_profile_reset()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            runner.step(resident=True)
        barrier()
        ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    prof = profile_json(L)
    lib.lqcov_profile_enable(0)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / a.steps
    value = n_bases_job / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: host buffers, copies inside the timed region ----
    runner.step(resident=False)
    lib.lqcov_profile_reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        runner.step(resident=False)
    barrier()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / a.steps
    prof_e2e = profile_json(L)
    e2e = {"value": n_bases_job / (e2e_ms * 1e-3) / 1e9, "unit": "Gbases/s",
           "h2d_bytes_per_step": int(prof_e2e["h2d_bytes"] // a.steps * world), "d2h_bytes_per_step": int(prof_e2e["d2h_bytes"] // a.steps * world),
           "ms_per_step": e2e_ms}

    # ---- parity spot-check of the benchmarked configuration (not timed) ----
    parity = runner.parity_note()
    if world > 1:   # the merged N-GPU table against ONE GPU running the same global workload alone
        same = runner.verify_against_one_gpu()
        if rank == 0:
            parity["vs_1gpu_same_global_workload"] = same

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (device time from CUDA events on the library's stream) ----
    peak, peak_src = load_peaks()
    kernels = sorted(prof["kernels"], key=lambda k: -k["ms"])
    ktot = sum(k["ms"] for k in kernels) or 1.0
    klist = []
    for k in kernels:
        ach = (k["bytes"] / (k["ms"] * 1e-3) / 1e9) if k["ms"] > 0 else 0.0
        klist.append({"name": k["name"], "ms_per_step": k["ms"] / a.steps, "launches_per_step": k["launches"] / a.steps,
                      "share": k["ms"] / ktot, "achieved_gbs": ach, "frac": ach / peak})
    top = kernels[0] if kernels else None
    roof = None
    if top:
        ach = top["bytes"] / (top["ms"] * 1e-3) / 1e9 if top["ms"] > 0 else 0.0
        traffic, traffic_src = load_traffic(top["name"]) if world == 1 else (None, None)   # the committed capture is of the 1-GPU job
        roof = {"kernel": top["name"], "bound": "hbm", "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes": top["bytes"] / max(1, a.steps), "ms_per_launch_group": top["ms"] / max(1, a.steps),
                "note": "algorithmic bytes / CUDA-event time on the launching stream; see DESIGN.md for the bytes per unit"}
    sk = [k for k in klist if k["name"] == "sketch"]
    cpu, cli, sd = None, None, None
    if world == 1 and not (a.no_cpu_baseline and a.no_cli):
        # the workload as the FASTQ files LongQC would hand the binaries (in /dev/shm): the reference binary and OUR drop-in executable
        # run on the same files, wall clock of the whole process each (exec, CUDA context, parsing, compute, table on stdout)
        files = FastqFiles(targets, queries)
        if not a.no_cpu_baseline:
            if a.cpu_sample:
                ct, cq = make_data(a, a.cpu_sample_reads, a.cpu_sample_queries)
                what = "%d target reads x %d b + %d queries of the same generator (bounded sample)" % (ct.n, a.read_len, cq.n)
                v, dt, kind, cores, ref_table = cpu_reference_run(a, ct, cq, host_threads())
            else:   # the whole benchmarked workload: ~20-30 s of CPU, and its table is the full-size parity check of ours
                what = "the whole workload (%d target reads, %d queries): same inputs as the GPU arm" % (targets.n, queries.n)
                v, dt, kind, cores, ref_table = cpu_reference_run(a, targets, queries, host_threads(), files)
                parity["vs_cpu_%s_full_size" % kind] = "identical (%d rows, byte for byte)" % ref_table.count(b"\n") if ref_table == runner.last_table else "DIFFERS"
            cpu = {"value": v, "unit": "Gbases/s", "cores": cores, "kind": kind, "seconds": dt, "sample": what}
        if not a.no_cli:
            exe = [L.bin_path("minimap2-coverage")] + FLAGS.split() + ["-t", str(host_threads())]
            files.run(exe)                                   # warm-up: page cache, CUDA driver's module cache
            secs, tab = [], b""
            for _ in range(3):
                dt, tab = files.run(exe)
                secs.append(dt)
            best = min(secs)
            cli = {"value": files.n_bases / (sum(secs) / len(secs)) / 1e9, "unit": "Gbases/s", "seconds": sum(secs) / len(secs), "seconds_best": best,
                   "runs": len(secs), "fastq_bytes": files.bytes, "fastq_gb_per_s": files.bytes / best / 1e9,
                   "what": "wall clock of longqc_b200/bin/minimap2-coverage %s -t %d <targets.fq> <queries.fq> > table, files in /dev/shm: process start, "
                           "CUDA context, FASTQ parsing, H2D, compute, table" % (FLAGS, host_threads()),
                   "table": "identical to the resident-path table" if tab == runner.last_table else "DIFFERS"}
            if cpu and not a.cpu_sample:
                cli["speedup_vs_cpu_reference_same_files"] = cpu["seconds"] / cli["seconds"]
        files.close()
    if world == 1 and not a.no_sdust:
        sd = sdust_line(a, L, targets)
    line = {"metric": METRIC,
            "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(a, world), "target_bases": int(runner.target_bases_all), "query_bases": int(runner.query_bases_all),
                       "parallelism": runner.parallelism(), "l2": "inputs (>= %.1f GB per rank) larger than the 126 MB L2" % (targets.n_bases / 1e9),
                       "host_wall_ms_per_step": 1e3 * wall / a.steps},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(prof["launches"] // a.steps),
            "roofline": roof, "sketch_kernel": sk[0] if sk else None, "kernels": klist, "cpu_baseline": cpu, "cli_e2e": cli, "sdust": sd, "parity": parity,
            "stats": runner.last_stats}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference_arm(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
