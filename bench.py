#!/usr/bin/env python
"""bench.py -- FP64 sincos throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W                 # N ranks, one contiguous slice of the array each
    python bench.py --impl reference ...                          # the reference's own CPU library on the host cores

Workload (default c3 = BASELINE configs[2]): N = 1e9 doubles uniform in [-1e4, 1e4], synthetic (counter-based
generator, same bits on host and device), slice-sharded contiguously over the ranks: total work is fixed, so
"scaling" is "strong".  A step = one pass of the hot path (fabe13_sincos_device on the rank's slice).
  value         whole-job Gelem/s with inputs and outputs resident in HBM (CUDA events, max over ranks)
  roofline      24 algorithmic bytes per element / the measured step time, against MEASURED_PEAKS.json hbm_gbs;
                .sustained = the same over >= 2 s of back-to-back steps after the timed region (power-capped clocks)
  e2e           the same metric through the host-pointer C-ABI entry (fabe13_sincos_host) with pinned HOST buffers, the
                rank's whole slice: H2D of the inputs and D2H of both outputs inside the timed region;
                .roofline = a bare cudaMemcpyAsync H2D || D2H probe in the path's 8:16 byte mix, run on the same
                buffers by all ranks at once right before it -- what the host link gives at this GPU count
  e2e_pageable  what an unmodified caller of the reference gets: malloc'd (numpy) arrays through void fabe13_sincos()
  cpu_baseline  the reference library (oracle/_ref, built by its own build.sh) on the host cores, bounded sample
The multi-rank harness (barrier, max over ranks) runs on gloo: the data path has no collective and no NCCL.
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, total elements, generator, args)
    "c1": ("sincos N=1e6 uniform [-pi,pi]", 1_000_000, "uniform", (-3.141592653589793, 3.141592653589793), 0xFABE1301),
    "c2": ("sincos N=1e8 uniform [-1e4,1e4]", 100_000_000, "uniform", (-1e4, 1e4), 0xFABE1302),
    "c3": ("sincos N=1e9 uniform [-1e4,1e4] slice-sharded", 1_000_000_000, "uniform", (-1e4, 1e4), 0xFABE1303),
    "c3b": ("sincos N=1e9 uniform [-1e5*pi,1e5*pi] (reference benchmark range)", 1_000_000_000, "uniform",
            (-314159.2653589793, 314159.2653589793), 0xFABE1303),
    "c4": ("sincos N=1e8 |x| log-uniform up to 2^1024 (Payne-Hanek mix)", 100_000_000, "loguniform", None, 0xFABE1304),
}
BYTES_PER_ELEM = 24  # 8 B read + 8 B sin + 8 B cos
METRIC = "fp64_sincos_throughput"
UNIT = "Gelem/s"
KERNEL_SOURCES = ("fabe13_kernels.cu", "fabe13_core.h", "fabe13_constants.h", "fabe13_ph_remainders.h", "fabe13_internal.h")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too, so move fd 1 to stderr for the whole run
    and keep the original for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def workload_config(workload, n_total):
    """The part of `config` that names the workload: identical in both arms (the CUDA arm and --impl reference)."""
    desc, _, gen, gargs, seed = WORKLOADS[workload]
    return {"workload": desc, "elements_total": n_total, "generator": f"{gen}, counter-based splitmix64, seed {seed:#x}",
            "l2": "every step reads and writes the whole array (24 B/element: 24 GB at N=1e9; per GPU >= 3 GB at 8 GPUs), "
                  "far beyond the 126 MB L2 and the host's last-level cache; no flush needed"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash():
    """sha256 over the sources the kernels are compiled from: what an ncu capture has to have been taken on."""
    h = hashlib.sha256()
    for name in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "fabe_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def ncu_traffic_per_elem():
    """(dram bytes per element of the streaming kernel, note) from the committed ncu capture -- only if that capture
    was taken on the kernel sources this run is built from."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            j = json.load(f)
    except Exception:
        return None, "no ncu capture committed (profiles/ncu_traffic.json)"
    now = kernel_source_hash()
    if j.get("kernel_source_sha256") != now:
        return None, (f"the committed ncu capture ({j.get('report')}) was taken on kernel sources "
                      f"{str(j.get('kernel_source_sha256'))[:12]}, this build is {now[:12]}: not carried over")
    return float(j["dram_bytes_per_elem"]), f"ncu --set full capture {j.get('report')} on kernel sources {now[:12]}"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.mem_samples = []
        self.power_w = []
        self.reason_mask = 0
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.mem_samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM))
                self.power_w.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.reason_mask |= int(r)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}
        s = sorted(self.samples)
        m = sorted(self.mem_samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "mem_mhz": m[len(m) // 2] if m else None,
                "samples": len(s), "sm_mhz_min": s[0] if s else None, "power_w_max": max(self.power_w) if self.power_w else None,
                "reason_mask": hex(self.reason_mask)}


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own library on the host cores
# --------------------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_cpu_runner(workload, total_elems, threads):
    """Returns (run_once, total_elems, info).  run_once() evaluates the reference's fabe13_sincos on `threads`
    disjoint contiguous slices of one `total_elems`-element array concurrently (the library itself is single-threaded:
    one call per slice) and returns the wall time."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor

    from oracle.loader import Oracle, Reference, cpu_model, ref_variant

    desc, _, gen, args, seed = WORKLOADS[workload]
    path = ref_variant("buildsh")
    if path is not None:
        ref = Reference(path=path)
        kind = "reference"
        call = lambda x, s, c: ref.sincos_raw(x, s, c)
        how = ("built by the reference's own build.sh in a scratch copy" if path.endswith("_own.so") else
               "the reference source compiled with build.sh's flags (oracle/Makefile)")
        name = f"{os.path.basename(path)} [{ref.name}], {how}"
    else:  # the compiled reference did not travel: fall back to the restated port, and say so
        orc = Oracle()
        kind = "port"
        call = lambda x, s, c: orc.lib.fabe13_oracle_sincos(x.ctypes.data, s.ctypes.data, c.ctypes.data, None, x.size)
        name = "oracle/fabe13_oracle.c (scalar port)"
    o = Oracle()
    per = (total_elems // threads) & ~7  # slices of whole SIMD vectors (the reference's tails take its scalar core)
    bounds = [(t * per, (t + 1) * per if t < threads - 1 else total_elems) for t in range(threads)]
    x = np.empty(total_elems)
    s = np.empty(total_elems)
    c = np.empty(total_elems)
    pool = ThreadPoolExecutor(max_workers=threads)

    def fill(t):
        b, e = bounds[t]
        if gen == "uniform":
            o.lib.fabe13_oracle_fill_uniform(x[b:e].ctypes.data, e - b, seed, b, args[0], args[1])
        else:
            o.lib.fabe13_oracle_fill_loguniform(x[b:e].ctypes.data, e - b, seed, b)
        s[b:e] = 0.0  # touched by the thread that will write them: no first-touch page faults inside the timed region
        c[b:e] = 0.0

    list(pool.map(fill, range(threads)))

    def run_once():
        t0 = time.perf_counter()
        list(pool.map(lambda t: call(x[bounds[t][0]:bounds[t][1]], s[bounds[t][0]:bounds[t][1]], c[bounds[t][0]:bounds[t][1]]),
                      range(threads)))  # ctypes releases the GIL
        return time.perf_counter() - t0

    info = {"kind": kind, "cores": threads, "library": name, "cpu": cpu_model(),
            "sample": f"{total_elems} elements of workload {workload} in {threads} contiguous slices, one fabe13_sincos call per "
                      f"host thread (the library is single-threaded), warm arrays"}
    return run_once, total_elems, info


def cpu_baseline(workload, seconds_budget=20.0):
    """Bounded CPU sample for the main JSON line: all host threads, and 1 thread (the reference as shipped)."""
    threads = host_threads()
    total = max(2_000_000, min(20_000_000, 640_000_000 // max(threads, 1))) * threads
    run_all, total, info = make_cpu_runner(workload, total, threads)
    run_all()  # warm-up
    reps, best = 0, float("inf")
    t_start = time.perf_counter()
    # about 10-30 s of CPU work in total: at least 5 repetitions, then until ~1 s of wall time on all threads
    while reps < 5 or (reps < 40 and time.perf_counter() - t_start < 1.0):
        best = min(best, run_all())
        reps += 1
        if time.perf_counter() - t_start > seconds_budget:
            break
    out = dict(info)
    out["value"] = total / best / 1e9
    out["unit"] = UNIT
    out["sample"] += f", best of {reps} repetitions"
    run_one, total1, _ = make_cpu_runner(workload, 20_000_000, 1)
    run_one()
    best1 = min(run_one() for _ in range(3))
    out["value_1core"] = total1 / best1 / 1e9
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    n_total = args.elems or WORKLOADS[args.workload][1]
    try:
        run_once, total, info = make_cpu_runner(args.workload, n_total, threads)
    except MemoryError:  # a host too small for 24 B x N: say so and time a quarter of the array
        log("bench.py: not enough host memory for the whole workload, timing a quarter of it")
        run_once, total, info = make_cpu_runner(args.workload, n_total // 4, threads)
    for _ in range(max(args.warmup, 1)):
        run_once()
    times = [run_once() for _ in range(args.steps)]
    tot = sum(times)
    value = total * args.steps / tot / 1e9
    info = dict(info)
    info["value"] = value
    info["unit"] = UNIT
    config = workload_config(args.workload, n_total)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "setup": {"elements_per_step": total, "host_threads": threads, "step_ms_min": min(times) * 1e3},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------------------------
# end to end: host buffers through the C ABI
# --------------------------------------------------------------------------------------------------------------
def link_probe(torch, dev, hx, hs, hc, x, s, c, n, seconds, barrier, max_over_ranks, sum_over_ranks, solo):
    """Bare copies in the path's mix -- 8 B in : 16 B out per element, chunked as the pipeline chunks them, no kernel --
    between the very buffers the end-to-end leg uses.  All ranks run inside one window (`solo`: rank 0 alone), so the
    result is what the host side gives this many GPUs at once.  Returns aggregate (H2D GB/s, D2H GB/s)."""
    chunk = 4 << 20  # elements: the pipeline's chunk (32 MB in, 2 x 32 MB out)
    rank = int(os.environ.get("RANK", "0"))
    active = (not solo) or rank == 0
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    thx, ths, thc = (torch.from_numpy(a) for a in (hx, hs, hc))
    nchunks = max(1, n // chunk)
    barrier()
    t0 = time.perf_counter()
    moved_in = moved_out = 0
    i = 0
    if active:
        while time.perf_counter() - t0 < seconds:
            b = (i % nchunks) * chunk
            e = min(b + chunk, n)
            with torch.cuda.stream(s_in):
                x[b:e].copy_(thx[b:e], non_blocking=True)
            with torch.cuda.stream(s_out):
                ths[b:e].copy_(s[b:e], non_blocking=True)
                thc[b:e].copy_(c[b:e], non_blocking=True)
            moved_in += 8 * (e - b)
            moved_out += 16 * (e - b)
            i += 1
            if i % 8 == 0:  # keep a bounded number of copies queued
                s_in.synchronize()
                s_out.synchronize()
        s_in.synchronize()
        s_out.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    tot_in, tot_out = sum_over_ranks(float(moved_in)), sum_over_ranks(float(moved_out))
    barrier()
    return tot_in / dt / 1e9, tot_out / dt / 1e9


def run_e2e(fb, torch, args, dev, x, s, c, n_local, world, barrier, max_over_ranks, sum_over_ranks):
    """End to end through the host-pointer entry: pinned host buffers, H2D + D2H inside the timed region; then the
    same with pageable arrays through the reference's own void entry."""
    import numpy as np

    n_e2e = min(n_local, args.e2e_elems) if args.e2e_elems else n_local
    # every rank must end up with the same sample size (the timed loop contains a collective): agree on success
    bufs = None
    for _ in range(4):
        try:
            bufs = (fb.PinnedArray(n_e2e), fb.PinnedArray(n_e2e), fb.PinnedArray(n_e2e))
            ok = 1.0
        except MemoryError:
            bufs, ok = None, 0.0
        if max_over_ranks(-ok) == -1.0:  # min over ranks of ok
            break
        bufs = None
        n_e2e //= 4
        log(f"bench.py: pinned allocation failed on some rank, retrying the end-to-end leg with {n_e2e} elements")
    if bufs is None:
        err = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "error": "could not allocate pinned host buffers for the end-to-end leg"}
        return err, None
    hx, hs, hc = (p.array for p in bufs)
    torch.from_numpy(hx).copy_(x[:n_e2e])
    first = s[:1000].cpu().numpy()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    fb.sincos(hx, out_sin=hs, out_cos=hc)  # warm-up: allocates the pipeline slots
    # the ceiling first: what the link (one GPU) and the host side (all GPUs at once) carry for this mix
    solo = link_probe(torch, dev, hx, hs, hc, x, s, c, n_e2e, 0.5, barrier, max_over_ranks, sum_over_ranks, True) if world > 1 else None
    both = link_probe(torch, dev, hx, hs, hc, x, s, c, n_e2e, 1.0, barrier, max_over_ranks, sum_over_ranks, False)
    hs[:1000] = 0.0
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fb.sincos(hx, out_sin=hs, out_cos=hc)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = n_e2e * world * e2e_steps / t_e2e / 1e9
    if float(abs(hs[:1000] - first).max()) != 0.0:
        raise SystemExit("bench.py: host-path results differ from the device-path results")
    peak_gelems = min(both[0] / 8, both[1] / 16)
    per_gpu = (both[0] + both[1]) / world
    if solo is None or per_gpu >= 0.9 * (solo[0] + solo[1]):
        bound = "pcie"
        why = "every GPU's copies run at the rate one GPU's link carries alone"
    else:
        bound = "host_dram"
        why = (f"{world} GPUs copying at once get {per_gpu:.1f} GB/s each where one alone gets {solo[0] + solo[1]:.1f}: the host side "
               f"(memory / IOMMU of this VM) is shared")
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n_e2e, "d2h_bytes_per_step": 16 * n_e2e,
           "elements_per_gpu": n_e2e, "steps": e2e_steps,
           "path": "fabe13_sincos_host, pinned host buffers, " + ("chunked H2D|kernel|D2H copy-engine pipeline"
                                                                  if n_e2e > fb.get_option("zero_copy_max") else
                                                                  "zero-copy kernels over PCIe"),
           "roofline": {"bound": bound, "unit": "GB/s", "peak_gbs": both[0] + both[1], "peak_h2d_gbs": both[0], "peak_d2h_gbs": both[1],
                        "achieved_gbs": BYTES_PER_ELEM * e2e_value, "peak": peak_gelems, "frac": e2e_value / peak_gelems,
                        "one_gpu_alone_gbs": (solo[0] + solo[1]) if solo else None,
                        "how": "bare cudaMemcpyAsync H2D || D2H (8 B : 16 B per element, 32 MB copies, no kernel) on the same pinned "
                               "buffers, all ranks inside one window, measured in this run right before the timed calls; " + why}}
    if n_e2e < n_local:
        e2e["cap"] = f"{n_e2e} of the rank's {n_local} elements (--e2e-elems, or pinned memory did not allow more)"
    # beside it, for the record and NOT the headline: the same call with option host_assist -- host threads run the
    # library's host core (same bits) on pieces from the array's tail while the copy-engine pipeline works from its
    # front.  A host-resident array is bound by PCIe and by the host's memory, so this is what the machine as a whole
    # can do for a caller who wants it; how the array was split is reported with it.
    helpers = max(0, host_threads() // world - 1)
    if helpers > 0 and n_e2e >= fb.get_option("host_assist_min_n") and not args.no_host_assist:
        last = s[n_e2e - 1000:n_e2e].cpu().numpy()
        # (a side measurement must not cost the run its headline: a failure is recorded, not raised -- and every rank goes
        # through the same collectives whatever happens on it)
        problem = None
        g0 = h0 = g1 = h1 = 0
        t_as = float("inf")
        try:
            fb.set_option("host_assist", helpers)
            fb.sincos(hx, out_sin=hs, out_cos=hc)  # warm-up
            hs[:1000] = 0.0
            hs[n_e2e - 1000:] = 0.0
            g0, h0 = fb.get_option("assist_gpu_elems"), fb.get_option("assist_host_elems")
        except Exception as ex:  # noqa: BLE001
            problem = repr(ex)
        barrier()
        t0 = time.perf_counter()
        if problem is None:
            try:
                for _ in range(e2e_steps):
                    fb.sincos(hx, out_sin=hs, out_cos=hc)
                torch.cuda.synchronize()
            except Exception as ex:  # noqa: BLE001
                problem = repr(ex)
        t_as = max_over_ranks(time.perf_counter() - t0)
        barrier()
        fb.set_option("host_assist", 0)
        if problem is None:
            g1, h1 = fb.get_option("assist_gpu_elems"), fb.get_option("assist_host_elems")
            if float(abs(hs[:1000] - first).max()) != 0.0 or float(abs(hs[n_e2e - 1000:] - last).max()) != 0.0:
                problem = "host-assisted results differ from the device-path results"
        gpu_elems, host_elems = sum_over_ranks(float(g1 - g0)), sum_over_ranks(float(h1 - h0))
        failed = max_over_ranks(0.0 if problem is None else 1.0)
        if failed:
            e2e["with_host_assist"] = {"value": None, "unit": UNIT, "error": problem or "failed on another rank"}
            log("bench.py: the host-assisted leg failed:", problem)
        else:
            e2e["with_host_assist"] = {
                "value": n_e2e * world * e2e_steps / t_as / 1e9, "unit": UNIT, "host_threads_per_rank": helpers,
                "gpu_share": gpu_elems / max(gpu_elems + host_elems, 1.0),
                "gpu_gelems": gpu_elems / t_as / 1e9, "host_gelems": host_elems / t_as / 1e9,
                "note": "NOT the headline and off by default (option host_assist): the copy-engine pipeline takes chunks from the front "
                        "of the pinned array, host threads run the library's host instantiation of the core (bit-identical) on pieces "
                        "from its back, both claiming from one cursor; gpu_share = fraction of the elements the GPU computed"}
    del hx, hs, hc
    for p in bufs:
        p.free()
    bufs = None

    # what an unmodified caller of the reference passes: pageable arrays, the void entry
    n_pg = min(n_local, args.pageable_elems) if args.pageable_elems else n_local
    try:
        px = np.empty(n_pg)
        torch.from_numpy(px).copy_(x[:n_pg])
        ps, pc = np.zeros(n_pg), np.zeros(n_pg)  # touched: no first-touch page faults inside the timed region
    except MemoryError:
        return e2e, {"value": None, "unit": UNIT, "error": "not enough host memory for pageable arrays"}
    fb.sincos_legacy(px, ps, pc)  # warm-up: allocates the staging buffers
    pg_steps = max(1, min(args.steps, args.e2e_steps))
    barrier()
    t0 = time.perf_counter()
    for _ in range(pg_steps):
        fb.sincos_legacy(px, ps, pc)
    t_pg = max_over_ranks(time.perf_counter() - t0)
    barrier()
    if float(abs(ps[:1000] - first).max()) != 0.0:
        raise SystemExit("bench.py: pageable host-path results differ from the device-path results")
    # beside it, for the record: the same call kept on the host (option min_n_pageable raised above n) -- the library's
    # own core on the host threads.  For pageable arrays the staging copies alone cost more CPU time than the
    # arithmetic, so this is what a caller without page-locked buffers may prefer; it is not the GPU path and not `e2e`.
    saved_cut = fb.get_option("min_n_pageable")
    host_core = None
    try:
        fb.set_option("min_n_pageable", 1 << 31)
        fb.sincos_legacy(px, ps, pc)
        barrier()
        t0 = time.perf_counter()
        for _ in range(pg_steps):
            fb.sincos_legacy(px, ps, pc)
        t_hc = max_over_ranks(time.perf_counter() - t0)
        barrier()
        if float(abs(ps[:1000] - first).max()) != 0.0:
            raise SystemExit("bench.py: host-core results differ from the device-path results")
        host_core = {"value": n_pg * world * pg_steps / t_hc / 1e9, "unit": UNIT,
                     "threads_per_rank": fb.get_option("host_threads") or min(16, os.cpu_count() or 1),
                     "note": "NOT the GPU path: the same call with option min_n_pageable raised, i.e. the library's host "
                             "instantiation of the core on the host threads (bit-identical results)"}
    finally:
        fb.set_option("min_n_pageable", saved_cut)
    pageable = {"value": n_pg * world * pg_steps / t_pg / 1e9, "unit": UNIT, "h2d_bytes_per_step": 8 * n_pg, "d2h_bytes_per_step": 16 * n_pg,
                "elements_per_gpu": n_pg, "steps": pg_steps,
                "path": "void fabe13_sincos(in, sin_out, cos_out, int n) on malloc'd (numpy) arrays: staged through pinned buffers by "
                        "worker threads (host-memory bound), as a relinked caller of the reference gets it",
                "stage_threads": fb.get_option("stage_threads") or "auto (3/4 of the hardware threads, at most 16)",
                "kept_on_the_host_instead": host_core}
    return e2e, pageable


# --------------------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------------------
def run_cuda_arm(args):
    import torch
    import torch.distributed as dist

    import fabe_b200 as fb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or fb.get_option("device_count") == 0:
        raise SystemExit("bench.py: no CUDA device -- this arm measures the CUDA path and nothing else")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # harness only (barrier, max over ranks) and on gloo: the data path has no collective, NCCL is not initialised
        dist.init_process_group("gloo")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return t.item()

    max_over_ranks = lambda v: reduce(v, dist.ReduceOp.MAX)
    min_over_ranks = lambda v: reduce(v, dist.ReduceOp.MIN)
    sum_over_ranks = lambda v: reduce(v, dist.ReduceOp.SUM)

    if args.core is not None:
        fb.set_option("core", args.core)
    desc, n_total, gen, gargs, seed = WORKLOADS[args.workload]
    if args.elems:
        n_total = args.elems
    b, e = fb.shard_bounds(n_total, world, rank)
    n_local = e - b
    x = torch.empty(n_local, dtype=torch.float64, device=dev)
    s = torch.empty_like(x)
    c = torch.empty_like(x)
    if gen == "uniform":
        fb.fill_uniform_device(x, seed, gargs[0], gargs[1], first_index=b)
    else:
        fb.fill_loguniform_device(x, seed, first_index=b)
    torch.cuda.synchronize()

    step = lambda: fb.sincos(x, out_sin=s, out_cos=c)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local).start() if rank == 0 else None
    launches0 = fb.get_option("launches")
    # the timed region: exactly K steps between two events on the launching stream, nothing else in the stream -- an event
    # recorded between two calls would keep the second call's kernel from being staged behind the first (every launch is
    # programmatic-dependent), which costs ~6 us per step: 1.5 % at the 8-GPU slice size, and not what a caller's
    # back-to-back calls look like
    e_first, e_last = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_first.record()
    for i in range(args.steps):
        step()
    e_last.record()
    barrier()
    launches = fb.get_option("launches") - launches0
    total_ms = e_first.elapsed_time(e_last)
    clocks = sampler.stop() if sampler else None
    # per-step spread (median, fastest step), from a second pass of K steps with an event behind every step; not part of `value`
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    barrier()
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    fastest_rank_ms = min_over_ranks(total_ms)  # the spread between the GPUs of the box: `value` is set by the slowest
    total_ms = max_over_ranks(total_ms)
    value = n_total * args.steps / (total_ms * 1e-3) / 1e9

    # sanity inside the bench: the outputs are a sine and a cosine (catches a kernel that did no work)
    m = min(n_local, 1_000_000)
    ident = (s[:m] * s[:m] + c[:m] * c[:m] - 1.0).abs().max().item()
    if not ident <= 1e-15:
        raise SystemExit(f"bench.py: sin^2+cos^2-1 = {ident} on the timed outputs")

    # sustained: the same step back to back for >= sustain_seconds on every rank at once (the 1 kW power cap pulls the
    # SM clock down after about a second; the burst figure above does not see that)
    sustained = None
    if args.sustain_seconds > 0:
        per_step = max(total_ms / args.steps * 1e-3, 1e-5)
        k = max(10, int(1.1 * args.sustain_seconds / per_step) + 1)  # (steps speed up once the caches and clocks settle)
        barrier()
        sus_sampler = ClockSampler(local).start() if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            step()
        e1.record()
        barrier()
        sus_ms = e0.elapsed_time(e1)
        sus_clocks = sus_sampler.stop() if sus_sampler else None
        sus_ms_max = max_over_ranks(sus_ms)
        sustained = (k, sus_ms, sus_ms_max, sus_clocks)

    e2e, pageable = None, None
    if not args.no_e2e:
        e2e, pageable = run_e2e(fb, torch, args, dev, x, s, c, n_local, world, barrier, max_over_ranks, sum_over_ranks)

    if rank == 0:
        peak, peak_src = measured_peak()
        avg_ms = e_first.elapsed_time(e_last) / args.steps  # rank 0's own timed region
        achieved = BYTES_PER_ELEM * n_local / (avg_ms * 1e-3) / 1e9
        tpe, traffic_note = ncu_traffic_per_elem()
        config = workload_config(args.workload, n_total)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (tpe * n_local) if tpe else None, "traffic_source": traffic_note,
                    "peak_source": peak_src, "frac_of_nominal_8TBps": achieved / 8000.0,
                    "algorithmic_bytes_per_launch": BYTES_PER_ELEM * n_local,
                    "step_ms_avg": avg_ms, "step_ms_median": sorted(step_ms)[len(step_ms) // 2], "step_ms_min": min(step_ms),
                    "note": "duration = rank 0's timed region (two CUDA events on the launching stream around the K steps) / K: one launch of "
                            "the fused streaming kernel per step; median / min from a second pass with an event behind every step"}
        if sustained:
            k, sus_ms, sus_ms_max, sus_clocks = sustained
            sus_achieved = BYTES_PER_ELEM * n_local * k / (sus_ms * 1e-3) / 1e9
            roofline["sustained"] = {"seconds": sus_ms * 1e-3, "steps": k, "value": n_total * k / (sus_ms_max * 1e-3) / 1e9, "unit": UNIT,
                                     "achieved": sus_achieved, "frac": sus_achieved / peak, "frac_of_nominal_8TBps": sus_achieved / 8000.0,
                                     "clocks": sus_clocks,
                                     "note": "back-to-back steps right after the timed region, all ranks at once; value = whole job, "
                                             "achieved/frac = rank 0's GPU"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config,
            "setup": {"elements_per_gpu": n_local, "sharding": f"contiguous slices, {world} rank(s), no collective on the data path "
                                                                "(harness barrier on gloo; NCCL not initialised)",
                      "core": ["poly", "psi"][fb.get_option("core")], "unroll": fb.get_option("unroll"),
                      "blocks_per_sm": fb.get_option("blocks_per_sm"), "worklist": fb.get_option("worklist"),
                      "hybrid_min_n": fb.get_option("hybrid_min_n"), "implementation": fb.implementation_name(),
                      "ms_per_step_slowest_rank": total_ms / args.steps, "ms_per_step_fastest_rank": fastest_rank_ms / args.steps},
            "roofline": roofline,
            "e2e": e2e,
            "e2e_pageable": pageable,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args.workload)
            except Exception as ex:  # never lose the GPU numbers to a CPU-side problem
                line["cpu_baseline"] = {"error": repr(ex)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--elems", type=int, default=0, help="override the total element count")
    ap.add_argument("--core", type=int, default=None, choices=[0, 1])
    ap.add_argument("--e2e-elems", type=int, default=0, help="per-GPU element cap of the end-to-end leg (0 = the rank's whole slice)")
    ap.add_argument("--pageable-elems", type=int, default=0, help="per-GPU element cap of the pageable end-to-end leg (0 = whole slice)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--sustain-seconds", type=float, default=2.0, help="length of the sustained run after the timed region (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end legs (profiling runs only)")
    ap.add_argument("--no-host-assist", action="store_true", help="skip the host-assisted variant of the end-to-end leg")
    args = ap.parse_args()
    claim_stdout()
    if args.warmup < 3 and args.impl == "cuda":
        log("bench.py: raising --warmup to 3 (timing rules)")
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_cuda_arm(args)


if __name__ == "__main__":
    sys.exit(main())
