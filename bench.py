#!/usr/bin/env python
"""bench.py -- FP64 sincos throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W                 # N ranks, one contiguous slice of the array each
    python bench.py --impl reference ...                          # the reference's own CPU library on the host cores

Workload (default c3 = BASELINE configs[2]): N = 1e9 doubles uniform in [-1e4, 1e4], synthetic (counter-based
generator, same bits on host and device), slice-sharded contiguously over the ranks: total work is fixed, so
"scaling" is "strong".  A step = one pass of the hot path (fabe13_sincos_device on the rank's slice).
  value      whole-job Gelem/s with inputs and outputs resident in HBM (CUDA events, max over ranks)
  e2e        the same metric through the host-pointer C-ABI entry (fabe13_sincos_host) with pinned HOST buffers:
             H2D of the inputs and D2H of both outputs inside the timed region
  roofline   24 algorithmic bytes per element / the measured step time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference library (oracle/_ref, its build.sh flags) on the host cores, bounded sample
"""
import argparse
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


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner, for one), so move
    fd 1 to stderr for the whole run and keep the original for emit()."""
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


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_elem():
    """dram bytes per element of the streaming kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            j = json.load(f)
        return float(j["dram_bytes_per_elem"])
    except Exception:
        return None


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


def sample_per_thread(threads):
    """Bounded CPU sample: 20M elements per thread (480 MB of arrays, far beyond the CPU caches), capped so that
    the whole sample stays under 640M elements (15 GB) on hosts with very many cores."""
    return max(2_000_000, min(20_000_000, 640_000_000 // max(threads, 1)))


def make_cpu_runner(workload, per_thread_elems, threads):
    """Returns (run_once, total_elems, description).  run_once() evaluates the reference's fabe13_sincos on
    `threads` disjoint slices concurrently (the library itself is single-threaded: one call per slice) and returns
    the wall time."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor

    from oracle.loader import Oracle, Reference, cpu_model, ref_variant

    desc, _, gen, args, seed = WORKLOADS[workload]
    path = ref_variant("buildsh")
    if path is not None:
        ref = Reference(path=path)
        kind = "reference"
        call = lambda x, s, c: ref.sincos_raw(x, s, c)
        name = f"{os.path.basename(path)} [{ref.name}]"
    else:  # the compiled reference did not travel: fall back to the restated port, and say so
        orc = Oracle()
        kind = "port"
        call = lambda x, s, c: orc.lib.fabe13_oracle_sincos(x.ctypes.data, s.ctypes.data, c.ctypes.data, None, x.size)
        name = "oracle/fabe13_oracle.c (scalar port)"
    o = Oracle()
    n = per_thread_elems
    xs, ss, cs = [], [], []
    for t in range(threads):
        if gen == "uniform":
            xs.append(o.fill_uniform(n, seed, args[0], args[1], first_index=t * n))
        else:
            xs.append(o.fill_loguniform(n, seed, first_index=t * n))
        ss.append(np.zeros(n))  # touched: no first-touch page faults inside the timed region
        cs.append(np.zeros(n))
    pool = ThreadPoolExecutor(max_workers=threads)

    def run_once():
        t0 = time.perf_counter()
        list(pool.map(lambda t: call(xs[t], ss[t], cs[t]), range(threads)))  # ctypes releases the GIL
        return time.perf_counter() - t0

    info = {"kind": kind, "cores": threads, "library": name, "cpu": cpu_model(),
            "sample": f"{threads} threads x {n} elements of workload {workload} (warm arrays, one fabe13_sincos call per thread)"}
    return run_once, n * threads, info


def cpu_baseline(workload, seconds_budget=20.0):
    """Bounded CPU sample for the main JSON line: all host threads, and 1 thread (the reference as shipped)."""
    threads = host_threads()
    per = sample_per_thread(threads)
    run_all, total, info = make_cpu_runner(workload, per, threads)
    run_all()  # warm-up
    reps, best = 0, float("inf")
    t_start = time.perf_counter()
    # about 10-30 s of CPU work in total: at least 5 repetitions, then until ~1 s of wall time on all threads
    while reps < 5 or (reps < 40 and time.perf_counter() - t_start < 1.0):
        best = min(best, run_all())
        reps += 1
        if time.perf_counter() - t_start > seconds_budget:
            break
    out_reps = reps
    out = dict(info)
    out["value"] = total / best / 1e9
    out["unit"] = UNIT
    out["sample"] += f", best of {out_reps} repetitions"
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
    per = sample_per_thread(threads)
    run_once, total, info = make_cpu_runner(args.workload, per, threads)
    for _ in range(max(args.warmup, 1)):
        run_once()
    times = [run_once() for _ in range(args.steps)]
    tot = sum(times)
    value = total * args.steps / tot / 1e9
    info = dict(info)
    info["value"] = value
    info["unit"] = UNIT
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][0], "sample_elems_per_step": total, "host_threads": threads},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def run_e2e(fb, torch, args, x, s, n_local, world, barrier, max_over_ranks):
    """End to end through the host-pointer entry: pinned host buffers, H2D + D2H inside the timed region."""
    n_e2e = min(n_local, args.e2e_elems)
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
        if bufs:
            for p in bufs:
                p.free()
        bufs = None
        n_e2e //= 4
        log(f"bench.py: pinned allocation failed on some rank, retrying the end-to-end leg with {n_e2e} elements")
    if bufs is None:
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "error": "could not allocate pinned host buffers for the end-to-end leg"}
    hx, hs, hc = bufs
    hx.array[:] = x[:n_e2e].cpu().numpy()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    fb.sincos(hx.array, out_sin=hs.array, out_cos=hc.array)  # warm-up: allocates the pipeline slots
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fb.sincos(hx.array, out_sin=hs.array, out_cos=hc.array)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = n_e2e * world * e2e_steps / t_e2e / 1e9
    if float(abs(hs.array[:1000] - s[:1000].cpu().numpy()).max()) != 0.0:
        raise SystemExit("bench.py: host-path results differ from the device-path results")
    for p in (hx, hs, hc):
        p.free()
    return {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n_e2e, "d2h_bytes_per_step": 16 * n_e2e,
            "elements_per_gpu": n_e2e, "steps": e2e_steps,
            "path": "fabe13_sincos_host, pinned host buffers, " + ("chunked H2D|kernel|D2H copy-engine pipeline"
                                                                   if n_e2e > fb.get_option("zero_copy_max") else
                                                                   "zero-copy kernels over PCIe")}


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
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CUDA path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)  # harness barrier / max-over-ranks only: no data-path collective

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

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
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = fb.get_option("launches")
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    barrier()
    launches = fb.get_option("launches") - launches0
    total_ms = evs[0].elapsed_time(evs[-1])
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if sampler else None
    total_ms = max_over_ranks(total_ms)
    value = n_total * args.steps / (total_ms * 1e-3) / 1e9

    # sanity inside the bench: the outputs are a sine and a cosine (catches a kernel that did no work)
    m = min(n_local, 1_000_000)
    ident = (s[:m] * s[:m] + c[:m] * c[:m] - 1.0).abs().max().item()
    if not ident <= 1e-15:
        raise SystemExit(f"bench.py: sin^2+cos^2-1 = {ident} on the timed outputs")

    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(fb, torch, args, x, s, n_local, world, barrier, max_over_ranks)

    if rank == 0:
        peak, peak_src = measured_peak()
        avg_ms = sum(step_ms) / len(step_ms)
        achieved = BYTES_PER_ELEM * n_local / (avg_ms * 1e-3) / 1e9
        tpe = ncu_traffic_per_elem()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "elements_total": n_total, "elements_per_gpu": n_local,
                       "sharding": f"contiguous slices, {world} rank(s), no collective",
                       "core": ["poly", "psi"][fb.get_option("core")], "unroll": fb.get_option("unroll"),
                       "blocks_per_sm": fb.get_option("blocks_per_sm"), "worklist": fb.get_option("worklist"),
                       "l2": "inputs+outputs per GPU (%.1f GB) far exceed the 126 MB L2; no flush needed" % (BYTES_PER_ELEM * n_local / 1e9),
                       "implementation": fb.implementation_name()},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (tpe * n_local) if tpe else None,
                         "peak_source": peak_src, "frac_of_nominal_8TBps": achieved / 8000.0,
                         "algorithmic_bytes_per_launch": BYTES_PER_ELEM * n_local,
                         "step_ms_avg": avg_ms, "step_ms_median": sorted(step_ms)[len(step_ms) // 2], "step_ms_min": min(step_ms),
                         "note": "duration = CUDA events around one step on the launching stream (the fused streaming kernel; "
                                 "from 2^26 elements per rank on also the counter reset and the second-pass launch, "
                                 "empty for this workload), rank 0"},
            "e2e": e2e,
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
    ap.add_argument("--e2e-elems", type=int, default=250_000_000, help="per-GPU element cap of the end-to-end leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs only)")
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
