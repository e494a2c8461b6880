#!/usr/bin/env python
"""Every float32 bit pattern through the HOST instantiation of fabe13_sincosf (the code the device kernel runs,
bit for bit -- tests/test_gpu_parity.py), against glibc's double sin/cos rounded once to float:
max error in float ulps and the number of results that are not the correctly rounded float.
    python tools/exhaustive_f32.py [first_chunk last_chunk] [--device] [--threads T]
256 chunks of 2^24 patterns; ~10 minutes on one thread.  --device also runs every chunk through fabe13_sincosf_device
and compares it with the host results bit for bit (needs a GPU)."""
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import fabe_b200 as fb  # noqa: E402


def one_chunk(chunk):
    bits = (np.arange(1 << 24, dtype=np.uint64) + (chunk << 24)).astype(np.uint32)
    x = bits.view(np.float32)
    s, c = fb.host_sincosf(x)
    fin = np.isfinite(x)
    # NaN / Inf -> canonical quiet NaN in both outputs
    special_bad = int(np.count_nonzero(s[~fin].view(np.uint32) != 0x7FC00000) + np.count_nonzero(c[~fin].view(np.uint32) != 0x7FC00000))
    xd = x[fin].astype(np.float64)
    out = {"special_bad": special_bad, "finite": int(fin.sum()), "x": x, "s": s, "c": c}
    with np.errstate(all="ignore"):
        for name, got, true in (("sin", s[fin], np.sin(xd)), ("cos", c[fin], np.cos(xd))):
            want = true.astype(np.float32)
            # ulp of the true value's binade, floored at the smallest subnormal
            ulp = np.maximum(np.spacing(np.abs(want)).astype(np.float64), 2.0 ** -149)
            err = np.abs(got.astype(np.float64) - true) / ulp
            i = int(np.argmax(err)) if err.size else 0
            out[name] = (float(err[i]) if err.size else 0.0, float(xd[i]) if err.size else None,
                         int(np.count_nonzero(got.view(np.uint32) != want.view(np.uint32))))
    return out


def main():
    from concurrent.futures import ThreadPoolExecutor

    argv = sys.argv[1:]
    device = "--device" in argv
    threads = 1
    if "--threads" in argv:
        i = argv.index("--threads")
        threads = int(argv[i + 1])
        del argv[i:i + 2]
    args = [a for a in argv if not a.startswith("--")]
    lo = int(args[0]) if len(args) > 0 else 0
    hi = int(args[1]) if len(args) > 1 else 256
    worst = {"sin": 0.0, "cos": 0.0}
    where = {"sin": None, "cos": None}
    off = {"sin": 0, "cos": 0}
    special_bad = 0
    device_bad = 0
    total = 0
    t0 = time.time()
    if device:
        import torch
    pool = ThreadPoolExecutor(max_workers=threads)
    for chunk, r in zip(range(lo, hi), pool.map(one_chunk, range(lo, hi))):
        special_bad += r["special_bad"]
        total += r["finite"]
        for name in ("sin", "cos"):
            e, xw, n_off = r[name]
            if e > worst[name]:
                worst[name], where[name] = e, xw
            off[name] += n_off
        if device:
            xt = torch.from_numpy(r["x"]).cuda()
            ds, dc = fb.sincosf(xt)
            torch.cuda.synchronize()
            device_bad += int(np.count_nonzero(ds.cpu().numpy().view(np.uint32) != r["s"].view(np.uint32)))
            device_bad += int(np.count_nonzero(dc.cpu().numpy().view(np.uint32) != r["c"].view(np.uint32)))
        if chunk % 16 == 15 or chunk == hi - 1:
            print(f"chunks {lo}..{chunk}: {total} finite inputs, {time.time() - t0:.0f} s; max error sin {worst['sin']:.6f} ulp at x={where['sin']!r}, "
                  f"cos {worst['cos']:.6f} ulp at x={where['cos']!r}; not correctly rounded: sin {off['sin']} cos {off['cos']}; "
                  f"bad special results {special_bad}" + (f"; device results differing from the host's: {device_bad}" if device else ""), flush=True)


if __name__ == "__main__":
    main()
