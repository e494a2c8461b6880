#!/usr/bin/env python
"""Randomised soak: random sizes, alignments, option settings, input mixes and entry points; every result compared
bit for bit with the host instantiation of the core.  python tools/soak.py [iterations] [seed]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def make_input(rng, n):
    kind = rng.integers(0, 5)
    if kind == 0:
        x = rng.uniform(-1e4, 1e4, n)
    elif kind == 1:
        x = rng.integers(0, 1 << 64, size=n, dtype=np.uint64).view(np.float64).copy()
    elif kind == 2:
        x = rng.uniform(-1e6, 1e6, n)
        m = rng.random(n) < rng.choice([1e-4, 1e-2, 0.2, 0.9])
        x[m] = np.ldexp(rng.uniform(1, 2, m.sum()), rng.integers(45, 1024, m.sum())) * rng.choice([-1.0, 1.0], m.sum())
    elif kind == 3:
        x = np.ldexp(rng.uniform(-2, 2, n), rng.integers(-1074, 1024, n))
    else:
        x = rng.uniform(-10, 10, n)
        x[rng.random(n) < 0.01] = np.nan
        x[rng.random(n) < 0.01] = np.inf
        x[rng.random(n) < 0.01] = -0.0
    return x


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda:0")
    bad = 0
    for it in range(iters):
        n = int(rng.choice([rng.integers(1, 100), rng.integers(100, 100_000), rng.integers(100_000, 6_000_000), rng.integers(4_000_000, 12_000_000)]))
        opts = {"small_n": int(rng.choice([0, 1 << 22, 1 << 30])), "worklist_shift": int(rng.choice([0, 3, 20])),
                "unroll": int(rng.choice([1, 2])), "blocks_per_sm": int(rng.choice([0, 0, 2, 7])), "tiles_per_block": int(rng.choice([1, 3])),
                "pdl": int(rng.choice([0, 1])), "core": int(rng.choice([0, 0, 1])), "second_pass_blocks_per_sm": int(rng.choice([1, 8, 16])),
                "worklist": int(rng.choice([0, 0, 1, 2, 3])), "hybrid_min_n": int(rng.choice([1 << 26, 1 << 20])),
                "local_min": int(rng.choice([2, 2, 1, 4, 129])), "dense_hint": int(rng.choice([16, 16, 1, 24, 33])), "direct_max": int(rng.choice([4, 4, 0, 1, 16, 128])),
                "min_n": int(rng.choice([0, 0, 0, 8192])), "min_n_pageable": int(rng.choice([0, 0, 0, 131072]))}  # mostly: every host-pointer call on the GPU, whatever its size
        for k, v in opts.items():
            fb.set_option(k, v)
        x = make_input(rng, n)
        hs, hc, hk = fb.host_core(x, opts["core"])
        mode = int(rng.integers(0, 8))
        off_in, off_s, off_c = (int(v) for v in rng.integers(0, 4, 3))
        if rng.random() < 0.5:
            off_s = off_c = off_in
        ok = True
        if mode in (0, 1, 2):  # device entries
            bx = torch.zeros(n + 8, dtype=torch.float64, device=dev)
            bx[off_in:off_in + n] = torch.from_numpy(x).to(dev)
            xv = bx[off_in:off_in + n]
            if mode == 0:
                bs = torch.zeros(n + 8, dtype=torch.float64, device=dev)
                bc = torch.zeros(n + 8, dtype=torch.float64, device=dev)
                fb.sincos(xv, out_sin=bs[off_s:off_s + n], out_cos=bc[off_c:off_c + n])
                torch.cuda.synchronize()
                ok = np.array_equal(bits(bs[off_s:off_s + n].cpu().numpy()), bits(hs)) and np.array_equal(bits(bc[off_c:off_c + n].cpu().numpy()), bits(hc))
            elif mode == 1:
                s, c, k = fb.sincos_k(xv.contiguous() if off_in == 0 else xv)
                torch.cuda.synchronize()
                ok = np.array_equal(bits(s.cpu().numpy()), bits(hs)) and np.array_equal(bits(c.cpu().numpy()), bits(hc)) and np.array_equal(k.cpu().numpy(), hk)
            else:
                so = fb.sin_array(xv)
                co = fb.cos_array(xv)
                fb.sincos(xv, out_sin=xv, out_cos=bx.new_empty(n))  # in place
                torch.cuda.synchronize()
                ok = np.array_equal(bits(so.cpu().numpy()), bits(hs)) and np.array_equal(bits(co.cpu().numpy()), bits(hc)) and np.array_equal(bits(xv.cpu().numpy()), bits(hs))
        elif mode == 6:  # derived functions: device (misaligned, in place) and host
            op, fn = [(fb.OP_TAN, fb.tan_array), (fb.OP_COT, fb.cot_array), (fb.OP_SINC, fb.sinc_array)][int(rng.integers(0, 3))]
            want = fb.host_derived(x, op, opts["core"])
            bx = torch.zeros(n + 8, dtype=torch.float64, device=dev)
            bx[off_in:off_in + n] = torch.from_numpy(x).to(dev)
            xv = bx[off_in:off_in + n]
            o = fn(xv)
            fn(xv, out=xv)
            torch.cuda.synchronize()
            ok = np.array_equal(bits(o.cpu().numpy()), bits(want)) and np.array_equal(bits(xv.cpu().numpy()), bits(want))
            if not ok:
                bad_i = np.nonzero(bits(o.cpu().numpy()) != bits(want))[0]
                for i in bad_i[:5]:
                    print(f"  derived op {op}: i={i} x={x[i]!r} ({x[i].hex()}) got {o[i].item()!r} want {want[i]!r}", flush=True)
            if n < 3_000_000:
                ok = ok and np.array_equal(bits(fn(x)), bits(want))
        elif mode == 7:  # FP32 arrays: device (every mutual alignment, in place) and host
            xf = x.astype(np.float32)
            if rng.random() < 0.3:
                xf = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32).view(np.float32).copy()
            ws, wc = fb.host_sincosf(xf, opts["core"])
            bx = torch.zeros(n + 16, dtype=torch.float32, device=dev)
            bs = torch.zeros(n + 16, dtype=torch.float32, device=dev)
            bc = torch.zeros(n + 16, dtype=torch.float32, device=dev)
            o_in, o_s, o_c = (int(v) for v in rng.integers(0, 9, 3))
            if rng.random() < 0.6:
                o_s = o_c = o_in
            bx[o_in:o_in + n] = torch.from_numpy(xf).to(dev)
            fb.sincosf(bx[o_in:o_in + n], out_sin=bs[o_s:o_s + n], out_cos=bc[o_c:o_c + n])
            torch.cuda.synchronize()
            u32 = lambda a: np.ascontiguousarray(a).view(np.uint32)
            ok = np.array_equal(u32(bs[o_s:o_s + n].cpu().numpy()), u32(ws)) and np.array_equal(u32(bc[o_c:o_c + n].cpu().numpy()), u32(wc))
            ok = ok and float(bs[:o_s].abs().sum()) == 0 and float(bs[o_s + n:].abs().sum()) == 0
            if n < 3_000_000:
                hs_, hc_ = fb.sincosf(xf)
                ok = ok and np.array_equal(u32(hs_), u32(ws)) and np.array_equal(u32(hc_), u32(wc))
        elif mode == 3:  # pageable host
            s, c = fb.sincos(x)
            ok = np.array_equal(bits(s), bits(hs)) and np.array_equal(bits(c), bits(hc))
        else:  # pinned host (zero-copy or pipeline)
            px, ps, pc = fb.PinnedArray(n), fb.PinnedArray(n), fb.PinnedArray(n)
            px.array[:] = x
            if mode == 4:
                fb.sincos(px.array, out_sin=ps.array, out_cos=pc.array)
                ok = np.array_equal(bits(ps.array), bits(hs)) and np.array_equal(bits(pc.array), bits(hc))
            else:
                fb.cos_array(px.array, out=pc.array)
                fb.sin_array(px.array, out=px.array)
                ok = np.array_equal(bits(px.array), bits(hs)) and np.array_equal(bits(pc.array), bits(hc))
            for p in (px, ps, pc):
                p.free()
        if not ok:
            bad += 1
            print(f"MISMATCH it={it} n={n} mode={mode} offs=({off_in},{off_s},{off_c}) opts={opts}", flush=True)
    print(f"soak: {iters} iterations, {bad} mismatches, seed {seed}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
