#!/usr/bin/env python
"""Instruction counts of a kernel's SASS, per loop -- the number that bounds an issue-bound kernel, read off the built
library without a GPU.

    python tools/sass_stats.py [library.so] 'sincos_fused_kernelILi0ELi4ELi1ELb0ELi0ELb1'   (substring of the mangled name)

Prints the kernel's total instruction count, every loop (a backward branch and its target) with its length and its
instruction mix, and the call targets.  (Until the middle of round 2 the dense-warp rounds of the fused kernel were a
rolled loop -- the one that held the IMAD.WIDE.U32 of the Payne-Hanek product; they are now the unrolled out-of-line
routine dense_vector, one of the call targets, and the exact Payne-Hanek product another.)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_sass(lib, pattern):
    text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, name = [], None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        if name and pattern in name:
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip(), name))
    return out


def opcode(ins):
    ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
    return ins.split()[0]


def classify(op):
    base = op.split(".")[0]
    if base in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"):
        return "fp64"
    if base in ("IMAD", "IMAD.WIDE"):
        return "imad"
    if base in ("LDS", "STS", "LDG", "STG", "LDC", "LDCU", "LD", "ST", "ATOMS", "ATOMG", "RED", "LDL", "STL"):
        return "mem"
    if base in ("BRA", "CALL", "RET", "EXIT", "BSSY", "BSYNC", "WARPSYNC", "NOP", "BAR", "VOTE", "VOTEU", "SHFL", "BMOV"):
        return "ctrl"
    return "alu"


def main():
    args = sys.argv[1:]
    lib = os.path.join(ROOT, "fabe_b200", "lib", "libfabe13.so")
    if args and args[0].endswith(".so"):
        lib = args.pop(0)
    pattern = args[0] if args else "sincos_fused_kernelILi0ELi4ELi1ELb0ELi0ELb1"
    ins = kernel_sass(lib, pattern)
    if not ins:
        sys.exit(f"no kernel matching {pattern!r}")
    names = sorted({n for _, _, n in ins})
    if len(names) > 1:
        sys.exit("pattern matches several kernels:\n  " + "\n  ".join(names))
    addr = [a for a, _, _ in ins]
    index = {a: i for i, a in enumerate(addr)}
    print(f"{names[0]}\n  {len(ins)} instructions")
    mix = collections.Counter(classify(opcode(t)) for _, t, _ in ins)
    print("  mix:", dict(mix))
    loops = []
    for i, (a, t, _) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in index:
                loops.append((index[tgt], i))
    for lo, hi in loops:
        body = ins[lo:hi + 1]
        ops = collections.Counter(opcode(t) for _, t, _ in body)
        kinds = collections.Counter(classify(opcode(t)) for _, t, _ in body)
        wide = sum(v for k, v in ops.items() if k.startswith("IMAD.WIDE"))
        print(f"  loop {addr[lo]:#06x}..{addr[hi]:#06x}: {len(body)} instructions, {dict(kinds)}, IMAD.WIDE {wide}")
        if "-v" in args:
            for k, v in ops.most_common():
                print(f"      {v:4d} {k}")
    calls = collections.Counter(re.search(r"0x[0-9a-f]+", t).group(0) for _, t, _ in ins if opcode(t).startswith("CALL") and re.search(r"0x[0-9a-f]+", t))
    print("  calls:", dict(calls))


if __name__ == "__main__":
    main()
