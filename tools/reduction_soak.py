#!/usr/bin/env python
"""Host soak of the large-argument reduction: random doubles with 2^45 <= |x| < 2^1024 through the product's routine
(fast form, exact form where it hands over) and through the exact form alone (fabe13_debug_reduce_huge, no GPU):
quadrant mismatches, largest difference of the remainders in ulps, how many arguments the fast form handed over.
    python tools/reduction_soak.py <seed> <iterations of 2e6 arguments>"""
import os
import ctypes
import sys

import numpy as np

vp=ctypes.c_void_p
l=ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'fabe_b200', 'lib', 'libfabe13.so'))
l.fabe13_debug_reduce_huge.argtypes=[vp]*6+[ctypes.c_size_t]
seed=int(sys.argv[1]); iters=int(sys.argv[2])
rng=np.random.default_rng(seed)
tot=0; nok=0; qbad=0; rdiff=0; n=2_000_000
r=np.empty(n); re=np.empty(n); q=np.empty(n,dtype=np.int32); qe=np.empty(n,dtype=np.int32); ok=np.empty(n,dtype=np.int32)
for it in range(iters):
    e=rng.integers(1068,2047,n).astype(np.uint64); m=rng.integers(0,1<<52,n).astype(np.uint64); sg=rng.integers(0,2,n).astype(np.uint64)
    x=((sg<<np.uint64(63))|(e<<np.uint64(52))|m).view(np.float64)
    l.fabe13_debug_reduce_huge(x.ctypes.data,r.ctypes.data,q.ctypes.data,re.ctypes.data,qe.ctypes.data,ok.ctypes.data,n)
    tot+=n; nok+=int(ok.sum()); qbad+=int((q!=qe).sum())
    rdiff=max(rdiff,int(np.abs(r.view(np.int64)-re.view(np.int64)).max()))
print(f"seed {seed}: {tot} arguments, fast form vouched for {nok} ({tot-nok} handed over), quadrant mismatches {qbad}, max |r - r_exact| {rdiff} ulp", flush=True)
