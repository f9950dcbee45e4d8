#!/usr/bin/env python
"""Times pvb_uncertainty_pnp (batched LM refinement, one warp per problem) for a few batch sizes, next to the same
arithmetic core compiled for the host (tests/pnp_host_harness.cpp, one core) on the same problems."""
import ctypes
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import clean_pvnet_b200 as pvb  # noqa: E402
from util import pnp_case  # noqa: E402

cases = [pnp_case(1000 + s, pn=9, noise=2.0) for s in range(256)]
so = os.path.join(ROOT, "tests", "_build", "libpnp_host.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "tests", "pnp_host_harness.cpp"), "-o", so])
lib = ctypes.CDLL(so)
DP = ctypes.POINTER(ctypes.c_double)
t0 = time.perf_counter()
out = np.empty(6)
info = (ctypes.c_int * 2)()
for c in cases:
    a = [np.ascontiguousarray(x) for x in c[:5]]
    lib.pnp_host_solve(*[x.ctypes.data_as(DP) for x in a], out.ctypes.data_as(DP), info, ctypes.c_int(9), ctypes.c_int(50),
                       ctypes.c_double(1e-6), ctypes.c_double(1e-10), ctypes.c_double(1e-8))
cpu_us = (time.perf_counter() - t0) / len(cases) * 1e6
print(f"host build of the same core, 1 core: {cpu_us:.1f} us per problem (incl. ctypes call)")
for n in (1, 16, 256, 4096):
    reps = [cases[i % 256] for i in range(n)]
    f = lambda i: torch.from_numpy(np.stack([c[i] for c in reps])).cuda()  # noqa: E731
    uv, p3, W, K, init = f(0), torch.from_numpy(cases[0][1]).cuda(), f(2), torch.from_numpy(cases[0][3]).cuda(), f(4)
    p3 = f(1)
    K = f(3)
    for _ in range(3):
        rt, inf = pvb.uncertainty_pnp_batch(uv, W, p3, K, init, return_info=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        rt = pvb.uncertainty_pnp_batch(uv, W, p3, K, init)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"n={n:5d}: {ms*1e3:8.1f} us per call  {ms*1e3/n:8.2f} us per problem  mean iterations {inf[:,0].float().mean().item():.2f}")
