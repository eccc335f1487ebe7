"""Split one launch of the reduction kernels into ramp-up / streaming / reduction tail with the in-kernel
%globaltimer timeline (lkb_debug_ktime).  One GPU; sizes = the C2 per-GPU share at N = 8 and at N = 1.

    python profiles/ktime_probe.py [j ...]
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightkrylov_b200 as lk  # noqa: E402

js = [int(a) for a in sys.argv[1:]] or [32, 100]
ctx = lk.Context(0)
lib = ctx.lib
NW = 4 * 1024 + 8


def read():
    buf = (C.c_uint64 * NW)()
    lk._lib.check(lib.lkb_debug_ktime_read(ctx.h, buf, NW), "ktime_read")
    return np.frombuffer(buf, dtype=np.uint64).astype(np.int64)


def summarize(name, n, j, t, nctas=None):
    cta = t[:4096].reshape(1024, 4)
    if nctas:                                  # rows beyond this kernel's grid are stale entries of the previous launch
        cta = cta[:nctas]
    live = cta[:, 0] > 0
    s, m, k = cta[live, 0], cta[live, 1], cta[live, 2]
    t0 = s.min()
    last = t[4096:4099]
    out = {"kernel": name, "n": n, "j": j, "ctas": int(live.sum()),
           "start_spread_us": round((s.max() - t0) / 1e3, 2),
           "main_done_us": [round(float(x - t0) / 1e3, 2) for x in (m.min(), np.median(m), m.max())],
           "ticket_done_us_max": round((k.max() - t0) / 1e3, 2),
           "tail_after_last_main_us": round((last[1] - m.max()) / 1e3, 2),
           "total_us": round((last[1] - t0) / 1e3, 2)}
    print(json.dumps(out), flush=True)


for ny in (512, 4096):
    n = 4096 * ny
    for j in js:
        X = lk.Basis(ctx, "d", n, j + 1)
        for i in range(j + 1):
            X.col(i).fill_random("normal", i + 1)
        for rep in range(2):                     # second repetition = warm
            lk._lib.check(lib.lkb_debug_ktime(ctx.h, 1), "ktime")
            X.innerprod(j, X, wcol0=j, p=1)
            if rep:
                summarize("k_multidot", n, j, read())
            lk.double_gram_schmidt_step(X, j, 1, X, j, if_chk_orthonormal=False)
            if rep:
                summarize("k_axpy_dot", n, j, read(), nctas=148)
        # final pass (k_multiaxpy_fin) through a one-step arnoldi with j existing vectors: the last reduction-class
        # kernel of the step, so the buffer holds its timeline (grid <= 592 CTAs)
        A = lk.LinOp.stencil5(ctx, "d", 4096, ny, (4.0, -1.0, -1.0, -1.0, -1.0))
        H = np.zeros((j + 1, j), order="F")
        for rep in range(2):
            lk._lib.check(lib.lkb_debug_ktime(ctx.h, 1), "ktime")
            lk.arnoldi(A, X, H, kstart=j, kend=j)
            if rep:
                summarize("k_multiaxpy_fin", n, j, read(), nctas=592)
        del A
        del X
ctx.close()
