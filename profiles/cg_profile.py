import sys, time; sys.path.insert(0, "/root/repo")
import numpy as np, lightkrylov_b200 as lk
ctx = lk.Context(0)
for kind in ("d", "s"):
    m = 256; n = m**3
    A = lk.LinOp.stencil7(ctx, kind, m, m, m, (6.0, -1, -1, -1, -1, -1, -1))
    b = lk.Vector(ctx, kind, n).fill_random("uniform", 45); x = lk.Vector(ctx, kind, n)
    for prof in (False, True, False):
        x.zero(); ctx.set_profile(prof); ctx.sync(); t0 = time.perf_counter()
        info, meta = lk.cg(A, b, x, maxiter=2000); ctx.sync(); dt = time.perf_counter() - t0
        print(kind, prof, info, round(dt, 3), {k: (round(v[0], 1), v[1]) for k, v in ctx.get_profile().items()} if prof else "")
    ctx.set_profile(False)
