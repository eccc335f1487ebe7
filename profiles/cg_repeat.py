import sys, time; sys.path.insert(0, "/root/repo")
import numpy as np, lightkrylov_b200 as lk
ctx = lk.Context(0)
m = 256; n = m**3
A = lk.LinOp.stencil7(ctx, "d", m, m, m, (6.0, -1, -1, -1, -1, -1, -1))
b = lk.Vector(ctx, "d", n).fill_random("uniform", 45); x = lk.Vector(ctx, "d", n)
for i in range(10):
    x.zero(); ctx.sync(); t0 = time.perf_counter(); l0 = ctx.kernel_launches
    info, meta = lk.cg(A, b, x, maxiter=2000); ctx.sync(); dt = time.perf_counter() - t0
    print(i, info, round(dt, 3), ctx.kernel_launches - l0, flush=True)
