import sys, time; sys.path.insert(0, ".")
import numpy as np, lightkrylov_b200 as lk
ctx = lk.Context(0)
nx = ny = 2048; n = nx * ny
A = lk.LinOp.stencil5(ctx, "d", nx, ny, (4.0, -1.0, -1.0, -1.0, -1.0))
b = lk.Vector(ctx, "d", n).fill_random("uniform", 43); x = lk.Vector(ctx, "d", n)
for graphs in (1, 1, 1, 0, 0):
    ctx.set_graphs(bool(graphs)); x.zero(); ctx.sync(); t0 = time.perf_counter()
    info, meta = lk.gmres(A, b, x, kdim=50, maxiter=10); ctx.sync(); dt = time.perf_counter() - t0
    print("graphs", graphs, info, meta["n_inner"], round(dt, 3), "s", round(meta["n_inner"] / dt), "inner steps/s", flush=True)
