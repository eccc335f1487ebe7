"""A/B of the stencil kernel variants (LKB_STENCIL_VARIANT) on C2 (4096^2 fp64) and C4 (384^3)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import torch, lightkrylov_b200 as lk
    ctx = lk.Context(0)
    ext = torch.cuda.ExternalStream(ctx.stream)
    out = {}
    for name, mk, n in (("2d_4096", lambda: lk.LinOp.stencil5(ctx, "d", 4096, 4096, (4., -1., -1., -1., -1.)), 4096 * 4096),
                        ("3d_384", lambda: lk.LinOp.stencil7(ctx, "d", 384, 384, 384, (6., -1., -1., -1., -1., -1., -1.)), 384 ** 3),
                        ("3d_512", lambda: lk.LinOp.stencil7(ctx, "d", 512, 512, 512, (6., -1.3, -0.7, -1.2, -0.8, -1.1, -0.9)), 512 ** 3)):
        A = mk(); x = lk.Vector(ctx, "d", n).fill_random("uniform", 1); y = lk.Vector(ctx, "d", n)
        for _ in range(5): A.matvec(x, y)
        ctx.sync()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record(ext)
            for _ in range(100): A.matvec(x, y); A.matvec(y, x)
            e1.record(ext)
        ctx.sync(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 200
        out[name] = (round(us, 1), round(2 * n * 8 / us / 1e3))
    print(sys.argv[1], out)
else:
    for v in range(5):
        env = dict(os.environ, LKB_STENCIL_VARIANT=str(v))
        subprocess.run([sys.executable, __file__, str(v)], env=env)
