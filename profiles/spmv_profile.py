"""Per-class device time of the C5-shaped bidiagonalisation (random CSR, cdp) and stand-alone SpMV GB/s."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, lightkrylov_b200 as lk
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
m, n, per_row = int(50e6 * scale), int(40e6 * scale), 32
rng = np.random.default_rng(46)
col = np.sort(rng.integers(0, n, size=(m, per_row), dtype=np.int32), axis=1).ravel()
val = (rng.standard_normal(m * per_row) + 1j * rng.standard_normal(m * per_row)).astype(np.complex128)
rowptr = np.arange(0, (m + 1) * per_row, per_row, dtype=np.int64)
ctx = lk.Context(0)
A = lk.LinOp.csr(ctx, m, n, rowptr, col, val)
x = lk.Vector(ctx, "z", n).fill_random("normal", 1); y = lk.Vector(ctx, "z", m).fill_random("normal", 2)
ext = torch.cuda.ExternalStream(ctx.stream)
nnz = m * per_row
alg = nnz * (16 + 4) + 8 * (m + 1) + (n + m) * 16
for name, fn in (("matvec", lambda: A.matvec(x, y)), ("rmatvec", lambda: A.rmatvec(y, x))):
    for _ in range(3): fn()
    ctx.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ext):
        e0.record(ext)
        for _ in range(20): fn()
        e1.record(ext)
    ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(json.dumps({"op": name, "m": m, "n": n, "nnz": nnz, "ms": ms, "alg_GBps": alg / ms / 1e6,
                      "x_bytes": (n if name == "matvec" else m) * 16}))
kdim = 32
U = lk.Basis(ctx, "z", m, kdim + 1); V = lk.Basis(ctx, "z", n, kdim + 1)
u0 = U.col(0).fill_random("normal", 47); u0.scal(1.0 / u0.norm())
B = np.zeros((kdim + 1, kdim), dtype=np.complex128, order="F")
lk.bidiagonalization(A, U, V, B)
ctx.set_profile(True)
lk.bidiagonalization(A, U, V, B)
print(json.dumps({"bidiag_profile_ms": {k: (round(v[0], 2), v[1]) for k, v in ctx.get_profile().items()}}))
