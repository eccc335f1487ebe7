"""A/B of the L2-blocked CSR SpMV on the C5 matrix (50M x 40M, 32 nnz/row, cdp; scale argument shrinks it):
kernel variant (LKB_CSR_BLOCKED_VARIANT 0 thread-per-row, 1 CSR-stream) x slice size (csr_slice_kb).

    python profiles/spmv_ab.py [scale] [slice_mb ...]      (one process per variant: the switch is read once)
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("SPMV_AB_CHILD"):
    sys.path.insert(0, ROOT)
    import torch, lightkrylov_b200 as lk
    scale = float(sys.argv[1]); slices = [int(a) for a in sys.argv[2:]]
    m, n, per_row = int(50e6 * scale), int(40e6 * scale), 32
    ctx = lk.Context(0)
    ext = torch.cuda.ExternalStream(ctx.stream)
    nnz = m * per_row
    alg = nnz * (16 + 4) + 8 * (m + 1) + (n + m) * 16
    for sl in slices:
        ctx.set_option("csr_slice_kb", sl * 1024)
        A = lk.LinOp.csr_random(ctx, "z", m, n, per_row, 46)
        x = lk.Vector(ctx, "z", n).fill_random("normal", 1); y = lk.Vector(ctx, "z", m).fill_random("normal", 2)
        out = {"variant": os.environ.get("LKB_CSR_BLOCKED_VARIANT", "default"), "slice_mb": sl, "m": m, "n": n}
        for name, fn in (("matvec", lambda: A.matvec(x, y)), ("rmatvec", lambda: A.rmatvec(y, x))):
            for _ in range(2): fn()
            ctx.sync()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ext):
                e0.record(ext)
                for _ in range(8): fn()
                e1.record(ext)
            ctx.sync(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 8
            out[name + "_ms"] = round(ms, 3); out[name + "_alg_GBps"] = round(alg / ms / 1e6, 1)
        print(json.dumps(out), flush=True)
        del A, x, y
    ctx.close()
else:
    scale = sys.argv[1] if len(sys.argv) > 1 else "1.0"
    slices = sys.argv[2:] or ["48"]
    for v, cfg in (("2", "0"), ("1", "0"), ("3", "0")):
        env = dict(os.environ, LKB_CSR_BLOCKED_VARIANT=v, LKB_CSR_STREAM_CFG=cfg, SPMV_AB_CHILD="1")
        print("variant", v, "stream cfg", cfg, flush=True)
        subprocess.run([sys.executable, __file__, scale] + slices, env=env)
