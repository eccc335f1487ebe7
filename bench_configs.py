#!/usr/bin/env python
"""bench_configs.py -- the other BASELINE.json configs (C2 gmres, C3 eigs, C4 lanczos/eighs/cg,
C5 bidiagonalization/svds) run through the public API on one GPU, with property checks.

These are NOT the headline bench lines (bench.py measures configs[1]); they show the callers of
the hot path working at (or near) the named sizes and report steps/s plus achieved GB/s against
the official per-step algorithmic bytes.  One JSON line per config.

    python bench_configs.py [--full] [--only c2,c2b,c2e,c3,c4,c5]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import lightkrylov_b200 as lk  # noqa: E402

POISSON5 = (4.0, -1.0, -1.0, -1.0, -1.0)
CONVDIFF7 = (6.0, -1.3, -0.7, -1.2, -0.8, -1.1, -0.9)
LAPLACE7 = (6.0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0)
PEAK = 6553.6
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def timed(ctx, fn):
    ctx.sync(); t0 = time.perf_counter(); r = fn(); ctx.sync()
    return r, time.perf_counter() - t0


def emit(**kw):
    print(json.dumps(kw), flush=True)


def c2_gmres(ctx, full):
    """Config 2, second half: gmres(kdim=50, maxiter=10) on the 4096^2 Poisson operator, grid rows over the ranks."""
    nx = ny = 4096 if full else 2048
    n = nx * ny
    A = lk.LinOp.stencil5(ctx, "d", nx, ny, POISSON5)             # slab of this rank
    nloc, row0 = A.n, A.row0
    b = lk.Vector(ctx, "d", nloc, n_global=n, row0=row0).fill_random("uniform", 43)
    x = lk.Vector(ctx, "d", nloc, n_global=n, row0=row0)
    lk.gmres(A, b, x, kdim=50, maxiter=10)                        # warm-up (allocation pool, graph capture)
    x.zero()
    (info, meta), dt = timed(ctx, lambda: lk.gmres(A, b, x, kdim=50, maxiter=10))
    r = lk.Vector(ctx, "d", nloc, n_global=n, row0=row0); A.matvec(x, r); r.sub(b)
    rn = r.norm()
    if ctx.rank != 0:
        return
    steps = meta["n_inner"]
    # per inner step j: matvec + 4*j*n*s (official); j cycles 1..50
    byts = sum(2 * n * 8 + 4 * ((i % 50) + 1) * n * 8 for i in range(steps))
    emit(config="C2 gmres(kdim=50, maxiter=10) 5-pt Poisson %dx%d fp64, %d GPU(s)" % (nx, ny, ctx.world), n=n, info=info,
         n_inner=steps, n_outer=meta["n_outer"], converged=meta["converged"], seconds=dt,
         inner_steps_per_s=steps / dt, alg_GBps_per_gpu=byts / dt / 1e9 / ctx.world,
         frac_of_measured_hbm=byts / dt / 1e9 / PEAK / ctx.world,
         final_residual=rn, res_first=meta["res"][0], res_last=meta["res"][-1],
         res_history_sha=__import__("hashlib").sha1(np.asarray(meta["res"]).tobytes()).hexdigest()[:12],
         res_every_50=[float(v) for v in meta["res"][::50]])


def c2_block(ctx, full):
    """Block Arnoldi (blksize = 2, 4) on the C2 operator: 129-132 basis columns as in the headline config.
    Official bytes per block step with j = k p existing vectors: p matvecs + 4 j n s per block COLUMN (the reference's
    innerprod_matrix / linear_combination_matrix re-read X for every column of Y) + the in-block QR."""
    nx = ny = 4096 if full else 2048
    n = nx * ny
    A = lk.LinOp.stencil5(ctx, "d", nx, ny, POISSON5)
    if ctx.world > 1:
        return
    for p, kdim in ((2, 64), (4, 32)):
        X = lk.Basis(ctx, "d", n, (kdim + 1) * p)
        for i in range(p):
            X.col(i).fill_random("uniform", 42 + i)
        X.orthonormalize(0, p)
        H = np.zeros(((kdim + 1) * p, kdim * p), order="F")
        lk.arnoldi(A, X, H, blksize=p, kend=2)                                   # warm-up
        (info, dt) = timed(ctx, lambda: lk.arnoldi(A, X, H, blksize=p))
        byts = sum(p * 2 * n * 8 + p * 4 * (k * p) * n * 8 for k in range(1, kdim + 1))
        actual = sum(p * 2 * n * 8 + (p // 2) * 4 * (k * p + 2) * n * 8 for k in range(1, kdim + 1))
        G = X.innerprod((kdim + 1) * p, X, wcol0=(kdim + 1) * p - 4, p=4)
        E = np.eye((kdim + 1) * p)[:, -4:]
        emit(config="C2 block arnoldi(blksize=%d, kdim=%d) 5-pt Poisson %dx%d fp64" % (p, kdim, nx, ny), info=int(info), seconds=dt,
             block_steps_per_s=kdim / dt, vectors_per_s=kdim * p / dt, official_GBps=byts / dt / 1e9,
             frac_of_measured_hbm_official=byts / dt / 1e9 / PEAK, sweep_GBps=actual / dt / 1e9,
             orth_err_last4=float(np.abs(G - E).max()), H_block_hessenberg=bool(np.abs(np.tril(H, -p - 1)).max() == 0.0))
        del X


def c2_expm(ctx, full):
    """Block Krylov exponential kexpm_mat (p = 3) on the C2 operator (block Arnoldi's caller), checked against
    kexpm_vec column by column.  exp(-tau L) with L = the 5-point operator (spectrum in (0, 8)), tau = 0.25, tol = 1e-9."""
    nx = ny = 4096 if full else 2048
    n = nx * ny
    if ctx.world > 1:
        return
    A = lk.LinOp.stencil5(ctx, "d", nx, ny, POISSON5)
    p, tau, tol, kdim = 3, -0.25, 1e-9, 20
    B = lk.Basis(ctx, "d", n, p); Cb = lk.Basis(ctx, "d", n, p); Cv = lk.Basis(ctx, "d", n, p)
    for i in range(p):
        B.col(i).fill_random("uniform", 50 + i)
    lk.kexpm_mat(Cb, A, B, tau, tol, kdim=kdim)                                   # warm-up (allocation pool)
    info, dt = timed(ctx, lambda: lk.kexpm_mat(Cb, A, B, tau, tol, kdim=kdim))
    vinfo, dtv = timed(ctx, lambda: [lk.kexpm(Cv.col(i), A, B.col(i), tau, tol, kdim=kdim * p) for i in range(p)])
    num = den = 0.0
    for i in range(p):
        d = Cv.col(i); den += d.norm() ** 2
        d.sub(Cb.col(i)); num += d.norm() ** 2
    ksteps = info // p - 1 if info > 0 else kdim * p
    # official bytes of the block steps taken (as c2_block) + the final assembly X(:, :kpp) M (p sweeps of kpp vectors)
    byts = sum(p * 2 * n * 8 + p * 4 * (k * p) * n * 8 for k in range(1, ksteps + 1)) + p * (max(info, p) + 1) * n * 8
    emit(config="C2 kexpm_mat(p=%d, tau=%g, tol=%g) 5-pt Poisson %dx%d fp64" % (p, tau, tol, nx, ny), info=int(info),
         block_steps=int(ksteps), seconds=dt, official_GBps=byts / dt / 1e9, frac_of_measured_hbm_official=byts / dt / 1e9 / PEAK,
         kexpm_vec_infos=[int(v) for v in vinfo], kexpm_vec_seconds_3cols=dtv,
         rel_diff_block_vs_columnwise=float(np.sqrt(num / den)))


def c3_eigs(ctx, full):
    """Config 3: eigs (Krylov-Schur, kdim=128) on the 7-point convection-diffusion operator, z-slabs over the ranks."""
    m = 512 if full else 256
    n = m ** 3
    lk.set_lapack_from_scipy()
    A = lk.LinOp.stencil7(ctx, "d", m, m, m, CONVDIFF7)          # slab of this rank (partition over ctx.world)
    nloc, row0 = A.n, A.row0
    nev, kdim = 8, 128
    X = lk.Basis(ctx, "d", nloc, nev, n_global=n, row0=row0)
    x0 = lk.Vector(ctx, "d", nloc, n_global=n, row0=row0).fill_random("uniform", 44)
    (res, dt) = timed(ctx, lambda: lk.eigs(A, X, nev, x0=x0, kdim=kdim, tolerance=1e-6))
    ev, resid, info = res
    # residual check of the leading real-pair / real eigenpair on the device
    y = lk.Vector(ctx, "d", nloc, n_global=n, row0=row0)
    chk = None
    if abs(ev[0].imag) == 0:
        A.matvec(X.col(0), y); y.axpby(-ev[0].real, X.col(0), 1.0); chk = y.norm() / X.col(0).norm()
    if ctx.rank != 0:
        return
    emit(config="C3 eigs(nev=8, kdim=128) 7-pt convection-diffusion %d^3 fp64, %d GPU(s)" % (m, ctx.world), n=n, info_niter=int(info),
         seconds=dt, arnoldi_steps_per_s=info / dt, eigvals=[[float(z.real), float(z.imag)] for z in ev],
         residuals=[float(r) for r in resid], leading_pair_residual=chk, matvecs=A.counters()[0])


def c4_sym(ctx, full):
    m = 384 if full else 256
    n = m ** 3
    lk.set_lapack_from_scipy()
    for kind in ("d", "s"):
        es = 8 if kind == "d" else 4
        A = lk.LinOp.stencil7(ctx, kind, m, m, m, LAPLACE7)
        kdim = 128
        X = lk.Basis(ctx, kind, n, kdim + 1)
        x0 = X.col(0).fill_random("uniform", 45); x0.scal(1.0 / x0.norm())
        T = np.zeros((kdim + 1, kdim), dtype=lk.DTYPES[kind], order="F")
        lk.lanczos(A, X, T)                                             # warm-up (graph capture)
        (info, dt) = timed(ctx, lambda: lk.lanczos(A, X, T))
        G = X.innerprod(kdim + 1, X, wcol0=0, p=8)                      # first 8 columns of the Gram matrix
        orth = float(np.abs(G - np.eye(kdim + 1)[:, :8]).max())
        byts = sum(2 * n * es + 4 * j * n * es + 4 * n * es for j in range(1, kdim + 1))
        emit(config="C4 lanczos(kdim=128) 7-pt Laplacian %d^3 %s" % (m, "fp64" if kind == "d" else "fp32"), n=n,
             info=info, seconds=dt, steps_per_s=kdim / dt, alg_GBps=byts / dt / 1e9,
             frac_of_measured_hbm=byts / dt / 1e9 / PEAK, orth_err_first8=orth,
             T_sym_err=float(np.abs(np.diag(T, 1)[:kdim - 1] - np.diag(T, -1)[:kdim - 1]).max()))
        del X
        Xe = lk.Basis(ctx, kind, n, 8)
        x0 = lk.Vector(ctx, kind, n).fill_random("uniform", 45)
        (res, dt) = timed(ctx, lambda: lk.eighs(A, Xe, 8, x0=x0, kdim=128, tolerance=1e-6 if kind == "d" else 1e-3))
        ev, resid, info = res
        emit(config="C4 eighs(nev=8, kdim=128) %d^3 %s" % (m, "fp64" if kind == "d" else "fp32"), info_k=int(info),
             seconds=dt, eigvals=[float(v) for v in ev], residuals=[float(r) for r in resid])
        del Xe
        b = lk.Vector(ctx, kind, n).fill_random("uniform", 45); x = lk.Vector(ctx, kind, n)
        ((info, meta), dt) = timed(ctx, lambda: lk.cg(A, b, x, maxiter=2000))
        byts = meta["n_iter"] * (2 + 2 + 3 + 3 + 2 + 3) * n * es     # matvec, dot, 2 axpby(3), dot, axpby per iteration
        emit(config="C4 cg(maxiter=2000) %d^3 %s" % (m, "fp64" if kind == "d" else "fp32"), info=int(info),
             n_iter=meta["n_iter"], converged=meta["converged"], seconds=dt, iters_per_s=meta["n_iter"] / dt,
             alg_GBps=byts / dt / 1e9, res_last=meta["res"][-1])


def c5_svds(ctx, full):
    """Config 5: random rectangular CSR (50M x 40M, 32 nnz/row, cdp): bidiagonalization(kdim=32) + svds(nsv=8).
    The matrix is generated AND transposed on the device (lkb_csr_random_device / lkb_op_csr_create_device): at BASELINE
    size it holds 1.6e9 non-zeros = 64.7 GB with its explicit transpose, the two bases take 47.5 GB more."""
    m, n = (50_000_000, 40_000_000) if full else (5_000_000, 4_000_000)
    per_row = 32
    lk.set_lapack_from_scipy()
    if ctx.world > 1:
        return c5_sharded(ctx, m, n, per_row)
    t0 = time.perf_counter()
    A = lk.LinOp.csr_random(ctx, "z", m, n, per_row, 46)
    ctx.sync()
    tgen = time.perf_counter() - t0
    nnz = m * per_row
    es = 16
    spmv = nnz * (es + 4) + 8 * (m + 1) + (n + m) * es
    # stand-alone SpMV rates (CUDA-event-free: wall clock around 10 back-to-back launches + sync)
    x = lk.Vector(ctx, "z", n).fill_random("normal", 1); y = lk.Vector(ctx, "z", m).fill_random("normal", 2)
    rates = {}
    for name, fn in (("matvec", lambda: A.matvec(x, y)), ("rmatvec", lambda: A.rmatvec(y, x))):
        fn(); ctx.sync()
        t1 = time.perf_counter()
        for _ in range(10):
            fn()
        ctx.sync()
        rates[name] = spmv / ((time.perf_counter() - t1) / 10) / 1e9
    del x, y
    kdim, nsv = 32, 8
    U = lk.Basis(ctx, "z", m, kdim + 1); V = lk.Basis(ctx, "z", n, kdim + 1)
    u0 = U.col(0).fill_random("normal", 47); u0.scal(1.0 / u0.norm())
    B = np.zeros((kdim + 1, kdim), dtype=np.complex128, order="F")
    lk.bidiagonalization(A, U, V, B)
    (info, dt) = timed(ctx, lambda: lk.bidiagonalization(A, U, V, B))
    byts = sum(2 * spmv + 4 * (k - 1) * n * es + 4 * k * m * es for k in range(1, kdim + 1))
    # parity-style properties at full size: orthonormality of both bases (device Gram blocks), the Golub-Kahan relation
    # A v_k = alpha_k u_k + beta_k u_{k+1} evaluated on the device for k = 1, 16, 32, bidiagonal structure of B
    Gu = U.innerprod(kdim + 1, U, wcol0=0, p=4); Gv = V.innerprod(kdim, V, wcol0=0, p=4)
    orth = float(max(np.abs(Gu - np.eye(kdim + 1)[:, :4]).max(), np.abs(Gv - np.eye(kdim)[:, :4]).max()))
    rel = 0.0
    r = lk.Vector(ctx, "z", m)
    for k in (1, 16, 32):
        A.matvec(V.col(k - 1), r)
        r.axpby(-B[k - 1, k - 1], U.col(k - 1), 1.0)
        r.axpby(-B[k, k - 1], U.col(k), 1.0)
        rel = max(rel, r.norm() / abs(B[k - 1, k - 1]))
    del r
    offband = float(np.abs(np.triu(B, 1)).max() + np.abs(np.tril(B, -2)).max())
    emit(config="C5 bidiagonalization(kdim=32) random CSR %dx%d, 32 nnz/row, cdp" % (m, n), info=int(info), seconds=dt,
         steps_per_s=kdim / dt, alg_GBps=byts / dt / 1e9, frac_of_measured_hbm=byts / dt / 1e9 / PEAK,
         device_generation_and_transpose_s=tgen, spmv_alg_GBps=rates, nnz=nnz,
         B_diag_first=[float(abs(B[i, i])) for i in range(4)], orth_err_first4=orth, golub_kahan_relation_rel_err=rel,
         B_off_band_max=offband)
    del U, V
    Us = lk.Basis(ctx, "z", m, nsv); Vs = lk.Basis(ctx, "z", n, nsv)
    u0 = lk.Vector(ctx, "z", m).fill_random("normal", 47)
    (res, dt) = timed(ctx, lambda: lk.svds(A, Us, Vs, nsv, u0=u0, kdim=32, tolerance=1e-6))
    S, resid, info = res
    y = lk.Vector(ctx, "z", m); A.matvec(Vs.col(0), y); y.axpby(-S[0], Us.col(0), 1.0)
    # known answer for this random ensemble: the largest singular value of an m x n matrix with `per_row` i.i.d. complex
    # N(0, 2) entries per row concentrates at sqrt(2 per_row) (1 + sqrt(m / n))  (Marchenko-Pastur edge)
    mp_edge = float(np.sqrt(2.0 * per_row) * (1.0 + np.sqrt(m / n)))
    emit(config="C5 svds(nsv=8, kdim=32)", info_k=int(info), seconds=dt, S=[float(s) for s in S],
         residuals=[float(r) for r in resid], triplet0_residual=y.norm(), marchenko_pastur_edge=mp_edge,
         S0_over_edge=float(S[0] / mp_edge))


def c5_sharded(ctx, m, n, per_row):
    """Config 5 row-sharded over the ranks (lkb_op_csr_create_dist_device): every rank generates, transposes and L2-blocks its
    row block on its own GPU; matvec gathers x over NVLink (grouped ncclBroadcast), rmatvec reduces A_loc^H u_loc onto the
    owners (grouped ncclReduce)."""
    t0 = time.perf_counter()
    A = lk.LinOp.csr_random_dist(ctx, "z", m, n, per_row, 46)
    ctx.sync()
    tgen = time.perf_counter() - t0
    r0, ml = lk.partition(m, ctx.world, ctx.rank); c0, nl = lk.partition(n, ctx.world, ctx.rank)
    nnz, es = m * per_row, 16
    spmv = nnz * (es + 4) + 8 * (m + 1) + (n + m) * es
    x = lk.Vector(ctx, "z", nl, n_global=n, row0=c0).fill_random("normal", 1)
    y = lk.Vector(ctx, "z", ml, n_global=m, row0=r0).fill_random("normal", 2)
    rates = {}
    for name, fn in (("matvec", lambda: A.matvec(x, y)), ("rmatvec", lambda: A.rmatvec(y, x))):
        fn(); ctx.sync()
        t1 = time.perf_counter()
        for _ in range(10):
            fn()
        ctx.sync()
        rates[name] = spmv / ((time.perf_counter() - t1) / 10) / 1e9
    del x, y
    kdim, nsv = 32, 8
    U = lk.Basis(ctx, "z", ml, kdim + 1, n_global=m, row0=r0); V = lk.Basis(ctx, "z", nl, kdim + 1, n_global=n, row0=c0)
    u0 = U.col(0).fill_random("normal", 47); u0.scal(1.0 / u0.norm())
    B = np.zeros((kdim + 1, kdim), dtype=np.complex128, order="F")
    lk.bidiagonalization(A, U, V, B)
    (info, dt) = timed(ctx, lambda: lk.bidiagonalization(A, U, V, B))
    byts = sum(2 * spmv + 4 * (k - 1) * n * es + 4 * k * m * es for k in range(1, kdim + 1))
    Gu = U.innerprod(kdim + 1, U, wcol0=0, p=4); Gv = V.innerprod(kdim, V, wcol0=0, p=4)
    orth = float(max(np.abs(Gu - np.eye(kdim + 1)[:, :4]).max(), np.abs(Gv - np.eye(kdim)[:, :4]).max()))
    del U, V
    Us = lk.Basis(ctx, "z", ml, nsv, n_global=m, row0=r0); Vs = lk.Basis(ctx, "z", nl, nsv, n_global=n, row0=c0)
    u0 = lk.Vector(ctx, "z", ml, n_global=m, row0=r0).fill_random("normal", 47)
    (res, dts) = timed(ctx, lambda: lk.svds(A, Us, Vs, nsv, u0=u0, kdim=32, tolerance=1e-6))
    S, resid, sinfo = res
    if ctx.rank != 0:
        return
    emit(config="C5 bidiagonalization(kdim=32) random CSR %dx%d, 32 nnz/row, cdp, row-sharded over %d GPU(s)" % (m, n, ctx.world),
         info=int(info), seconds=dt, steps_per_s=kdim / dt, alg_GBps_total=byts / dt / 1e9,
         device_generation_transpose_blocking_s=tgen, spmv_alg_GBps_total=rates, nnz=nnz,
         B_diag_first=[float(abs(B[i, i])) for i in range(4)], orth_err_first4=orth,
         B_off_band_max=float(np.abs(np.triu(B, 1)).max() + np.abs(np.tril(B, -2)).max()))
    emit(config="C5 svds(nsv=8, kdim=32), row-sharded over %d GPU(s)" % ctx.world, info_k=int(sinfo), seconds=dts,
         S=[float(v) for v in S], residuals=[float(r) for r in resid])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="BASELINE.json sizes (C3 512^3, C4 384^3, C5 50Mx40M)")
    ap.add_argument("--only", default="c2,c2b,c2e,c3,c4,c5")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:                                   # torchrun: only the sharded config (C3) is meaningful
        import torch
        import torch.distributed as dist
        local = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ctx = lk.Context.from_torch_distributed(local)
        args.only = ",".join(c for c in args.only.split(",") if c in ("c2", "c3", "c5"))      # the row-sharded configs
    else:
        ctx = lk.Context(0)
    for name, fn in (("c2", c2_gmres), ("c2b", c2_block), ("c2e", c2_expm), ("c3", c3_eigs), ("c4", c4_sym), ("c5", c5_svds)):
        if name in args.only.split(","):
            try:
                fn(ctx, args.full)
            except Exception as e:                       # keep going: one JSON line per config either way
                emit(config=name, error=repr(e))
    ctx.sync()
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()
    else:
        ctx.close()


if __name__ == "__main__":
    main()
