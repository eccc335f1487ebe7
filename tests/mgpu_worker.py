"""Worker of tests/test_multi_gpu.py: run under torchrun with N ranks (one per GPU).
Row-sharded arnoldi / lanczos / gmres / cg through the C ABI against the single-process CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import lightkrylov_b200 as lk
    from oracle import lk_oracle as lo
    from helpers import CONVDIFF7, POISSON5, rel_normwise

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = lk.Context.from_torch_distributed(local)
    assert ctx.world == world and ctx.rank == rank
    assert ctx.p2p, "in-kernel NVLink allreduce should be active on a single NVSwitch box"

    def gather_rows(local_arr):
        parts = [None] * world
        dist.all_gather_object(parts, local_arr)
        return np.concatenate(parts, axis=0)

    # ---- 2-D Poisson, ragged slabs (ny not divisible by world) ----
    nx, ny, kdim = 128, 101, 40
    n = nx * ny
    for coef, trans in ((POISSON5, False), (CONVDIFF7[:5], True)):
        A = lk.LinOp.stencil5(ctx, "d", nx, ny, coef)
        X = lk.Basis(ctx, "d", A.n, kdim + 1, n_global=n, row0=A.row0)
        x0 = X.col(0).fill_random("uniform", 42)
        x0.scal(1.0 / x0.norm())
        H = np.zeros((kdim + 1, kdim), order="F")
        info = lk.arnoldi(A, X, H, transpose=trans)
        Xg = gather_rows(X.get())
        Hs = [None] * world
        dist.all_gather_object(Hs, H)
        assert all(np.array_equal(Hs[0], h) for h in Hs), "H must be identical on every rank"
        if rank == 0:
            Xo = np.zeros((n, kdim + 1), order="F"); Xo[:, 0] = lo.fill(n, "d", "uniform", 42); lo.normalize(Xo[:, 0])
            Ho = np.zeros_like(H)
            oinfo = lo.arnoldi(lo.Op.stencil("d", (nx, ny), coef), Xo, Ho, trans=trans)
            assert info == oinfo == 0
            assert rel_normwise(H, Ho) < 1e-10, rel_normwise(H, Ho)
            assert np.abs(Xg.T @ Xg - np.eye(kdim + 1)).max() < 1e-12
            assert rel_normwise(Xg, Xo) < 1e-8

    # ---- in-kernel p2p allreduce vs ncclAllReduce: same factorisation to rounding, both deterministic ----
    A = lk.LinOp.stencil5(ctx, "d", nx, ny, POISSON5)
    Hs = {}
    for mode in (1, 0, 1):
        ctx.set_option("p2p", mode)
        X = lk.Basis(ctx, "d", A.n, kdim + 1, n_global=n, row0=A.row0)
        x0 = X.col(0).fill_random("uniform", 42); x0.scal(1.0 / x0.norm())
        H = np.zeros((kdim + 1, kdim), order="F")
        assert lk.arnoldi(A, X, H) == 0
        if mode in Hs:
            assert np.array_equal(Hs[mode], H), "p2p allreduce must be run-to-run deterministic"
        Hs[mode] = H
    assert rel_normwise(Hs[1], Hs[0]) < 1e-12
    ctx.set_option("p2p", 1)

    # ---- round 2: halo push fused into the kernel that finishes the vector (k_multiaxpy_fin / k_scale_dev) vs the
    # separate k_halo_push kernel: identical data reaches the stencil, so H is BITWISE identical; and the fused final
    # pass (predicted norm) vs the round-1 tail agree to rounding ----
    Hmode = {}
    for name, opts in (("default", {}), ("halo_kernel", {"fused_halo": 0}), ("no_fin", {"fin": 0})):
        for k_, v_ in opts.items():
            ctx.set_option(k_, v_)
        X = lk.Basis(ctx, "d", A.n, kdim + 1, n_global=n, row0=A.row0)
        x0 = X.col(0).fill_random("uniform", 42); x0.scal(1.0 / x0.norm())
        H = np.zeros((kdim + 1, kdim), order="F")
        assert lk.arnoldi(A, X, H) == 0
        # one step at a time: every call starts with the generic push and never pushes from its last step
        X2 = lk.Basis(ctx, "d", A.n, kdim + 1, n_global=n, row0=A.row0)
        x2 = X2.col(0).fill_random("uniform", 42); x2.scal(1.0 / x2.norm())
        H2 = np.zeros((kdim + 1, kdim), order="F")
        for k_ in range(1, kdim + 1, 3):
            assert lk.arnoldi(A, X2, H2, kstart=k_, kend=min(k_ + 2, kdim)) == 0
        assert np.array_equal(H, H2), f"{name}: chunked resume must reproduce the one-shot factorisation bitwise"
        Hmode[name] = H
        for k_ in opts:
            ctx.set_option(k_, 1)
    assert np.array_equal(Hmode["default"], Hmode["halo_kernel"]), "fused halo push changed the result"
    assert rel_normwise(Hmode["default"], Hmode["no_fin"]) < 1e-13
    assert np.array_equal(Hmode["default"], Hs[1]) or rel_normwise(Hmode["default"], Hs[1]) < 1e-15

    # ---- 3-D 7-point, z-sharded: lanczos + cg + gmres ----
    dims = (20, 16, 13); n3 = int(np.prod(dims)); kd = 24
    L7 = (6.0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0)
    A = lk.LinOp.stencil7(ctx, "d", *dims, L7)
    X = lk.Basis(ctx, "d", A.n, kd + 1, n_global=n3, row0=A.row0)
    x0 = X.col(0).fill_random("uniform", 45); x0.scal(1.0 / x0.norm())
    T = np.zeros((kd + 1, kd), order="F")
    info = lk.lanczos(A, X, T)
    b = lk.Vector(ctx, "d", A.n, n_global=n3, row0=A.row0).fill_random("uniform", 43)
    x = lk.Vector(ctx, "d", A.n, n_global=n3, row0=A.row0)
    cinfo, cmeta = lk.cg(A, b, x, maxiter=500)
    xg = gather_rows(x.get())
    A2 = lk.LinOp.stencil7(ctx, "d", *dims, CONVDIFF7)
    x2 = lk.Vector(ctx, "d", A.n, n_global=n3, row0=A.row0)
    ginfo, gmeta = lk.gmres(A2, b, x2, kdim=20, maxiter=30)
    x2g = gather_rows(x2.get())
    # ---- eigs (Krylov-Schur restarts) and eighs on the sharded operator: config-3 shaped ----
    lk.set_lapack_from_scipy()
    nev = 4
    Xe = lk.Basis(ctx, "d", A.n, nev, n_global=n3, row0=A.row0)
    x0e = lk.Vector(ctx, "d", A.n, n_global=n3, row0=A.row0).fill_random("uniform", 44)
    ev, eres, einfo = lk.eigs(A2, Xe, nev, x0=x0e, kdim=24, tolerance=1e-8)
    evs = [None] * world
    dist.all_gather_object(evs, (ev, einfo))
    assert all(np.array_equal(evs[0][0], e[0]) and evs[0][1] == e[1] for e in evs), "eigs must agree on all ranks"
    Xh = lk.Basis(ctx, "d", A.n, nev, n_global=n3, row0=A.row0)
    evh, hres, hinfo = lk.eighs(A, Xh, nev, x0=x0e, kdim=40, tolerance=1e-8)
    if rank == 0:
        evo, reso, Xo_, infoo = lo.eigs(lo.Op.stencil("d", dims, CONVDIFF7), n3, nev, lo.fill(n3, "d", "uniform", 44), kdim=24, tolerance=1e-8)
        assert einfo == infoo, (einfo, infoo)
        assert np.abs(ev[:, None] - evo[None, :]).min(axis=1).max() < 1e-10 * np.abs(evo).max()
        evho, _, _, kho = lo.eighs(lo.Op.stencil("d", dims, L7), n3, nev, lo.fill(n3, "d", "uniform", 44), kdim=40, tolerance=1e-8)
        assert hinfo == kho and np.abs(evh - evho).max() < 1e-10 * np.abs(evho).max()
    # ---- row-sharded random CSR (config-5 shaped): matvec gathers x, rmatvec reduces onto the owners ----
    from helpers import random_csr
    for kind in ("d", "z"):
        dt = lk.DTYPES[kind]
        m5, n5, kd5 = 903, 701, 16                   # ragged slabs on purpose
        S = random_csr(np.random.default_rng(46), m5, n5, 12, dt)
        r0, ml = lk.partition(m5, world, rank); c0, nl = lk.partition(n5, world, rank)
        ip = S.indptr
        Ac = lk.LinOp.csr_dist(ctx, m5, n5, ip[r0:r0 + ml + 1] - ip[r0], S.indices[ip[r0]:ip[r0 + ml]],
                               S.data[ip[r0]:ip[r0 + ml]].astype(dt))
        x5g = lo.fill(n5, kind, "normal", 3); u5g = lo.fill(m5, kind, "normal", 4)
        x5v = lk.Vector(ctx, kind, nl, n_global=n5, row0=c0).put(x5g[c0:c0 + nl])
        u5v = lk.Vector(ctx, kind, ml, n_global=m5, row0=r0).put(u5g[r0:r0 + ml])
        y5v = lk.Vector(ctx, kind, ml, n_global=m5, row0=r0); v5v = lk.Vector(ctx, kind, nl, n_global=n5, row0=c0)
        Ac.matvec(x5v, y5v); Ac.rmatvec(u5v, v5v)
        assert np.allclose(y5v.get(), (S @ x5g)[r0:r0 + ml], rtol=1e-12, atol=1e-12)
        assert np.allclose(v5v.get(), (S.conj().T @ u5g)[c0:c0 + nl], rtol=1e-12, atol=1e-12)
        U5 = lk.Basis(ctx, kind, ml, kd5 + 1, n_global=m5, row0=r0); V5 = lk.Basis(ctx, kind, nl, kd5 + 1, n_global=n5, row0=c0)
        u0 = U5.col(0).fill_random("normal", 47); u0.scal(1.0 / u0.norm())
        B5 = np.zeros((kd5 + 1, kd5), dtype=dt, order="F")
        binfo = lk.bidiagonalization(Ac, U5, V5, B5)
        if rank == 0:
            Uo = np.zeros((m5, kd5 + 1), dtype=dt, order="F"); Uo[:, 0] = lo.fill(m5, kind, "normal", 47); lo.normalize(Uo[:, 0])
            Vo = np.zeros((n5, kd5 + 1), dtype=dt, order="F"); Bo = np.zeros_like(B5)
            assert binfo == lo.bidiag(lo.Op.csr(m5, n5, S.indptr, S.indices, S.data.astype(dt)), Uo, Vo, Bo) == 0
            assert rel_normwise(B5, Bo) < 1e-10
    # ---- the same operator type generated ON THE DEVICE per rank (lkb_csr_random_device + lkb_op_csr_create_dist_device), with
    # the L2-blocked forward layout forced: must equal the oracle's twin of the global matrix ----
    for kind in ("z",):
        dt = lk.DTYPES[kind]
        m6, n6, pr6 = 2003, 1501, 9
        ctx.set_option("csr_slice_kb", 4); ctx.set_option("csr_block_min_kb", 0)
        Ad = lk.LinOp.csr_random_dist(ctx, kind, m6, n6, pr6, 46)
        ctx.set_option("csr_slice_kb", 48 * 1024); ctx.set_option("csr_block_min_kb", 96 * 1024)
        rp6, ci6, va6 = lo.csr_random(kind, m6, n6, pr6, 46)
        Ao6 = lo.Op.csr(m6, n6, rp6, ci6, va6)
        r0, ml = lk.partition(m6, world, rank); c0, nl = lk.partition(n6, world, rank)
        x6 = lo.fill(n6, kind, "normal", 5); u6 = lo.fill(m6, kind, "normal", 6)
        xv = lk.Vector(ctx, kind, nl, n_global=n6, row0=c0).put(x6[c0:c0 + nl]); uv = lk.Vector(ctx, kind, ml, n_global=m6, row0=r0).put(u6[r0:r0 + ml])
        yv = lk.Vector(ctx, kind, ml, n_global=m6, row0=r0); vv = lk.Vector(ctx, kind, nl, n_global=n6, row0=c0)
        Ad.matvec(xv, yv); Ad.rmatvec(uv, vv)
        assert np.allclose(yv.get(), Ao6.apply(x6)[r0:r0 + ml], rtol=1e-11, atol=1e-11)
        assert np.allclose(vv.get(), Ao6.apply(u6, trans=True)[c0:c0 + nl], rtol=1e-11, atol=1e-11)
    if rank == 0:
        Xo = np.zeros((n3, kd + 1), order="F"); Xo[:, 0] = lo.fill(n3, "d", "uniform", 45); lo.normalize(Xo[:, 0])
        To = np.zeros_like(T)
        assert info == lo.lanczos(lo.Op.stencil("d", dims, L7), Xo, To) == 0
        assert rel_normwise(T, To) < 1e-10
        bh = lo.fill(n3, "d", "uniform", 43)
        xo = np.zeros(n3); oinfo, ometa = lo.cg(lo.Op.stencil("d", dims, L7), bh, xo, maxiter=500)
        assert cinfo == oinfo > 0
        assert np.linalg.norm(xg - xo) / np.linalg.norm(xo) < 1e-8
        xo2 = np.zeros(n3); oinfo2, ometa2 = lo.gmres(lo.Op.stencil("d", dims, CONVDIFF7), bh, xo2, kdim=20, maxiter=30)
        assert ginfo == oinfo2 > 0
        assert np.linalg.norm(x2g - xo2) / np.linalg.norm(xo2) < 1e-7
        print(f"MGPU_OK world={world}", flush=True)
    ctx.sync()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
