"""2-GPU parity of the row-sharded path (NCCL halo exchange + all-reduce) against the oracle.  Needs >= 2 CUDA
devices: skipped on single-GPU boxes (the host logic is covered by tests/test_distributed_cpu.py with gloo)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist

    import asgfem_b200 as A
    from asgfem_b200 import distributed as D
    from oracle import problem as oproblem, solver as osolver

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        _worker_body(rank, world, q, dist, A, D, oproblem, osolver)
    except Exception as e:  # the parent must not wait for the timeout
        import traceback
        q.put(("error", rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


def _worker_body(rank, world, q, dist, A, D, oproblem, osolver):
    if True:
        P = oproblem.poisson_simple(nrefs=3, order=2)
        A0 = sp.csr_matrix(P.A0)
        owner = D.partition_rows(P.n, world)  # index blocks (dof coordinates of P2 face dofs are not needed)
        L = D.LocalProblem(rank, owner, A0.indptr, A0.indices)
        ctx = A.Context(rank)
        D.setup_context(ctx, L, [P.A0] + P.Am, P.bdofs, P.family, P.multi_indices)
        ctx.vec_alloc(6)
        be = D.ContextBackend(ctx, P.N)
        op = D.DistributedOperator(L, be, dist)
        xg = np.random.default_rng(0).standard_normal(P.n * P.N)
        ref = osolver.SystemPrimal(P.A0, P.Am, P.G, P.bdofs, P.N).mul(xg).reshape(P.N, P.n)
        xl = np.zeros((P.N, L.n_local))
        xl[:, :L.n_owned] = xg.reshape(P.N, P.n)[:, L.owned]  # halo rows zero: the exchange must fill them
        ctx.vec_upload(0, xl.reshape(-1))
        op.apply(0, 1)
        got = ctx.vec_download(1).reshape(P.N, L.n_local)[:, :L.n_owned]
        err_apply = np.abs(got - ref[:, L.owned]).max() / np.abs(ref).max()
        # distributed PCG (rank-local mean preconditioner) against the oracle's GMRES solution
        refsol = np.zeros(P.n * P.N)
        osolver.solve_primal(refsol, P.A0, P.Am, P.b0, P.G, P.N, P.bdofs)
        refsol = refsol.reshape(P.N, P.n)
        b = np.zeros((P.N, L.n_local))
        b[0, :L.n_owned] = P.b0[L.owned]
        g2l = L.global_to_local[P.bdofs]
        b[:, g2l[g2l >= 0]] = 0
        ctx.vec_zero(0)
        ctx.vec_upload(1, b.reshape(-1))
        st = D.pcg(op, dict(x=0, b=1, r=2, z=3, p=4, q=5), atol=1e-14, rtol=1e-13, itmax=500)
        sol = ctx.vec_download(0).reshape(P.N, L.n_local)[:, :L.n_owned]
        err_sol = np.abs(sol - refsol[:, L.owned]).max() / np.abs(refsol).max()
        # ---- the same through the library's own NCCL path (asgfem_comm_init / asgfem_set_halo): one C call per operator
        # application and per solve, what the Julia shim would use
        ids = [A.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(world, rank, ids[0])
        ctx.set_halo(L.send, L.recv, 0, L.n_interior)
        ctx.vec_upload(0, xl.reshape(-1))
        ctx.apply(0, 1)
        got2 = ctx.vec_download(1).reshape(P.N, L.n_local)[:, :L.n_owned]
        err_apply2 = np.abs(got2 - ref[:, L.owned]).max() / np.abs(ref).max()
        same = bool(np.array_equal(got2, got))
        ctx.vec_zero(0)
        b0l = np.zeros(L.n_local)
        b0l[:L.n_owned] = P.b0[L.owned]
        st2 = ctx.pcg(b0l, 0, 1e-14, 1e-13, 500)
        sol2 = ctx.vec_download(0).reshape(P.N, L.n_local)[:, :L.n_owned]
        err_sol2 = np.abs(sol2 - refsol[:, L.owned]).max() / np.abs(refsol).max()
        gd = ctx.vec_dot_global(0, 0)
        # ---- exact (global) mean preconditioner through the mode-shard transposition: the iteration count of the
        # single-GPU solve, whatever the number of ranks
        Ls = [D.LocalProblem(r_, owner, A0.indptr, A0.indices) for r_ in range(world)]
        perm = np.concatenate([l.owned for l in Ls])       # rank-major global numbering
        inv = np.empty(P.n, dtype=np.int64)
        inv[perm] = np.arange(P.n)
        A0g = sp.csc_matrix(sp.csr_matrix(P.A0)[perm][:, perm])
        A0g.sort_indices()
        offs = np.concatenate([[0], np.cumsum([l.n_owned for l in Ls])])
        ctx.precond_setup_global(P.n, A0g.indptr + 1, A0g.indices + 1, A0g.data, inv[P.bdofs] + 1, offs)
        ctx.vec_zero(0)
        st3 = ctx.pcg(b0l, 0, 1e-14, 1e-13, 500)
        sol3 = ctx.vec_download(0).reshape(P.N, L.n_local)[:, :L.n_owned]
        err_sol3 = np.abs(sol3 - refsol[:, L.owned]).max() / np.abs(refsol).max()
        # reference iteration count: the same solve on one GPU (rank 0 only)
        nit1 = -1
        if rank == 0:
            c1 = A.Context(rank)
            c1.set_multiindices(P.family, np.array(P.multi_indices, dtype=np.int64))
            A0c = sp.csc_matrix(P.A0)
            A0c.sort_indices()
            c1.set_pattern_csc(P.n, A0c.indptr.astype(np.int64) + 1, A0c.indices.astype(np.int64) + 1)
            c1.set_num_stiffness(P.M)
            c1.set_stiffness(0, A0c.data)
            for m_, Am_ in enumerate(P.Am, start=1):
                Am_ = sp.csc_matrix(Am_)
                Am_.sort_indices()
                c1.set_stiffness(m_, Am_.data)
            c1.set_bdofs(P.bdofs + 1)
            c1.vec_alloc(1)
            c1.vec_zero(0)
            nit1 = int(c1.pcg(P.b0, 0, 1e-14, 1e-13, 500)["niter"])
            c1.close()
        q.put((rank, err_apply, st["niter"], bool(st["solved"]), err_sol, err_apply2, same, int(st2["niter"]), bool(st2["solved"]),
               err_sol2, int(st3["niter"]), bool(st3["solved"]), err_sol3, nit1, gd))
        dist.barrier()
        ctx.comm_destroy()
        ctx.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_operator_and_pcg():
    world = 2
    c = mp.get_context("spawn")
    q = c.Queue()
    procs = [c.Process(target=_worker, args=(r, world, 29650 + os.getpid() % 300, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    for _ in range(world):
        try:
            r = q.get(timeout=240)
        except Exception:
            r = ("error", -1, "timeout waiting for a rank")
        if r[0] == "error":
            for p in procs:  # the other rank may be stuck in a collective
                p.terminate()
            raise AssertionError(r)
        res.append(r)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err_apply, niter, solved, err_sol, err_apply2, same, niter2, solved2, err_sol2, niter3, solved3, err_sol3, nit1, gd in res:
        assert err_apply < 1e-12, (rank, err_apply)
        assert solved and niter < 300
        assert err_sol < 1e-10, (rank, err_sol)
        assert err_apply2 < 1e-12 and same, (rank, err_apply2, same)
        assert solved2 and niter2 == niter, (niter2, niter)
        assert err_sol2 < 1e-10, (rank, err_sol2)
        assert solved3 and err_sol3 < 1e-10 and niter3 < niter, (niter3, niter, err_sol3)
        if nit1 >= 0:
            assert abs(niter3 - nit1) <= 1, (niter3, nit1)  # exact mean preconditioner: the single-GPU iteration count
    assert res[0][-1] == res[1][-1]  # the all-reduced inner product is the same number on both ranks


# ---- row-sharded estimator (cells / faces sharded with the rows, totals all-reduced) -------------
def _est_worker(rank, world, port, q, order):
    import torch.distributed as dist

    import asgfem_b200 as A
    from asgfem_b200 import distributed as D, multiindices as MI, grids as G

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ncx, ncy = 12, 6
        modes = A.graded_lex_multiindices(3, 10)
        Cf = A.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=8)
        mi_ext = np.array(MI.add_boundary_modes(modes, tail_extension=(5, 2)), dtype=np.int64)
        quadorder = 2 * (order - 1) + 1
        xref, w = G.quadrature_rule(quadorder)
        sf, wf = G.quadrature_rule_1d(quadorder)

        def field(S):  # the same smooth "solution" on every rank: a function of the global dof position and the mode
            k = np.arange(len(modes))[:, None]
            return np.sin(0.37 * S.px[None, :] + 0.11 * k) * np.cos(0.23 * S.py[None, :] - 0.05 * k) / (1.0 + k)

        def cell_keys(S):
            c = S.grid.cellnodes
            xy = S.grid.coords[c].sum(axis=1)  # 3 x centroid
            return np.rint(xy[:, 0] * 1e6).astype(np.int64) * 10_000_000 + np.rint(xy[:, 1] * 1e6).astype(np.int64)

        def run(S, ctx, sharded):
            ctx.set_multiindices(A.LEGENDRE, np.array(modes, dtype=np.int64))
            D.setup_strip_context(ctx, S, Cf, 3, G.quadrature_rule(2 * order))
            ctx.vec_alloc(1)
            u = field(S)
            if sharded:
                u[:, S.n_owned:] = 0.0  # the exchange must fill the halo rows
            ctx.vec_upload(0, u.reshape(-1))
            if sharded:
                ids = [A.Context.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                ctx.comm_init(world, rank, ids[0])
                ctx.set_halo(S.send, S.recv, *S.interior)
                ctx.halo_exchange(0)
                ctx.set_owned_cells(S.cell_owned)
            c = S.grid.cellnodes
            x1, x2, x3 = S.grid.coords[c[:, 0]], S.grid.coords[c[:, 1]], S.grid.coords[c[:, 2]]
            xq = x1[:, None, :] + xref[None, :, 0:1] * (x2 - x1)[:, None, :] + xref[None, :, 1:2] * (x3 - x1)[:, None, :]
            fq = 1.0 + xq[:, :, 0] * xq[:, :, 1]
            return ctx.estimate_poisson_primal(0, mi_ext, xref, w, sf, wf, S.grid.ncells, fq)

        S = D.strip_shard(rank, world, order, ncx, ncy, upper_halo_rows=2)
        ctx = A.Context(rank)
        em, ec = run(S, ctx, True)
        own = S.cell_owned.astype(bool)
        out = (rank, em.copy(), cell_keys(S)[own], np.array(ec)[own].copy(), None, None, None)
        if rank == 0:  # reference: the whole mesh on one GPU
            S1 = D.strip_shard(0, 1, order, ncx, ncy * world)
            c1 = A.Context(rank)
            em1, ec1 = run(S1, c1, False)
            out = out[:4] + (em1.copy(), cell_keys(S1), np.array(ec1).copy())
            c1.close()
        q.put(out)
        dist.barrier()
        ctx.comm_destroy()
        ctx.close()
    except Exception:
        import traceback
        q.put(("error", rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("order", [1, 2])
def test_two_gpu_estimator_matches_single_gpu(order):
    world = 2
    c = mp.get_context("spawn")
    q = c.Queue()
    procs = [c.Process(target=_est_worker, args=(r, world, 29350 + os.getpid() % 300 + order, q, order)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        try:
            r = q.get(timeout=240)
        except Exception:
            r = ("error", -1, "timeout waiting for a rank")
        if r[0] == "error":
            for p in procs:
                p.terminate()
            raise AssertionError(r)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    em_ref, keys_ref, ec_ref = res[0][4], res[0][5], res[0][6]
    pos = {int(k): i for i, k in enumerate(keys_ref)}
    seen = 0
    for r in range(world):
        _, em, keys, ec = res[r][:4]
        # the all-reduced mode totals are the single-GPU ones on every rank (tolerance of the estimator: 1e-10)
        assert np.max(np.abs(em - em_ref)) <= 1e-10 * np.max(np.abs(em_ref)), (r, em, em_ref)
        rows = np.array([pos[int(k)] for k in keys])
        assert np.max(np.abs(ec - ec_ref[rows])) <= 1e-10 * np.max(np.abs(ec_ref))
        seen += len(rows)
    assert seen == len(keys_ref)  # every cell of the mesh is owned by exactly one rank
