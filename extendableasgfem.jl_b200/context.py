"""Thin numpy-facing wrapper around the C ABI (one `Context` = one asgfem_ctx = one GPU, one refinement level).

Everything is converted to exactly what a Julia caller would pass: 1-based Int64/Int32 index arrays,
column-major matrices, flat `entries` vectors in the reference layout.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

LEGENDRE, HERMITE = 0, 1


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Context:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.asgfem_create(C.byref(h), device)
        if rc != 0:
            raise _lib.AsgfemError(rc, self.lib.asgfem_last_error(None).decode())
        self.h = h
        self.n = 0
        self.N = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.asgfem_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise _lib.AsgfemError(rc, self.lib.asgfem_last_error(self.h).decode())

    # ---- stochastic discretisation ---------------------------------------------------------------
    def set_multiindices(self, family, multi_indices):
        mi = _i64(np.asarray(multi_indices))  # (N, M) C-order == M x N column-major
        self.N, self.Mlen = mi.shape
        self._ck(self.lib.asgfem_set_multiindices(self.h, family, self.N, self.Mlen, _ptr(mi)))

    def coupling_csc(self):
        """G as (colptr, rowval, nzval), 1-based, of shape (M*N) x N (TensorizedBasis.G.cscmatrix)."""
        nnz = C.c_int64()
        self._ck(self.lib.asgfem_get_coupling_nnz(self.h, C.byref(nnz)))
        colptr = np.zeros(self.N + 1, dtype=np.int64)
        rowval = np.zeros(nnz.value, dtype=np.int64)
        nzval = np.zeros(nnz.value)
        self._ck(self.lib.asgfem_get_coupling_csc(self.h, _ptr(colptr), _ptr(rowval), _ptr(nzval)))
        return colptr, rowval, nzval

    def neighbours(self):
        plus = np.zeros((self.N, self.Mlen), dtype=np.int64)
        minus = np.zeros((self.N, self.Mlen), dtype=np.int64)
        self._ck(self.lib.asgfem_get_neighbours(self.h, _ptr(plus), _ptr(minus)))
        return plus.T.copy(), minus.T.copy()  # M x N like the Julia matrices

    # ---- matrices --------------------------------------------------------------------------------
    def set_pattern_csc(self, n, colptr, rowval):
        self.n = int(n)
        colptr, rowval = _i64(colptr), _i64(rowval)
        self._ck(self.lib.asgfem_set_pattern_csc(self.h, n, _ptr(colptr), _ptr(rowval)))

    def set_num_stiffness(self, M):
        self._ck(self.lib.asgfem_set_num_stiffness(self.h, M))

    def set_stiffness(self, m, nzval):
        nzval = _f64(nzval)
        self._ck(self.lib.asgfem_set_stiffness(self.h, m, _ptr(nzval)))

    def set_stiffness_csc(self, m, colptr, rowval, nzval):
        colptr, rowval, nzval = _i64(colptr), _i64(rowval), _f64(nzval)
        self._ck(self.lib.asgfem_set_stiffness_csc(self.h, m, _ptr(colptr), _ptr(rowval), _ptr(nzval)))

    def pattern_csc(self):
        nnz = C.c_int64()
        self._ck(self.lib.asgfem_get_pattern_nnz(self.h, C.byref(nnz)))
        colptr = np.zeros(self.n + 1, dtype=np.int64)
        rowval = np.zeros(nnz.value, dtype=np.int64)
        self._ck(self.lib.asgfem_get_pattern_csc(self.h, _ptr(colptr), _ptr(rowval)))
        return colptr, rowval

    def get_stiffness(self, m):
        nnz = C.c_int64()
        self._ck(self.lib.asgfem_get_pattern_nnz(self.h, C.byref(nnz)))
        v = np.zeros(nnz.value)
        self._ck(self.lib.asgfem_get_stiffness(self.h, m, _ptr(v)))
        return v

    def set_bdofs(self, bdofs_1based):
        b = _i64(bdofs_1based)
        self._ck(self.lib.asgfem_set_bdofs(self.h, len(b), _ptr(b)))

    # ---- mesh / space / coefficient ----------------------------------------------------------------
    def set_mesh(self, coords_2xn, cellnodes_3xnc_1based):
        """coords: (nnodes, 2) C-order == 2 x nnodes column-major; cellnodes: (ncells, 3), 1-based Int32."""
        co = _f64(coords_2xn)
        cn = _i32(cellnodes_3xnc_1based)
        self._ck(self.lib.asgfem_set_mesh(self.h, co.shape[0], cn.shape[0], _ptr(co), _ptr(cn)))

    def set_space(self, order, ndofs, celldofs_1based):
        cd = _i32(celldofs_1based)  # (ncells, ndofs4cell) C-order == ndofs4cell x ncells column-major
        self._ck(self.lib.asgfem_set_space(self.h, order, ndofs, cd.shape[1], _ptr(cd)))
        self.n = int(ndofs)

    def set_coefficient_cosinus(self, mean, decay_factors, b1, b2):
        d, b1, b2 = _f64(decay_factors), _i64(b1), _i64(b2)
        self._ck(self.lib.asgfem_set_coefficient_cosinus(self.h, len(d), float(mean), _ptr(d), _ptr(b1), _ptr(b2)))

    def assemble_stiffness(self, M, xref, w):
        xref, w = _f64(xref), _f64(w)  # xref (nq, 2) C-order == 2 x nq column-major
        self._ck(self.lib.asgfem_assemble_stiffness(self.h, M, len(w), _ptr(xref), _ptr(w)))

    # ---- vectors ---------------------------------------------------------------------------------
    def vec_alloc(self, nslots):
        """Sets the number of device vector slots; slots with index >= nslots are FREED (asgfem_vec_alloc)."""
        self._ck(self.lib.asgfem_vec_alloc(self.h, nslots))
        self._nslots = nslots

    def vec_ensure(self, nslots):
        """Grow-only variant for helpers that need scratch slots: existing vectors of the caller survive."""
        if getattr(self, "_nslots", 0) < nslots:
            self.vec_alloc(nslots)

    def vec_upload(self, slot, host):
        host = _f64(host)
        assert host.size == self.n * self.N, "host vector must have n*N entries (reference layout)"
        self._ck(self.lib.asgfem_vec_upload(self.h, slot, _ptr(host)))

    def vec_download(self, slot, out=None):
        out = np.empty(self.n * self.N) if out is None else out
        self._ck(self.lib.asgfem_vec_download(self.h, slot, _ptr(out)))
        return out

    def vec_zero(self, slot):
        self._ck(self.lib.asgfem_vec_zero(self.h, slot))

    def vec_fill_random(self, slot, seed=20240):
        self._ck(self.lib.asgfem_vec_fill_random(self.h, slot, seed))

    def vec_dot(self, a, b):
        out = C.c_double()
        self._ck(self.lib.asgfem_vec_dot(self.h, a, b, C.byref(out)))
        return out.value

    def vec_dot_owned(self, a, b):
        out = C.c_double()
        self._ck(self.lib.asgfem_vec_dot_owned(self.h, a, b, C.byref(out)))
        return out.value

    def precond_setup_global(self, n_global, colptr, rowval, nzval, bdofs1, row_offsets, coords=None):
        """Global K_0 (CSC, 1-based, rank-major global numbering) for the exact mean preconditioner of a sharded solve.
        Collective: rank 0 factorises and broadcasts the sweep tasks, so the matrix (and coords) may be None elsewhere."""
        cp = _i64(colptr) if colptr is not None else None
        rv = _i64(rowval) if rowval is not None else None
        nz = _f64(nzval) if nzval is not None else None
        bd, ro = _i64(bdofs1), _i64(row_offsets)
        xy = _f64(coords) if coords is not None else None
        opt = lambda a: _ptr(a) if a is not None else None  # noqa: E731
        self._ck(self.lib.asgfem_precond_setup_global(self.h, int(n_global), opt(cp), opt(rv), opt(nz), len(bd), _ptr(bd),
                                                      opt(xy), _ptr(ro)))

    def vec_dot_global(self, a, b):
        out = C.c_double()
        self._ck(self.lib.asgfem_vec_dot_global(self.h, a, b, C.byref(out)))
        return out.value

    # ---- NCCL inside the library (one process per GPU) ----------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id (rank 0 creates it, the host layer broadcasts it)."""
        lib = _lib.load()
        buf = (C.c_char * 128)()
        rc = lib.asgfem_comm_unique_id(buf)
        if rc != 0:
            raise _lib.AsgfemError(rc, lib.asgfem_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, nranks, rank, uid):
        self._ck(self.lib.asgfem_comm_init(self.h, nranks, rank, C.c_char_p(uid)))

    def comm_destroy(self):
        self._ck(self.lib.asgfem_comm_destroy(self.h))

    def set_halo(self, send, recv, interior0, interior1):
        """send / recv: dict neighbour rank -> 0-based local row ids (owned rows to send, halo rows to fill)."""
        ranks = sorted(set(send) | set(recv))
        sp, rp, srows, rrows = [0], [0], [], []
        for q in ranks:
            srows.extend(int(r) + 1 for r in send.get(q, []))
            rrows.extend(int(r) + 1 for r in recv.get(q, []))
            sp.append(len(srows))
            rp.append(len(rrows))
        a = lambda v, t: np.ascontiguousarray(np.array(v, dtype=t))  # noqa: E731
        rk, spa, rpa, sra, rra = a(ranks, np.int32), a(sp, np.int64), a(rp, np.int64), a(srows, np.int64), a(rrows, np.int64)
        self._ck(self.lib.asgfem_set_halo(self.h, len(ranks), rk.ctypes.data, spa.ctypes.data, sra.ctypes.data, rpa.ctypes.data,
                                          rra.ctypes.data, int(interior0), int(interior1)))

    def vec_axpy(self, alpha, x, y):
        self._ck(self.lib.asgfem_vec_axpy(self.h, alpha, x, y))

    def vec_xpay(self, x, beta, y):
        self._ck(self.lib.asgfem_vec_xpay(self.h, x, beta, y))

    def vec_copy(self, src, dst):
        self._ck(self.lib.asgfem_vec_copy(self.h, src, dst))

    # ---- operator / preconditioner / solver ----------------------------------------------------------
    def set_apply_variant(self, v):
        self._ck(self.lib.asgfem_set_apply_variant(self.h, v))

    def apply(self, sx, sy):
        self._ck(self.lib.asgfem_apply(self.h, sx, sy))

    def apply_rows(self, sx, sy, row0, row1):
        """Operator on the local rows [row0, row1) (0-based); see asgfem_apply_rows."""
        self._ck(self.lib.asgfem_apply_rows(self.h, sx, sy, row0, row1))

    def last_apply_ms(self):
        out = C.c_double()
        self._ck(self.lib.asgfem_last_apply_ms(self.h, C.byref(out)))
        return out.value

    def last_estimate_ms(self):
        out = C.c_double()
        self._ck(self.lib.asgfem_last_estimate_ms(self.h, C.byref(out)))
        return out.value

    def apply_host(self, x, out=None):
        x = _f64(x)
        out = np.empty_like(x) if out is None else out
        self._ck(self.lib.asgfem_apply_host(self.h, _ptr(x), _ptr(out)))
        return out

    def apply_host_ptr(self, x_ptr, y_ptr):
        """Raw host pointers (e.g. pinned torch tensors) - used by bench.py's end-to-end leg."""
        self._ck(self.lib.asgfem_apply_host(self.h, C.c_void_p(x_ptr), C.c_void_p(y_ptr)))

    def precond_setup(self):
        self._ck(self.lib.asgfem_precond_setup(self.h))

    def precond_apply(self, sr, sz):
        self._ck(self.lib.asgfem_precond_apply(self.h, sr, sz))

    def precond_apply_host(self, b, out=None):
        b = _f64(b)
        out = np.empty_like(b) if out is None else out
        self._ck(self.lib.asgfem_precond_apply_host(self.h, _ptr(b), _ptr(out)))
        return out

    def pcg(self, b0, slot_x, atol=1e-14, rtol=1e-14, itmax=0):
        b0 = _f64(b0)
        st = _lib.Stats()
        self._ck(self.lib.asgfem_pcg(self.h, _ptr(b0), slot_x, atol, rtol, itmax, C.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_ if k != "_pad"}

    def assemble_logprimal(self, M, xref, w):
        xref, w = _f64(xref), _f64(w)
        self._ck(self.lib.asgfem_assemble_logprimal(self.h, M, len(w), _ptr(xref), _ptr(w)))

    def assemble_logprimal_rhs(self, xref, w, f_at_qp, ntrunc, slot_b):
        xref, w, fq = _f64(xref), _f64(w), _f64(f_at_qp)
        self._ck(self.lib.asgfem_assemble_logprimal_rhs(self.h, len(w), _ptr(xref), _ptr(w), _ptr(fq), int(ntrunc), slot_b))

    def set_samples(self, samples):
        """samples: (Msamples, nsamples) - the columns of the device vectors become the samples (asgfem_set_samples)."""
        S = np.asfortranarray(np.asarray(samples, dtype=np.float64))
        Ms, ns = S.shape
        self._ck(self.lib.asgfem_set_samples(self.h, ns, Ms, S.ctypes.data))
        self.nsamples = ns

    def solve_samples_host(self, b, atol=1e-14, rtol=1e-14, itmax=0):
        """Deterministic solutions of all samples: returns (u, stats), u of shape (n, nsamples)."""
        b = _f64(b)
        out = np.zeros((self.nsamples, len(b)))  # C-order (nsamples, n) == n x nsamples column-major
        st = _lib.Stats()
        self._ck(self.lib.asgfem_solve_samples_host(self.h, _ptr(out), _ptr(b), atol, rtol, itmax, C.byref(st)))
        return out.T, {k: getattr(st, k) for k, _ in st._fields_ if k != "_pad"}

    def solve_primal_host(self, sol, b0, atol=1e-14, rtol=1e-14, itmax=0):
        assert sol.dtype == np.float64 and sol.flags.c_contiguous
        b0 = _f64(b0)
        st = _lib.Stats()
        self._ck(self.lib.asgfem_solve_primal_host(self.h, _ptr(sol), _ptr(b0), atol, rtol, itmax, C.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_ if k != "_pad"}

    # ---- evaluation at samples ----------------------------------------------------------------------
    def evaluate_samples(self, slot_u, vals):
        """vals[s, m, :] = TB.vals[m] after set_sample!(TB, xi_s); returns the (nsamples, n) array of the evaluated
        spatial coefficient vectors (row s = FEV entries for sample s)."""
        vals = _f64(vals)
        S, M, nvals = vals.shape
        out = np.zeros((S, self.n))
        self._ck(self.lib.asgfem_evaluate_samples(self.h, slot_u, S, M, nvals, _ptr(vals), _ptr(out)))
        return out

    # ---- log-transformed primal problem ------------------------------------------------------------
    def set_precond_matrix_csc(self, colptr=None, rowval=None, nzval=None):
        """SPD matrix for the mean preconditioner (default: matrix 0); None returns to the default."""
        if colptr is None:
            self._ck(self.lib.asgfem_set_precond_matrix_csc(self.h, None, None, None))
            return
        colptr, rowval, nzval = _i64(colptr), _i64(rowval), _f64(nzval)
        self._ck(self.lib.asgfem_set_precond_matrix_csc(self.h, _ptr(colptr), _ptr(rowval), _ptr(nzval)))

    def bicgstab(self, slot_b, slot_x, atol=1e-14, rtol=1e-14, itmax=0):
        st = _lib.Stats()
        self._ck(self.lib.asgfem_bicgstab(self.h, slot_b, slot_x, atol, rtol, itmax, C.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_ if k != "_pad"}

    def solve_logprimal_host(self, sol, b, atol=1e-14, rtol=1e-14, itmax=0):
        assert sol.dtype == np.float64 and sol.flags.c_contiguous
        b = _f64(b)
        assert b.size == sol.size
        st = _lib.Stats()
        self._ck(self.lib.asgfem_solve_logprimal_host(self.h, _ptr(sol), _ptr(b), atol, rtol, itmax, C.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_ if k != "_pad"}

    # ---- estimator -------------------------------------------------------------------------------
    def estimate_poisson_primal(self, slot_u, mi_ext, xref, w, sf, wf, ncells, f_at_qp=None):
        mi = _i64(np.asarray(mi_ext))
        N_ext, M_ext = mi.shape
        xref, w, sf, wf = _f64(xref), _f64(w), _f64(sf), _f64(wf)
        fq = None if f_at_qp is None else _f64(f_at_qp)
        eta4cell = np.zeros((N_ext, ncells))  # C-order (N_ext, ncells) == ncells x N_ext column-major
        eta4modes = np.zeros(N_ext)
        self._ck(self.lib.asgfem_estimate_poisson_primal(
            self.h, slot_u, N_ext, M_ext, _ptr(mi), len(w), _ptr(xref), _ptr(w), _ptr(fq), len(wf), _ptr(sf),
            _ptr(wf), _ptr(eta4cell), _ptr(eta4modes)))
        return eta4modes, eta4cell.T  # view as ncells x N_ext

    def estimate_logpoisson_primal(self, slot_u, mi_ext, xref, w, sf, wf, ncells, f_at_qp, ntrunc, lam_at_qp=None):
        mi = _i64(np.asarray(mi_ext))
        N_ext, M_ext = mi.shape
        xref, w, sf, wf, fq = _f64(xref), _f64(w), _f64(sf), _f64(wf), _f64(f_at_qp)
        lam = None if lam_at_qp is None else _f64(lam_at_qp)
        eta4cell = np.zeros((N_ext, ncells))
        eta4modes = np.zeros(N_ext)
        zeta = np.zeros(3)
        self._ck(self.lib.asgfem_estimate_logpoisson_primal(
            self.h, slot_u, N_ext, M_ext, _ptr(mi), len(w), _ptr(xref), _ptr(w), _ptr(fq), _ptr(lam) if lam is not None else None,
            int(ntrunc), len(wf), _ptr(sf), _ptr(wf), _ptr(eta4cell), _ptr(eta4modes), _ptr(zeta)))
        return eta4modes, eta4cell.T, zeta

    def estimate_poisson_primal_marking(self, slot_u, mi_ext, xref, w, sf, wf, ncells, sel_cols_1based, f_at_qp=None):
        """eta4modes and the per-cell sum of eta4cell over the selected columns (no ncells x N_ext transfer)."""
        mi = _i64(np.asarray(mi_ext))
        N_ext, M_ext = mi.shape
        xref, w, sf, wf = _f64(xref), _f64(w), _f64(sf), _f64(wf)
        fq = None if f_at_qp is None else _f64(f_at_qp)
        sel = _i64(np.asarray(sel_cols_1based))
        cellsum = np.zeros(ncells)
        eta4modes = np.zeros(N_ext)
        self._ck(self.lib.asgfem_estimate_poisson_primal_marking(
            self.h, slot_u, N_ext, M_ext, _ptr(mi), len(w), _ptr(xref), _ptr(w), _ptr(fq), len(wf), _ptr(sf),
            _ptr(wf), len(sel), _ptr(sel), _ptr(cellsum), _ptr(eta4modes)))
        return eta4modes, cellsum

    # ---- multi-GPU helpers -----------------------------------------------------------------------
    def set_owned_rows(self, n_owned):
        self._ck(self.lib.asgfem_set_owned_rows(self.h, n_owned))

    def set_owned_cells(self, owned):
        """Row-sharded estimator: flags (one per cell of the rank's mesh) of the cells this rank owns; None = all."""
        if owned is None:
            self._ck(self.lib.asgfem_set_owned_cells(self.h, 0, None))
            return
        f = np.ascontiguousarray(np.asarray(owned, dtype=np.uint8))
        self._ck(self.lib.asgfem_set_owned_cells(self.h, len(f), f.ctypes.data))

    def halo_exchange(self, slot):
        self._ck(self.lib.asgfem_halo_exchange(self.h, slot))

    def vec_device_ptr(self, slot):
        p, ld = C.c_void_p(), C.c_int64()
        self._ck(self.lib.asgfem_vec_device_ptr(self.h, slot, C.byref(p), C.byref(ld)))
        return p.value, ld.value

    def pack_rows(self, slot, rows_1based, dbuf_ptr):
        r = _i64(rows_1based)
        self._ck(self.lib.asgfem_pack_rows(self.h, slot, len(r), _ptr(r), C.c_void_p(dbuf_ptr)))

    def unpack_rows(self, slot, rows_1based, dbuf_ptr):
        r = _i64(rows_1based)
        self._ck(self.lib.asgfem_unpack_rows(self.h, slot, len(r), _ptr(r), C.c_void_p(dbuf_ptr)))


# ---- context-free index helpers ------------------------------------------------------------------
def coupling_weights(family, maxdeg):
    lib = _lib.load()
    gp = np.zeros(maxdeg + 1)
    gm = np.zeros(maxdeg + 1)
    rc = lib.asgfem_coupling_weights(family, maxdeg, _ptr(gp), _ptr(gm))
    if rc:
        raise _lib.AsgfemError(rc, "asgfem_coupling_weights")
    return gp, gm


def host_factor_solve(indptr, indices, data, bdofs_1based, b, coords_2xn=None):
    """Host-only diagnostic of the factorisation behind Context.precond_setup (asgfem_host_factor_solve): K (CSR, 0-based,
    symmetric) without the rows/columns bdofs, multifrontal Cholesky on the host cores, one solve.  Returns (x, lnz)."""
    lib = _lib.load()
    ip, ix, dv = _i64(indptr), _i32(indices), _f64(data)
    n = len(ip) - 1
    mask = np.zeros(n, dtype=np.uint8)
    mask[np.asarray(bdofs_1based, dtype=np.int64) - 1] = 1
    bb = _f64(b)
    assert bb.shape == (n,)
    xy = None if coords_2xn is None else _f64(np.asarray(coords_2xn, dtype=np.float64).T.copy())  # (n, 2): x0 y0 x1 y1 ...
    x = np.zeros(n)
    lnz = C.c_int64(0)
    err = C.create_string_buffer(256)
    rc = lib.asgfem_host_factor_solve(n, _ptr(ip), _ptr(ix), _ptr(dv), _ptr(mask), None if xy is None else _ptr(xy), _ptr(bb),
                                      _ptr(x), C.addressof(lnz), C.addressof(err), 256)
    if rc:
        raise _lib.AsgfemError(rc, err.value.decode() or "asgfem_host_factor_solve")
    return x, lnz.value


def add_boundary_modes(multi_indices, p_extension=1, tail_extension=(10, 2)):
    lib = _lib.load()
    mi = _i64(np.asarray(multi_indices))
    N, M = mi.shape
    Next, Mext = C.c_int64(), C.c_int64()
    rc = lib.asgfem_add_boundary_modes(N, M, _ptr(mi), p_extension, tail_extension[0], tail_extension[1],
                                       C.byref(Next), C.byref(Mext), None, 0)
    if rc:
        raise _lib.AsgfemError(rc, "asgfem_add_boundary_modes")
    out = np.zeros((Next.value, Mext.value), dtype=np.int64)
    rc = lib.asgfem_add_boundary_modes(N, M, _ptr(mi), p_extension, tail_extension[0], tail_extension[1],
                                       C.byref(Next), C.byref(Mext), _ptr(out), out.size)
    if rc:
        raise _lib.AsgfemError(rc, "asgfem_add_boundary_modes")
    return out


def classify_modes(mi_ext, n_active):
    """Returns (inactive_else, inactive_bnd, inactive_bnd2, active_bnd, active_int) as 1-based index lists."""
    lib = _lib.load()
    mi = _i64(np.asarray(mi_ext))
    cls = np.zeros(mi.shape[0], dtype=np.int32)
    rc = lib.asgfem_classify_modes(mi.shape[0], mi.shape[1], _ptr(mi), n_active, _ptr(cls))
    if rc:
        raise _lib.AsgfemError(rc, "asgfem_classify_modes")
    return tuple((np.where(cls == c)[0] + 1).tolist() for c in range(5))
