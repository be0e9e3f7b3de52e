"""(1) The "libstr" twins on BLASFEO containers (include/hpmpc_blasfeo_compat.h) and (2) the high-level
re-solve pair {c,fortran}_order_d_solve_kkt_new_rhs_ocp_hard_tv (include/c_interface.h:63,67).  GPU tests through the C ABI.

libstr conventions followed (the reference's libstr flavour cannot be built here -- BLASFEO is absent -- so these are checked
against the oracle / the lib4 twins, which are pinned): panel-major bs = 4 with cn = cols rounded up to 4, hspi[n+1] / hsPb[n+1]
node-indexed (lqcp_solvers/d_back_ric_rec_libstr.c:145,196), d / lam / t = [lb lg ub ug] unpadded
(interfaces/c/fortran_order_interface_libstr.c:408-415,751-755)."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_err
from hpmpc_b200 import capi, problems
from hpmpc_b200.capi import aligned_zeros, int_array, ptr_array
from oracle import api as oracle

pytestmark = pytest.mark.gpu
TOL = 1e-9


class DMat(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("pm", C.c_int), ("cn", C.c_int), ("pA", C.c_void_p), ("dA", C.c_void_p), ("use_dA", C.c_int), ("memsize", C.c_int)]


class DVec(C.Structure):
    _fields_ = [("m", C.c_int), ("pm", C.c_int), ("pa", C.c_void_p), ("memsize", C.c_int)]


def _rup(x, m):
    return (x + m - 1) // m * m


def dmat(M, keep):
    rows, cols = M.shape
    pm, cn = _rup(max(rows, 1), 4), _rup(max(cols, 1), 4)
    buf = aligned_zeros(pm * cn + 8)
    if rows and cols:
        I, J = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
        buf[(I // 4) * 4 * cn + I % 4 + 4 * J] = M
    dA = aligned_zeros(max(rows, cols, 1) + 4)
    keep += [buf, dA]
    return DMat(rows, cols, pm, cn, buf.ctypes.data, dA.ctypes.data, 0, 8 * (pm * cn + max(rows, cols, 1)))


def dvec(v, keep, size=None):
    n = len(v) if size is None else size
    buf = aligned_zeros(_rup(max(n, 1), 4) + 4)
    buf[:len(v)] = v
    keep.append(buf)
    return DVec(n, _rup(max(n, 1), 4), buf.ctypes.data, 8 * _rup(max(n, 1), 4)), buf


def _structs(p, keep):
    N = p.N
    ngl = p.ng_list()
    Cg, Dg, lgg, ugg = p.general_arrays()
    BAbt, RSQ, DCt = (DMat * (N + 1))(), (DMat * (N + 1))(), (DMat * (N + 1))()
    for n in range(N + 1):
        nu, nx = p.nu[n], p.nx[n]
        nux = nu + nx
        if n < N:
            BAbt[n] = dmat(np.vstack([p.B[n].T.reshape(nu, p.nx[n + 1]), p.A[n].T.reshape(nx, p.nx[n + 1]), p.b[n].reshape(1, -1)]), keep)
        H = np.zeros((nux + 1, nux))
        H[:nu, :nu] = p.R[n]; H[nu:nux, :nu] = p.S[n].T; H[:nu, nu:nux] = p.S[n]; H[nu:nux, nu:nux] = p.Q[n]; H[nux, :nu] = p.r[n]; H[nux, nu:] = p.q[n]
        RSQ[n] = dmat(H, keep)
        G = np.vstack([np.asarray(Dg[n]).reshape(ngl[n], nu).T, np.asarray(Cg[n]).reshape(ngl[n], nx).T]) if ngl[n] else np.zeros((max(nux, 1), 1))
        DCt[n] = dmat(G, keep)
    return BAbt, RSQ, DCt, ngl, lgg, ugg


@pytest.fixture(scope="module")
def lib():
    import torch
    assert torch.cuda.is_available()
    return C.CDLL(capi.PRODUCT_LIB)


@pytest.mark.parametrize("shape", [(8, 3, 10), (12, 5, 30)])
def test_libstr_riccati_sv_and_trf_trs(lib, shape):
    p = problems.mass_spring_ocp(*shape, xi=(0.3, 0.1, -0.2, 0.5))
    o = oracle.ric(p, "sv")
    N = p.N
    for mode in ("sv", "trf_trs"):
        keep = []
        BAbt, RSQ, DCt, ngl, _, _ = _structs(p, keep)
        vec = lambda lens: ((DVec * (N + 2))(), [])
        ux, pi, Pb, b, rq, Qx, qx = (DVec * (N + 2))(), (DVec * (N + 2))(), (DVec * (N + 2))(), (DVec * (N + 2))(), (DVec * (N + 2))(), (DVec * (N + 2))(), (DVec * (N + 2))()
        bufs = {}
        for n in range(N + 1):
            ux[n], bufs["ux", n] = dvec(np.zeros(p.nu[n] + p.nx[n]), keep)
            pi[n], bufs["pi", n] = dvec(np.zeros(p.nx[n]), keep); Pb[n], _ = dvec(np.zeros(p.nx[n]), keep)
            rq[n], _ = dvec(np.concatenate([p.r[n], p.q[n]]), keep); Qx[n], _ = dvec(np.zeros(1), keep); qx[n], _ = dvec(np.zeros(1), keep)
            if n < N:
                b[n], _ = dvec(p.b[n], keep)
        L = (DMat * (N + 1))()
        for n in range(N + 1):
            nux = p.nu[n] + p.nx[n]
            L[n] = dmat(np.zeros((nux + 1, max(nux, 1))), keep)
        nx, nu, nb, ng = int_array(p.nx), int_array(p.nu), int_array([0] * (N + 1)), int_array([0] * (N + 1))
        idxb = ptr_array([np.zeros(1, dtype=np.int32) for _ in range(N + 1)])
        if mode == "sv":
            lib.d_back_ric_rec_sv_libstr(N, nx, nu, nb, idxb, ng, 0, BAbt, b, 0, RSQ, rq, DCt, Qx, qx, ux, 1, pi, 1, Pb, L, None)
        else:
            lib.d_back_ric_rec_trf_libstr(N, nx, nu, nb, idxb, ng, BAbt, RSQ, DCt, Qx, L, None)
            lib.d_back_ric_rec_trs_libstr(N, nx, nu, nb, idxb, ng, BAbt, b, rq, DCt, qx, ux, 1, pi, 1, Pb, L, None)
        u = [bufs["ux", n][:p.nu[n]].copy() for n in range(N)]
        x = [bufs["ux", n][p.nu[n]:p.nu[n] + p.nx[n]].copy() for n in range(N + 1)]
        pis = [bufs["pi", n + 1][:p.nx[n + 1]].copy() for n in range(N)]          # node-indexed
        assert rel_err(u, o["u"]) < TOL and rel_err(x, o["x"]) < TOL and rel_err(pis, o["pi"]) < TOL, mode


@pytest.mark.parametrize("general", [False, True])
def test_libstr_ipm_lam_ordering(lib, general):
    p = problems.general_test_problem(8, 3, 10, xi=(0.1, 0.2, 0.3, -0.4)) if general else problems.mass_spring_ocp(8, 3, 10, bounds=True, xi=(0.1, 0.2, 0.3, -0.4))
    o = oracle.ipm(p, k_max=30)
    N = p.N
    keep = []
    BAbt, RSQ, DCt, ngl, lgg, ugg = _structs(p, keep)
    d, ux, pi, lam, t = (DVec * (N + 2))(), (DVec * (N + 2))(), (DVec * (N + 2))(), (DVec * (N + 2))(), (DVec * (N + 2))()
    bufs = {}
    for n in range(N + 1):
        dn = np.concatenate([p.lb[n], lgg[n], p.ub[n], ugg[n]]) if general else np.concatenate([p.lb[n], p.ub[n]])
        d[n], _ = dvec(dn, keep)
        ux[n], bufs["ux", n] = dvec(np.zeros(p.nu[n] + p.nx[n]), keep)
        pi[n], bufs["pi", n] = dvec(np.zeros(p.nx[n]), keep)
        lam[n], bufs["lam", n] = dvec(np.zeros(len(dn)), keep); t[n], bufs["t", n] = dvec(np.zeros(len(dn)), keep)
    idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
    stat = np.zeros(5 * 30 + 5); kk = C.c_int(0)
    fn = lib.d_ip2_res_mpc_hard_libstr
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 10 + [C.c_int] + [C.c_void_p] * 4
    status = fn(C.byref(kk), 30, 2.0, 1e-8, 1e-8, 0, stat.ctypes.data, N, int_array(p.nx), int_array(p.nu), int_array(p.nb), ptr_array(idxb), int_array(ngl),
                BAbt, RSQ, DCt, d, ux, 1, pi, lam, t, None)
    assert (kk.value, status) == (o["kk"], o["status"])
    u = [bufs["ux", n][:p.nu[n]].copy() for n in range(N)]
    x = [bufs["ux", n][p.nu[n]:p.nu[n] + p.nx[n]].copy() for n in range(N + 1)]
    pis = [bufs["pi", n + 1][:p.nx[n + 1]].copy() for n in range(N)]
    assert rel_err(u, o["u"]) < TOL and rel_err(x, o["x"]) < TOL and rel_err(pis, o["pi"]) < TOL
    for n in range(N + 1):
        nb, ng = p.nb[n], ngl[n]
        v = bufs["lam", n][:2 * nb + 2 * ng]
        mine = np.concatenate([v[:nb], v[nb + ng:2 * nb + ng], v[nb:nb + ng], v[2 * nb + ng:]])      # [lb lg ub ug] -> [lb ub lg ug]
        assert rel_err([mine], [o["lam"][n]]) < TOL, n


@pytest.mark.parametrize("order", ["fortran", "c"])
def test_high_level_ip_then_solve_kkt_new_rhs(order):
    """{c,fortran}_order_d_ip_ocp_hard_tv followed by {c,fortran}_order_d_solve_kkt_new_rhs_ocp_hard_tv on the same work0, against the
    oracle's restatement of the low-level pair (pinned on the reference: tests/test_kkt_new_rhs.py)."""
    prod = capi.HpmpcLib(capi.PRODUCT_LIB)
    p = problems.mass_spring_ocp(8, 3, 10, bounds=True, xi=(0.3, -0.1, 0.2, 0.1))
    p2 = problems.mass_spring_ocp(8, 3, 10, bounds=True, xi=(0.35, -0.05, 0.2, 0.1))
    for n in range(p.N + 1):                     # same matrices, new vectors
        p2.Q[n], p2.R[n], p2.S[n] = p.Q[n], p.R[n], p.S[n]
        p2.q[n] = p.q[n] * 1.1; p2.r[n] = p.r[n] * 0.9
    o = oracle.ipm_then_kkt_new_rhs(p, p2)
    r = prod.ip_then_solve_kkt_new_rhs_high_level(p, p2, order=order)
    assert r["kk"] == o["kk"]
    for f in ("u", "x", "pi"):
        assert rel_err(r[f], o[f]) < TOL, f
    assert rel_err(r["lam"], o["lam"]) < 1e-6          # as in tests/test_kkt_new_rhs.py (t_inv ~ 1e8 amplifies rounding in dux)
    assert np.all(np.isfinite(r["inf_norm_res"]))
