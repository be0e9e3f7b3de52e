"""ctypes access to libraries that export HPMPC's C symbols, and to the batched hpmpc_b200 C ABI.

The same `HpmpcLib` class drives
  * the product  : hpmpc_b200/lib/libhpmpc_b200.so (CUDA; raises if it is missing -- there is no CPU fallback),
  * the reference: oracle/_ref/libhpmpc_ref_{c99,avx2}.so (test infrastructure only),
so a parity test calls the *same function name with the same arguments* on both.
This module is harness code: it moves pointers around and never computes a solution.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence

import numpy as np

from .problems import Ocp

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.environ.get("HPMPC_B200_LIB") or os.path.join(_HERE, "lib", "libhpmpc_b200.so")   # override: A/B builds in tools/

c_dpp = C.POINTER(C.c_void_p)
BS, NCL = 4, 2


def _rup(x, m):
    return (x + m - 1) // m * m


def aligned_zeros(n: int, align: int = 64) -> np.ndarray:
    raw = np.zeros(n + align // 8 + 1, dtype=np.float64)
    off = (-raw.ctypes.data % align) // 8
    return raw[off:off + n]


def ptr_array(arrs: Sequence[np.ndarray]):
    """double** (or int**) from a list of numpy arrays; the list must outlive the call."""
    n = max(len(arrs), 1)
    out = (C.c_void_p * n)()
    for i, a in enumerate(arrs):
        out[i] = a.ctypes.data if a is not None else None
    return out


def int_array(v: Sequence[int]):
    return (C.c_int * max(len(v), 1))(*[int(x) for x in v])


def to_pmat(M: np.ndarray) -> np.ndarray:
    """Dense matrix -> the reference's lib4 panel-major storage (auxiliary/d_aux_lib4.c:1310), 64-byte aligned."""
    rows, cols = M.shape
    pr, sda = _rup(max(rows, 1), BS), _rup(max(cols, 1), NCL)
    p = aligned_zeros(pr * sda + 8)
    if rows and cols:
        I, J = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
        p[(I // BS) * BS * sda + I % BS + BS * J] = M
    return p


class HpmpcLib:
    """A shared library exporting HPMPC's hot-path symbols (reference build or libhpmpc_b200.so)."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not found -- build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path = path
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        L = self.lib
        for name in ("fortran_order_d_ip_ocp_hard_tv", "c_order_d_ip_ocp_hard_tv"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_int] + [C.c_void_p] * 5 \
                + [C.c_int, C.c_int] + [C.c_void_p] * 18 + [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hpmpc_d_ip_ocp_hard_tv_work_space_size_bytes.restype = C.c_int
        L.hpmpc_d_ip_ocp_hard_tv_work_space_size_bytes.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int]
        for name in ("d_back_ric_rec_sv_tv_work_space_size_bytes", "d_back_ric_rec_sv_tv_memory_space_size_bytes",
                     "d_ip2_res_mpc_hard_tv_work_space_size_bytes"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_int] + [C.c_void_p] * 4
        L.d_back_ric_rec_sv_tv_res.restype = None
        L.d_back_ric_rec_sv_tv_res.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int] \
            + [C.c_void_p] * 7 + [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.d_back_ric_rec_trf_tv_res.restype = None
        L.d_back_ric_rec_trf_tv_res.argtypes = [C.c_int] + [C.c_void_p] * 12
        L.d_back_ric_rec_trs_tv_res.restype = None
        L.d_back_ric_rec_trs_tv_res.argtypes = [C.c_int] + [C.c_void_p] * 11 + [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.d_ip2_res_mpc_hard_tv.restype = C.c_int
        L.d_ip2_res_mpc_hard_tv.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p,
                                            C.c_int] + [C.c_void_p] * 10 + [C.c_int] + [C.c_void_p] * 4

    # ---------------------------------------------------------------- high level
    def ip_ocp_hard_tv(self, p: Ocp, *, order: str = "fortran", k_max: int = 40, mu0: float = 2.0, mu_tol: float = 1e-8,
                       warm_start: int = 0, x_init=None, u_init=None, N2: int = None):
        """{c,fortran}_order_d_ip_ocp_hard_tv (reference include/c_interface.h:62,65); N2 < N: partial condensing into N2 blocks."""
        N = p.N
        N2 = N if N2 is None else N2
        conv = (lambda M: np.ascontiguousarray(M)) if order == "c" else (lambda M: np.asfortranarray(M))
        A = [conv(M) for M in p.A]; B = [conv(M) for M in p.B]
        Q = [conv(M) for M in p.Q]; S = [conv(M) for M in p.S]; R = [conv(M) for M in p.R]
        # numpy keeps 0-sized arrays valid; HPMPC never dereferences them
        b = [np.ascontiguousarray(v) for v in p.b]; q = [np.ascontiguousarray(v) for v in p.q]; r = [np.ascontiguousarray(v) for v in p.r]
        lb = [np.ascontiguousarray(v) for v in p.lb]; ub = [np.ascontiguousarray(v) for v in p.ub]
        x = [np.zeros(max(n, 1)) for n in p.nx]; u = [np.zeros(max(n, 1)) for n in p.nu[:N]]
        if x_init is not None:
            for n in range(N + 1): x[n][:p.nx[n]] = x_init[n]
            for n in range(N): u[n][:p.nu[n]] = u_init[n]
        pi = [np.zeros(max(p.nx[n + 1], 1)) for n in range(N)]
        lam = [np.zeros(max(2 * nb, 1)) for nb in p.nb]
        idxb = [np.ascontiguousarray(v, dtype=np.int32) for v in p.idxb]
        ngl = p.ng_list()
        nx, nu, nb, ng = int_array(p.nx), int_array(p.nu), int_array(p.nb), int_array(ngl)
        pad = lambda M: M if M.size else np.zeros(1)
        Cg, Dg, lgg, ugg = p.general_arrays()
        Cg = [pad(conv(M)) for M in Cg]; Dg = [pad(conv(M)) for M in Dg]
        lgg = [pad(np.ascontiguousarray(v)) for v in lgg]; ugg = [pad(np.ascontiguousarray(v)) for v in ugg]
        lam = [np.zeros(max(2 * p.nb[n] + 2 * ngl[n], 1)) for n in range(N + 1)]
        wsz = self.lib.hpmpc_d_ip_ocp_hard_tv_work_space_size_bytes(N, nx, nu, nb, ptr_array(idxb), ng, N2)
        work = aligned_zeros(wsz // 8 + 16)
        res = np.zeros(8); stat = np.zeros(5 * k_max + 5)
        kk = C.c_int(0)
        keep = [A, B, b, Q, S, R, q, r, lb, ub, x, u, pi, lam, idxb, Cg, Dg, lgg, ugg]
        fn = self.lib.c_order_d_ip_ocp_hard_tv if order == "c" else self.lib.fortran_order_d_ip_ocp_hard_tv
        pa = ptr_array
        arrs = [pa(A), pa(B), pa(b), pa(Q), pa(S), pa(R), pa(q), pa(r), pa(lb), pa(ub), pa(Cg), pa(Dg), pa(lgg), pa(ugg),
                pa(x), pa(u), pa(pi), pa(lam)]
        pidx = pa(idxb)
        status = fn(C.byref(kk), k_max, mu0, mu_tol, N, nx, nu, nb, pidx, ng, N2, warm_start, *arrs,
                    res.ctypes.data, work.ctypes.data, stat.ctypes.data)
        del keep
        return dict(status=status, kk=kk.value, x=[x[n][:p.nx[n]].copy() for n in range(N + 1)],
                    u=[u[n][:p.nu[n]].copy() for n in range(N)], pi=[pi[n][:p.nx[n + 1]].copy() for n in range(N)],
                    lam=[lam[n][:2 * p.nb[n] + 2 * ngl[n]].copy() for n in range(N + 1)], inf_norm_res=res[:4].copy(),
                    stat=stat[:5 * kk.value].reshape(-1, 5).copy())

    # ---------------------------------------------------------------- low level (panel-major)
    def _pm_problem(self, p: Ocp):
        N = p.N
        BAbt, RSQ = [], []
        for n in range(N + 1):
            nx, nu = p.nx[n], p.nu[n]
            nux = nx + nu
            if n < N:
                nx1 = p.nx[n + 1]
                M = np.vstack([p.B[n].T.reshape(nu, nx1), p.A[n].T.reshape(nx, nx1), p.b[n].reshape(1, nx1)])
                BAbt.append(to_pmat(M))
            H = np.zeros((nux + 1, nux))
            H[:nu, :nu] = p.R[n]; H[nu:nux, :nu] = p.S[n].T; H[:nu, nu:nux] = p.S[n]; H[nu:nux, nu:nux] = p.Q[n]
            H[nux, :nu] = p.r[n]; H[nux, nu:] = p.q[n]
            RSQ.append(to_pmat(H))
        return BAbt, RSQ

    def _sizes(self, p: Ocp):
        N = p.N
        nx, nu, nb, ng = int_array(p.nx), int_array(p.nu), int_array(p.nb), int_array([0] * (N + 1))
        return nx, nu, nb, ng

    def ric(self, p: Ocp, mode: str = "sv"):
        """d_back_ric_rec_sv_tv_res, or trf followed by trs (reference include/lqcp_solvers.h:41-45); unconstrained."""
        N = p.N
        nx, nu, nb0, ng = self._sizes(p)
        nb = int_array([0] * (N + 1))
        BAbt, RSQ = self._pm_problem(p)
        wsz = self.lib.d_back_ric_rec_sv_tv_work_space_size_bytes(N, nx, nu, nb, ng)
        msz = self.lib.d_back_ric_rec_sv_tv_memory_space_size_bytes(N, nx, nu, nb, ng)
        work, mem = aligned_zeros(wsz // 8 + 16), aligned_zeros(msz // 8 + 16)
        hux = [aligned_zeros(_rup(p.nx[n] + p.nu[n] + 1, BS) + 4) for n in range(N + 1)]
        hpi = [aligned_zeros(_rup(p.nx[n + 1], BS) + 4) for n in range(N)]
        hPb = [aligned_zeros(_rup(p.nx[n + 1], BS) + 4) for n in range(N)]
        idxb = [np.zeros(1, dtype=np.int32) for _ in range(N + 1)]
        dummy = [aligned_zeros(8) for _ in range(N + 1)]
        pa = ptr_array
        pBAbt, pRSQ, pux, ppi, pPb, pidx, pd = pa(BAbt), pa(RSQ), pa(hux), pa(hpi), pa(hPb), pa(idxb), pa(dummy)
        if mode == "sv":
            self.lib.d_back_ric_rec_sv_tv_res(N, nx, nu, nb, pidx, ng, 0, pBAbt, pd, 0, pRSQ, pd, pd, pd, pd, pd,
                                              pux, 1, ppi, 1, pPb, mem.ctypes.data, work.ctypes.data)
        else:
            hb = [aligned_zeros(_rup(p.nx[n + 1], BS) + 4) for n in range(N)]
            hq = [aligned_zeros(_rup(p.nx[n] + p.nu[n] + 1, BS) + 4) for n in range(N + 1)]
            for n in range(N):
                hb[n][:p.nx[n + 1]] = p.b[n]
            for n in range(N + 1):
                hq[n][:p.nu[n]] = p.r[n]; hq[n][p.nu[n]:p.nu[n] + p.nx[n]] = p.q[n]
            phb, phq = pa(hb), pa(hq)
            self.lib.d_back_ric_rec_trf_tv_res(N, nx, nu, nb, pidx, ng, pBAbt, pRSQ, pd, pd, pd, mem.ctypes.data, work.ctypes.data)
            self.lib.d_back_ric_rec_trs_tv_res(N, nx, nu, nb, pidx, ng, pBAbt, phb, phq, pd, pd, pux, 1, ppi, 1, pPb,
                                               mem.ctypes.data, work.ctypes.data)
        return dict(u=[hux[n][:p.nu[n]].copy() for n in range(N)],
                    x=[hux[n][p.nu[n]:p.nu[n] + p.nx[n]].copy() for n in range(N + 1)],
                    pi=[hpi[n][:p.nx[n + 1]].copy() for n in range(N)],
                    Pb=[hPb[n][:p.nx[n + 1]].copy() for n in range(N)])

    def _pm_general(self, p: Ocp):
        """hpDCt[n] = [D C]'_n panel-major (nux x ng), and the bound-like vector layout [lb(pnb) ub(pnb) lg(png) ug(png)]
        (interfaces/c/fortran_order_interface.c:276-283, :345-378)."""
        N = p.N
        ngl = p.ng_list()
        Cg, Dg, lgg, ugg = p.general_arrays()
        DCt = []
        for n in range(N + 1):
            M = np.vstack([np.asarray(Dg[n]).reshape(ngl[n], p.nu[n]).T, np.asarray(Cg[n]).reshape(ngl[n], p.nx[n]).T]) if ngl[n] else np.zeros((1, 1))
            DCt.append(to_pmat(M))
        pnb = [_rup(v, BS) for v in p.nb]; png = [_rup(v, BS) for v in ngl]
        d = [aligned_zeros(2 * pnb[n] + 2 * png[n] + 4) for n in range(N + 1)]
        for n in range(N + 1):
            d[n][:p.nb[n]] = p.lb[n]; d[n][pnb[n]:pnb[n] + p.nb[n]] = p.ub[n]
            if ngl[n]:
                d[n][2 * pnb[n]:2 * pnb[n] + ngl[n]] = lgg[n]; d[n][2 * pnb[n] + png[n]:2 * pnb[n] + png[n] + ngl[n]] = ugg[n]
        return DCt, d, pnb, png, ngl

    @staticmethod
    def _split_bound_like(v, nb, pnb, ng, png):
        """[lb(pnb) ub(pnb) lg(png) ug(png)] -> [lb(nb) ub(nb) lg(ng) ug(ng)] (the high-level ordering, c_order_interface.c:662-681)."""
        return np.concatenate([v[:nb], v[pnb:pnb + nb], v[2 * pnb:2 * pnb + ng], v[2 * pnb + png:2 * pnb + png + ng]])

    def ip2_res_mpc_hard_tv(self, p: Ocp, *, k_max=40, mu0=2.0, mu_tol=1e-8, alpha_min=1e-8, warm_start=0):
        """d_ip2_res_mpc_hard_tv on panel-major data (reference include/mpc_solvers.h:42), general constraints included."""
        N = p.N
        nx, nu, nb, _ = self._sizes(p)
        BAbt, RSQ = self._pm_problem(p)
        DCt, d, pnb, png, ngl = self._pm_general(p)
        ng = int_array(ngl)
        ux = [aligned_zeros(_rup(p.nx[n] + p.nu[n] + 1, BS) + 4) for n in range(N + 1)]
        pi = [aligned_zeros(_rup(p.nx[n + 1], BS) + 4) for n in range(N)]
        lam = [aligned_zeros(2 * pnb[n] + 2 * png[n] + 4) for n in range(N + 1)]
        t = [aligned_zeros(2 * pnb[n] + 2 * png[n] + 4) for n in range(N + 1)]
        idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
        wsz = self.lib.d_ip2_res_mpc_hard_tv_work_space_size_bytes(N, nx, nu, nb, ng)
        work = aligned_zeros(wsz // 8 + 16)
        stat = np.zeros(5 * k_max + 5)
        kk = C.c_int(0)
        pa = ptr_array
        keep = (pa(BAbt), pa(RSQ), pa(DCt), pa(d), pa(ux), pa(pi), pa(lam), pa(t), pa(idxb))
        status = self.lib.d_ip2_res_mpc_hard_tv(C.byref(kk), k_max, mu0, mu_tol, alpha_min, warm_start, stat.ctypes.data, N,
                                                nx, nu, nb, keep[8], ng, keep[0], keep[1], keep[2], keep[3], keep[4], 1,
                                                keep[5], keep[6], keep[7], work.ctypes.data)
        sp = self._split_bound_like
        return dict(status=status, kk=kk.value, u=[ux[n][:p.nu[n]].copy() for n in range(N)],
                    x=[ux[n][p.nu[n]:p.nu[n] + p.nx[n]].copy() for n in range(N + 1)],
                    pi=[pi[n][:p.nx[n + 1]].copy() for n in range(N)],
                    lam=[sp(lam[n], p.nb[n], pnb[n], ngl[n], png[n]) for n in range(N + 1)],
                    t=[sp(t[n], p.nb[n], pnb[n], ngl[n], png[n]) for n in range(N + 1)],
                    stat=stat[:5 * kk.value].reshape(-1, 5).copy())

    def part_cond(self, p: Ocp, N2: int, expand_from=None):
        """d_part_cond_compute_problem_size + d_part_cond on panel-major data (reference include/lqcp_solvers.h:86-92): the condensed
        problem as dense arrays per block.  expand_from = (ux2, pi2, lam2, t2) lists in the condensed problem's sizes additionally runs
        d_part_expand_solution (:95) and returns the full-space (u, x, pi, lam, t)."""
        N = p.N
        nx, nu, nb, ng = self._sizes(p)
        BAbt, RSQ = self._pm_problem(p)
        DCt, d, pnb, png, ngl = self._pm_general(p)
        idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
        o = [(C.c_int * (N2 + 1))() for _ in range(4)]
        L = self.lib
        L.d_part_cond_compute_problem_size.restype = None
        L.d_part_cond_compute_problem_size(N, nx, nu, nb, ptr_array(idxb), ng, N2, *o)
        nx2, nu2, nb2, ng2 = (list(v) for v in o)
        for f in (L.d_part_cond_memory_space_size_bytes, L.d_part_cond_work_space_size_bytes, L.d_part_expand_work_space_size_bytes):
            f.restype = C.c_int
        msz = L.d_part_cond_memory_space_size_bytes(N, nx, nu, nb, ptr_array(idxb), ng, N2, *o)
        wsz = L.d_part_cond_work_space_size_bytes(N, nx, nu, nb, ptr_array(idxb), ng, N2, *o)
        mem, work = aligned_zeros(msz // 8 + 64), aligned_zeros(wsz // 8 + 64)
        P = lambda n: (C.c_void_p * n)()
        pB2, pQ2, pD2, pd2, pi2_ = P(N2 + 1), P(N2 + 1), P(N2 + 1), P(N2 + 1), P(N2 + 1)
        pa = ptr_array
        keep = (pa(BAbt), pa(RSQ), pa(DCt), pa(d), pa(idxb))
        L.d_part_cond.restype = None
        L.d_part_cond(N, nx, nu, nb, keep[4], ng, keep[0], keep[1], keep[2], keep[3], N2, o[0], o[1], o[2], pi2_, o[3], pB2, pQ2, pD2, pd2,
                      C.c_void_p(mem.ctypes.data), C.c_void_p(work.ctypes.data))

        def from_pmat(ptr, rows, cols):
            pr, sda = _rup(max(rows, 1), BS), _rup(max(cols, 1), NCL)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(pr * sda,))
            I, J = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
            return a[(I // BS) * BS * sda + I % BS + BS * J].copy() if rows and cols else np.zeros((rows, cols))
        out = dict(nx2=nx2, nu2=nu2, nb2=nb2, ng2=ng2, BAbt=[], RSQrq=[], DCt=[], lb=[], ub=[], lg=[], ug=[], idxb=[])
        for k in range(N2):
            nux2 = nu2[k] + nx2[k]
            out["BAbt"].append(from_pmat(pB2[k], nux2 + 1, nx2[k + 1]))
            out["RSQrq"].append(np.tril(from_pmat(pQ2[k], nux2 + 1, nux2)))
            out["DCt"].append(from_pmat(pD2[k], nux2, ng2[k]))
            pnb2, png2 = _rup(nb2[k], BS), _rup(ng2[k], BS)
            dd = np.ctypeslib.as_array(C.cast(pd2[k], C.POINTER(C.c_double)), shape=(2 * pnb2 + 2 * png2 + 1,))
            out["lb"].append(dd[:nb2[k]].copy()); out["ub"].append(dd[pnb2:pnb2 + nb2[k]].copy())
            out["lg"].append(dd[2 * pnb2:2 * pnb2 + ng2[k]].copy()); out["ug"].append(dd[2 * pnb2 + png2:2 * pnb2 + png2 + ng2[k]].copy())
            out["idxb"].append(np.ctypeslib.as_array(C.cast(pi2_[k], C.POINTER(C.c_int)), shape=(max(nb2[k], 1),))[:nb2[k]].copy())
        if expand_from is not None:
            ux2l, pi2l, lam2l, t2l = expand_from
            pnb2 = [_rup(v, BS) for v in nb2]; png2 = [_rup(v, BS) for v in ng2]

            def bound_like(v, k):     # [lb ub lg ug] of a condensed stage -> the padded lib4 arrangement
                a = aligned_zeros(2 * pnb2[k] + 2 * png2[k] + 4)
                a[:nb2[k]] = v[:nb2[k]]; a[pnb2[k]:pnb2[k] + nb2[k]] = v[nb2[k]:2 * nb2[k]]
                a[2 * pnb2[k]:2 * pnb2[k] + ng2[k]] = v[2 * nb2[k]:2 * nb2[k] + ng2[k]]
                a[2 * pnb2[k] + png2[k]:2 * pnb2[k] + png2[k] + ng2[k]] = v[2 * nb2[k] + ng2[k]:]
                return a
            hux2 = [np.concatenate([np.asarray(ux2l[k], dtype=np.float64), np.zeros(4)]) for k in range(N2 + 1)]
            hpi2 = [np.concatenate([np.asarray(pi2l[k], dtype=np.float64), np.zeros(4)]) for k in range(N2)]
            hlam2 = [bound_like(np.asarray(lam2l[k]), k) for k in range(N2 + 1)]; ht2 = [bound_like(np.asarray(t2l[k]), k) for k in range(N2 + 1)]
            hux = [aligned_zeros(_rup(p.nx[n] + p.nu[n] + 1, BS) + 4) for n in range(N + 1)]
            hpi = [aligned_zeros(_rup(p.nx[n + 1], BS) + 4) for n in range(N)]
            hlam = [aligned_zeros(2 * pnb[n] + 2 * png[n] + 4) for n in range(N + 1)]; ht = [aligned_zeros(2 * pnb[n] + 2 * png[n] + 4) for n in range(N + 1)]
            hb = [np.ascontiguousarray(v) for v in p.b]
            hrq = [np.concatenate([p.r[n], p.q[n]]) if n < N else np.ascontiguousarray(p.q[n]) for n in range(N + 1)]
            ew = aligned_zeros(L.d_part_expand_work_space_size_bytes(N, nx, nu, nb, ng) // 8 + 64)
            k2 = (pa(hux2), pa(hpi2), pa(hlam2), pa(ht2), pa(hux), pa(hpi), pa(hlam), pa(ht), pa(hb), pa(hrq))
            L.d_part_expand_solution.restype = None
            L.d_part_expand_solution(N, nx, nu, nb, keep[4], ng, keep[0], k2[8], keep[1], k2[9], keep[2], k2[4], k2[5], k2[6], k2[7], N2,
                                     o[0], o[1], o[2], pi2_, o[3], k2[0], k2[1], k2[2], k2[3], C.c_void_p(ew.ctypes.data))
            sp = self._split_bound_like
            out["expanded"] = dict(u=[hux[n][:p.nu[n]].copy() for n in range(N)], x=[hux[n][p.nu[n]:p.nu[n] + p.nx[n]].copy() for n in range(N + 1)],
                                   pi=[hpi[n][:p.nx[n + 1]].copy() for n in range(N)],
                                   lam=[sp(hlam[n], p.nb[n], pnb[n], ngl[n], png[n]) for n in range(N + 1)],
                                   t=[sp(ht[n], p.nb[n], pnb[n], ngl[n], png[n]) for n in range(N + 1)])
        del keep
        return out

    def ip_then_solve_kkt_new_rhs_high_level(self, p: Ocp, p2: Ocp, *, order="fortran", k_max=40, mu0=2.0, mu_tol=1e-8):
        """{c,fortran}_order_d_ip_ocp_hard_tv on p, then {c,fortran}_order_d_solve_kkt_new_rhs_ocp_hard_tv (reference
        include/c_interface.h:63,67) on the SAME work0 with the vectors b, q, r, lb, ub of p2."""
        N = p.N
        conv = (lambda M: np.ascontiguousarray(M)) if order == "c" else (lambda M: np.asfortranarray(M))
        A = [conv(M) for M in p.A]; B = [conv(M) for M in p.B]; Q = [conv(M) for M in p.Q]; S = [conv(M) for M in p.S]; R = [conv(M) for M in p.R]
        c = np.ascontiguousarray
        v1 = [[c(v) for v in arr] for arr in (p.b, p.q, p.r, p.lb, p.ub)]
        v2 = [[c(v) for v in arr] for arr in (p2.b, p2.q, p2.r, p2.lb, p2.ub)]
        x = [np.zeros(max(n, 1)) for n in p.nx]; u = [np.zeros(max(n, 1)) for n in p.nu[:N]]
        pi = [np.zeros(max(p.nx[n + 1], 1)) for n in range(N)]; lam = [np.zeros(max(2 * nb, 1)) for nb in p.nb]
        idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
        nx, nu, nb, ng = int_array(p.nx), int_array(p.nu), int_array(p.nb), int_array([0] * (N + 1))
        empty = [np.zeros(1) for _ in range(N + 1)]
        wsz = self.lib.hpmpc_d_ip_ocp_hard_tv_work_space_size_bytes(N, nx, nu, nb, ptr_array(idxb), ng, N)
        work = aligned_zeros(wsz // 8 + 16)
        res = np.zeros(8); stat = np.zeros(5 * k_max + 5); kk = C.c_int(0)
        pa = ptr_array
        pidx = pa(idxb)
        mats = [pa(A), pa(B)]
        outs = [pa(x), pa(u), pa(pi), pa(lam)]
        e4 = [pa(empty), pa(empty), pa(empty), pa(empty)]
        b, q, r, lb, ub = (pa(v) for v in v1)
        fn = self.lib.c_order_d_ip_ocp_hard_tv if order == "c" else self.lib.fortran_order_d_ip_ocp_hard_tv
        status = fn(C.byref(kk), k_max, mu0, mu_tol, N, nx, nu, nb, pidx, ng, N, 0, mats[0], mats[1], b, pa(Q), pa(S), pa(R), q, r, lb, ub, *e4,
                    *outs, res.ctypes.data, work.ctypes.data, stat.ctypes.data)
        f2 = self.lib.c_order_d_solve_kkt_new_rhs_ocp_hard_tv if order == "c" else self.lib.fortran_order_d_solve_kkt_new_rhs_ocp_hard_tv
        f2.restype = None
        f2.argtypes = [C.c_int] + [C.c_void_p] * 25
        b2, q2, r2, lb2, ub2 = (pa(v) for v in v2)
        res2 = np.zeros(8)
        f2(N, nx, nu, nb, pidx, ng, mats[0], mats[1], b2, pa(Q), pa(S), pa(R), q2, r2, lb2, ub2, *e4, *outs, res2.ctypes.data, work.ctypes.data)
        return dict(status=status, kk=kk.value, x=[x[n][:p.nx[n]].copy() for n in range(N + 1)], u=[u[n][:p.nu[n]].copy() for n in range(N)],
                    pi=[pi[n][:p.nx[n + 1]].copy() for n in range(N)], lam=[lam[n][:2 * p.nb[n]].copy() for n in range(N + 1)],
                    inf_norm_res=res2[:4].copy())

    def single_newton_step(self, p: Ocp, ux0, pi0, lam0, t0, *, k_max=1, mu0=1e-3):
        """fortran_order_d_ip_ocp_hard_tv_single_newton_step (reference include/c_interface.h:66): k_max Newton steps from the
        iterate ux0[n] = [u_n ; x_n], pi0[n], lam0[n] / t0[n] = [lb(nb) ub(nb)]."""
        N = p.N
        f = np.asfortranarray
        A = [f(M) for M in p.A]; B = [f(M) for M in p.B]; Q = [f(M) for M in p.Q]; S = [f(M) for M in p.S]; R = [f(M) for M in p.R]
        c = np.ascontiguousarray
        b = [c(v) for v in p.b]; q = [c(v) for v in p.q]; r = [c(v) for v in p.r]; lb = [c(v) for v in p.lb]; ub = [c(v) for v in p.ub]
        x = [np.zeros(max(n, 1)) for n in p.nx]; u = [np.zeros(max(n, 1)) for n in p.nu[:N]]
        pi = [np.zeros(max(p.nx[n + 1], 1)) for n in range(N)]
        lam = [np.zeros(max(2 * nb, 1)) for nb in p.nb]; t = [np.zeros(max(2 * nb, 1)) for nb in p.nb]
        pad = lambda v: c(np.asarray(v, dtype=np.float64)) if len(v) else np.zeros(1)
        a_ux0 = [pad(v) for v in ux0]; a_pi0 = [pad(v) for v in pi0] + [np.zeros(1)]; a_lam0 = [pad(v) for v in lam0]; a_t0 = [pad(v) for v in t0]
        idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
        nx, nu, nb, ng = int_array(p.nx), int_array(p.nu), int_array(p.nb), int_array([0] * (N + 1))
        empty = [np.zeros(1) for _ in range(N + 1)]
        wsz = self.lib.hpmpc_d_ip_ocp_hard_tv_work_space_size_bytes(N, nx, nu, nb, ptr_array(idxb), ng, N)
        work = aligned_zeros(wsz // 8 + 16)
        res = np.zeros(8); stat = np.zeros(5 * k_max + 5); kk = C.c_int(0)
        fn = self.lib.fortran_order_d_ip_ocp_hard_tv_single_newton_step
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int] + [C.c_void_p] * 26
        pa = ptr_array
        arrs = [pa(A), pa(B), pa(b), pa(Q), pa(S), pa(R), pa(q), pa(r), pa(lb), pa(ub), pa(empty), pa(empty), pa(empty), pa(empty),
                pa(x), pa(u), pa(pi), pa(lam), pa(t)]
        tail = [pa(a_ux0), pa(a_pi0), pa(a_lam0), pa(a_t0)]
        pidx = pa(idxb)
        status = fn(C.byref(kk), k_max, mu0, 1e-8, N, nx, nu, nb, pidx, ng, N, 0, *arrs, res.ctypes.data, work.ctypes.data, stat.ctypes.data, *tail)
        return dict(status=status, kk=kk.value, x=[x[n][:p.nx[n]].copy() for n in range(N + 1)], u=[u[n][:p.nu[n]].copy() for n in range(N)],
                    pi=[pi[n][:p.nx[n + 1]].copy() for n in range(N)], lam=[lam[n][:2 * p.nb[n]].copy() for n in range(N + 1)],
                    t=[t[n][:2 * p.nb[n]].copy() for n in range(N + 1)], inf_norm_res=res[:4].copy(), stat=stat[:5 * kk.value].reshape(-1, 5).copy())

    def residuals(self, p: Ocp, u, x, pi, lam, t, which: str = "res_res"):
        """d_res_res_mpc_hard_tv (reference include/mpc_solvers.h:47) or d_res_mpc_hard_tv (:36) on panel-major data; lam / t are
        [lb ub lg ug] per stage.  Returns rq, rb, rd (and rm) in the same orderings, and mu."""
        N = p.N
        nx, nu, nb, _ = self._sizes(p)
        BAbt, RSQ = self._pm_problem(p)
        DCt, d, pnb, png, ngl = self._pm_general(p)
        ng = int_array(ngl)
        az = aligned_zeros
        hb = [az(_rup(p.nx[n + 1], BS) + 4) for n in range(N)]
        hq = [az(_rup(p.nx[n] + p.nu[n] + 1, BS) + 4) for n in range(N + 1)]
        hux = [az(_rup(p.nx[n] + p.nu[n] + 1, BS) + 4) for n in range(N + 1)]
        hpi = [az(_rup(p.nx[n + 1], BS) + 4) for n in range(N)]
        mk = lambda: [az(2 * pnb[n] + 2 * png[n] + 4) for n in range(N + 1)]
        hlam, ht, hrd, hrm = mk(), mk(), mk(), mk()
        hrq = [az(_rup(p.nx[n] + p.nu[n] + 1, BS) + 4) for n in range(N + 1)]; hrb = [az(_rup(p.nx[n + 1], BS) + 4) for n in range(N)]
        for n in range(N):
            hb[n][:p.nx[n + 1]] = p.b[n]; hpi[n][:p.nx[n + 1]] = pi[n]
        for n in range(N + 1):
            nun, nxn, nbn, ngn = p.nu[n], p.nx[n], p.nb[n], ngl[n]
            hq[n][:nun] = p.r[n]; hq[n][nun:nun + nxn] = p.q[n]
            if n < N: hux[n][:nun] = u[n]
            hux[n][nun:nun + nxn] = x[n]
            for src, dst in ((lam[n], hlam[n]), (t[n], ht[n])):
                dst[:nbn] = src[:nbn]; dst[pnb[n]:pnb[n] + nbn] = src[nbn:2 * nbn]
                dst[2 * pnb[n]:2 * pnb[n] + ngn] = src[2 * nbn:2 * nbn + ngn]; dst[2 * pnb[n] + png[n]:2 * pnb[n] + png[n] + ngn] = src[2 * nbn + ngn:]
        idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
        mu = C.c_double(0.0)
        work = az(4096)
        pa = ptr_array
        k = [pa(BAbt), pa(hb), pa(RSQ), pa(hq), pa(hux), pa(DCt), pa(d), pa(hpi), pa(hlam), pa(ht), pa(hrq), pa(hrb), pa(hrd), pa(hrm), pa(idxb)]
        if which == "res_res":
            fn = self.lib.d_res_res_mpc_hard_tv
            fn.restype = None; fn.argtypes = [C.c_int] + [C.c_void_p] * 21
            fn(N, nx, nu, nb, k[14], ng, k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7], k[8], k[9], work.ctypes.data, k[10], k[11], k[12], k[13], C.byref(mu))
        else:
            fn = self.lib.d_res_mpc_hard_tv
            fn.restype = None; fn.argtypes = [C.c_int] + [C.c_void_p] * 19
            fn(N, nx, nu, nb, k[14], ng, k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7], k[8], k[9], k[10], k[11], k[12], C.byref(mu))
        sp = self._split_bound_like
        return dict(rq=[hrq[n][:p.nu[n] + p.nx[n]].copy() for n in range(N + 1)], rb=[hrb[n][:p.nx[n + 1]].copy() for n in range(N)],
                    rd=[sp(hrd[n], p.nb[n], pnb[n], ngl[n], png[n]) for n in range(N + 1)],
                    rm=[sp(hrm[n], p.nb[n], pnb[n], ngl[n], png[n]) for n in range(N + 1)], mu=mu.value)

    def ric_upd(self, p: Ocp, Qx, qx, mode: str = "sv"):
        """d_back_ric_rec_sv_tv_res (or trf + trs) WITH the IPM's per-constraint updates (reference lqcp_solvers/d_back_ric_rec.c:112):
        Qx[n], qx[n] hold nb[n] + ng[n] entries (bounds first); bounds add Qx to the Hessian diagonal / qx to the gradient at idxb,
        general constraints add [D C]' diag(Qx) [D C] and [D C]' qx."""
        N = p.N
        nx, nu, nb, _ = self._sizes(p)
        BAbt, RSQ = self._pm_problem(p)
        DCt, _, pnb, png, ngl = self._pm_general(p)
        ng = int_array(ngl)
        wsz = self.lib.d_back_ric_rec_sv_tv_work_space_size_bytes(N, nx, nu, nb, ng)
        msz = self.lib.d_back_ric_rec_sv_tv_memory_space_size_bytes(N, nx, nu, nb, ng)
        work, mem = aligned_zeros(wsz // 8 + 64), aligned_zeros(msz // 8 + 64)
        hux = [aligned_zeros(_rup(p.nx[n] + p.nu[n] + 1, BS) + 4) for n in range(N + 1)]
        hpi = [aligned_zeros(_rup(p.nx[n + 1], BS) + 4) for n in range(N)]
        hPb = [aligned_zeros(_rup(p.nx[n + 1], BS) + 4) for n in range(N)]
        idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
        pQx = [aligned_zeros(pnb[n] + png[n] + 4) for n in range(N + 1)]; pqx = [aligned_zeros(pnb[n] + png[n] + 4) for n in range(N + 1)]
        bd = [aligned_zeros(pnb[n] + 4) for n in range(N + 1)]
        for n in range(N + 1):
            nbn = p.nb[n]
            pQx[n][:nbn] = Qx[n][:nbn]; pQx[n][pnb[n]:pnb[n] + ngl[n]] = Qx[n][nbn:]
            pqx[n][:nbn] = qx[n][:nbn]; pqx[n][pnb[n]:pnb[n] + ngl[n]] = qx[n][nbn:]
            H = np.zeros((p.nu[n] + p.nx[n],)); H[:p.nu[n]] = np.diag(p.R[n]); H[p.nu[n]:] = np.diag(p.Q[n])
            bd[n][:nbn] = H[np.asarray(p.idxb[n], dtype=int)] if nbn else 0.0
        hb = [aligned_zeros(_rup(p.nx[n + 1], BS) + 4) for n in range(N)]
        hq = [aligned_zeros(_rup(p.nx[n] + p.nu[n] + 1, BS) + 4) for n in range(N + 1)]
        for n in range(N):
            hb[n][:p.nx[n + 1]] = p.b[n]
        for n in range(N + 1):
            hq[n][:p.nu[n]] = p.r[n]; hq[n][p.nu[n]:p.nu[n] + p.nx[n]] = p.q[n]
        pa = ptr_array
        k = dict(B=pa(BAbt), Q=pa(RSQ), G=pa(DCt), ux=pa(hux), pi=pa(hpi), Pb=pa(hPb), idx=pa(idxb), Qx=pa(pQx), qx=pa(pqx), bd=pa(bd),
                 hb=pa(hb), hq=pa(hq))
        if mode == "sv":
            self.lib.d_back_ric_rec_sv_tv_res(N, nx, nu, nb, k["idx"], ng, 0, k["B"], k["hb"], 0, k["Q"], k["hq"], k["bd"], k["G"], k["Qx"], k["qx"],
                                              k["ux"], 1, k["pi"], 1, k["Pb"], mem.ctypes.data, work.ctypes.data)
        else:
            self.lib.d_back_ric_rec_trf_tv_res(N, nx, nu, nb, k["idx"], ng, k["B"], k["Q"], k["G"], k["Qx"], k["bd"], mem.ctypes.data, work.ctypes.data)
            self.lib.d_back_ric_rec_trs_tv_res(N, nx, nu, nb, k["idx"], ng, k["B"], k["hb"], k["hq"], k["G"], k["qx"], k["ux"], 1, k["pi"], 1, k["Pb"],
                                               mem.ctypes.data, work.ctypes.data)
        return dict(u=[hux[n][:p.nu[n]].copy() for n in range(N)], x=[hux[n][p.nu[n]:p.nu[n] + p.nx[n]].copy() for n in range(N + 1)],
                    pi=[hpi[n][:p.nx[n + 1]].copy() for n in range(N)])

    def ip2_then_kkt_new_rhs(self, p: Ocp, p2: Ocp, *, k_max=40, mu0=2.0, mu_tol=1e-8, alpha_min=1e-8):
        """d_ip2_res_mpc_hard_tv on p, then d_kkt_solve_new_rhs_res_mpc_hard_tv (reference mpc_solvers/d_ip2_res_hard.c:1922, called as in
        test_problems/test_d_ip_hard.c:1040) on the SAME work memory with the vectors b, q, r, lb, ub of p2: the IPM's last KKT system
        solved again for a new right-hand side.  (The reference's high-level fortran_order_d_solve_kkt_new_rhs_ocp_hard_tv lays its work
        space out differently from fortran_order_d_ip_ocp_hard_tv -- fortran_order_interface.c:1193 against :459 -- and cannot follow it.)"""
        N = p.N
        nx, nu, nb, ng = self._sizes(p)
        BAbt, RSQ = self._pm_problem(p)
        pnb = [_rup(v, BS) for v in p.nb]
        mk = lambda: [aligned_zeros(2 * pnb[n] + 4) for n in range(N + 1)]
        d, d2, lam, t = mk(), mk(), mk(), mk()
        for n in range(N + 1):
            d[n][:p.nb[n]] = p.lb[n]; d[n][pnb[n]:pnb[n] + p.nb[n]] = p.ub[n]
            d2[n][:p.nb[n]] = p2.lb[n]; d2[n][pnb[n]:pnb[n] + p.nb[n]] = p2.ub[n]
        ux = [aligned_zeros(_rup(p.nx[n] + p.nu[n] + 1, BS) + 4) for n in range(N + 1)]
        pi = [aligned_zeros(_rup(p.nx[n + 1], BS) + 4) for n in range(N)]
        hb = [aligned_zeros(_rup(p.nx[n + 1], BS) + 4) for n in range(N)]
        hq = [aligned_zeros(_rup(p.nx[n] + p.nu[n] + 1, BS) + 4) for n in range(N + 1)]
        for n in range(N):
            hb[n][:p.nx[n + 1]] = p2.b[n]
        for n in range(N + 1):
            hq[n][:p.nu[n]] = p2.r[n]; hq[n][p.nu[n]:p.nu[n] + p.nx[n]] = p2.q[n]
        idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
        dummy = [aligned_zeros(8) for _ in range(N + 1)]
        wsz = self.lib.d_ip2_res_mpc_hard_tv_work_space_size_bytes(N, nx, nu, nb, ng)
        work = aligned_zeros(wsz // 8 + 16)
        stat = np.zeros(5 * k_max + 5)
        kk = C.c_int(0)
        pa = ptr_array
        k = dict(BAbt=pa(BAbt), RSQ=pa(RSQ), dm=pa(dummy), d=pa(d), d2=pa(d2), ux=pa(ux), pi=pa(pi), lam=pa(lam), t=pa(t), idxb=pa(idxb),
                 hb=pa(hb), hq=pa(hq))
        status = self.lib.d_ip2_res_mpc_hard_tv(C.byref(kk), k_max, mu0, mu_tol, alpha_min, 0, stat.ctypes.data, N,
                                                nx, nu, nb, k["idxb"], ng, k["BAbt"], k["RSQ"], k["dm"], k["d"], k["ux"], 1,
                                                k["pi"], k["lam"], k["t"], work.ctypes.data)
        fn = self.lib.d_kkt_solve_new_rhs_res_mpc_hard_tv
        fn.restype = None
        fn.argtypes = [C.c_int] + [C.c_void_p] * 12 + [C.c_int] + [C.c_void_p] * 4
        fn(N, nx, nu, nb, k["idxb"], ng, k["BAbt"], k["hb"], k["RSQ"], k["hq"], k["dm"], k["d2"], k["ux"], 1, k["pi"], k["lam"], k["t"],
           work.ctypes.data)
        return dict(status=status, kk=kk.value, u=[ux[n][:p.nu[n]].copy() for n in range(N)],
                    x=[ux[n][p.nu[n]:p.nu[n] + p.nx[n]].copy() for n in range(N + 1)],
                    pi=[pi[n][:p.nx[n + 1]].copy() for n in range(N)],
                    lam=[np.concatenate([lam[n][:p.nb[n]], lam[n][pnb[n]:pnb[n] + p.nb[n]]]) for n in range(N + 1)],
                    t=[np.concatenate([t[n][:p.nb[n]], t[n][pnb[n]:pnb[n] + p.nb[n]]]) for n in range(N + 1)])



# ------------------------------------------------------------------------------------------- batched C ABI
class Sizes(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in ("in_stride", "ux_stride", "pi_stride", "lam_stride", "L_stride", "ipm_work_stride")] \
        + [(n, C.c_int) for n in ("N", "nzM", "nxM", "nbtot", "grid", "warps_per_cta", "n_slots", "smem_per_cta", "fast_variant",
                                    "ipm_grid", "ipm_warps_per_cta", "ipm_fast_variant")]


_product = None


def product() -> C.CDLL:
    """libhpmpc_b200.so, loaded once.  Raises if the library is missing: there is no fallback path."""
    global _product
    if _product is None:
        if not os.path.exists(PRODUCT_LIB):
            raise RuntimeError(f"{PRODUCT_LIB} is missing; run __graft_entry__.build() -- hpmpc_b200 has no CPU fallback")
        L = C.CDLL(PRODUCT_LIB, mode=C.RTLD_LOCAL)
        L.hpmpc_b200_ocp_create.restype = C.c_int
        L.hpmpc_b200_ocp_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hpmpc_b200_ocp_create_gen.restype = C.c_int
        L.hpmpc_b200_ocp_create_gen.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hpmpc_b200_ocp_create_padded.restype = C.c_int
        L.hpmpc_b200_ocp_create_padded.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hpmpc_b200_ocp_padded_shape.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hpmpc_b200_pack_general.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.hpmpc_b200_d_back_ric_rec_sv_upd_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 7
        L.hpmpc_b200_d_back_ric_rec_trf_upd_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 4
        L.hpmpc_b200_d_back_ric_rec_trs_upd_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 6
        L.hpmpc_b200_ocp_generic_factor_layout.argtypes = [C.c_void_p]
        L.hpmpc_b200_ocp_destroy.argtypes = [C.c_void_p]
        L.hpmpc_b200_ocp_set_launch.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hpmpc_b200_ocp_sizes.argtypes = [C.c_void_p, C.POINTER(Sizes)]
        L.hpmpc_b200_ocp_stage_offsets.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 7
        L.hpmpc_b200_pack_instance.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 11
        L.hpmpc_b200_d_back_ric_rec_sv_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 5
        L.hpmpc_b200_d_back_ric_rec_trf_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 3
        L.hpmpc_b200_d_back_ric_rec_trs_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 5
        L.hpmpc_b200_d_ip2_res_mpc_hard_batch.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                                         C.c_double, C.c_int] + [C.c_void_p] * 6
        L.hpmpc_b200_d_back_ric_rec_sv_batch_host.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 3
        L.hpmpc_b200_d_ip2_res_mpc_hard_batch_host.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                                              C.c_double, C.c_int] + [C.c_void_p] * 5
        L.hpmpc_b200_part_cond_compute_problem_size.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 5
        L.hpmpc_b200_pcond_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.hpmpc_b200_pcond_destroy.argtypes = [C.c_void_p]
        L.hpmpc_b200_pcond_full.restype = C.c_void_p
        L.hpmpc_b200_pcond_full.argtypes = [C.c_void_p]
        L.hpmpc_b200_pcond_cond.restype = C.c_void_p
        L.hpmpc_b200_pcond_cond.argtypes = [C.c_void_p]
        L.hpmpc_b200_d_part_cond_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 3
        L.hpmpc_b200_d_part_expand_solution_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 10
        L.hpmpc_b200_d_ip2_res_mpc_hard_part_cond_batch.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                                                   C.c_double] + [C.c_void_p] * 6
        L.hpmpc_b200_fp64_peak_tflops.restype = C.c_double
        L.hpmpc_b200_fp64_peak_tflops.argtypes = [C.c_int]
        L.hpmpc_b200_version.restype = C.c_char_p
        _product = L
    return _product


class BatchOcp:
    """Handle on a size pattern (hpmpc_b200_ocp_create) plus numpy-side packing helpers."""

    def __init__(self, p: Ocp, device: int = 0, handle=None, padded: bool = False):
        """handle: wrap an existing hpmpc_b200_ocp* owned by someone else (a partial-condensing handle) instead of creating one.
        padded: hpmpc_b200_ocp_create_padded -- a uniform shape without kernels of its own is embedded in the next compiled shape."""
        L = product()
        self.L, self.p, self.device = L, p, device
        self.h = C.c_void_p()
        self.owned = handle is None
        if handle is not None:
            self.h = C.c_void_p(handle)
        else:
            idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
            self._keep = idxb
            if padded:
                assert not p.ng
                rc = L.hpmpc_b200_ocp_create_padded(C.byref(self.h), p.N, int_array(p.nx), int_array(p.nu), int_array(p.nb), ptr_array(idxb), device)
            else:
                rc = L.hpmpc_b200_ocp_create_gen(C.byref(self.h), p.N, int_array(p.nx), int_array(p.nu), int_array(p.nb), ptr_array(idxb),
                                                 int_array(p.ng_list()), device)
            if rc != 0:
                raise RuntimeError(f"hpmpc_b200_ocp_create failed ({rc})")
        self.refresh()
        NX, NU = C.c_int(0), C.c_int(0)
        self.padded = bool(L.hpmpc_b200_ocp_padded_shape(self.h, C.byref(NX), C.byref(NU)))
        # slots of the u-part of a stage in the handle's frame (the x-part starts behind them)
        self.nu_frame = [(NU.value if n < p.N else 0) if self.padded else p.nu[n] for n in range(p.N + 1)]
        self.off = []
        for n in range(p.N + 1):
            v = [C.c_int() for _ in range(7)]
            L.hpmpc_b200_ocp_stage_offsets(self.h, n, *[C.byref(x) for x in v])
            self.off.append(dict(zip(("BAbt", "RSQ", "d", "ux", "pi", "lam", "L"), [x.value for x in v])))

    def refresh(self):
        self.sz = Sizes()
        self.L.hpmpc_b200_ocp_sizes(self.h, C.byref(self.sz))

    def set_launch(self, ctas_per_sm: int, warps: int):
        rc = self.L.hpmpc_b200_ocp_set_launch(self.h, ctas_per_sm, warps)
        self.refresh()
        return rc

    def close(self):
        if self.h and self.owned:
            self.L.hpmpc_b200_ocp_destroy(self.h)
        self.h = C.c_void_p()

    def pack(self, p: Ocp) -> np.ndarray:
        """One instance -> native packed block (hpmpc_b200_pack_instance, row-major inputs)."""
        blk = np.zeros(self.sz.in_stride)
        c = np.ascontiguousarray
        arrs = [[c(M) for M in L] for L in (p.A, p.B, p.b, p.Q, p.S, p.R, p.q, p.r, p.lb, p.ub)]
        ptrs = [ptr_array(a) for a in arrs]
        rc = self.L.hpmpc_b200_pack_instance(self.h, 1, *ptrs, blk.ctypes.data)
        assert rc == 0
        if p.ng:
            g = [[c(M) if M.size else np.zeros(1) for M in L] for L in p.general_arrays()]
            rc = self.L.hpmpc_b200_pack_general(self.h, 1, *[ptr_array(a) for a in g], blk.ctypes.data)
            assert rc == 0
        return blk

    def split_ux(self, ux: np.ndarray):
        p = self.p
        u = [ux[self.off[n]["ux"]:self.off[n]["ux"] + p.nu[n]].copy() for n in range(p.N)]
        x = [ux[self.off[n]["ux"] + self.nu_frame[n]:self.off[n]["ux"] + self.nu_frame[n] + p.nx[n]].copy() for n in range(p.N + 1)]
        return u, x

    def split_pi(self, pi: np.ndarray):
        p = self.p
        return [pi[self.off[n]["pi"]:self.off[n]["pi"] + p.nx[n + 1]].copy() for n in range(p.N)]

    def split_pi_real(self, pi: np.ndarray):
        """pi of every edge, the caller's real entries (they come first in a padded frame's slots)."""
        return self.split_pi(pi)

    def split_lam(self, lam: np.ndarray):
        p = self.p
        ng = p.ng_list()
        return [lam[self.off[n]["lam"]:self.off[n]["lam"] + 2 * p.nb[n] + 2 * ng[n]].copy() for n in range(p.N + 1)]


def part_cond_sizes(p: Ocp, N2: int):
    """hpmpc_b200_part_cond_compute_problem_size: (nx2, nu2, nb2, ng2, idxb2) of the condensed problem (host only)."""
    L = product()
    N = p.N
    idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
    out = [(C.c_int * (N2 + 1))() for _ in range(4)]
    nbt = sum(p.nb) + 1
    idxb2 = [np.zeros(nbt, dtype=np.int32) for _ in range(N2 + 1)]
    rc = L.hpmpc_b200_part_cond_compute_problem_size(N, int_array(p.nx), int_array(p.nu), int_array(p.nb), ptr_array(idxb),
                                                     int_array(p.ng_list()) if p.ng else None, N2,
                                                     *out, ptr_array(idxb2))
    if rc != 0:
        raise RuntimeError(f"hpmpc_b200_part_cond_compute_problem_size failed ({rc})")
    nx2, nu2, nb2, ng2 = (list(v) for v in out)
    return nx2, nu2, nb2, ng2, [idxb2[k][:nb2[k]].copy() for k in range(N2 + 1)]


class PartCond:
    """hpmpc_b200_pcond: the full and the condensed size pattern of a partially condensed problem (N2 blocks)."""

    def __init__(self, p: Ocp, N2: int, device: int = 0):
        L = product()
        self.L, self.p, self.N2 = L, p, N2
        self.h = C.c_void_p()
        idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
        rc = L.hpmpc_b200_pcond_create(C.byref(self.h), p.N, int_array(p.nx), int_array(p.nu[:p.N]), int_array(p.nb), ptr_array(idxb),
                                       int_array(p.ng_list()) if p.ng else None, N2, device)
        if rc != 0:
            raise RuntimeError(f"hpmpc_b200_pcond_create failed ({rc})")
        nx2, nu2, nb2, ng2, idxb2 = part_cond_sizes(p, N2)
        self.p2 = Ocp(N=N2, nx=nx2, nu=nu2, nb=nb2, idxb=idxb2, ng=ng2)
        self.full = BatchOcp(p, device, handle=L.hpmpc_b200_pcond_full(self.h))
        self.cond = BatchOcp(self.p2, device, handle=L.hpmpc_b200_pcond_cond(self.h))

    def close(self):
        if self.h:
            self.L.hpmpc_b200_pcond_destroy(self.h)
            self.h = C.c_void_p()
