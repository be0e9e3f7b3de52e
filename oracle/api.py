"""ctypes access to oracle/_ref/liboracle.so (our C restatement) and to the compiled reference (checker only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from hpmpc_b200.capi import HpmpcLib, int_array, ptr_array
from hpmpc_b200.problems import Ocp

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ORACLE_LIB = os.path.join(REF_DIR, "liboracle.so")
REF_C99 = os.path.join(REF_DIR, "libhpmpc_ref_c99.so")
REF_AVX2 = os.path.join(REF_DIR, "libhpmpc_ref_avx2.so")


def build(quiet: bool = True):
    """make -C oracle : liboracle.so always, the reference .so files when /root/reference is present."""
    subprocess.run(["make", "-C", HERE, "-j8"], check=True, stdout=subprocess.DEVNULL if quiet else None,
                   stderr=subprocess.DEVNULL if quiet else None)


def have_reference() -> bool:
    return os.path.exists(REF_C99)


def reference(kind: str = "c99") -> HpmpcLib:
    return HpmpcLib(REF_C99 if kind == "c99" else REF_AVX2)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            build()
        L = C.CDLL(ORACLE_LIB)
        L.orc_fortran_order_d_ip_ocp_hard_tv.restype = C.c_int
        L.orc_fortran_order_d_ip_ocp_hard_tv.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_int] + [C.c_void_p] * 5 \
            + [C.c_int, C.c_int] + [C.c_void_p] * 18 + [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_fortran_order_d_ric.restype = None
        L.orc_fortran_order_d_ric.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 13
        _lib = L
    return _lib


def _f(M):
    return np.asfortranarray(M, dtype=np.float64)


def ipm(p: Ocp, *, k_max=40, mu0=2.0, mu_tol=1e-8, warm_start=0, N2=None):
    """orc_fortran_order_d_ip_ocp_hard_tv: same arguments as the reference's fortran_order_d_ip_ocp_hard_tv."""
    L, N = lib(), p.N
    A = [_f(M) for M in p.A]; B = [_f(M) for M in p.B]; Q = [_f(M) for M in p.Q]; S = [_f(M) for M in p.S]; R = [_f(M) for M in p.R]
    c = np.ascontiguousarray
    b = [c(v) for v in p.b]; q = [c(v) for v in p.q]; r = [c(v) for v in p.r]; lb = [c(v) for v in p.lb]; ub = [c(v) for v in p.ub]
    x = [np.zeros(max(n, 1)) for n in p.nx]; u = [np.zeros(max(n, 1)) for n in p.nu[:N]]
    pi = [np.zeros(max(p.nx[n + 1], 1)) for n in range(N)]
    ngl = p.ng_list()
    lam = [np.zeros(max(2 * p.nb[n] + 2 * ngl[n], 1)) for n in range(N + 1)]
    idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
    pad = lambda M: M if M.size else np.zeros(1)
    Cg, Dg, lgg, ugg = p.general_arrays()
    Cg = [pad(_f(M)) for M in Cg]; Dg = [pad(_f(M)) for M in Dg]; lgg = [pad(c(v)) for v in lgg]; ugg = [pad(c(v)) for v in ugg]
    res = np.zeros(8); stat = np.zeros(5 * k_max + 5); kk = C.c_int(0)
    pa = ptr_array
    arrs = [pa(A), pa(B), pa(b), pa(Q), pa(S), pa(R), pa(q), pa(r), pa(lb), pa(ub), pa(Cg), pa(Dg), pa(lgg), pa(ugg),
            pa(x), pa(u), pa(pi), pa(lam)]
    pidx = pa(idxb)
    status = L.orc_fortran_order_d_ip_ocp_hard_tv(C.byref(kk), k_max, mu0, mu_tol, N, int_array(p.nx), int_array(p.nu), int_array(p.nb),
                                                   pidx, int_array(ngl), N if N2 is None else N2, warm_start, *arrs, res.ctypes.data, None, stat.ctypes.data)
    return dict(status=status, kk=kk.value, x=[x[n][:p.nx[n]].copy() for n in range(N + 1)],
                u=[u[n][:p.nu[n]].copy() for n in range(N)], pi=[pi[n][:p.nx[n + 1]].copy() for n in range(N)],
                lam=[lam[n][:2 * p.nb[n] + 2 * ngl[n]].copy() for n in range(N + 1)], inf_norm_res=res[:4].copy(),
                stat=stat[:5 * kk.value].reshape(-1, 5).copy())


def ipm_then_kkt_new_rhs(p: Ocp, p2: Ocp, *, k_max=40, mu0=2.0, mu_tol=1e-8):
    """orc_fortran_order_d_ip_then_kkt_new_rhs: the IPM on p, then its last KKT system solved again for the vectors
    (b, q, r, lb, ub) of p2 (reference: fortran_order_d_solve_kkt_new_rhs_ocp_hard_tv after fortran_order_d_ip_ocp_hard_tv)."""
    L, N = lib(), p.N
    A = [_f(M) for M in p.A]; B = [_f(M) for M in p.B]; Q = [_f(M) for M in p.Q]; S = [_f(M) for M in p.S]; R = [_f(M) for M in p.R]
    c = np.ascontiguousarray
    v1 = [[c(v) for v in arr] for arr in (p.b, p.q, p.r, p.lb, p.ub)]
    v2 = [[c(v) for v in arr] for arr in (p2.b, p2.q, p2.r, p2.lb, p2.ub)]
    x = [np.zeros(max(n, 1)) for n in p.nx]; u = [np.zeros(max(n, 1)) for n in p.nu[:N]]
    pi = [np.zeros(max(p.nx[n + 1], 1)) for n in range(N)]
    lam = [np.zeros(max(2 * nb, 1)) for nb in p.nb]; t = [np.zeros(max(2 * nb, 1)) for nb in p.nb]
    idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
    kk = C.c_int(0)
    pa = ptr_array
    b, q, r, lb, ub = v1
    arrs = [pa(A), pa(B), pa(b), pa(Q), pa(S), pa(R), pa(q), pa(r), pa(lb), pa(ub)] + [pa(v) for v in v2] + [pa(x), pa(u), pa(pi), pa(lam), pa(t)]
    fn = L.orc_fortran_order_d_ip_then_kkt_new_rhs
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_int] + [C.c_void_p] * 4 + [C.c_void_p] * 20
    status = fn(C.byref(kk), k_max, mu0, mu_tol, N, int_array(p.nx), int_array(p.nu), int_array(p.nb), pa(idxb), *arrs)
    return dict(status=status, kk=kk.value, x=[x[n][:p.nx[n]].copy() for n in range(N + 1)],
                u=[u[n][:p.nu[n]].copy() for n in range(N)], pi=[pi[n][:p.nx[n + 1]].copy() for n in range(N)],
                lam=[lam[n][:2 * p.nb[n]].copy() for n in range(N + 1)], t=[t[n][:2 * p.nb[n]].copy() for n in range(N + 1)])


def ric(p: Ocp, mode: str = "sv"):
    """Unconstrained LQCP by the oracle: mode 'sv' (factor+solve) or 'trf_trs'."""
    L, N = lib(), p.N
    A = [_f(M) for M in p.A]; B = [_f(M) for M in p.B]; Q = [_f(M) for M in p.Q]; S = [_f(M) for M in p.S]; R = [_f(M) for M in p.R]
    c = np.ascontiguousarray
    b = [c(v) for v in p.b]; q = [c(v) for v in p.q]; r = [c(v) for v in p.r]
    x = [np.zeros(max(n, 1)) for n in p.nx]; u = [np.zeros(max(n, 1)) for n in p.nu[:N]]
    pi = [np.zeros(max(p.nx[n + 1], 1)) for n in range(N)]
    pa = ptr_array
    arrs = [pa(A), pa(B), pa(b), pa(Q), pa(S), pa(R), pa(q), pa(r), pa(x), pa(u), pa(pi)]
    L.orc_fortran_order_d_ric(0 if mode == "sv" else 1, N, int_array(p.nx), int_array(p.nu), *arrs)
    return dict(x=[x[n][:p.nx[n]].copy() for n in range(N + 1)], u=[u[n][:p.nu[n]].copy() for n in range(N)],
                pi=[pi[n][:p.nx[n + 1]].copy() for n in range(N)])


def tree_ric(t):
    """orc_tree_ric_sv (oracle/ric_oracle.c): Riccati factor+solve over a scenario tree; t is a hpmpc_b200.tree.TreeOcp."""
    L = lib()
    topo = t.topo
    Nn = topo["Nn"]
    L.orc_tree_ric_sv.restype = None
    L.orc_tree_ric_sv.argtypes = [C.c_int] + [C.c_void_p] * 9
    BAbt, RSQ = [], []
    for n in range(Nn):
        nx, nu = t.nx[n], t.nu[n]
        nz = nx + nu + 1
        H = np.zeros((nz, nx + nu + 1), order="F")
        H[:nu, :nu] = t.R[n]; H[nu:nu + nx, :nu] = t.S[n].T; H[nu:nu + nx, nu:nu + nx] = t.Q[n]
        H[nu + nx, :nu] = t.r[n]; H[nu + nx, nu:nu + nx] = t.q[n]
        RSQ.append(np.asfortranarray(H[:, :max(nx + nu, 1)]))
        if n == 0:
            BAbt.append(np.zeros((1, 1), order="F"))
        else:
            d = topo["dad"][n]
            nzd = t.nx[d] + t.nu[d] + 1
            M = np.zeros((nzd, max(nx, 1)), order="F")
            M[:t.nu[d], :nx] = t.B[n].T; M[t.nu[d]:t.nu[d] + t.nx[d], :nx] = t.A[n].T; M[nzd - 1, :nx] = t.b[n]
            BAbt.append(M)
    ux = [np.zeros(max(t.nx[n] + t.nu[n], 1)) for n in range(Nn)]
    pi = [np.zeros(max(t.nx[n], 1)) for n in range(Nn)]
    pa = ptr_array
    L.orc_tree_ric_sv(Nn, int_array(topo["dad"]), int_array(topo["first_kid"]), int_array(topo["nkids"]), int_array(t.nx), int_array(t.nu),
                      pa(BAbt), pa(RSQ), pa(ux), pa(pi))
    return dict(u=[ux[n][:t.nu[n]].copy() for n in range(Nn)], x=[ux[n][t.nu[n]:t.nu[n] + t.nx[n]].copy() for n in range(Nn)],
                pi=[pi[n][:t.nx[n]].copy() if n > 0 else np.zeros(0) for n in range(Nn)])


def tree_ric_trf_trs(t):
    """orc_tree_ric_trf_trs: factorize the tree, then solve with the stored factors for the b, [r q] of the problem."""
    L = lib()
    Nn = t.topo["Nn"]
    L.orc_tree_ric_trf_trs.restype = None
    L.orc_tree_ric_trf_trs.argtypes = [C.c_int] + [C.c_void_p] * 7
    BAbt, RSQ = _tree_dense(t)
    ux = [np.zeros(max(t.nx[n] + t.nu[n], 1)) for n in range(Nn)]
    pi = [np.zeros(max(t.nx[n], 1)) for n in range(Nn)]
    L.orc_tree_ric_trf_trs(Nn, int_array(t.topo["dad"]), int_array(t.nx), int_array(t.nu), ptr_array(BAbt), ptr_array(RSQ), ptr_array(ux), ptr_array(pi))
    return dict(u=[ux[n][:t.nu[n]].copy() for n in range(Nn)], x=[ux[n][t.nu[n]:t.nu[n] + t.nx[n]].copy() for n in range(Nn)],
                pi=[pi[n][:t.nx[n]].copy() if n > 0 else np.zeros(0) for n in range(Nn)])


def _tree_dense(t):
    """Node-indexed dense matrices of a TreeOcp in the oracle's formats (BAbt[k]: edge into node k)."""
    topo = t.topo
    BAbt, RSQ = [], []
    for n in range(topo["Nn"]):
        nx, nu = t.nx[n], t.nu[n]
        nz = nx + nu + 1
        H = np.zeros((nz, nx + nu + 1), order="F")
        H[:nu, :nu] = t.R[n]; H[nu:nu + nx, :nu] = t.S[n].T; H[nu:nu + nx, nu:nu + nx] = t.Q[n]
        H[nu + nx, :nu] = t.r[n]; H[nu + nx, nu:nu + nx] = t.q[n]
        RSQ.append(np.asfortranarray(H[:, :max(nx + nu, 1)]))
        if n == 0:
            BAbt.append(np.zeros((1, 1), order="F"))
        else:
            d = topo["dad"][n]
            nzd = t.nx[d] + t.nu[d] + 1
            M = np.zeros((nzd, max(nx, 1)), order="F")
            M[:t.nu[d], :nx] = t.B[n].T; M[t.nu[d]:t.nu[d] + t.nx[d], :nx] = t.A[n].T; M[nzd - 1, :nx] = t.b[n]
            BAbt.append(M)
    return BAbt, RSQ


def tree_ipm(t, *, k_max=40, mu0=2.0, mu_tol=1e-8, alpha_min=1e-8, warm_start=0):
    """orc_tree_ip2_res_mpc_hard (oracle/ric_oracle.c): box-constrained IPM over a scenario tree (TreeOcp with bounds)."""
    L = lib()
    topo = t.topo
    Nn = topo["Nn"]
    L.orc_tree_ip2_res_mpc_hard.restype = C.c_int
    L.orc_tree_ip2_res_mpc_hard.argtypes = [C.c_int] + [C.c_void_p] * 8 + [C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_double, C.c_int] + [C.c_void_p] * 5
    BAbt, RSQ = _tree_dense(t)
    nb = list(t.nb) if t.nb else [0] * Nn
    idxb = [np.ascontiguousarray(t.idxb[n], dtype=np.int32) if nb[n] else np.zeros(1, dtype=np.int32) for n in range(Nn)]
    d = [np.concatenate([t.lb[n], t.ub[n]]).astype(np.float64) if nb[n] else np.zeros(1) for n in range(Nn)]
    ux = [np.zeros(max(t.nx[n] + t.nu[n], 1)) for n in range(Nn)]
    pi = [np.zeros(max(t.nx[n], 1)) for n in range(Nn)]
    lam = [np.zeros(max(2 * nb[n], 1)) for n in range(Nn)]; tt = [np.zeros(max(2 * nb[n], 1)) for n in range(Nn)]
    stat = np.zeros(5 * k_max + 5); kk = C.c_int(0)
    pa = ptr_array
    status = L.orc_tree_ip2_res_mpc_hard(Nn, int_array(topo["dad"]), int_array(t.nx), int_array(t.nu), int_array(nb), pa(idxb), pa(BAbt), pa(RSQ), pa(d),
                                         C.byref(kk), k_max, mu0, mu_tol, alpha_min, warm_start, stat.ctypes.data, pa(ux), pa(pi), pa(lam), pa(tt))
    return dict(status=status, kk=kk.value, u=[ux[n][:t.nu[n]].copy() for n in range(Nn)],
                x=[ux[n][t.nu[n]:t.nu[n] + t.nx[n]].copy() for n in range(Nn)],
                pi=[pi[n][:t.nx[n]].copy() if n > 0 else np.zeros(0) for n in range(Nn)],
                lam=[lam[n][:2 * nb[n]].copy() for n in range(Nn)], t=[tt[n][:2 * nb[n]].copy() for n in range(Nn)],
                stat=stat[:5 * kk.value].reshape(-1, 5).copy())


class TreeTimed:
    """A tree problem packed once into the oracle's formats, so that a timing loop measures only the C solver
    (bench.py --impl reference --workload tree / tree_ipm; ctypes releases the GIL during the call)."""

    def __init__(self, t, ipm: bool, k_max=40):
        self.L = lib()
        topo = t.topo
        self.Nn = Nn = topo["Nn"]
        self.ipm, self.k_max = ipm, k_max
        self.BAbt, self.RSQ = _tree_dense(t)
        nb = list(t.nb) if (ipm and t.nb) else [0] * Nn
        self.idxb = [np.ascontiguousarray(t.idxb[n], dtype=np.int32) if nb[n] else np.zeros(1, dtype=np.int32) for n in range(Nn)]
        self.d = [np.concatenate([t.lb[n], t.ub[n]]).astype(np.float64) if nb[n] else np.zeros(1) for n in range(Nn)]
        self.ux = [np.zeros(max(t.nx[n] + t.nu[n], 1)) for n in range(Nn)]
        self.pi = [np.zeros(max(t.nx[n], 1)) for n in range(Nn)]
        self.lam = [np.zeros(max(2 * nb[n], 1)) for n in range(Nn)]; self.tt = [np.zeros(max(2 * nb[n], 1)) for n in range(Nn)]
        self.stat = np.zeros(5 * k_max + 5); self.kk = C.c_int(0)
        pa = ptr_array
        self.a = dict(dad=int_array(topo["dad"]), fk=int_array(topo["first_kid"]), nk=int_array(topo["nkids"]), nx=int_array(t.nx), nu=int_array(t.nu),
                      nb=int_array(nb), idxb=pa(self.idxb), B=pa(self.BAbt), Q=pa(self.RSQ), d=pa(self.d), ux=pa(self.ux), pi=pa(self.pi),
                      lam=pa(self.lam), tt=pa(self.tt))
        self.L.orc_tree_ric_sv.restype = None
        self.L.orc_tree_ric_sv.argtypes = [C.c_int] + [C.c_void_p] * 9
        self.L.orc_tree_ip2_res_mpc_hard.restype = C.c_int
        self.L.orc_tree_ip2_res_mpc_hard.argtypes = [C.c_int] + [C.c_void_p] * 8 + [C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_double, C.c_int] + [C.c_void_p] * 5

    def run(self):
        a = self.a
        if self.ipm:
            return self.L.orc_tree_ip2_res_mpc_hard(self.Nn, a["dad"], a["nx"], a["nu"], a["nb"], a["idxb"], a["B"], a["Q"], a["d"], C.byref(self.kk), self.k_max,
                                                    2.0, 1e-8, 1e-8, 0, self.stat.ctypes.data, a["ux"], a["pi"], a["lam"], a["tt"])
        self.L.orc_tree_ric_sv(self.Nn, a["dad"], a["fk"], a["nk"], a["nx"], a["nu"], a["B"], a["Q"], a["ux"], a["pi"])
        return 0


# ------------------------------------------------------------------------------------------- CPU timing harness
def _harness():
    L = lib()
    L.ref_harness_ric_sv.restype = C.c_double
    L.ref_harness_ric_sv.argtypes = [C.c_char_p, C.c_int, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long,
                                     C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long]
    L.ref_harness_ipm.restype = C.c_double
    L.ref_harness_ipm.argtypes = [C.c_char_p, C.c_int, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
    return L


class RefSample:
    """A bounded sample of the batch workload in the REFERENCE's own input formats, for timing it on the host cores:
    panel-major hpBAbt/hpRSQrq for d_back_ric_rec_sv_tv_res, column-major stage arrays for fortran_order_d_ip_ocp_hard_tv.
    Instance i is the same problem as instance i of hpmpc_b200.batchgen.BatchSpec."""

    def __init__(self, spec, n_inst: int, first: int = 0, want=("pm", "cm")):
        from hpmpc_b200.capi import to_pmat, _rup
        self.spec, self.n_inst = spec, n_inst
        p = spec.base
        N = p.N
        x01, x02, qs, rs = spec.scalars(n_inst, first)
        if "pm" in want:
            self._build_pm(p, N, spec, n_inst, x01, x02, qs, rs)
        if "cm" in want:
            self._build_cm(p, N, spec, n_inst, x01, x02, qs, rs)

    def _build_pm(self, p, N, spec, n_inst, x01, x02, qs, rs):
        from hpmpc_b200.capi import to_pmat, _rup
        # ---- panel-major (Riccati)
        parts, self.off_pm, pos = [], [], 0
        iq, ir, ib0 = [], [], []
        href = HpmpcLib.__new__(HpmpcLib)
        BAbt, RSQ = HpmpcLib._pm_problem(href, p)
        for n in range(N):
            a = BAbt[n]; self.off_pm.append(pos)
            if n == 0 and spec.x0_elim:
                nux, nx1 = p.nu[0] + p.nx[0], p.nx[1]
                sda = _rup(nx1, 2)
                ib0 = [pos + (nux // 4) * 4 * sda + nux % 4 + 4 * j for j in range(nx1)]
            parts.append(a); pos += _rup(a.size, 8)
        for n in range(N + 1):
            a = RSQ[n]; self.off_pm.append(pos)
            nu, nx = p.nu[n], p.nx[n]; sda = _rup(nu + nx, 2)
            ir += [pos + (i // 4) * 4 * sda + i % 4 + 4 * i for i in range(nu)]
            iq += [pos + (i // 4) * 4 * sda + i % 4 + 4 * i for i in range(nu, nu + nx)]
            parts.append(a); pos += _rup(a.size, 8)
        self.pm_stride = pos
        base = np.zeros(pos)
        for o, a in zip(self.off_pm, parts):
            base[o:o + a.size] = a
        from hpmpc_b200.capi import aligned_zeros
        pm = aligned_zeros(n_inst * pos).reshape(n_inst, pos)
        pm[:] = base[None, :]
        pm[:, iq] = qs[:, None]; pm[:, ir] = rs[:, None]
        if spec.x0_elim:
            pm[:, ib0] = x01[:, None] * spec.A_cols[None, :, 0] + x02[:, None] * spec.A_cols[None, :, 1] + 0.1
        self.pm = pm

    def _build_cm(self, p, N, spec, n_inst, x01, x02, qs, rs):
        from hpmpc_b200.capi import _rup, aligned_zeros
        # ---- column-major stage arrays (IPM)
        parts, offs, pos = [], np.zeros((10, N + 1), dtype=np.int64), 0
        iq, ir, ib0 = [], [], []
        lists = (p.A, p.B, p.b, p.Q, p.S, p.R, p.q, p.r, p.lb, p.ub)
        for k, Lk in enumerate(lists):
            for n, M in enumerate(Lk):
                a = np.asfortranarray(M, dtype=np.float64).ravel(order="F")
                offs[k, n] = pos
                if k == 3: iq += [pos + i * (p.nx[n] + 1) for i in range(p.nx[n])]
                if k == 5: ir += [pos + i * (p.nu[n] + 1) for i in range(p.nu[n])]
                if k == 2 and n == 0 and spec.x0_elim: ib0 = [pos + j for j in range(p.nx[1])]
                parts.append((pos, a)); pos += _rup(max(a.size, 1), 8)
        self.cm_stride, self.off_cm = pos, offs
        base = np.zeros(pos)
        for o, a in parts:
            base[o:o + a.size] = a
        cm = aligned_zeros(n_inst * pos).reshape(n_inst, pos)
        cm[:] = base[None, :]
        cm[:, iq] = qs[:, None]; cm[:, ir] = rs[:, None]
        if spec.x0_elim:
            cm[:, ib0] = x01[:, None] * spec.A_cols[None, :, 0] + x02[:, None] * spec.A_cols[None, :, 1] + 0.1
        self.cm = cm

    def time_ric_sv(self, kind="avx2", n_threads=None, n_pass=1, want_out=False):
        p = self.spec.base
        L = _harness()
        n_threads = n_threads or os.cpu_count()
        off = np.asarray(self.off_pm, dtype=np.int64)
        n_ux, n_pi = sum(p.nx) + sum(p.nu), sum(p.nx[1:])
        ux = np.zeros((self.n_inst, n_ux)) if want_out else None
        pi = np.zeros((self.n_inst, n_pi)) if want_out else None
        sec = L.ref_harness_ric_sv((REF_AVX2 if kind == "avx2" else REF_C99).encode(), n_threads, self.n_inst, n_pass, p.N,
                                   int_array(p.nx), int_array(p.nu), self.pm.ctypes.data, self.pm_stride, off.ctypes.data,
                                   ux.ctypes.data if want_out else None, n_ux, pi.ctypes.data if want_out else None, n_pi)
        if sec < 0:
            raise RuntimeError("reference harness failed")
        return sec, ux, pi

    def time_ipm(self, kind="avx2", n_threads=None, n_pass=1, k_max=40, mu0=2.0, mu_tol=1e-8, want_out=False):
        p = self.spec.base
        L = _harness()
        n_threads = n_threads or os.cpu_count()
        idxb = np.concatenate([np.asarray(v, dtype=np.int32) for v in p.idxb] + [np.zeros(1, dtype=np.int32)])
        n_ux = sum(p.nx) + sum(p.nu)
        kk = np.zeros(self.n_inst, dtype=np.int32); st = np.zeros(self.n_inst, dtype=np.int32)
        ux = np.zeros((self.n_inst, n_ux)) if want_out else None
        sec = L.ref_harness_ipm((REF_AVX2 if kind == "avx2" else REF_C99).encode(), n_threads, self.n_inst, n_pass, p.N,
                                int_array(p.nx), int_array(p.nu), int_array(p.nb), idxb.ctypes.data, k_max, mu0, mu_tol,
                                self.cm.ctypes.data, self.cm_stride, np.ascontiguousarray(self.off_cm).ctypes.data,
                                kk.ctypes.data, st.ctypes.data, ux.ctypes.data if want_out else None, n_ux)
        if sec < 0:
            raise RuntimeError("reference harness failed")
        return sec, kk, st, ux

    def solve_ipm_full(self, kind="c99", n_threads=None, k_max=40, mu0=2.0, mu_tol=1e-8):
        """Every instance of the sample through the reference's fortran_order_d_ip_ocp_hard_tv on the host cores; returns
        kk, status, ux, pi, lam ([lb ub] per stage) and inf_norm_res of every instance (the checker of the full-batch parity tests)."""
        p = self.spec.base
        L = lib()
        fn = L.ref_harness_ipm_full
        fn.restype = C.c_double
        fn.argtypes = [C.c_char_p, C.c_int, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double,
                       C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p]
        n_threads = n_threads or os.cpu_count()
        idxb = np.concatenate([np.asarray(v, dtype=np.int32) for v in p.idxb] + [np.zeros(1, dtype=np.int32)])
        n_ux, n_pi, n_lam = sum(p.nx) + sum(p.nu), sum(p.nx[1:]), 2 * sum(p.nb)
        n = self.n_inst
        kk = np.zeros(n, dtype=np.int32); st = np.zeros(n, dtype=np.int32)
        ux, pi, lam, res = np.zeros((n, n_ux)), np.zeros((n, n_pi)), np.zeros((n, max(n_lam, 1))), np.zeros((n, 4))
        sec = fn((REF_AVX2 if kind == "avx2" else REF_C99).encode(), n_threads, n, p.N, int_array(p.nx), int_array(p.nu), int_array(p.nb),
                 idxb.ctypes.data, k_max, mu0, mu_tol, self.cm.ctypes.data, self.cm_stride, np.ascontiguousarray(self.off_cm).ctypes.data,
                 kk.ctypes.data, st.ctypes.data, ux.ctypes.data, n_ux, pi.ctypes.data, n_pi, lam.ctypes.data, max(n_lam, 1), res.ctypes.data)
        if sec < 0:
            raise RuntimeError("reference harness failed")
        return dict(sec=sec, kk=kk, status=st, ux=ux, pi=pi, lam=lam, inf_norm_res=res)


def single_newton_step(p: Ocp, ux0, pi0, lam0, t0, *, k_max=1, mu0=1e-3):
    """orc_fortran_order_single_newton_step (oracle/ric_oracle.c): k_max Newton steps from the given iterate."""
    L, N = lib(), p.N
    A = [_f(M) for M in p.A]; B = [_f(M) for M in p.B]; Q = [_f(M) for M in p.Q]; S = [_f(M) for M in p.S]; R = [_f(M) for M in p.R]
    c = np.ascontiguousarray
    b = [c(v) for v in p.b]; q = [c(v) for v in p.q]; r = [c(v) for v in p.r]; lb = [c(v) for v in p.lb]; ub = [c(v) for v in p.ub]
    x = [np.zeros(max(n, 1)) for n in p.nx]; u = [np.zeros(max(n, 1)) for n in p.nu[:N]]
    pi = [np.zeros(max(p.nx[n + 1], 1)) for n in range(N)]
    lam = [np.zeros(max(2 * nb, 1)) for nb in p.nb]; t = [np.zeros(max(2 * nb, 1)) for nb in p.nb]
    pad = lambda v: c(np.asarray(v, dtype=np.float64)) if len(v) else np.zeros(1)
    a0 = [[pad(v) for v in ux0], [pad(v) for v in pi0] + [np.zeros(1)], [pad(v) for v in lam0], [pad(v) for v in t0]]
    idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
    res = np.zeros(8); stat = np.zeros(5 * k_max + 5); kk = C.c_int(0)
    fn = L.orc_fortran_order_single_newton_step
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_double, C.c_int] + [C.c_void_p] * 25
    pa = ptr_array
    arrs = [pa(A), pa(B), pa(b), pa(Q), pa(S), pa(R), pa(q), pa(r), pa(lb), pa(ub), pa(x), pa(u), pa(pi), pa(lam), pa(t)]
    status = fn(C.byref(kk), k_max, mu0, N, int_array(p.nx), int_array(p.nu), int_array(p.nb), pa(idxb), *arrs, res.ctypes.data, stat.ctypes.data,
                *[pa(a) for a in a0])
    return dict(status=status, kk=kk.value, x=[x[n][:p.nx[n]].copy() for n in range(N + 1)], u=[u[n][:p.nu[n]].copy() for n in range(N)],
                pi=[pi[n][:p.nx[n + 1]].copy() for n in range(N)], lam=[lam[n][:2 * p.nb[n]].copy() for n in range(N + 1)],
                t=[t[n][:2 * p.nb[n]].copy() for n in range(N + 1)], inf_norm_res=res[:4].copy(), stat=stat[:5 * kk.value].reshape(-1, 5).copy())
