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


def ipm(p: Ocp, *, k_max=40, mu0=2.0, mu_tol=1e-8, warm_start=0):
    """orc_fortran_order_d_ip_ocp_hard_tv: same arguments as the reference's fortran_order_d_ip_ocp_hard_tv."""
    L, N = lib(), p.N
    A = [_f(M) for M in p.A]; B = [_f(M) for M in p.B]; Q = [_f(M) for M in p.Q]; S = [_f(M) for M in p.S]; R = [_f(M) for M in p.R]
    c = np.ascontiguousarray
    b = [c(v) for v in p.b]; q = [c(v) for v in p.q]; r = [c(v) for v in p.r]; lb = [c(v) for v in p.lb]; ub = [c(v) for v in p.ub]
    x = [np.zeros(max(n, 1)) for n in p.nx]; u = [np.zeros(max(n, 1)) for n in p.nu[:N]]
    pi = [np.zeros(max(p.nx[n + 1], 1)) for n in range(N)]
    lam = [np.zeros(max(2 * nb, 1)) for nb in p.nb]
    idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
    empty = [np.zeros(1) for _ in range(N + 1)]
    res = np.zeros(8); stat = np.zeros(5 * k_max + 5); kk = C.c_int(0)
    pa = ptr_array
    arrs = [pa(A), pa(B), pa(b), pa(Q), pa(S), pa(R), pa(q), pa(r), pa(lb), pa(ub), pa(empty), pa(empty), pa(empty), pa(empty),
            pa(x), pa(u), pa(pi), pa(lam)]
    pidx = pa(idxb)
    status = L.orc_fortran_order_d_ip_ocp_hard_tv(C.byref(kk), k_max, mu0, mu_tol, N, int_array(p.nx), int_array(p.nu), int_array(p.nb),
                                                   pidx, int_array([0] * (N + 1)), N, warm_start, *arrs, res.ctypes.data, None, stat.ctypes.data)
    return dict(status=status, kk=kk.value, x=[x[n][:p.nx[n]].copy() for n in range(N + 1)],
                u=[u[n][:p.nu[n]].copy() for n in range(N)], pi=[pi[n][:p.nx[n + 1]].copy() for n in range(N)],
                lam=[lam[n][:2 * p.nb[n]].copy() for n in range(N + 1)], inf_norm_res=res[:4].copy(),
                stat=stat[:5 * kk.value].reshape(-1, 5).copy())


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
