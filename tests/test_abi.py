"""CPU tests: the C-ABI library loads and exports every symbol include/*.h declares; host-only layout and packing
logic (no GPU compute is called here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from hpmpc_b200 import capi, problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in ("hpmpc_b200.h", "hpmpc_compat.h", "hpmpc_blasfeo_compat.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", " ".join(l for l in src.splitlines() if not l.strip().startswith("#")))
    return sorted(set(n for n in names if n.startswith(("hpmpc_", "d_", "c_order", "fortran_order"))))


def test_every_declared_symbol_is_exported():
    L = capi.product()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ but not exported by libhpmpc_b200.so"
    assert b"sm_100a" in L.hpmpc_b200_version()


def test_reference_signature_symbols_present():
    for s in ("c_order_d_ip_ocp_hard_tv", "fortran_order_d_ip_ocp_hard_tv", "hpmpc_d_ip_ocp_hard_tv_work_space_size_bytes",
              "d_back_ric_rec_sv_tv_res", "d_back_ric_rec_trf_tv_res", "d_back_ric_rec_trs_tv_res", "d_ip2_res_mpc_hard_tv",
              "d_back_ric_rec_sv_tv_work_space_size_bytes", "d_back_ric_rec_sv_tv_memory_space_size_bytes",
              "d_ip2_res_mpc_hard_tv_work_space_size_bytes"):
        assert s in declared_symbols()


def test_host_only_handle_layout_and_packing():
    p = problems.make("cfg2", xi=(0.1, 0.2, 0.3, 0.4))
    h = capi.BatchOcp(p, device=-1)
    nx, nu, N = 12, 5, 30
    # algorithmic doubles per instance: sum (nux+1) nx1 + sum tri(nux)+nux ; padded to even per block
    expect = 0
    for n in range(N + 1):
        nux = p.nx[n] + p.nu[n]
        if n < N:
            expect += ((nux + 1) * p.nx[n + 1] + 1) // 2 * 2
        expect += (nux * (nux + 1) // 2 + nux + 1) // 2 * 2
    assert h.sz.in_stride == expect
    assert h.sz.in_stride * 8 <= 1.01 * 97968          # within 1 % of the algorithmic byte count of SURVEY.md 8d
    assert h.sz.nzM == nx + nu + 1 and h.sz.nbtot == 0
    blk = h.pack(p)
    o = h.off[3]
    M = blk[o["BAbt"]:o["BAbt"] + (nx + nu + 1) * nx].reshape(nx + nu + 1, nx)
    np.testing.assert_array_equal(M[:nu], p.B[3].T)
    np.testing.assert_array_equal(M[nu:nu + nx], p.A[3].T)
    np.testing.assert_array_equal(M[nu + nx], p.b[3])
    H = blk[o["RSQ"]:]
    for i in range(nx + nu):
        for j in range(i + 1):
            full = np.block([[p.R[3], p.S[3]], [p.S[3].T, p.Q[3]]])
            assert H[i * (i + 1) // 2 + j] == full[i, j]
    # compute entry points refuse a host-only handle (no CPU fallback)
    rc = capi.product().hpmpc_b200_d_back_ric_rec_sv_batch(h.h, 1, None, None, None, None, None)
    assert rc != 0
    h.close()


def test_bounds_layout():
    p = problems.mass_spring_ocp(8, 3, 10, bounds=True)
    h = capi.BatchOcp(p, device=-1)
    assert h.sz.nbtot == sum(p.nb) and h.sz.lam_stride == 2 * sum(p.nb)
    blk = h.pack(p)
    o = h.off[2]
    np.testing.assert_array_equal(blk[o["d"]:o["d"] + p.nb[2]], p.lb[2])
    np.testing.assert_array_equal(blk[o["d"] + p.nb[2]:o["d"] + 2 * p.nb[2]], p.ub[2])
    h.close()


def test_bad_sizes_are_rejected():
    p = problems.mass_spring_ocp(4, 2, 3, bounds=True)
    p.nb[1] = 99
    with pytest.raises(RuntimeError):
        capi.BatchOcp(p, device=-1)
