"""Re-solve of the IPM's last KKT system with a new right-hand side.

Reference: d_kkt_solve_new_rhs_res_mpc_hard_tv (mpc_solvers/d_ip2_res_hard.c:1922) called after d_ip2_res_mpc_hard_tv on the
same work memory (test_problems/test_d_ip_hard.c:1040).  CPU: the oracle (oracle/ric_oracle.c: orc_kkt_solve_new_rhs) against
golden vectors the REAL reference produced (tests/golden/make_golden_kkt.py) and, when it is compiled here, against the
reference itself.  GPU: hpmpc_b200_d_ip2_res_mpc_hard_kkt_batch + hpmpc_b200_d_kkt_solve_new_rhs_batch through the C ABI
against the oracle and the golden vectors.

Tolerances: 1e-9 relative on u, x, pi, t.  lam is compared at 1e-6: the step is dlam = -t_inv (lam dt + res_m) with t_inv up to
1e8 on active constraints at the IPM's last iterate, so rounding differences of 1e-16 in dux show up as 1e-8 in lam (the
oracle and the reference's own C99 build already differ by 1.5e-8 there).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_kkt as G  # noqa: E402
from hpmpc_b200 import capi  # noqa: E402
from oracle import api  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "golden_kkt_v1.npz"))
TOL, TOL_LAM = 1e-9, 1e-6
cat = G.cat


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b))))) if a.size else 0.0


def check(got, want, what=""):
    for f in ("u", "x", "pi", "t"):
        assert rel(cat(got[f]), cat(want[f])) < TOL, (what, f)
    assert rel(cat(got["lam"]), cat(want["lam"])) < TOL_LAM, (what, "lam")


@pytest.mark.parametrize("case", list(G.CASES))
def test_oracle_new_rhs_matches_reference_golden(case):
    p = G.build(case)
    o = api.ipm_then_kkt_new_rhs(p, G.perturbed(p), k_max=G.K_MAX, mu0=G.MU0, mu_tol=G.MU_TOL)
    assert [o["kk"], o["status"]] == list(GOLD[f"{case}/kk"])
    check(o, {f: GOLD[f"{case}/{f}"] for f in ("u", "x", "pi", "lam", "t")}, case)


@pytest.mark.skipif(not api.have_reference(), reason="reference not compiled here")
def test_oracle_new_rhs_matches_live_reference():
    ref = api.reference("c99")
    for case, seed in (("ms_8_3_10", 11), ("ms_12_5_30", 12)):
        p = G.build(case)
        p2 = G.perturbed(p, seed=seed, scale=0.2)
        check(api.ipm_then_kkt_new_rhs(p, p2), ref.ip2_then_kkt_new_rhs(p, p2), case)


def test_oracle_same_rhs_reproduces_the_ipm_solution():
    """with the old right-hand side the Newton step from the backed-up iterate lands on the IPM's solution (to the IPM's own accuracy)"""
    p = G.build("ms_8_3_10")
    o, s = api.ipm_then_kkt_new_rhs(p, p), api.ipm(p)
    assert rel(cat(o["u"]), cat(s["u"])) < 1e-4 and rel(cat(o["x"]), cat(s["x"])) < 1e-4


def test_oracle_new_rhs_needs_a_phase2_factor():
    """mu_tol above 1e-5: the IPM stops in phase 1 and leaves no backup (the reference would read unset memory): status -10"""
    p = G.build("ms_4_2_5")
    o = api.ipm_then_kkt_new_rhs(p, G.perturbed(p), mu_tol=1e-3)
    assert o["status"] == -10


def test_kkt_state_stride_and_no_cpu_fallback():
    """host-only handle (device < 0): the state layout can be queried, the solve entry points refuse (there is no CPU path)"""
    import ctypes as C
    L = capi.product()
    L.hpmpc_b200_kkt_state_stride.restype = C.c_longlong
    L.hpmpc_b200_kkt_state_stride.argtypes = [C.c_void_p]
    L.hpmpc_b200_d_ip2_res_mpc_hard_kkt_batch.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                                         C.c_double, C.c_int] + [C.c_void_p] * 7
    L.hpmpc_b200_d_kkt_solve_new_rhs_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 8
    for case in ("ms_8_3_10", "ms_4_2_5"):
        p = G.build(case)
        h = capi.BatchOcp(p, device=-1)
        try:
            even = lambda v: (v + 1) & ~1
            ks = L.hpmpc_b200_kkt_state_stride(h.h)
            assert ks == even(h.sz.ipm_work_stride) + h.sz.ux_stride + h.sz.pi_stride + 4 * even(h.sz.nbtot) + 2
            buf = np.zeros(8)
            a = buf.ctypes.data
            assert L.hpmpc_b200_d_ip2_res_mpc_hard_kkt_batch(h.h, 1, a, 10, 2.0, 1e-8, 1e-8, 0, a, a, a, a, a, a, None) == -4
            assert L.hpmpc_b200_d_kkt_solve_new_rhs_batch(h.h, 1, a, a, a, a, a, a, a, None) == -4
        finally:
            h.close()


# ------------------------------------------------------------------------------------------------------------ GPU
def _gpu_pair(h, ps, p2s, k_max=G.K_MAX, mu_tol=G.MU_TOL):
    import ctypes as C
    import torch
    L = capi.product()
    L.hpmpc_b200_kkt_state_stride.restype = C.c_longlong
    L.hpmpc_b200_kkt_state_stride.argtypes = [C.c_void_p]
    L.hpmpc_b200_d_ip2_res_mpc_hard_kkt_batch.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                                         C.c_double, C.c_int] + [C.c_void_p] * 7
    L.hpmpc_b200_d_kkt_solve_new_rhs_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 8
    n = len(ps)
    ks = L.hpmpc_b200_kkt_state_stride(h.h)
    d_in = torch.from_numpy(np.stack([h.pack(p) for p in ps])).cuda()
    d_in2 = torch.from_numpy(np.stack([h.pack(p) for p in p2s])).cuda()
    z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
    ux, pi, lam, t, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * k_max)
    kkt = torch.full((n, ks), float("nan"), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = L.hpmpc_b200_d_ip2_res_mpc_hard_kkt_batch(h.h, n, d_in.data_ptr(), k_max, G.MU0, mu_tol, 1e-8, 0, ux.data_ptr(), pi.data_ptr(),
                                                   lam.data_ptr(), t.data_ptr(), info.data_ptr(), kkt.data_ptr(), st)
    assert rc == 0
    outs = []
    for _ in range(2):          # the state is reusable: the second solve must give the same bits
        ux2, pi2, lam2, t2, info2 = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6)
        rc = L.hpmpc_b200_d_kkt_solve_new_rhs_batch(h.h, n, d_in2.data_ptr(), kkt.data_ptr(), ux2.data_ptr(), pi2.data_ptr(), lam2.data_ptr(),
                                                    t2.data_ptr(), info2.data_ptr(), st)
        assert rc == 0
        torch.cuda.synchronize()
        outs.append([a.cpu().numpy() for a in (ux2, pi2, lam2, t2, info2)])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    return info.cpu().numpy(), outs[0]


def _split(h, out, i):
    u, x = h.split_ux(out[0][i])
    return dict(u=u, x=x, pi=h.split_pi(out[1][i]), lam=h.split_lam(out[2][i]), t=h.split_lam(out[3][i]))


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(G.CASES))
def test_gpu_new_rhs_matches_reference_golden(case):
    p = G.build(case)
    h = capi.BatchOcp(p)
    try:
        info, out = _gpu_pair(h, [p], [G.perturbed(p)])
        assert [int(info[0, 0]), int(info[0, 1])] == list(GOLD[f"{case}/kk"]) and out[4][0, 1] == 0
        check(_split(h, out, 0), {f: GOLD[f"{case}/{f}"] for f in ("u", "x", "pi", "lam", "t")}, case)
    finally:
        h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,n_inst", [((8, 3, 10), 70), ((24, 11, 50), 24), ((6, 2, 7), 300)])
def test_gpu_new_rhs_batch_vs_oracle(shape, n_inst):
    """size-specialised sweeps (8,3) and (24,11), generic sweeps (6,2); batches larger than one wave of warps are ragged"""
    from hpmpc_b200 import problems
    xis = problems.instance_xi(n_inst, first=9)
    ps = [problems.mass_spring_ocp(*shape, bounds=True, xi=tuple(x)) for x in xis]
    p2s = [G.perturbed(p, seed=100 + i) for i, p in enumerate(ps)]
    h = capi.BatchOcp(ps[0])
    try:
        info, out = _gpu_pair(h, ps, p2s)
        assert np.all(out[4][:, 1] == 0)
        for i in range(0, n_inst, max(1, n_inst // 24)):
            o = api.ipm_then_kkt_new_rhs(ps[i], p2s[i])
            assert int(info[i, 0]) == o["kk"]
            check(_split(h, out, i), o, i)
    finally:
        h.close()


@pytest.mark.gpu
def test_gpu_new_rhs_without_phase2_factor_reports_it():
    p = G.build("ms_4_2_5")
    h = capi.BatchOcp(p)
    try:
        _, out = _gpu_pair(h, [p], [G.perturbed(p)], mu_tol=1e-3)
        assert out[4][0, 1] == -10
    finally:
        h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ms_8_3_10", "ms_24_11_50", "ms_8_3_10_free_x0"])
def test_gpu_compat_symbols_chain_like_the_reference(case):
    """the reference's own pair of calls (d_ip2_res_mpc_hard_tv, then d_kkt_solve_new_rhs_res_mpc_hard_tv on the same work memory,
    test_problems/test_d_ip_hard.c:1040) against libhpmpc_b200.so: same symbols, same arguments, reference golden vectors"""
    prod = capi.HpmpcLib(capi.PRODUCT_LIB)
    p = G.build(case)
    r = prod.ip2_then_kkt_new_rhs(p, G.perturbed(p), k_max=G.K_MAX, mu0=G.MU0, mu_tol=G.MU_TOL)
    assert [r["kk"], r["status"]] == list(GOLD[f"{case}/kk"])
    check(r, {f: GOLD[f"{case}/{f}"] for f in ("u", "x", "pi", "lam", "t")}, case)
