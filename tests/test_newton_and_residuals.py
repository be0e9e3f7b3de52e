"""Single Newton step and stand-alone residuals (row a6).

Reference: d_ip2_res_mpc_hard_tv_single_newton_step, mpc_solvers/d_ip2_res_hard.c:1348 (high level
fortran_order_d_ip_ocp_hard_tv_single_newton_step, interfaces/c/fortran_order_interface.c:695);
d_res_res_mpc_hard_tv, mpc_solvers/c99/d_res_ip_res_hard.c:39; d_res_mpc_hard_tv, mpc_solvers/d_res_ip_hard.c:38.
Golden vectors: tests/golden/make_golden_newton.py (the compiled reference)."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import rel_err
from hpmpc_b200 import capi, problems
from oracle import api as oracle

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_newton", os.path.join(HERE, "golden", "make_golden_newton.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)
GOLD = np.load(os.path.join(HERE, "golden", "golden_newton_v1.npz"))
TOL = 1e-9


def _start(name):
    p = mg.build_problem(name)
    g = lambda f: GOLD[f"{name}/{f}"]
    N = p.N
    cut = lambda v, lens: [v[sum(lens[:n]):sum(lens[:n + 1])] for n in range(len(lens))]
    ux0 = cut(g("ux0"), [p.nu[n] + p.nx[n] for n in range(N + 1)])
    pi0 = cut(g("pi0"), [p.nx[n + 1] for n in range(N)])
    lam0 = cut(g("lam0"), [2 * p.nb[n] for n in range(N + 1)])
    t0 = cut(g("t0"), [2 * p.nb[n] for n in range(N + 1)])
    return p, g, ux0, pi0, lam0, t0


@pytest.mark.parametrize("name", list(mg.CASES))
def test_oracle_single_newton_step_matches_reference_golden(name):
    p, g, ux0, pi0, lam0, t0 = _start(name)
    c = mg.CASES[name]
    r = oracle.single_newton_step(p, ux0, pi0, lam0, t0, k_max=c["k_max"], mu0=c["mu0"])
    assert (r["kk"], r["status"]) == (int(g("kk")), int(g("status")))
    for f in ("u", "x", "pi", "lam", "t"):
        assert rel_err([mg.cat(r[f])], [g(f)]) < 1e-12, f
    assert rel_err([r["stat"]], [g("stat")]) < 1e-12


@pytest.fixture(scope="module")
def prod():
    import torch
    assert torch.cuda.is_available()
    return capi.HpmpcLib(capi.PRODUCT_LIB)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mg.CASES))
def test_compat_single_newton_step_matches_golden(prod, name):
    p, g, ux0, pi0, lam0, t0 = _start(name)
    c = mg.CASES[name]
    r = prod.single_newton_step(p, ux0, pi0, lam0, t0, k_max=c["k_max"], mu0=c["mu0"])
    assert (r["kk"], r["status"]) == (int(g("kk")), int(g("status")))
    for f in ("u", "x", "pi", "lam", "t"):
        assert rel_err([mg.cat(r[f])], [g(f)]) < TOL, f
    assert rel_err([r["stat"]], [g("stat")]) < 1e-9
    gi = g("inf_norm_res")
    for j in range(4):
        assert abs(r["inf_norm_res"][j] - gi[j]) <= 1e-9 * max(1.0, abs(gi[j])) + 1e-12, (j, r["inf_norm_res"], gi)


@pytest.mark.gpu
def test_batched_single_newton_step_vs_oracle():
    import torch
    from hpmpc_b200.batchgen import BatchSpec
    L = capi.product()
    L.hpmpc_b200_d_ip2_res_mpc_hard_single_newton_step_batch.argtypes = [capi.C.c_void_p, capi.C.c_longlong, capi.C.c_void_p, capi.C.c_int,
                                                                        capi.C.c_double, capi.C.c_double] + [capi.C.c_void_p] * 6
    for cfg, n in ((dict(nx=8, nu=3, N=10, bounds=True), 300), ("cfg3", 12)):
        spec = BatchSpec(cfg); h = spec.h
        blk = spec.torch_batch(n, first=21)
        k0, k_max, mu0 = 3, 2, 1e-3
        z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
        ux, pi, lam, t, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * k0)
        st = torch.cuda.current_stream().cuda_stream
        assert L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, blk.data_ptr(), k0, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(), pi.data_ptr(), lam.data_ptr(),
                                                     t.data_ptr(), info.data_ptr(), st) == 0
        torch.cuda.synchronize()
        s_ux, s_pi, s_lam, s_t = (v.cpu().numpy().copy() for v in (ux, pi, lam, t))
        info = z(6 + 5 * k_max)                    # one record of 6 + 5 k_max doubles per instance
        assert L.hpmpc_b200_d_ip2_res_mpc_hard_single_newton_step_batch(h.h, n, blk.data_ptr(), k_max, mu0, 1e-8, ux.data_ptr(), pi.data_ptr(),
                                                                        lam.data_ptr(), t.data_ptr(), info.data_ptr(), st) == 0
        torch.cuda.synchronize()
        uxh, pih, lamh, th, infoh = (v.cpu().numpy() for v in (ux, pi, lam, t, info))
        for i in (0, 1, n // 2, n - 1):
            p = spec.problem(21 + i)
            u0, x0 = h.split_ux(s_ux[i])
            ux0 = [np.concatenate([u0[k] if k < p.N else np.zeros(0), x0[k]]) for k in range(p.N + 1)]
            o = oracle.single_newton_step(p, ux0, h.split_pi(s_pi[i]), h.split_lam(s_lam[i]), h.split_lam(s_t[i]), k_max=k_max, mu0=mu0)
            assert (int(infoh[i, 0]), int(infoh[i, 1])) == (o["kk"], o["status"])
            u, x = h.split_ux(uxh[i])
            assert rel_err(u, o["u"]) < TOL and rel_err(x, o["x"]) < TOL and rel_err(h.split_pi(pih[i]), o["pi"]) < TOL
            assert rel_err(h.split_lam(lamh[i]), o["lam"]) < TOL and rel_err(h.split_lam(th[i]), o["t"]) < TOL
        h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("general", [False, True])
def test_compat_residual_symbols_match_reference(prod, general):
    """d_res_res_mpc_hard_tv and d_res_mpc_hard_tv at a non-converged iterate, against the reference's own routines."""
    ref = oracle.reference("c99")
    p = problems.general_test_problem(8, 3, 10, xi=(0.1, 0.3, -0.2, 0.5)) if general else problems.make("cfg3", xi=(0.2, -0.1, 0.4, 0.3))
    s = ref.ip2_res_mpc_hard_tv(p, k_max=3)
    for which in ("res_res", "res"):
        a = prod.residuals(p, s["u"], s["x"], s["pi"], s["lam"], s["t"], which)
        b = ref.residuals(p, s["u"], s["x"], s["pi"], s["lam"], s["t"], which)
        assert max(np.max(np.abs(v)) for v in b["rq"]) > 1e-6            # really a non-converged point
        for f in ("rq", "rb", "rd") + (("rm",) if which == "res_res" else ()):
            assert rel_err(a[f], b[f]) < 1e-12, (which, f)
        assert abs(a["mu"] - b["mu"]) <= 1e-14 * max(1.0, abs(b["mu"]))
