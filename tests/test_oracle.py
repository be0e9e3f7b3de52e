"""CPU tests: pin the oracle (oracle/ric_oracle.c) against (a) the golden vectors the real reference produced and
(b) the real reference itself when oracle/_ref/libhpmpc_ref_c99.so is available (always in the build container)."""
import numpy as np
import pytest

from conftest import rel_err
from golden_util import case_ids, cat, gold, problem
from hpmpc_b200 import problems
from oracle import api as oracle

TOL = 1e-9      # BASELINE.json: (u, x, pi, lam) within 1e-9 relative, identical iteration counts


@pytest.mark.parametrize("name,inst", case_ids())
def test_oracle_matches_golden(name, inst):
    kind, p = problem(name, inst)
    if kind == "ric":
        o = oracle.ric(p, "sv")
        o2 = oracle.ric(p, "trf_trs")
        for f in ("u", "x", "pi"):
            assert rel_err([cat(o[f])], [gold(name, inst, f)]) < TOL
            assert rel_err([cat(o2[f])], [gold(name, inst, f)]) < TOL
    else:
        o = oracle.ipm(p, k_max=40, mu0=2.0, mu_tol=1e-8)
        assert o["kk"] == int(gold(name, inst, "kk"))
        assert o["status"] == int(gold(name, inst, "status"))
        for f in ("u", "x", "pi", "lam"):
            assert rel_err([cat(o[f])], [gold(name, inst, f)]) < TOL
        assert rel_err([o["stat"]], [gold(name, inst, "stat")]) < 1e-7
        g = gold(name, inst, "inf_norm_res")
        assert abs(o["inf_norm_res"][3] - g[3]) <= 1e-9 * max(1.0, abs(g[3]))


def test_published_probe_trace():
    """SURVEY.md appendix C.2 / reference test_problems/test_d_ip_hard.c: nx=8 nu=3 N=10, kk=7, mu=9.562630e-09."""
    p = problems.mass_spring_ocp(8, 3, 10, bounds=True)
    o = oracle.ipm(p, k_max=20)
    assert o["status"] == 0 and o["kk"] == 7
    assert abs(o["stat"][-1, 4] - 9.562630e-09) < 1e-14
    np.testing.assert_allclose(o["u"][0], [-0.4083206620583, -0.4999999997310, -0.4999999999328], atol=1e-12)
    np.testing.assert_allclose(o["stat"][0], [9.176421e-02, 5.287043e-01, 9.020995e-01, 6.989131e-01, 8.102921e-01], rtol=2e-6)


def test_mass_spring_generator_matches_reference_constants():
    """A[0,0], A[0,4], B[0,0], B[4,0] of the nx=8 nu=3 Ts=0.5 system."""
    A, B = problems.mass_spring_AB(8, 3)
    np.testing.assert_allclose([A[0, 0], A[0, 4], B[0, 0], B[4, 0]], [0.762721, 0.459614, 0.119899, 0.459614], atol=1e-6)


needs_ref = pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref reference build not present")


@needs_ref
@pytest.mark.parametrize("shape", [(4, 2, 5), (8, 3, 10), (12, 5, 30), (6, 3, 3), (10, 1, 7)])
def test_oracle_vs_reference_riccati(shape):
    nx, nu, N = shape
    ref = oracle.reference("c99")
    for inst in range(3):
        xi = tuple(problems.instance_xi(1, first=100 + inst)[0])
        p = problems.mass_spring_ocp(nx, nu, N, xi=xi)
        for mode in ("sv", "trf_trs"):
            r, o = ref.ric(p, mode), oracle.ric(p, mode)
            for f in ("u", "x", "pi"):
                assert rel_err(o[f], r[f]) < TOL, (shape, mode, f)


@needs_ref
@pytest.mark.parametrize("shape", [(4, 2, 5), (8, 3, 10), (12, 5, 15), (24, 11, 50)])
def test_oracle_vs_reference_ipm(shape):
    nx, nu, N = shape
    ref = oracle.reference("c99")
    for inst in range(3 if nx < 24 else 1):
        xi = tuple(problems.instance_xi(1, first=200 + inst)[0])
        p = problems.mass_spring_ocp(nx, nu, N, bounds=True, xi=xi)
        r = ref.ip_ocp_hard_tv(p, order="fortran")
        rc = ref.ip_ocp_hard_tv(p, order="c")
        o = oracle.ipm(p)
        assert (o["kk"], o["status"]) == (r["kk"], r["status"]) == (rc["kk"], rc["status"])
        for f in ("u", "x", "pi", "lam"):
            assert rel_err(o[f], r[f]) < TOL, (shape, f)
            assert rel_err(rc[f], r[f]) < 1e-12


@needs_ref
def test_oracle_vs_reference_edge_cases():
    ref = oracle.reference("c99")
    # k_max reached -> status 1 ; free initial state ; variable state size ; warm start is covered on the GPU side
    p = problems.mass_spring_ocp(8, 3, 10, bounds=True)
    r, o = ref.ip_ocp_hard_tv(p, k_max=3), oracle.ipm(p, k_max=3)
    assert r["status"] == o["status"] == 1 and r["kk"] == o["kk"] == 3
    p = problems.make("cfg4")
    r, o = ref.ip_ocp_hard_tv(p), oracle.ipm(p)
    assert (r["kk"], r["status"]) == (o["kk"], o["status"])
    assert rel_err(o["x"], r["x"]) < TOL and rel_err(o["lam"], r["lam"]) < TOL
    # mu0 <= 0 -> estimated from the cost
    p = problems.mass_spring_ocp(8, 3, 10, bounds=True)
    r, o = ref.ip_ocp_hard_tv(p, mu0=0.0), oracle.ipm(p, mu0=0.0)
    assert (r["kk"], r["status"]) == (o["kk"], o["status"])
    assert rel_err(o["u"], r["u"]) < TOL
