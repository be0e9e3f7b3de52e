"""Partial condensing -- SURVEY.md section 8f row f3.

Reference: lqcp_solvers/d_part_cond.c (d_part_cond_compute_problem_size :694, d_cond_BAbt :214, d_cond_RSQrq :312, d_cond_DCtd :579,
d_part_cond :926, d_part_expand_solution :1103) and the N2 < N branch of interfaces/c/fortran_order_interface.c:389-528.

CPU part: the oracle's restatement (orc_part_cond / orc_part_expand behind orc_fortran_order_d_ip_ocp_hard_tv with N2 < N) against golden
vectors produced by the compiled reference (tests/golden/make_golden_part_cond.py), the size bookkeeping against the reference's own
routine, and the reference's defect for nu > 4 characterised (it returns a non-stationary point there; the oracle does not).
GPU part: the product through the C ABI -- the drop-in high-level symbols with N2 < N and the batched condense / solve / expand entry
points -- against the same golden vectors and the oracle."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from conftest import rel_err
from hpmpc_b200 import capi, problems
from hpmpc_b200.capi import int_array, ptr_array
from oracle import api as oracle

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_part_cond", os.path.join(HERE, "golden", "make_golden_part_cond.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)
GOLD = np.load(os.path.join(HERE, "golden", "golden_part_cond_v1.npz"))
IDS = [(name, N2, inst) for name, c in mg.CASES.items() for (N2, inst) in c["runs"]]
TOL = 1e-9


def _case(name, N2, inst):
    p = mg.build_problem(name, inst)
    key = f"{name}/{N2}/{inst}"
    assert abs(mg.checksum(p) - float(GOLD[key + "/checksum"])) < 1e-9, "input generator drifted"
    return p, (lambda f: GOLD[key + "/" + f])


def _check(r, g, tol=TOL, stat_tol=1e-6):
    assert r["kk"] == int(g("kk")) and r["status"] == int(g("status")), (r["kk"], int(g("kk")), r["status"])
    for f in ("u", "x", "pi", "lam"):
        e = rel_err([mg.cat(r[f])], [g(f)])
        assert e < tol, (f, e)
    assert rel_err([r["stat"]], [g("stat")]) < max(stat_tol, tol)
    assert np.all(np.abs(r["inf_norm_res"][:3] - g("inf_norm_res")[:3]) < 1e-9)


# ------------------------------------------------------------------------------------------- CPU: the oracle is pinned
@pytest.mark.parametrize("name,N2,inst", IDS)
def test_oracle_matches_reference_golden(name, N2, inst):
    p, g = _case(name, N2, inst)
    _check(oracle.ipm(p, k_max=mg.K_MAX, N2=N2), g)


def test_condensing_changes_the_iterates():
    """The golden vectors pin the CONDENSED solve: the same problem solved without condensing differs at the 1e-8 level (the IPM
    stops at mu_tol, on different iterates), so a wrapper that ignored N2 would fail the 1e-9 bar."""
    p, g = _case("pc_8_3_10", 3, 1)
    r = oracle.ipm(p, k_max=mg.K_MAX)
    assert rel_err([mg.cat(r["lam"])], [g("lam")]) > 1e-9


@pytest.mark.parametrize("shape,N2", [((8, 3, 10), 3), ((8, 3, 10), 4), ((12, 5, 30), 7), ((4, 2, 9), 2), ((24, 11, 50), 13)])
def test_condensed_sizes_match_reference_routine(shape, N2):
    """hpmpc_b200_part_cond_compute_problem_size (host only) against the reference's d_part_cond_compute_problem_size when the
    reference build is present, and against the block arithmetic of d_part_cond.c:694-735 always."""
    nx, nu, N = shape
    p = problems.mass_spring_ocp(nx, nu, N, bounds=True)
    nx2, nu2, nb2, ng2, idxb2 = capi.part_cond_sizes(p, N2)
    N1, R1 = N // N2, N - N2 * (N // N2)
    lens = [N1 + 1 if k < R1 else N1 for k in range(N2)]
    assert sum(lens) == N
    n0 = 0
    for k, T in enumerate(lens):
        assert nx2[k] == p.nx[n0] and nu2[k] == sum(p.nu[n0:n0 + T])
        states = sum(int(np.sum(np.asarray(p.idxb[n]) >= p.nu[n])) for n in range(n0 + 1, n0 + T))
        assert ng2[k] == states and nb2[k] == sum(p.nb[n0:n0 + T]) - states
        assert len(set(idxb2[k].tolist())) == nb2[k] and (nb2[k] == 0 or idxb2[k].max() < nu2[k] + nx2[k])
        n0 += T
    assert (nx2[N2], nu2[N2], nb2[N2], ng2[N2]) == (p.nx[N], 0, p.nb[N], 0)
    if oracle.have_reference():
        R = oracle.reference("c99").lib
        idxb = [np.ascontiguousarray(v, dtype=np.int32) if len(v) else np.zeros(1, dtype=np.int32) for v in p.idxb]
        out = [(C.c_int * (N2 + 1))() for _ in range(4)]
        R.d_part_cond_compute_problem_size.restype = None
        R.d_part_cond_compute_problem_size(N, int_array(p.nx), int_array(p.nu), int_array(p.nb), ptr_array(idxb), int_array([0] * (N + 1)),
                                           N2, *out)
        assert [list(v) for v in out] == [nx2, nu2, nb2, ng2]


@pytest.mark.skipif(not oracle.have_reference(), reason="reference build not present")
def test_reference_defect_for_more_than_four_inputs_is_not_reproduced():
    """For nu > 4 and blocks of three or more stages the reference's lib4 routine returns a point whose stationarity residual (its
    own inf_norm_res[0]) is O(1); the oracle's condensed solve is a KKT point and agrees with the un-condensed solve to the IPM's
    tolerance.  Parity for such shapes is therefore judged on the KKT residuals, not against the reference."""
    p = problems.mass_spring_ocp(12, 5, 30, bounds=True, xi=(0.1, 0.2, -0.5, 0.7))
    ref = oracle.reference("c99")
    full = oracle.ipm(p)
    for N2 in (10, 5):
        r, o = ref.ip_ocp_hard_tv(p, N2=N2), oracle.ipm(p, N2=N2)
        assert r["inf_norm_res"][0] > 1e-2, "the reference's defect is gone: move cfg-2 shapes into the golden set"
        assert o["status"] == 0 and np.all(o["inf_norm_res"][:3] < 1e-8)
        assert rel_err([mg.cat(o["u"])], [mg.cat(full["u"])]) < 1e-4
    # with blocks of two stages the reference is right for nu = 5 too, and the oracle agrees with it
    r, o = ref.ip_ocp_hard_tv(p, N2=15), oracle.ipm(p, N2=15)
    assert r["kk"] == o["kk"] and rel_err([mg.cat(o[f]) for f in ("u", "x", "pi", "lam")], [mg.cat(r[f]) for f in ("u", "x", "pi", "lam")]) < TOL


def test_abi_exports_partial_condensing_symbols():
    L = capi.product()
    for s in ("hpmpc_b200_part_cond_compute_problem_size", "hpmpc_b200_pcond_create", "hpmpc_b200_pcond_destroy", "hpmpc_b200_pcond_full",
              "hpmpc_b200_pcond_cond", "hpmpc_b200_d_part_cond_batch", "hpmpc_b200_d_part_expand_solution_batch",
              "hpmpc_b200_d_ip2_res_mpc_hard_part_cond_batch"):
        assert hasattr(L, s), s


def test_host_only_handle_refuses_to_solve():
    """device < 0: sizes and packing only; a solve is refused (no CPU fallback)."""
    p = problems.mass_spring_ocp(8, 3, 10, bounds=True)
    h = capi.PartCond(p, 3, device=-1)
    assert h.cond.sz.N == 3 and h.full.sz.N == 10
    rc = h.L.hpmpc_b200_d_ip2_res_mpc_hard_part_cond_batch(h.h, 1, None, 10, 2.0, 1e-8, 1e-8, None, None, None, None, None, None)
    assert rc == -4
    h.close()


# ------------------------------------------------------------------------------------------- GPU: the product
@pytest.mark.gpu
@pytest.mark.parametrize("order", ["fortran", "c"])
@pytest.mark.parametrize("name,N2,inst", IDS)
def test_gpu_high_level_symbols_match_reference_golden(name, N2, inst, order):
    """{c,fortran}_order_d_ip_ocp_hard_tv with N2 < N through the drop-in library."""
    p, g = _case(name, N2, inst)
    lib = capi.HpmpcLib(capi.PRODUCT_LIB)
    _check(lib.ip_ocp_hard_tv(p, order=order, k_max=mg.K_MAX, N2=N2), g)


def _solve_batch(p_list, N2, k_max=30):
    import torch
    p0 = p_list[0]
    h = capi.PartCond(p0, N2)
    n = len(p_list)
    F = h.full
    blk = np.stack([F.pack(p) for p in p_list])
    d_in = torch.from_numpy(blk).cuda()
    lam_len = max(F.sz.lam_stride, 2)
    ux = torch.zeros((n, F.sz.ux_stride), dtype=torch.float64, device="cuda"); pi = torch.zeros((n, F.sz.pi_stride), dtype=torch.float64, device="cuda")
    lam = torch.zeros((n, lam_len), dtype=torch.float64, device="cuda"); t = torch.zeros_like(lam)
    info = torch.zeros((n, 6 + 5 * k_max), dtype=torch.float64, device="cuda")
    rc = h.L.hpmpc_b200_d_ip2_res_mpc_hard_part_cond_batch(h.h, n, d_in.data_ptr(), k_max, 2.0, 1e-8, 1e-8, ux.data_ptr(), pi.data_ptr(),
                                                           lam.data_ptr(), t.data_ptr(), info.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    out = []
    uxh, pih, lamh, infoh = ux.cpu().numpy(), pi.cpu().numpy(), lam.cpu().numpy(), info.cpu().numpy()
    for i in range(n):
        u, x = F.split_ux(uxh[i])
        out.append(dict(u=u, x=x, pi=F.split_pi(pih[i]), lam=F.split_lam(lamh[i]), kk=int(infoh[i, 0]), status=int(infoh[i, 1]),
                        inf_norm_res=infoh[i, 2:6].copy()))
    h.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("shape,N2,n_inst", [((8, 3, 10), 3, 40), ((12, 5, 30), 6, 24), ((12, 5, 30), 10, 16), ((24, 11, 12), 4, 8),
                                             ((4, 2, 9), 1, 33)])
def test_gpu_batched_condense_solve_expand_vs_oracle(shape, N2, n_inst):
    """hpmpc_b200_d_ip2_res_mpc_hard_part_cond_batch on a batch of distinct instances against the oracle (which is pinned on the
    reference for nu <= 4 and is a verified KKT point beyond): iteration counts identical, u / x / pi / lam within 1e-9, and the exit
    norms are those of the FULL problem."""
    nx, nu, N = shape
    ps = [problems.mass_spring_ocp(nx, nu, N, bounds=True, xi=tuple(xi)) for xi in problems.instance_xi(n_inst, first=100)]
    got = _solve_batch(ps, N2)
    for p, r in zip(ps, got):
        o = oracle.ipm(p, k_max=30, N2=N2)
        assert (r["kk"], r["status"]) == (o["kk"], o["status"])
        for f in ("u", "x", "pi", "lam"):
            assert rel_err([mg.cat(r[f])], [mg.cat(o[f])]) < TOL, f
        assert np.all(np.abs(r["inf_norm_res"][:3] - o["inf_norm_res"][:3]) < 1e-9) and np.all(r["inf_norm_res"][:2] < 1e-7)


@pytest.mark.gpu
def test_gpu_condense_then_expand_of_the_uncondensed_solution_is_consistent():
    """Size-independent property at a large batch: the condensed problem built on the device, solved, and expanded satisfies the
    FULL problem's KKT conditions (exit norms from the full-problem residual kernel) for every instance of a 4096 batch."""
    ps = [problems.mass_spring_ocp(12, 5, 30, bounds=True, xi=tuple(xi)) for xi in problems.instance_xi(64, first=7)]
    got = _solve_batch(ps * 64, 5)
    res = np.array([r["inf_norm_res"] for r in got])
    assert all(r["status"] == 0 for r in got)
    assert res[:, 0].max() < 1e-8 and res[:, 1].max() < 1e-8 and res[:, 2].max() < 1e-10
    # the 64 copies of every instance give bit-identical results (no cross-instance interference in the scratch slots)
    for i in range(64):
        for c in range(1, 64):
            assert np.array_equal(mg.cat(got[i]["u"]), mg.cat(got[i + 64 * c]["u"]))


# ------------------------------------------------------------------------------------------- the lib4 routines themselves
def _cmp_cond(a, b, tol=1e-11):
    for k in ("nx2", "nu2", "nb2", "ng2"):
        assert a[k] == b[k], k
    for k in ("BAbt", "RSQrq", "DCt", "lb", "ub", "lg", "ug"):
        for x, y in zip(a[k], b[k]):
            assert x.shape == y.shape and (x.size == 0 or np.max(np.abs(x - y) / np.maximum(1.0, np.abs(y))) < tol), k
    for x, y in zip(a["idxb"], b["idxb"]):
        assert np.array_equal(x, y)


@pytest.mark.gpu
@pytest.mark.skipif(not oracle.have_reference(), reason="reference build not present")
@pytest.mark.parametrize("shape,N2", [((8, 3, 10), 3), ((8, 3, 10), 1), ((12, 4, 10), 4), ((4, 2, 9), 2), ((6, 1, 8), 3), ((12, 5, 10), 5)])
def test_gpu_lib4_d_part_cond_matches_reference_routine(shape, N2):
    """d_part_cond through the drop-in library (device kernel behind panel-major arguments) against the reference's own routine:
    every condensed matrix and vector, entry by entry.  (12,5) with two-stage blocks is the largest nu the reference gets right.)"""
    nx, nu, N = shape
    p = problems.mass_spring_ocp(nx, nu, N, bounds=True, xi=(0.3, -0.2, 0.5, 0.1))
    ref = oracle.reference("c99").part_cond(p, N2)
    got = capi.HpmpcLib(capi.PRODUCT_LIB).part_cond(p, N2)
    _cmp_cond(got, ref)


@pytest.mark.gpu
@pytest.mark.skipif(not oracle.have_reference(), reason="reference build not present")
def test_gpu_lib4_d_part_cond_where_the_reference_is_wrong():
    """nu = 5, blocks of four stages: the reference's condensed Hessian differs from the product's -- and it is the reference's that
    does not reproduce the un-condensed optimum (tools/repro_part_cond_nu5.py); the dynamics and the constraints agree."""
    p = problems.mass_spring_ocp(12, 5, 10, bounds=True, xi=(0.3, -0.2, 0.5, 0.1))
    ref = oracle.reference("c99").part_cond(p, 3)
    got = capi.HpmpcLib(capi.PRODUCT_LIB).part_cond(p, 3)
    for k in ("BAbt", "DCt", "lb", "ub", "lg", "ug"):
        for x, y in zip(got[k], ref[k]):
            assert np.max(np.abs(x - y)) < 1e-11, k
    assert max(np.max(np.abs(x - y)) for x, y in zip(got["RSQrq"], ref["RSQrq"])) > 1e-3


@pytest.mark.gpu
@pytest.mark.skipif(not oracle.have_reference(), reason="reference build not present")
@pytest.mark.parametrize("shape,N2", [((8, 3, 10), 3), ((12, 4, 10), 4), ((4, 2, 9), 2)])
def test_gpu_lib4_d_part_expand_solution_matches_reference_routine(shape, N2):
    """d_part_expand_solution on an arbitrary condensed point (not a solution: the map itself is compared), both libraries."""
    nx, nu, N = shape
    p = problems.mass_spring_ocp(nx, nu, N, bounds=True, xi=(0.1, 0.4, -0.5, 0.2))
    R, G = oracle.reference("c99"), capi.HpmpcLib(capi.PRODUCT_LIB)
    nx2, nu2, nb2, ng2, _ = capi.part_cond_sizes(p, N2)
    rng = np.random.default_rng(5)
    ux2 = [rng.standard_normal(nu2[k] + nx2[k]) for k in range(N2 + 1)]
    pi2 = [rng.standard_normal(nx2[k + 1]) for k in range(N2)]
    lam2 = [rng.random(2 * nb2[k] + 2 * ng2[k]) for k in range(N2 + 1)]
    t2 = [rng.random(2 * nb2[k] + 2 * ng2[k]) + 0.5 for k in range(N2 + 1)]
    a = G.part_cond(p, N2, expand_from=(ux2, pi2, lam2, t2))["expanded"]
    b = R.part_cond(p, N2, expand_from=(ux2, pi2, lam2, t2))["expanded"]
    for f in ("u", "x", "pi", "lam", "t"):
        assert rel_err([mg.cat(a[f])], [mg.cat(b[f])]) < 1e-11, f
