"""Scenario-tree box IPM (BASELINE config 5: d_tree_ip2_res_hard).

CPU: the oracle's tree IPM (oracle/ric_oracle.c: orc_tree_ip2_res_mpc_hard) against golden vectors produced by the REAL
reference solving the level-stacked chain problem (tests/golden/make_golden_tree_ipm.py), and against the oracle's chain
IPM on the same stacked problem.  GPU: hpmpc_b200_d_tree_ip2_res_mpc_hard_batch through the C ABI against the oracle
(tolerance 1e-9 relative on u, x, pi, lam; identical iteration counts)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_tree_ipm as G  # noqa: E402
from hpmpc_b200 import tree as T  # noqa: E402
from oracle import api  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "golden_tree_ipm_v1.npz"))
TOL = 1e-9


def cat(v):
    return np.concatenate([np.asarray(a, dtype=np.float64).ravel() for a in v])


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b))))) if a.size else 0.0


@pytest.mark.parametrize("case", list(G.CASES))
def test_tree_ipm_oracle_matches_reference_golden(case):
    t = G.build(case)
    r = api.tree_ipm(t, k_max=G.K_MAX, mu0=G.MU0, mu_tol=G.MU_TOL)
    assert [r["kk"], r["status"]] == list(GOLD[f"{case}/kk"])
    for f in ("u", "x", "pi", "lam"):
        assert rel(cat(r[f]), GOLD[f"{case}/{f}"]) < TOL, f


def test_tree_ipm_oracle_matches_stacked_chain_oracle():
    for shape in [(4, 2, 2, 2, 5), (8, 3, 3, 2, 6), (12, 5, 1, 0, 8), (12, 5, 2, 3, 8)]:
        t = T.mass_spring_tree(*shape, xi=(0.3, -0.2, 0.1, 0.4), bounds=True)
        r = api.tree_ipm(t)
        p, maps = T.stacked_chain(t)
        c = api.ipm(p, mu0=2.0)
        u, x, pi = T.unstack(t, maps, c)
        assert (r["kk"], r["status"]) == (c["kk"], c["status"]) and r["kk"] >= 5      # the input bounds are active
        assert max(rel(cat(r["u"]), cat(u)), rel(cat(r["x"]), cat(x)), rel(cat(r["pi"]), cat(pi))) < TOL
        assert rel(cat(r["lam"]), cat(G.node_lam(t, maps, p, c["lam"]))) < TOL


def test_tree_ipm_without_bounds_is_the_tree_riccati():
    t = T.mass_spring_tree(6, 2, 3, 2, 5, xi=(0.1, 0.2, -0.3, 0.0), bounds=False)
    r, s = api.tree_ipm(t), api.tree_ric(t)
    assert r["kk"] == 0 and r["status"] == 0
    assert max(rel(cat(r[f]), cat(s[f])) for f in ("u", "x", "pi")) < 1e-12


# ------------------------------------------------------------------------------------------------------------ GPU
def _gpu_ipm(tb, blocks, k_max=G.K_MAX, mu0=G.MU0, mu_tol=G.MU_TOL):
    import ctypes as C
    import torch
    n = len(blocks)
    sz = tb.sz
    d_in = torch.from_numpy(np.stack(blocks)).cuda()
    z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
    ux, pi, lam, tt, info = z(sz.ux_stride), z(sz.pi_stride), z(2 * tb.nbtot), z(2 * tb.nbtot), z(6 + 5 * k_max)
    rc = tb.L.hpmpc_b200_d_tree_ip2_res_mpc_hard_batch(tb.h, n, d_in.data_ptr(), k_max, C.c_double(mu0), C.c_double(mu_tol), C.c_double(1e-8), 0,
                                                      ux.data_ptr(), pi.data_ptr(), lam.data_ptr(), tt.data_ptr(), info.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    return ux.cpu().numpy(), pi.cpu().numpy(), lam.cpu().numpy(), info.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(G.CASES))
def test_tree_ipm_gpu_matches_reference_golden(case):
    t = G.build(case)
    tb = T.TreeBatch(t)
    try:
        ux, pi, lam, info = _gpu_ipm(tb, [tb.pack(t)])
        assert [int(info[0, 0]), int(info[0, 1])] == list(GOLD[f"{case}/kk"])
        u, x, p = tb.split(ux[0], pi[0])
        got = dict(u=u, x=x, pi=p, lam=tb.split_lam(lam[0]))
        for f in ("u", "x", "pi", "lam"):
            assert rel(cat(got[f]), GOLD[f"{case}/{f}"]) < TOL, f
    finally:
        tb.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 2, 2, 2, 5), (8, 3, 3, 2, 6), (12, 5, 1, 0, 8), (12, 5, 4, 3, 20)])
def test_tree_ipm_gpu_batch_vs_oracle(shape):
    """a ragged batch of different instances (more trees than one warp slot sees), every tree against the oracle"""
    from hpmpc_b200 import problems
    n = 5 if shape[-1] < 20 else 2
    xis = problems.instance_xi(n, first=11)
    ts = [T.mass_spring_tree(*shape, xi=tuple(xis[i]), bounds=True) for i in range(n)]
    tb = T.TreeBatch(ts[0])
    try:
        ux, pi, lam, info = _gpu_ipm(tb, [tb.pack(t) for t in ts])
        for i, t in enumerate(ts):
            r = api.tree_ipm(t, k_max=G.K_MAX, mu0=G.MU0, mu_tol=G.MU_TOL)
            assert (int(info[i, 0]), int(info[i, 1])) == (r["kk"], r["status"])
            u, x, p = tb.split(ux[i], pi[i])
            got = dict(u=u, x=x, pi=p, lam=tb.split_lam(lam[i]))
            for f in ("u", "x", "pi", "lam"):
                assert rel(cat(got[f]), cat(r[f])) < TOL, (i, f)
            assert abs(info[i, 5] - r["stat"][-1, 4]) <= 1e-9 * max(1.0, abs(r["stat"][-1, 4]))      # final mu
    finally:
        tb.close()


@pytest.mark.gpu
def test_tree_ipm_gpu_unconstrained_tree_is_one_riccati_solve():
    t = T.mass_spring_tree(6, 2, 3, 2, 5, xi=(0.1, 0.2, -0.3, 0.0), bounds=False)
    tb = T.TreeBatch(t)
    try:
        ux, pi, lam, info = _gpu_ipm(tb, [tb.pack(t)])
        s = api.tree_ric(t)
        u, x, p = tb.split(ux[0], pi[0])
        assert int(info[0, 0]) == 0 and int(info[0, 1]) == 0
        assert max(rel(cat(u), cat(s["u"])), rel(cat(x), cat(s["x"])), rel(cat(p), cat(s["pi"]))) < TOL
    finally:
        tb.close()


@pytest.mark.gpu
def test_tree_ipm_gpu_single_kernel_path_agrees(monkeypatch):
    """HPMPC_B200_TREE_IPM_FUSED=1 forces the one-kernel path (the one trees without size-specialised tails always take);
    both paths must give the oracle's iteration counts and agree with each other far below the parity tolerance."""
    t = G.build("ipm_cfg5_small")
    tb = T.TreeBatch(t)
    try:
        blk = [tb.pack(t)]
        ux0, pi0, lam0, info0 = _gpu_ipm(tb, blk)
        monkeypatch.setenv("HPMPC_B200_TREE_IPM_FUSED", "1")
        ux1, pi1, lam1, info1 = _gpu_ipm(tb, blk)
        assert list(info0[0, :2]) == list(info1[0, :2]) == [float(v) for v in GOLD["ipm_cfg5_small/kk"]]
        (u0, x0, p0), (u1, x1, p1) = tb.split(ux0[0], pi0[0]), tb.split(ux1[0], pi1[0])      # defined outputs, not the stride padding
        assert rel(cat(u0), cat(u1)) < 1e-11 and rel(cat(x0), cat(x1)) < 1e-11 and rel(cat(p0), cat(p1)) < 1e-11
        assert rel(cat(tb.split_lam(lam0[0])), cat(tb.split_lam(lam1[0]))) < 1e-10
    finally:
        tb.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_tree_ipm_gpu_iteration_limit(monkeypatch, fused):
    """k_max reached -> status 1 (mpc_solvers/d_ip2_res_hard.c:1331-1343), same iterate as the oracle after 3 iterations"""
    if fused:
        monkeypatch.setenv("HPMPC_B200_TREE_IPM_FUSED", "1")
    t = G.build("ipm_2x2")
    tb = T.TreeBatch(t)
    try:
        ux, pi, lam, info = _gpu_ipm(tb, [tb.pack(t)], k_max=3)
        r = api.tree_ipm(t, k_max=3, mu0=G.MU0, mu_tol=G.MU_TOL)
        assert (int(info[0, 0]), int(info[0, 1])) == (3, 1) == (r["kk"], r["status"])
        u, x, p = tb.split(ux[0], pi[0])
        assert max(rel(cat(u), cat(r["u"])), rel(cat(x), cat(r["x"])), rel(cat(p), cat(r["pi"])), rel(cat(tb.split_lam(lam[0])), cat(r["lam"]))) < TOL
        assert np.allclose(info[0, 6:6 + 15].reshape(3, 5), r["stat"], rtol=1e-9, atol=1e-12)      # sigma, alpha_aff, mu_aff, alpha, mu
    finally:
        tb.close()


@pytest.mark.gpu
def test_tree_ipm_gpu_chunked_batch_equals_one_pass(monkeypatch):
    """large batches go through the multi-kernel driver in chunks (bounded private buffers): same results bit for bit"""
    from hpmpc_b200 import problems
    xis = problems.instance_xi(7, first=40)
    ts = [T.mass_spring_tree(8, 3, 2, 2, 6, xi=tuple(x), bounds=True) for x in xis]
    tb = T.TreeBatch(ts[0])
    try:
        blocks = [tb.pack(t) for t in ts]
        a = _gpu_ipm(tb, blocks)
        monkeypatch.setenv("HPMPC_B200_TREE_IPM_CHUNK", "3")
        b = _gpu_ipm(tb, blocks)
        # strides are padded to an even count: compare what the layout defines (u, x, pi, lam per node) and the info rows
        for i in range(len(ts)):
            ua, xa, pa = tb.split(a[0][i], a[1][i]); ub, xb, pb = tb.split(b[0][i], b[1][i])
            assert np.array_equal(cat(ua), cat(ub)) and np.array_equal(cat(xa), cat(xb)) and np.array_equal(cat(pa), cat(pb))
            assert np.array_equal(cat(tb.split_lam(a[2][i])), cat(tb.split_lam(b[2][i])))
        assert np.array_equal(a[3], b[3])
    finally:
        tb.close()
