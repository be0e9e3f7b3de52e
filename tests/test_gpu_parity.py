"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
hpmpc_b200/lib/libhpmpc_b200.so; the oracle (oracle/) and the golden vectors are only the checker.

Bar (BASELINE.json): u, x, pi, lam within 1e-9 relative in FP64, IPM iteration counts identical."""
import numpy as np
import pytest

from conftest import rel_err
from golden_util import case_ids, cat, gold, problem
from hpmpc_b200 import capi, problems
from hpmpc_b200.batchgen import BatchSpec
from oracle import api as oracle

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def prod():
    import torch
    assert torch.cuda.is_available()
    return capi.HpmpcLib(capi.PRODUCT_LIB)


# ---------------------------------------------------------------- drop-in symbols, batch of one
@pytest.mark.parametrize("name,inst", case_ids())
def test_compat_symbols_match_golden(prod, name, inst):
    kind, p = problem(name, inst)
    if kind == "ric":
        for mode in ("sv", "trf_trs"):
            r = prod.ric(p, mode)
            for f in ("u", "x", "pi"):
                assert rel_err([cat(r[f])], [gold(name, inst, f)]) < TOL, (mode, f)
    else:
        for order in ("fortran", "c"):
            r = prod.ip_ocp_hard_tv(p, order=order, k_max=40, mu0=2.0, mu_tol=1e-8)
            assert r["kk"] == int(gold(name, inst, "kk")) and r["status"] == int(gold(name, inst, "status"))
            for f in ("u", "x", "pi", "lam"):
                assert rel_err([cat(r[f])], [gold(name, inst, f)]) < TOL, (order, f)
            assert rel_err([r["stat"]], [gold(name, inst, "stat")]) < 1e-6
            g = gold(name, inst, "inf_norm_res")
            assert abs(r["inf_norm_res"][3] - g[3]) <= 1e-9 * max(1.0, abs(g[3]))
            assert np.all(r["inf_norm_res"][:3] < 1e-6)


@pytest.mark.parametrize("shape", [(4, 2, 5), (8, 3, 10), (12, 5, 30), (2, 1, 1), (10, 1, 7), (30, 15, 4)])
def test_compat_riccati_vs_oracle(prod, shape):
    nx, nu, N = shape
    xi = tuple(problems.instance_xi(1, first=300)[0])
    p = problems.mass_spring_ocp(nx, nu, N, xi=xi)
    for mode in ("sv", "trf_trs"):
        r, o = prod.ric(p, mode), oracle.ric(p, mode)
        for f in ("u", "x", "pi"):
            assert rel_err(r[f], o[f]) < TOL, (shape, mode, f)


def test_compat_low_level_ipm_and_edge_cases(prod):
    p = problems.mass_spring_ocp(8, 3, 10, bounds=True)
    o = oracle.ipm(p)
    r = prod.ip2_res_mpc_hard_tv(p)              # d_ip2_res_mpc_hard_tv on panel-major data
    assert (r["kk"], r["status"]) == (o["kk"], o["status"])
    for f in ("u", "x", "pi", "lam"):
        assert rel_err(r[f], o[f]) < TOL
    # k_max reached -> 1
    r, o = prod.ip_ocp_hard_tv(p, k_max=3), oracle.ipm(p, k_max=3)
    assert r["status"] == o["status"] == 1 and r["kk"] == o["kk"] == 3
    assert rel_err(r["u"], o["u"]) < TOL
    # mu0 <= 0 -> estimated from the cost entries
    r, o = prod.ip_ocp_hard_tv(p, mu0=0.0), oracle.ipm(p, mu0=0.0)
    assert (r["kk"], r["status"]) == (o["kk"], o["status"]) and rel_err(r["u"], o["u"]) < TOL
    # no bounds at all -> kk = 0, plain Riccati solve
    p0 = problems.mass_spring_ocp(8, 3, 5, bounds=False)
    r, o = prod.ip_ocp_hard_tv(p0), oracle.ipm(p0)
    assert r["kk"] == o["kk"] == 0 and r["status"] == 0 and rel_err(r["x"], o["x"]) < TOL
    # warm start from the solution of a neighbouring problem: same answer as the oracle given the same start
    p1 = problems.mass_spring_ocp(8, 3, 10, bounds=True, xi=(0.5, -0.5, 0.2, 0.1))
    s = oracle.ipm(p)
    import ctypes  # noqa: F401
    r = prod.ip_ocp_hard_tv(p1, warm_start=1, x_init=s["x"], u_init=s["u"])
    ref = oracle.reference("c99").ip_ocp_hard_tv(p1, warm_start=1, x_init=s["x"], u_init=s["u"]) if oracle.have_reference() else None
    if ref is not None:
        assert (r["kk"], r["status"]) == (ref["kk"], ref["status"])
        for f in ("u", "x", "pi", "lam"):
            assert rel_err(r[f], ref[f]) < TOL


# ---------------------------------------------------------------- batched entry points
def _run_sv(spec, n_inst, first=0, launch=None):
    import torch
    L = capi.product()
    h = spec.h
    if launch:
        assert h.set_launch(*launch) == 0
    d_in = spec.torch_batch(n_inst, first)
    ux = torch.full((n_inst, h.sz.ux_stride), float("nan"), dtype=torch.float64, device="cuda")
    pi = torch.full((n_inst, h.sz.pi_stride), float("nan"), dtype=torch.float64, device="cuda")
    rc = L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n_inst, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), None,
                                              torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    return d_in, ux, pi


def _run_ipm(spec, n_inst, first=0, k_max=40, launch=None):
    import torch
    L = capi.product()
    h = spec.h
    if launch:
        assert h.set_launch(*launch) == 0
    d_in = spec.torch_batch(n_inst, first)
    z = lambda m: torch.zeros((n_inst, max(m, 2)), dtype=torch.float64, device="cuda")
    ux, pi, lam, t, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * k_max)
    rc = L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n_inst, d_in.data_ptr(), k_max, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(), pi.data_ptr(),
                                               lam.data_ptr(), t.data_ptr(), info.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    return d_in, ux, pi, lam, t, info


@pytest.mark.parametrize("cfg,n_inst", [("cfg1", 1), ("cfg2", 300), (dict(nx=4, nu=2, N=3, bounds=False), 1000)])
def test_batched_sv_vs_oracle(cfg, n_inst):
    spec = BatchSpec(cfg)
    _, ux, pi = _run_sv(spec, n_inst, first=11)
    uxh, pih = ux.cpu().numpy(), pi.cpu().numpy()
    assert np.all(np.isfinite(uxh[:, :sum(spec.base.nx) + sum(spec.base.nu)]))
    for i in list(range(min(n_inst, 24))) + [n_inst - 1]:
        o = oracle.ric(spec.problem(11 + i), "sv")
        u, x = spec.h.split_ux(uxh[i])
        assert rel_err(u, o["u"]) < TOL and rel_err(x, o["x"]) < TOL
        assert rel_err(spec.h.split_pi(pih[i]), o["pi"]) < TOL
    spec.h.close()


@pytest.mark.parametrize("cfg,n_inst", [("cfg2", 1), ("cfg2", 7), ("cfg2", 301), (dict(nx=4, nu=2, N=3, bounds=False), 13),
                                        (dict(nx=8, nu=3, N=10, bounds=False), 1027)])
def test_batched_sv_ragged_batches(cfg, n_inst):
    """Batch sizes that do not fill the last group of lanes / the last warp: every instance must still be solved, and
    nothing outside the batch may be written."""
    import torch
    spec = BatchSpec(cfg)
    h = spec.h
    d_in = spec.torch_batch(n_inst + 3, 17)
    ux = torch.full((n_inst + 3, h.sz.ux_stride), -7.0, dtype=torch.float64, device="cuda")
    pi = torch.full((n_inst + 3, h.sz.pi_stride), -7.0, dtype=torch.float64, device="cuda")
    rc = capi.product().hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n_inst, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), None, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert float((ux[n_inst:] + 7.0).abs().max()) == 0.0 and float((pi[n_inst:] + 7.0).abs().max()) == 0.0
    uxh, pih = ux.cpu().numpy(), pi.cpu().numpy()
    for i in sorted({0, n_inst // 2, n_inst - 1}):
        o = oracle.ric(spec.problem(17 + i), "sv")
        u, x = h.split_ux(uxh[i])
        assert rel_err(u, o["u"]) < TOL and rel_err(x, o["x"]) < TOL and rel_err(h.split_pi(pih[i]), o["pi"]) < TOL
    h.close()


def test_batched_ipm_waves_and_ragged_batch():
    """The IPM runs in waves of one instance per resident warp: a batch of one wave + 5 must give the same answers as the same
    instances solved as single-instance batches, and must not touch rows outside the batch."""
    import torch
    spec = BatchSpec(dict(nx=8, nu=3, N=10, bounds=True))
    h = spec.h
    wave = h.sz.ipm_grid * h.sz.ipm_warps_per_cta
    n = wave + 5
    _, ux, pi, lam, t, info = _run_ipm(spec, n, first=3)
    assert np.all(info.cpu().numpy()[:, 1] == 0)
    for i in (0, wave - 1, wave, n - 1):
        _, ux1, pi1, lam1, t1, info1 = _run_ipm(spec, 1, first=3 + i)
        assert torch.equal(ux1[0], ux[i]) and torch.equal(lam1[0], lam[i]) and torch.equal(info1[0], info[i])
    h.close()


def test_batched_trf_trs_equals_sv():
    import torch
    spec = BatchSpec("cfg2")
    n_inst = 200
    d_in, ux, pi = _run_sv(spec, n_inst)
    L, h = capi.product(), spec.h
    Lf = torch.zeros((n_inst, h.sz.L_stride), dtype=torch.float64, device="cuda")
    ux2, pi2 = torch.zeros_like(ux), torch.zeros_like(pi)
    st = torch.cuda.current_stream().cuda_stream
    assert L.hpmpc_b200_d_back_ric_rec_trf_batch(h.h, n_inst, d_in.data_ptr(), Lf.data_ptr(), st) == 0
    assert L.hpmpc_b200_d_back_ric_rec_trs_batch(h.h, n_inst, d_in.data_ptr(), Lf.data_ptr(), ux2.data_ptr(), pi2.data_ptr(), st) == 0
    torch.cuda.synchronize()
    n_ux = sum(spec.base.nx) + sum(spec.base.nu)
    assert float((ux2[:, :n_ux] - ux[:, :n_ux]).abs().max()) < 1e-10
    assert float((pi2 - pi).abs().max()) < 1e-9
    h.close()


@pytest.mark.parametrize("cfg,n_inst", [(dict(nx=8, nu=3, N=10, bounds=True), 600), ("cfg3", 40)])
def test_batched_ipm_vs_oracle(cfg, n_inst):
    spec = BatchSpec(cfg)
    _, ux, pi, lam, t, info = _run_ipm(spec, n_inst, first=5)
    uxh, pih, lamh, infoh = ux.cpu().numpy(), pi.cpu().numpy(), lam.cpu().numpy(), info.cpu().numpy()
    mism = 0
    for i in range(n_inst if n_inst <= 64 else 48):
        o = oracle.ipm(spec.problem(5 + i))
        kk, status = int(infoh[i, 0]), int(infoh[i, 1])
        if kk != o["kk"]:
            mism += 1
            continue
        assert status == o["status"]
        u, x = spec.h.split_ux(uxh[i])
        assert rel_err(u, o["u"]) < TOL and rel_err(x, o["x"]) < TOL
        assert rel_err(spec.h.split_pi(pih[i]), o["pi"]) < TOL
        assert rel_err(spec.h.split_lam(lamh[i]), o["lam"]) < TOL
        assert rel_err([infoh[i, 6:6 + 5 * kk].reshape(-1, 5)], [o["stat"]]) < 1e-6
        assert abs(infoh[i, 5] - o["inf_norm_res"][3]) < 1e-9 * max(1.0, o["inf_norm_res"][3])
    assert mism == 0, f"{mism} iteration-count mismatches"
    assert np.all(infoh[:, 1] == 0)
    spec.h.close()


def test_cfg4_variable_state_size_batch():
    """Config 4: nx shrinking 40 -> 4 over the horizon, free x0, input bounds only (reference test_d_ip_diag_box.c:87-95 shape)."""
    import torch
    p0 = problems.make("cfg4")
    h = capi.BatchOcp(p0, device=0)
    xis = problems.instance_xi(12, first=40)
    probs = [problems.make("cfg4", xi=tuple(x)) for x in xis]
    blk = torch.from_numpy(np.stack([h.pack(p) for p in probs])).cuda()
    n, k_max = len(probs), 40
    z = lambda m: torch.zeros((n, max(m, 2)), dtype=torch.float64, device="cuda")
    ux, pi, lam, t, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * k_max)
    rc = capi.product().hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, blk.data_ptr(), k_max, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(), pi.data_ptr(),
                                                            lam.data_ptr(), t.data_ptr(), info.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    for i, p in enumerate(probs):
        o = oracle.ipm(p)
        assert int(info[i, 0]) == o["kk"] and int(info[i, 1]) == o["status"]
        u, x = h.split_ux(ux[i].cpu().numpy())
        assert rel_err(u, o["u"]) < TOL and rel_err(x, o["x"]) < TOL
        assert rel_err(h.split_lam(lam[i].cpu().numpy()), o["lam"]) < TOL
    h.close()


def test_host_buffer_entry_points_match_device_entry_points():
    import torch
    L = capi.product()
    spec = BatchSpec("cfg2")
    h, n = spec.h, 700
    _, ux, pi = _run_sv(spec, n)
    h_in = spec.numpy_batch(n)
    hux, hpi = np.zeros((n, h.sz.ux_stride)), np.zeros((n, h.sz.pi_stride))
    assert L.hpmpc_b200_d_back_ric_rec_sv_batch_host(h.h, n, h_in.ctypes.data, hux.ctypes.data, hpi.ctypes.data) == 0
    n_ux = sum(spec.base.nx) + sum(spec.base.nu)
    np.testing.assert_array_equal(hux[:, :n_ux], ux.cpu().numpy()[:, :n_ux])
    np.testing.assert_array_equal(hpi, pi.cpu().numpy())
    h.close()
    spec = BatchSpec(dict(nx=8, nu=3, N=10, bounds=True))
    h, n, k_max = spec.h, 500, 30
    _, ux, pi, lam, t, info = _run_ipm(spec, n, k_max=k_max)
    h_in = spec.numpy_batch(n)
    z = lambda m: np.zeros((n, max(m, 2)))
    hux, hpi, hlam, ht, hinfo = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * k_max)
    assert L.hpmpc_b200_d_ip2_res_mpc_hard_batch_host(h.h, n, h_in.ctypes.data, k_max, 2.0, 1e-8, 1e-8, 0, hux.ctypes.data,
                                                      hpi.ctypes.data, hlam.ctypes.data, ht.ctypes.data, hinfo.ctypes.data) == 0
    np.testing.assert_array_equal(hinfo[:, :2], info.cpu().numpy()[:, :2])
    np.testing.assert_allclose(hlam, lam.cpu().numpy(), rtol=0, atol=0)
    h.close()


# ---------------------------------------------------------------- full size, size-independent properties
def _kkt_residual_sv(spec, n_inst, ux, pi, first=0):
    """Unconstrained KKT residuals of the whole batch, computed with torch from the generator's own A, B, Q_i, R_i
    (independent of the packed layout and of the solver): dynamics and stationarity, max-abs over the batch."""
    import torch
    p, h = spec.base, spec.h
    nx, nu, N = p.nx[1], p.nu[0], p.N
    A = torch.from_numpy(spec.A0).cuda(); B = torch.from_numpy(spec.B0).cuda()
    x01, x02, qs, rs = (torch.from_numpy(v).cuda() for v in spec.scalars(n_inst, first))
    U = torch.stack([ux[:, h.off[n]["ux"]:h.off[n]["ux"] + nu] for n in range(N)], 1)                    # n_inst x N x nu
    X = torch.stack([ux[:, h.off[n]["ux"] + p.nu[n]:h.off[n]["ux"] + p.nu[n] + nx] for n in range(1, N + 1)], 1)   # x_1..x_N
    PI = torch.stack([pi[:, h.off[n]["pi"]:h.off[n]["pi"] + nx] for n in range(N)], 1)                    # multiplier of x_{n+1}
    x0 = torch.zeros(n_inst, nx, dtype=torch.float64, device="cuda"); x0[:, 0] = x01; x0[:, 1] = x02
    Xprev = torch.cat([x0[:, None, :], X[:, :-1, :]], 1)
    r_dyn = Xprev @ A.T + U @ B.T + 0.1 - X
    r_u = rs[:, None, None] * U + 0.2 + PI @ B
    PInext = torch.cat([PI[:, 1:, :] @ A, torch.zeros(n_inst, 1, nx, dtype=torch.float64, device="cuda")], 1)
    r_x = qs[:, None, None] * X + 0.1 + PInext - PI
    return float(r_dyn.abs().max()), float(r_u.abs().max()), float(r_x.abs().max())


def test_full_size_cfg2_kkt_and_determinism():
    """BASELINE config 2 at full size (65 536 instances): KKT residuals of every instance, a checksum-of-results that must
    not depend on the launch shape, and spot parity against the oracle."""
    import torch
    spec = BatchSpec("cfg2")
    n_inst = 65536
    _, ux, pi = _run_sv(spec, n_inst)
    rd, ru, rx = _kkt_residual_sv(spec, n_inst, ux, pi)
    assert rd < 1e-11 and ru < 1e-10 and rx < 1e-10, (rd, ru, rx)
    ux_a = ux.clone()
    _, ux_b, pi_b = _run_sv(spec, n_inst, launch=(2, 3))
    n_ux = sum(spec.base.nx) + sum(spec.base.nu)
    assert torch.equal(ux_a[:, :n_ux], ux_b[:, :n_ux]) and torch.equal(pi, pi_b)
    uxh, pih = ux.cpu().numpy(), pi.cpu().numpy()
    for i in (0, 1, 4097, 65535):
        o = oracle.ric(spec.problem(i), "sv")
        u, x = spec.h.split_ux(uxh[i])
        assert rel_err(u, o["u"]) < TOL and rel_err(x, o["x"]) < TOL and rel_err(spec.h.split_pi(pih[i]), o["pi"]) < TOL
    spec.h.close()


def test_full_size_cfg3_ipm_properties():
    """BASELINE config 3 (16 384 box-constrained instances): all converge, bounds respected, complementarity at mu_tol,
    exit residuals small, iteration counts of a sample identical to the oracle."""
    import torch
    spec = BatchSpec("cfg3")
    n_inst = 16384
    _, ux, pi, lam, t, info = _run_ipm(spec, n_inst)
    assert int((info[:, 1] != 0).sum()) == 0
    assert float(info[:, 5].max()) <= 1e-8                       # mu
    assert float(info[:, 2:5].max()) < 1e-6                      # ||rq||, ||rb||, ||rd||
    assert float(lam.min()) > 0 and float(t.min()) > 0
    p, h = spec.base, spec.h
    for n in (0, 1, 25, 50):
        for j, id_ in enumerate(p.idxb[n]):
            v = ux[:, h.off[n]["ux"] + int(id_)]
            assert float(v.min()) >= p.lb[n][j] - 1e-9 and float(v.max()) <= p.ub[n][j] + 1e-9
    infoh = info.cpu().numpy()
    for i in (0, 1, 777, 16383):
        o = oracle.ipm(spec.problem(i))
        assert int(infoh[i, 0]) == o["kk"]
        u, x = h.split_ux(ux[i].cpu().numpy())
        assert rel_err(u, o["u"]) < TOL and rel_err(x, o["x"]) < TOL
    h.close()
