"""Four-warps-per-instance sweeps of the any-size kernels (hpmpc_b200/csrc/ric_team.cuh) against the one-warp-per-instance sweeps
they replace (HPMPC_B200_TEAM=0).  The team factorisation accumulates on FP64 tensor-core tiles (mma.sync.m8n8k4.f64) and the solve
sweeps split every dot product over the four warps, so sums are ordered differently: factors and solutions agree to rounding (1e-10 /
1e-11 here, 1e-9 is the bar) and the IPM must take the same number of iterations; and against the oracle (the reference's algorithm,
lqcp_solvers/d_back_ric_rec.c:236-397)."""
import os

import numpy as np
import pytest

from conftest import rel_err
from hpmpc_b200 import capi, problems
from oracle import api as oracle

pytestmark = pytest.mark.gpu


def _with_team(flag, fn):
    if flag:
        os.environ["HPMPC_B200_TEAM"] = "1"        # forced: by default stages with fewer than 20 rows stay on one warp
    else:
        os.environ["HPMPC_B200_TEAM"] = "0"
    try:
        return fn()
    finally:
        os.environ.pop("HPMPC_B200_TEAM", None)


def _close(x, y):
    xa, ya = x.cpu().numpy(), y.cpu().numpy()
    return float(np.max(np.abs(xa - ya) / np.maximum(1.0, np.abs(ya))))


def _batch(mk, n, first):
    import torch
    probs = [mk(tuple(x)) for x in problems.instance_xi(n, first=first)]
    h = capi.BatchOcp(probs[0], device=0)
    blk = torch.from_numpy(np.stack([h.pack(p) for p in probs])).cuda()
    return probs, h, blk


@pytest.mark.parametrize("name,mk,n", [
    ("cfg4", lambda xi: problems.make("cfg4", xi=xi), 300),                                        # rows 49 -> 13: all four warps, ragged panels
    ("small", lambda xi: problems.mass_spring_ocp(10, 3, 7, bounds=False, xi=xi), 64),             # nu+nx+1 = 14: one warp of the team
    ("wide", lambda xi: problems.mass_spring_ocp(30, 12, 6, bounds=False, xi=xi), 40),             # 43 rows
    ("tiny", lambda xi: problems.mass_spring_ocp(2, 1, 3, bounds=False, xi=xi), 17),
    ("limit", lambda xi: problems.mass_spring_ocp(50, 13, 4, bounds=False, xi=xi), 24),           # nu+nx+1 = 64: eight row tiles, eight column tiles
])
def test_team_sv_trf_trs_equal_one_warp(name, mk, n):
    import torch
    L = capi.product()
    probs, h, blk = _batch(mk, n, 900)
    z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")

    def sv():
        ux, pi = z(h.sz.ux_stride), z(h.sz.pi_stride)
        assert L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n, blk.data_ptr(), ux.data_ptr(), pi.data_ptr(), None, None) == 0
        torch.cuda.synchronize()
        return ux, pi

    def trf():
        Lf = z(h.sz.L_stride)
        assert L.hpmpc_b200_d_back_ric_rec_trf_batch(h.h, n, blk.data_ptr(), Lf.data_ptr(), None) == 0
        torch.cuda.synchronize()
        return Lf

    a, b = _with_team(True, sv), _with_team(False, sv)
    assert _close(a[0], b[0]) < 1e-11 and _close(a[1], b[1]) < 1e-11
    fa, fb = _with_team(True, trf), _with_team(False, trf)
    # a stage's block of the factor is [triangle | gradient row | inverse diagonal]; trf leaves the gradient row undefined
    p0 = probs[0]
    mask = np.zeros(fa.shape[1], dtype=bool)
    for s in range(p0.N + 1):
        nux = p0.nx[s] + (p0.nu[s] if s < p0.N else 0)
        o = h.off[s]["L"]
        mask[o:o + nux * (nux + 1) // 2] = True
        mask[o + nux * (nux + 1) // 2 + nux:o + nux * (nux + 1) // 2 + 2 * nux] = True
    mk_t = torch.from_numpy(mask).cuda()
    assert _close(fa[:, mk_t], fb[:, mk_t]) < 1e-10       # the team sweep sums on FP64 tensor-core tiles (DMMA), the one-warp sweep by FMA

    # solve with the stored factor (trs): b, q, r taken from the block
    def trs():
        ux, pi = z(h.sz.ux_stride), z(h.sz.pi_stride)
        assert L.hpmpc_b200_d_back_ric_rec_trs_batch(h.h, n, blk.data_ptr(), fa.data_ptr(), ux.data_ptr(), pi.data_ptr(), None) == 0
        torch.cuda.synchronize()
        return ux, pi
    ta, tb = _with_team(True, trs), _with_team(False, trs)
    assert _close(ta[0], tb[0]) < 1e-11 and _close(ta[1], tb[1]) < 1e-11
    assert _close(ta[0], a[0]) < 1e-9 and _close(ta[1], a[1]) < 1e-9
    for i in (0, n // 2, n - 1):
        o = oracle.ric(probs[i], mode="sv")
        u, x = h.split_ux(a[0][i].cpu().numpy())
        assert rel_err(u, o["u"]) < 1e-9 and rel_err(x, o["x"]) < 1e-9
        assert rel_err(h.split_pi(a[1][i].cpu().numpy()), o["pi"]) < 1e-9
    h.close()


@pytest.mark.parametrize("name,mk,n", [
    ("cfg4", lambda xi: problems.make("cfg4", xi=xi), 96),
    ("general", lambda xi: problems.general_test_problem(8, 3, 10, xi=xi), 80),
])
def test_team_ipm_equals_one_warp(name, mk, n):
    """The multi-kernel IPM driver with the team kernels (default for any-size patterns) against the fused one-warp kernel: same
    iteration counts and status, every output within 1e-9."""
    import torch
    L = capi.product()
    probs, h, blk = _batch(mk, n, 1300)
    k_max = 30
    z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")

    def ipm():
        out = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * k_max)
        assert L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, blk.data_ptr(), k_max, 2.0, 1e-8, 1e-8, 0, *[t.data_ptr() for t in out], None) == 0
        torch.cuda.synchronize()
        return out

    a, b = _with_team(True, ipm), _with_team(False, ipm)
    info, info_b = a[4].cpu().numpy(), b[4].cpu().numpy()
    assert np.array_equal(info[:, :2], info_b[:, :2])                   # iteration counts and status
    ok = torch.from_numpy(info[:, 1] == 0).cuda()                       # an instance that stops short of convergence is ill-conditioned by then
    for x, y, nm in zip(a[:4], b[:4], ("ux", "pi", "lam", "t")):
        assert _close(x[ok], y[ok]) < 1e-9, nm
    assert np.mean(info[:, 1] == 0) > 0.9
    o = oracle.ipm(probs[0], k_max=k_max)
    assert int(info[0, 0]) == o["kk"]
    u, x = h.split_ux(a[0][0].cpu().numpy())
    assert rel_err(u, o["u"]) < 1e-9 and rel_err(x, o["x"]) < 1e-9
    h.close()


def test_team_full_config4_batch_matches_one_warp():
    """BASELINE config 4 at its full batch size through both kernel sets (every CTA slot of the GPU busy for several rounds: a race
    between the warps of a team shows up here, not in a handful of instances): identical iteration counts, solutions within 1e-9."""
    import torch
    L = capi.product()
    n, k_max = 8192, 40
    p0 = problems.make("cfg4")
    h = capi.BatchOcp(p0, device=0)
    blk = torch.from_numpy(h.pack(p0)).cuda()[None, :].repeat(n, 1)
    xi = torch.from_numpy(problems.instance_xi(n)[:, 2].copy()).cuda()
    for s in range(p0.N + 1):                      # distinct instances: the gradient rows scaled per instance
        nux = p0.nx[s] + p0.nu[s]
        o = h.off[s]["RSQ"] + nux * (nux + 1) // 2
        blk[:, o:o + nux] *= (1.0 + 0.3 * xi[:, None])
    z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")

    def ipm():
        out = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * k_max)
        assert L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, blk.data_ptr(), k_max, 2.0, 1e-8, 1e-8, 0, *[t.data_ptr() for t in out], None) == 0
        torch.cuda.synchronize()
        return out

    a, b = _with_team(True, ipm), _with_team(False, ipm)
    ia, ib = a[4].cpu().numpy(), b[4].cpu().numpy()
    assert np.all(ia[:, 1] == 0) and np.array_equal(ia[:, :2], ib[:, :2])
    for x, y, nm in zip(a[:4], b[:4], ("ux", "pi", "lam", "t")):
        assert _close(x, y) < 1e-9, nm
    a2 = _with_team(True, ipm)                      # and the team kernels give the same bits twice
    for x, y in zip(a, a2):
        assert torch.equal(x, y)
    h.close()
