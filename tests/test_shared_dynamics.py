"""Shared-dynamics batch mode (include/hpmpc_b200.h: one set of matrices for the whole batch, per-instance b, q, r): factor once +
batched solve with the stored factor (reference d_back_ric_rec_trf_tv_res + d_back_ric_rec_trs_tv_res, lqcp_solvers/d_back_ric_rec.c:403,564),
against the oracle's factor+solve of every instance's full problem."""
import copy

import numpy as np
import pytest

from conftest import rel_err
from hpmpc_b200 import capi, problems
from oracle import api as oracle

pytestmark = pytest.mark.gpu
TOL = 1e-9


def shared_batch(base, n, seed=3):
    """Instances that share the matrices of `base` and differ in b, q, r (x0 through b_0, references through q, r)."""
    rng = np.random.default_rng(seed)
    probs = []
    for _ in range(n):
        p = copy.deepcopy(base)
        for s in range(p.N + 1):
            if s < p.N:
                p.b[s] = base.b[s] + 0.3 * rng.standard_normal(base.b[s].shape)
                p.r[s] = base.r[s] * (1.0 + 0.5 * rng.standard_normal())
            p.q[s] = base.q[s] + 0.2 * rng.standard_normal(base.q[s].shape)
        probs.append(p)
    return probs


def vec_of(h, p):
    """[r q] of every stage in the ux layout, then b of every stage in the pi layout."""
    v = np.zeros(h.sz.ux_stride + h.sz.pi_stride)
    for s in range(p.N + 1):
        o = h.off[s]["ux"]
        v[o:o + p.nu[s]] = p.r[s]; v[o + p.nu[s]:o + p.nu[s] + p.nx[s]] = p.q[s]
        if s < p.N:
            v[h.sz.ux_stride + h.off[s]["pi"]:h.sz.ux_stride + h.off[s]["pi"] + p.nx[s + 1]] = p.b[s]
    return v


def _api():
    L = capi.product()
    C = capi.C
    L.hpmpc_b200_shared_vec_stride.restype = C.c_longlong; L.hpmpc_b200_shared_vec_stride.argtypes = [C.c_void_p]
    L.hpmpc_b200_shared_factor_doubles.restype = C.c_longlong; L.hpmpc_b200_shared_factor_doubles.argtypes = [C.c_void_p]
    L.hpmpc_b200_d_back_ric_rec_trf_shared.argtypes = [C.c_void_p] * 4
    L.hpmpc_b200_d_back_ric_rec_trs_shared_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 6
    L.hpmpc_b200_d_back_ric_rec_sv_shared_batch_host.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 4
    L.hpmpc_b200_d_back_ric_rec_trs_shared_batch_stage_major.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 6
    return L


@pytest.mark.parametrize("cfg", ["cfg2", dict(nx=8, nu=3, N=10), dict(nx=40, nu=8, N=20, nx_profile=[40 - (9 * n) // 5 for n in range(21)]),
                                 dict(nx=30, nu=15, N=60)])
def test_shared_dynamics_device_and_host_paths(cfg):
    import torch
    L = _api()
    base = problems.make(cfg) if isinstance(cfg, str) else problems.mass_spring_ocp(cfg["nx"], cfg["nu"], cfg["N"], nx_profile=cfg.get("nx_profile"))
    h = capi.BatchOcp(base, device=0)
    n = 777
    probs = shared_batch(base, n)
    assert L.hpmpc_b200_shared_vec_stride(h.h) == h.sz.ux_stride + h.sz.pi_stride
    blk = h.pack(base)
    vec = np.stack([vec_of(h, p) for p in probs])
    d_blk, d_vec = torch.from_numpy(blk).cuda(), torch.from_numpy(vec).cuda()
    d_L = torch.zeros(L.hpmpc_b200_shared_factor_doubles(h.h) + 8, dtype=torch.float64, device="cuda")
    ux = torch.zeros((n, h.sz.ux_stride), dtype=torch.float64, device="cuda"); pi = torch.zeros((n, h.sz.pi_stride), dtype=torch.float64, device="cuda")
    assert L.hpmpc_b200_d_back_ric_rec_trf_shared(h.h, d_blk.data_ptr(), d_L.data_ptr(), None) == 0
    assert L.hpmpc_b200_d_back_ric_rec_trs_shared_batch(h.h, n, d_blk.data_ptr(), d_L.data_ptr(), d_vec.data_ptr(), ux.data_ptr(), pi.data_ptr(), None) == 0
    torch.cuda.synchronize()
    uxh, pih = ux.cpu().numpy(), pi.cpu().numpy()
    for i in (0, 1, 2, n // 2, n - 1):
        o = oracle.ric(probs[i], "sv")
        u, x = h.split_ux(uxh[i])
        assert rel_err(u, o["u"]) < TOL and rel_err(x, o["x"]) < TOL and rel_err(h.split_pi(pih[i]), o["pi"]) < TOL, i
    hux, hpi = np.zeros_like(uxh), np.zeros_like(pih)
    assert L.hpmpc_b200_d_back_ric_rec_sv_shared_batch_host(h.h, n, blk.ctypes.data, vec.ctypes.data, hux.ctypes.data, hpi.ctypes.data) == 0
    n_ux = sum(base.nx) + sum(base.nu)
    np.testing.assert_array_equal(hux[:, :n_ux], uxh[:, :n_ux])
    np.testing.assert_array_equal(hpi, pih)
    h.close()


@pytest.mark.parametrize("shape", [dict(nx=12, nu=5, N=30), dict(nx=12, nu=5, N=7, free_x0=True), dict(nx=8, nu=3, N=10),
                                   dict(nx=8, nu=3, N=4, free_x0=True)])
@pytest.mark.parametrize("n", [32 * 70 + 13, 64 * 9 + 45])
def test_thread_per_instance_kernel_equals_warp_per_instance_kernel(shape, n, monkeypatch):
    """The size-specialised shared-dynamics kernel (one thread per instance, csrc/ric_shared_tpi.cuh) against the any-size one
    (one warp per instance, HPMPC_B200_SHARED_GENERIC=1) and the oracle: nx_0 = 0 and nx_0 = nx, a ragged last warp."""
    import torch
    L = _api()
    base = problems.mass_spring_ocp(shape["nx"], shape["nu"], shape["N"], free_x0=shape.get("free_x0", False))
    h = capi.BatchOcp(base, device=0)
    probs = shared_batch(base, n, seed=11)
    blk = h.pack(base)
    vec = np.stack([vec_of(h, p) for p in probs])
    d_blk, d_vec = torch.from_numpy(blk).cuda(), torch.from_numpy(vec).cuda()
    d_L = torch.zeros(L.hpmpc_b200_shared_factor_doubles(h.h) + 8, dtype=torch.float64, device="cuda")
    assert L.hpmpc_b200_d_back_ric_rec_trf_shared(h.h, d_blk.data_ptr(), d_L.data_ptr(), None) == 0
    out = {}
    for mode in ("tpi", "generic"):
        if mode == "generic":
            monkeypatch.setenv("HPMPC_B200_SHARED_GENERIC", "1")
        else:
            monkeypatch.delenv("HPMPC_B200_SHARED_GENERIC", raising=False)
        ux = torch.full((n, h.sz.ux_stride), np.nan, dtype=torch.float64, device="cuda")
        pi = torch.full((n, h.sz.pi_stride), np.nan, dtype=torch.float64, device="cuda")
        assert L.hpmpc_b200_d_back_ric_rec_trs_shared_batch(h.h, n, d_blk.data_ptr(), d_L.data_ptr(), d_vec.data_ptr(), ux.data_ptr(), pi.data_ptr(), None) == 0
        torch.cuda.synchronize()
        out[mode] = (ux.cpu().numpy(), pi.cpu().numpy())
    n_ux, n_pi = sum(base.nx) + sum(base.nu), sum(base.nx[1:])
    a, b = out["tpi"], out["generic"]
    assert np.isfinite(a[0][:, :n_ux]).all() and np.isfinite(a[1][:, :n_pi]).all()
    assert rel_err(a[0][:, :n_ux], b[0][:, :n_ux]) < 1e-12 and rel_err(a[1][:, :n_pi], b[1][:, :n_pi]) < 1e-12
    for i in (0, 31, 32, n - 1):
        o = oracle.ric(probs[i], "sv")
        u, x = h.split_ux(a[0][i])
        assert rel_err(u, o["u"]) < TOL and rel_err(x, o["x"]) < TOL and rel_err(h.split_pi(a[1][i]), o["pi"]) < TOL, i
    h.close()


def _to_stage_major(a, parts, n):
    """a: (n, stride) instance-major; parts: [(offset, K)] -> flat array with part (o, K) as an [n][K] array at o * n."""
    out = np.zeros(a.shape[0] * a.shape[1])
    for o, K in parts:
        out[o * n:o * n + n * K] = a[:, o:o + K].reshape(-1)
    return out


@pytest.mark.parametrize("shape", [dict(nx=12, nu=5, N=30), dict(nx=8, nu=3, N=6, free_x0=True)])
def test_stage_major_vectors_give_the_same_bits(shape):
    """hpmpc_b200_d_back_ric_rec_trs_shared_batch_stage_major: the same kernel arithmetic on vectors stored stage by stage."""
    import torch
    L = _api()
    base = problems.mass_spring_ocp(shape["nx"], shape["nu"], shape["N"], free_x0=shape.get("free_x0", False))
    h = capi.BatchOcp(base, device=0)
    n = 32 * 21 + 7
    probs = shared_batch(base, n, seed=5)
    blk = h.pack(base)
    vec = np.stack([vec_of(h, p) for p in probs])
    us, ps = h.sz.ux_stride, h.sz.pi_stride
    ux_parts = [(h.off[s]["ux"], base.nu[s] + base.nx[s]) for s in range(base.N + 1)]
    pi_parts = [(h.off[s]["pi"], base.nx[s + 1]) for s in range(base.N)]
    vec_sm = np.concatenate([_to_stage_major(vec[:, :us], ux_parts, n), _to_stage_major(vec[:, us:], pi_parts, n)])
    d_blk = torch.from_numpy(blk).cuda()
    d_L = torch.zeros(L.hpmpc_b200_shared_factor_doubles(h.h) + 8, dtype=torch.float64, device="cuda")
    assert L.hpmpc_b200_d_back_ric_rec_trf_shared(h.h, d_blk.data_ptr(), d_L.data_ptr(), None) == 0
    d_vec, d_vec_sm = torch.from_numpy(vec).cuda(), torch.from_numpy(vec_sm).cuda()
    ux = torch.zeros((n, us), dtype=torch.float64, device="cuda"); pi = torch.zeros((n, ps), dtype=torch.float64, device="cuda")
    ux_sm = torch.zeros(n * us, dtype=torch.float64, device="cuda"); pi_sm = torch.zeros(n * ps, dtype=torch.float64, device="cuda")
    assert L.hpmpc_b200_d_back_ric_rec_trs_shared_batch(h.h, n, d_blk.data_ptr(), d_L.data_ptr(), d_vec.data_ptr(), ux.data_ptr(), pi.data_ptr(), None) == 0
    assert L.hpmpc_b200_d_back_ric_rec_trs_shared_batch_stage_major(h.h, n, d_blk.data_ptr(), d_L.data_ptr(), d_vec_sm.data_ptr(), ux_sm.data_ptr(),
                                                                   pi_sm.data_ptr(), None) == 0
    torch.cuda.synchronize()
    np.testing.assert_array_equal(ux_sm.cpu().numpy(), _to_stage_major(ux.cpu().numpy(), ux_parts, n))
    np.testing.assert_array_equal(pi_sm.cpu().numpy(), _to_stage_major(pi.cpu().numpy(), pi_parts, n))
    o = oracle.ric(probs[n - 1], "sv")
    u, x = h.split_ux(ux.cpu().numpy()[n - 1])
    assert rel_err(u, o["u"]) < TOL and rel_err(x, o["x"]) < TOL
    h.close()
    # a pattern without a thread-per-instance kernel is refused, not served by something else
    h2 = capi.BatchOcp(problems.mass_spring_ocp(10, 4, 8), device=0)
    assert L.hpmpc_b200_d_back_ric_rec_trs_shared_batch_stage_major(h2.h, 4, d_blk.data_ptr(), d_L.data_ptr(), d_vec_sm.data_ptr(), ux_sm.data_ptr(),
                                                                    pi_sm.data_ptr(), None) == -2
    h2.close()
