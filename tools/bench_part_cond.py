"""Partial condensing on the device: time of the condense and expand kernels and of the whole condense -> IPM -> expand call against
the un-condensed IPM on the same batch (box-constrained mass-spring, nx=12 nu=5 N=30).  usage: python tools/bench_part_cond.py [n_inst]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from hpmpc_b200 import capi, problems

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nx, nu, N = 12, 5, 30
k_max = 30
ps = [problems.mass_spring_ocp(nx, nu, N, bounds=True, xi=tuple(xi)) for xi in problems.instance_xi(256, first=0)]
L = capi.product()

def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))[reps // 2]

out = []
for N2 in (30, 15, 10, 6, 5, 3):
    if N2 == N:
        h = capi.BatchOcp(ps[0]); F = h
    else:
        pc = capi.PartCond(ps[0], N2); F = pc.full
    blk = np.stack([F.pack(p) for p in ps])
    d_in = torch.from_numpy(np.tile(blk, (n // 256, 1))).cuda()
    lam_len = max(F.sz.lam_stride, 2)
    z = lambda c: torch.zeros((n, c), dtype=torch.float64, device="cuda")
    ux, pi, lam, t, info = z(F.sz.ux_stride), z(F.sz.pi_stride), z(lam_len), z(lam_len), z(6 + 5 * k_max)
    if N2 == N:
        run = lambda: L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, d_in.data_ptr(), k_max, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(), pi.data_ptr(),
                                                            lam.data_ptr(), t.data_ptr(), info.data_ptr(), None)
        assert run() == 0
        ms = timed(run)
        rec = dict(N2=N2, kind="un-condensed IPM (size-specialised sweeps)", ms=ms, solves_per_s=n / ms * 1e3)
    else:
        Cc = pc.cond
        run = lambda: L.hpmpc_b200_d_ip2_res_mpc_hard_part_cond_batch(pc.h, n, d_in.data_ptr(), k_max, 2.0, 1e-8, 1e-8, ux.data_ptr(), pi.data_ptr(),
                                                                      lam.data_ptr(), t.data_ptr(), info.data_ptr(), None)
        assert run() == 0
        ms = timed(run)
        in2 = z(Cc.sz.in_stride); lam2_len = max(Cc.sz.lam_stride, 2)
        ux2, pi2, lam2, t2 = z(Cc.sz.ux_stride), z(Cc.sz.pi_stride), z(lam2_len), z(lam2_len)
        ms_c = timed(lambda: L.hpmpc_b200_d_part_cond_batch(pc.h, n, d_in.data_ptr(), in2.data_ptr(), None))
        ms_e = timed(lambda: L.hpmpc_b200_d_part_expand_solution_batch(pc.h, n, d_in.data_ptr(), ux2.data_ptr(), pi2.data_ptr(), lam2.data_ptr(),
                                                                       t2.data_ptr(), ux.data_ptr(), pi.data_ptr(), lam.data_ptr(), t.data_ptr(), None))
        rec = dict(N2=N2, kind="condense -> IPM (any-size kernels, ng > 0) -> expand", ms=ms, solves_per_s=n / ms * 1e3, condense_ms=ms_c,
                   expand_ms=ms_e, cond_stage=dict(nu=Cc.p.nu[0], nx=Cc.p.nx[1], ng=Cc.p.ng[1], nb=Cc.p.nb[1]),
                   cond_in_bytes=int(Cc.sz.in_stride * 8), full_in_bytes=int(F.sz.in_stride * 8))
    ih = info.cpu().numpy()
    rec.update(mean_iterations=float(ih[:, 0].mean()), converged=int((ih[:, 1] == 0).sum()), n_inst=n)
    out.append(rec)
    print(json.dumps(rec), flush=True)
