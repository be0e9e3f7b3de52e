"""Small run of the any-size team kernels (config 4 and a general-constraint pattern: sv, trf, trs, multi-kernel IPM), written for
compute-sanitizer (racecheck / memcheck); the tool is closed on the GPU pool this was developed on, so the race guard that actually
runs is tests/test_team.py::test_team_full_config4_batch_matches_one_warp (full batch, both kernel sets, twice).
usage: [compute-sanitizer --tool racecheck] python tools/sanitize_team.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["HPMPC_B200_TEAM"] = "1"
import numpy as np, torch
from hpmpc_b200 import capi, problems
L = capi.product()
for mk, n in ((lambda xi: problems.make("cfg4", xi=xi), 80), (lambda xi: problems.general_test_problem(8, 3, 10, xi=xi), 72)):
    probs = [mk(tuple(x)) for x in problems.instance_xi(n, first=5)]
    h = capi.BatchOcp(probs[0], device=0)
    blk = torch.from_numpy(np.stack([h.pack(p) for p in probs])).cuda()
    z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
    ux, pi, lam, t, info, Lf = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * 20), z(h.sz.L_stride)
    if not probs[0].ng:
        assert L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n, blk.data_ptr(), ux.data_ptr(), pi.data_ptr(), None, None) == 0
        assert L.hpmpc_b200_d_back_ric_rec_trf_batch(h.h, n, blk.data_ptr(), Lf.data_ptr(), None) == 0
        assert L.hpmpc_b200_d_back_ric_rec_trs_batch(h.h, n, blk.data_ptr(), Lf.data_ptr(), ux.data_ptr(), pi.data_ptr(), None) == 0
    assert L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, blk.data_ptr(), 20, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(), pi.data_ptr(), lam.data_ptr(), t.data_ptr(), info.data_ptr(), None) == 0
    torch.cuda.synchronize()
    print("kk", float(info[:, 0].mean()), "converged", int((info[:, 1] == 0).sum()), "of", n)
    h.close()
print("done")
