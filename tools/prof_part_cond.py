"""One condense -> IPM -> expand call (partial condensing, N2 given) for an ncu launch list.  usage: python tools/prof_part_cond.py [N2] [n_inst]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from hpmpc_b200 import capi, problems
N2 = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
k_max = 30
ps = [problems.mass_spring_ocp(12, 5, 30, bounds=True, xi=tuple(xi)) for xi in problems.instance_xi(256, first=0)]
L = capi.product()
pc = capi.PartCond(ps[0], N2); F = pc.full
blk = np.stack([F.pack(p) for p in ps])
d_in = torch.from_numpy(np.tile(blk, (n // 256, 1))).cuda()
lam_len = max(F.sz.lam_stride, 2)
z = lambda c: torch.zeros((n, c), dtype=torch.float64, device="cuda")
ux, pi, lam, t, info = z(F.sz.ux_stride), z(F.sz.pi_stride), z(lam_len), z(lam_len), z(6 + 5 * k_max)
rc = L.hpmpc_b200_d_ip2_res_mpc_hard_part_cond_batch(pc.h, n, d_in.data_ptr(), k_max, 2.0, 1e-8, 1e-8, ux.data_ptr(), pi.data_ptr(), lam.data_ptr(), t.data_ptr(), info.data_ptr(), None)
torch.cuda.synchronize()
print("rc", rc, "kk", float(info[:, 0].mean()))
