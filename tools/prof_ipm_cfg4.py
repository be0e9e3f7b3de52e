"""One wave of the generic IPM kernel on BASELINE config 4 (variable nx 40 -> 4), for ncu.  usage: python tools/prof_ipm_cfg4.py [n_inst]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch
from hpmpc_b200 import capi, problems
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
L = capi.product()
p0 = problems.make("cfg4"); h = capi.BatchOcp(p0, device=0)
d_in = torch.from_numpy(h.pack(p0)).cuda()[None, :].repeat(n, 1)
z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
ux, pi, lam, t, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * 40)
for r in range(2):
    L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, d_in.data_ptr(), 40, C.c_double(2.0), C.c_double(1e-8), C.c_double(1e-8), 0, ux.data_ptr(), pi.data_ptr(),
                                          lam.data_ptr(), t.data_ptr(), info.data_ptr(), None)
torch.cuda.synchronize()
print("kk", float(info[:, 0].mean()), "grid", h.sz.ipm_grid, "warps", h.sz.ipm_warps_per_cta)
