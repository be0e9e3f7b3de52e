import os, sys
sys.path.insert(0, "/root/repo")
import ctypes as C, torch
from hpmpc_b200 import capi
from hpmpc_b200.batchgen import BatchSpec
n = 1184
L = capi.product(); spec = BatchSpec("cfg3"); h = spec.h
d_in = spec.torch_batch(n)
z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
ux, pi, lam, t, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * 40)
for r in range(2):
    L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, d_in.data_ptr(), 40, C.c_double(2.0), C.c_double(1e-8), C.c_double(1e-8), 0, ux.data_ptr(), pi.data_ptr(), lam.data_ptr(), t.data_ptr(), info.data_ptr(), None)
torch.cuda.synchronize()
print("kk", float(info[:,0].mean()))
