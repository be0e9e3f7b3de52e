"""One launch of the any-size Riccati sv kernel on BASELINE config 4, for ncu.  usage: python tools/prof_sv_cfg4.py [n_inst]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hpmpc_b200 import capi, problems
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1480
L = capi.product()
p0 = problems.make("cfg4"); h = capi.BatchOcp(p0, device=0)
d_in = torch.from_numpy(h.pack(p0)).cuda()[None, :].repeat(n, 1)
z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
ux, pi = z(h.sz.ux_stride), z(h.sz.pi_stride)
for r in range(2):
    L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), None, None)
torch.cuda.synchronize()
print("ok", float(ux.abs().sum()))
