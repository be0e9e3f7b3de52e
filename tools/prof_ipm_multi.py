"""One cfg-3 IPM solve of n instances through the default path (multi-kernel driver for large batches); for ncu launch lists.
usage: python tools/prof_ipm_multi.py [n_inst] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hpmpc_b200 import capi
from hpmpc_b200.batchgen import BatchSpec
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = sys.argv[3] if len(sys.argv) > 3 else "cfg3"
L = capi.product(); spec = BatchSpec(cfg); h = spec.h
d_in = spec.torch_batch(n)
z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
ux, pi, lam, t, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * 40)
st = torch.cuda.current_stream().cuda_stream
run = lambda: L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, d_in.data_ptr(), 40, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(), pi.data_ptr(), lam.data_ptr(), t.data_ptr(), info.data_ptr(), st)
assert run() == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    assert run() == 0
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{cfg} n={n} {ms:.2f} ms  {n / ms * 1e3:.0f} solves/s  mean kk {float(info[:, 0].mean()):.2f} converged {int((info[:, 1] == 0).sum())}")
