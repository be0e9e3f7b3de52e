import os, sys
sys.path.insert(0, "/root/repo")
import torch
from hpmpc_b200 import capi
from hpmpc_b200.batchgen import BatchSpec
from oracle import api as oracle
import numpy as np
n = 16384
L = capi.product(); spec = BatchSpec("cfg3"); h = spec.h
if len(sys.argv) > 2: h.set_launch(int(sys.argv[1]), int(sys.argv[2]))
d_in = spec.torch_batch(n)
z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
ux, pi = z(h.sz.ux_stride), z(h.sz.pi_stride)
st = torch.cuda.current_stream().cuda_stream
f = lambda: L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), None, st)
assert f() == 0; torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record()
for i in range(3): f(); ev[i+1].record()
torch.cuda.synchronize()
ms = min(ev[i].elapsed_time(ev[i+1]) for i in range(3))
# parity of instance 0 and last (unconstrained solve of the cfg3 matrices)
import copy
def rel(a, b): return max(float(np.max(np.abs(x-y)/np.maximum(1, np.abs(y)))) if len(y) else 0.0 for x, y in zip(a, b))
errs = []
for i in (0, n-1):
    p = spec.problem(i); p.nb = [0]*(p.N+1)
    o = oracle.ric(p, "sv"); u, x = h.split_ux(ux[i].cpu().numpy())
    errs.append(max(rel(u, o["u"]), rel(x, o["x"]), rel(h.split_pi(pi[i].cpu().numpy()), o["pi"])))
print(f"cfg3 sv n={n} grid={h.sz.grid} warps={h.sz.warps_per_cta} smem={h.sz.smem_per_cta} ms={ms:.2f} solves/s={n/ms*1e3:.3e} maxerr={max(errs):.1e}")
