"""Phase clocks of the four-warp team backward sweep (hpmpc_b200/csrc/ric_team.cuh built with -DHBT_TIMING into
hpmpc_b200/lib/variants/libhpmpc_b200_tm.so): cycles of block 0, lane 0 of each warp, summed over the stages of the instances that
block solved.  usage: HPMPC_B200_LIB=hpmpc_b200/lib/variants/libhpmpc_b200_tm.so python tools/phase_timing_team.py [n_inst]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hpmpc_b200 import capi, problems
n = int(sys.argv[1]) if len(sys.argv) > 1 else 740
L = capi.product()
p0 = problems.make("cfg4"); h = capi.BatchOcp(p0, device=0)
d_in = torch.from_numpy(h.pack(p0)).cuda()[None, :].repeat(n, 1)
z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
Lf = z(h.sz.L_stride)
buf = (C.c_longlong * 48)()
names = ["load", "fixups", "trmm", "Pb+grad", "panel: acc", "panel: barrier 1", "panel: diag+scale", "panel: barrier 2", "panel: loop head", "syrk (W W^T pass)"]
for rep in range(2):
    L.hbt_timing_read(buf)
    L.hpmpc_b200_d_back_ric_rec_trf_batch(h.h, n, d_in.data_ptr(), Lf.data_ptr(), None)
    L.hbt_timing_read(buf)
tot = [sum(buf[12 * w + k] for k in range(12)) for w in range(4)]
print(f"trf (GRAD=false), {n} instances, block 0 solved {max(1, (n + 739) // 740)} instance(s); cycles per warp (lane 0):")
print(f"{'phase':22s}" + "".join(f"  warp{w}: cyc    %" for w in range(4)))
for k, nm in enumerate(names):
    print(f"{nm:22s}" + "".join(f"  {buf[12 * w + k]:10d} {100.0 * buf[12 * w + k] / max(tot[w], 1):5.1f}" for w in range(4)))
print(f"{'total':22s}" + "".join(f"  {tot[w]:10d}      " for w in range(4)))
