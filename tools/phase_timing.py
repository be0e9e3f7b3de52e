"""In-kernel phase timing of the fast sv kernel (debug library built by `make -C hpmpc_b200/csrc dbg`)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from hpmpc_b200 import capi
capi.PRODUCT_LIB = os.path.join(os.path.dirname(capi.PRODUCT_LIB), "libhpmpc_b200_dbg.so")
from hpmpc_b200.batchgen import BatchSpec
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
cps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
warps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
L = capi.product()
spec = BatchSpec("cfg2"); h = spec.h
h.set_launch(cps, warps)
d_in = spec.torch_batch(n)
ux = torch.zeros((n, h.sz.ux_stride), dtype=torch.float64, device="cuda"); pi = torch.zeros((n, h.sz.pi_stride), dtype=torch.float64, device="cuda")
dbg = torch.zeros(8000, dtype=torch.int64, device="cuda")
for rep in range(2):
    dbg.zero_()
    L.hb_debug_timing(C.c_void_p(dbg.data_ptr()))
    L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), None, None)
    torch.cuda.synchronize()
v = dbg.cpu().numpy().reshape(-1, 2)
v = v[v[:, 0] > 0]
agg = collections.OrderedDict()
for i in range(len(v) - 1):
    key = (int(v[i, 0]), int(v[i + 1, 0]))
    agg.setdefault(key, []).append(int(v[i + 1, 1] - v[i, 1]))
names = {110: "bwd:W computed", 111: "bwd:W stored", 120: "bwd:chol done", 100: "bwd:top", 101: "bwd:inputs ready", 102: "bwd:assembled", 103: "bwd:pre-factor", 104: "bwd:factored", 200: "fwd:top", 201: "fwd:ready", 202: "fwd:done"}
tot = 0
for (a, b), ds in agg.items():
    print(f"{names.get(a,a):>18s} -> {names.get(b,b):<18s} n={len(ds):4d} mean={sum(ds)/len(ds):9.0f} min={min(ds):7d} max={max(ds):7d}")
    tot += sum(ds)
print("total cycles", tot, "stamps", len(v))
