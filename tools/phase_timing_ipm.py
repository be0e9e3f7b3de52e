"""In-kernel sweep timing of the fused IPM kernel (debug library, make -C hpmpc_b200/csrc dbg).  usage: [n_inst]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from hpmpc_b200 import capi
capi.PRODUCT_LIB = os.path.join(os.path.dirname(capi.PRODUCT_LIB), "libhpmpc_b200_dbg.so")
from hpmpc_b200.batchgen import BatchSpec
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L = capi.product()
spec = BatchSpec("cfg3"); h = spec.h
d_in = spec.torch_batch(n)
z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
ux, pi, lam, t, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * 40)
dbg = torch.zeros(8000, dtype=torch.int64, device="cuda")
for rep in range(2):
    dbg.zero_()
    L.hb_debug_timing(C.c_void_p(dbg.data_ptr()))
    L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, d_in.data_ptr(), 40, C.c_double(2.0), C.c_double(1e-8), C.c_double(1e-8), 0, ux.data_ptr(), pi.data_ptr(),
                                          lam.data_ptr(), t.data_ptr(), info.data_ptr(), None)
    torch.cuda.synchronize()
v = dbg.cpu().numpy().reshape(-1, 2)
v = v[v[:, 0] > 0]
names = {310: "bwd stage top", 311: "inputs ready", 312: "hooks done", 313: "Pb done", 314: "assembled", 315: "pre-factor", 316: "factored", 300: "iter top", 301: "factor done", 302: "fwd(sv) done", 303: "step A done", 304: "trs done", 305: "step B done", 306: "residuals done"}
agg = collections.OrderedDict()
for i in range(len(v) - 1):
    key = (int(v[i, 0]), int(v[i + 1, 0]))
    agg.setdefault(key, []).append(int(v[i + 1, 1] - v[i, 1]))
for (a, b), ds in agg.items():
    print(f"{names.get(a, a):>16s} -> {names.get(b, b):<16s} n={len(ds):4d} mean={sum(ds) / len(ds):10.0f}")
print("kk of instance 0:", int(info[0, 0]))
