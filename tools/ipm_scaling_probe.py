import os, sys
sys.path.insert(0, "/root/repo")
import torch
from hpmpc_b200 import capi
from hpmpc_b200.batchgen import BatchSpec
n = 16384
L = capi.product(); spec = BatchSpec("cfg3"); h = spec.h
d_in = spec.torch_batch(n)
z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
ux, pi, lam, t, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * 40)
st = torch.cuda.current_stream().cuda_stream
def run(first, m):
    o = lambda T: T.data_ptr() + first * T.shape[1] * 8
    return L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, m, o(d_in), 40, 2.0, 1e-8, 1e-8, 0, o(ux), o(pi), o(lam), o(t), o(info), st)
def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
print("one call n=16384: %.1f ms" % timeit(lambda: run(0, n)))
print("8 calls n=2048 (distinct slices): %.1f ms" % timeit(lambda: [run(k * 2048, 2048) for k in range(8)]))
print("8 calls n=2048 (same slice): %.1f ms" % timeit(lambda: [run(0, 2048) for k in range(8)]))
print("4 calls n=4096 (distinct slices): %.1f ms" % timeit(lambda: [run(k * 4096, 4096) for k in range(4)]))
print("14 calls n=1184 (distinct): %.1f ms" % timeit(lambda: [run(k * 1184, 1184) for k in range(13)]))
