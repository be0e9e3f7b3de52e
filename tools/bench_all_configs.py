"""One line per BASELINE.json configuration that runs on the GPU (2: batched Riccati sv, 3: box IPM, 4: variable-size IPM,
5: scenario tree), kernels timed with CUDA events on data resident in HBM.  usage: python tools/bench_all_configs.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from hpmpc_b200 import capi, problems, tree as T
from hpmpc_b200.batchgen import BatchSpec

L = capi.product()
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))


z = lambda n, m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
# config 2
spec = BatchSpec("cfg2"); h = spec.h; n = 65536
d_in = spec.torch_batch(n); ux, pi = z(n, h.sz.ux_stride), z(n, h.sz.pi_stride)
ms = timeit(lambda: L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), None, st))
print(f"cfg2  Riccati sv   nx=12 nu=5 N=30          n={n:6d}  {ms:8.2f} ms  {n / ms * 1e3:12.0f} solves/s")
del d_in, ux, pi; h.close(); torch.cuda.empty_cache()
# configs 3 and 4: IPM
for name, n in (("cfg3", 16384), ("cfg4", 8192)):
    if name == "cfg3":
        spec = BatchSpec(name); h = spec.h; d_in = spec.torch_batch(n)
    else:
        p0 = problems.make("cfg4"); h = capi.BatchOcp(p0, device=0)
        base = torch.from_numpy(h.pack(p0)).cuda()
        d_in = base[None, :].repeat(n, 1)
        # distinct instances: scale the gradient rows (q, r) per instance (keeps every problem well posed)
        xi = torch.from_numpy(problems.instance_xi(n)[:, 2].copy()).cuda()
        for s in range(p0.N + 1):
            nux = p0.nx[s] + p0.nu[s]
            o = h.off[s]["RSQ"] + nux * (nux + 1) // 2
            d_in[:, o:o + nux] *= (1.0 + 0.3 * xi[:, None])
    ux, pi, lam, t, info = z(n, h.sz.ux_stride), z(n, h.sz.pi_stride), z(n, h.sz.lam_stride), z(n, h.sz.lam_stride), z(n, 6 + 5 * 40)
    ms = timeit(lambda: L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, d_in.data_ptr(), 40, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(), pi.data_ptr(),
                                                              lam.data_ptr(), t.data_ptr(), info.data_ptr(), st), reps=2)
    print(f"{name}  box IPM      {'nx=24 nu=11 N=50' if name == 'cfg3' else 'nx 40->4 nu=8 N=20'}      n={n:6d}  {ms:8.2f} ms  {n / ms * 1e3:12.0f} solves/s   "
          f"mean kk {float(info[:, 0].mean()):.2f}, converged {int((info[:, 1] == 0).sum())}")
    del d_in, ux, pi, lam, t, info; h.close(); torch.cuda.empty_cache()
# config 5: tree
t0 = T.mass_spring_tree(12, 5, 4, 3, 20); h = T.TreeBatch(t0, device=0); n = 1024
d_in = torch.from_numpy(h.pack(t0)).cuda()[None, :].repeat(n, 1)
ux, pi, Lst = z(n, h.sz.ux_stride), z(n, h.sz.pi_stride), z(n, h.sz.L_stride)
ms = timeit(lambda: h.L.hpmpc_b200_d_tree_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), Lst.data_ptr(), None))
print(f"cfg5  tree Riccati md=4 Nr=3 Nh=20 (1173 nodes) n={n:6d}  {ms:8.2f} ms  {n / ms * 1e3:12.0f} trees/s")
