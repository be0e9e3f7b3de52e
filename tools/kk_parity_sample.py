"""Iteration-count and solution parity of the batched IPM on BASELINE config 3 against the oracle on a large sample
(the GPU test-suite checks a handful of instances; this checks `n_check` of them with the host cores in parallel).
usage: python tools/kk_parity_sample.py [n_inst] [n_check] [cfg]"""
import os, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from hpmpc_b200 import capi
from hpmpc_b200.batchgen import BatchSpec
from oracle import api as oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n_check = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
cfg = sys.argv[3] if len(sys.argv) > 3 else "cfg3"
L = capi.product()
spec = BatchSpec(cfg); h = spec.h
d_in = spec.torch_batch(n)
z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
ux, pi, lam, t, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * 40)
assert L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, d_in.data_ptr(), 40, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(), pi.data_ptr(), lam.data_ptr(),
                                             t.data_ptr(), info.data_ptr(), None) == 0
torch.cuda.synchronize()
uxh, pih, lamh, infoh = ux.cpu().numpy(), pi.cpu().numpy(), lam.cpu().numpy(), info.cpu().numpy()
idx = np.unique(np.linspace(0, n - 1, n_check).astype(int))


def rel(a, b):
    return max((float(np.max(np.abs(x - y) / np.maximum(1.0, np.abs(y)))) if len(y) else 0.0) for x, y in zip(a, b))


def check(i):
    o = oracle.ipm(spec.problem(int(i)))
    kk, status = int(infoh[i, 0]), int(infoh[i, 1])
    u, x = h.split_ux(uxh[i])
    e = max(rel(u, o["u"]), rel(x, o["x"]), rel(h.split_pi(pih[i]), o["pi"]), rel(h.split_lam(lamh[i]), o["lam"])) if kk == o["kk"] else float("nan")
    return kk == o["kk"], status == o["status"], e, kk


with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
    res = list(ex.map(check, idx))
kk_ok = sum(r[0] for r in res); st_ok = sum(r[1] for r in res)
errs = np.array([r[2] for r in res if r[0]])
hist = np.bincount([r[3] for r in res])
print(f"{cfg}: {len(idx)} of {n} instances checked against the oracle (oracle/ric_oracle.c, pinned on the reference's C99 build)")
print(f"  iteration count identical: {kk_ok}/{len(idx)}   exit status identical: {st_ok}/{len(idx)}")
print(f"  max relative error over u, x, pi, lam (instances with identical kk): {errs.max():.3e}   (bar: 1e-9)")
print("  iteration-count histogram (GPU):", {k: int(v) for k, v in enumerate(hist) if v})
