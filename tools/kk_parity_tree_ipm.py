"""Iteration-count and solution parity of the scenario-tree IPM (BASELINE config 5 shape) against the oracle on a sample of
distinct trees: x0, Q and R scaling vary per tree (problems.instance_xi), u-bounds at every node.
usage: python tools/kk_parity_tree_ipm.py [n_trees]"""
import os, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
from hpmpc_b200 import problems, tree as T
from oracle import api as oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
xis = problems.instance_xi(n, first=1000)
trees = [T.mass_spring_tree(12, 5, 4, 3, 20, xi=tuple(xis[i]), bounds=True) for i in range(n)]
tb = T.TreeBatch(trees[0])
d_in = torch.from_numpy(np.stack([tb.pack(t) for t in trees])).cuda()
k_max = 40
z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
ux, pi, lam, tt, info = z(tb.sz.ux_stride), z(tb.sz.pi_stride), z(2 * tb.nbtot), z(2 * tb.nbtot), z(6 + 5 * k_max)
assert tb.L.hpmpc_b200_d_tree_ip2_res_mpc_hard_batch(tb.h, n, d_in.data_ptr(), k_max, C.c_double(2.0), C.c_double(1e-8), C.c_double(1e-8), 0,
                                                    ux.data_ptr(), pi.data_ptr(), lam.data_ptr(), tt.data_ptr(), info.data_ptr(), None) == 0
torch.cuda.synchronize()
uxh, pih, lamh, infoh = ux.cpu().numpy(), pi.cpu().numpy(), lam.cpu().numpy(), info.cpu().numpy()
cat = lambda v: np.concatenate([np.asarray(a).ravel() for a in v])


def check(i):
    o = oracle.tree_ipm(trees[i], k_max=k_max)
    kk, status = int(infoh[i, 0]), int(infoh[i, 1])
    u, x, p = tb.split(uxh[i], pih[i])
    got = dict(u=u, x=x, pi=p, lam=tb.split_lam(lamh[i]))
    e = max(float(np.max(np.abs(cat(got[f]) - cat(o[f])) / np.maximum(1.0, np.abs(cat(o[f]))))) for f in ("u", "x", "pi", "lam")) if kk == o["kk"] else float("nan")
    return kk == o["kk"], status == o["status"], e, kk


with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
    res = list(ex.map(check, range(n)))
same_kk = sum(r[0] for r in res); same_st = sum(r[1] for r in res)
errs = [r[2] for r in res if r[0]]
hist = {}
for r in res:
    hist[r[3]] = hist.get(r[3], 0) + 1
print(f"scenario-tree IPM, md=4 Nr=3 Nh=20 nx=12 nu=5 (1173 nodes, {tb.nbtot} bounds), {n} distinct trees vs oracle/ric_oracle.c orc_tree_ip2_res_mpc_hard")
print(f"  iteration count identical: {same_kk}/{n}   exit status identical: {same_st}/{n}")
print(f"  max relative error over u, x, pi, lam (trees with identical kk): {max(errs) if errs else float('nan'):.3e}   (bar: 1e-9)")
print(f"  iteration-count histogram (GPU): {dict(sorted(hist.items()))}")
