"""Config-4 bench batch: iteration counts of the team kernels (FP64 tensor-core factorisation) against the one-warp FMA kernels
(HPMPC_B200_TEAM=0); for every instance whose count differs, the last mu values of both runs next to the tolerance."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hpmpc_b200 import capi, problems
n = 8192
L = capi.product()
p0 = problems.make("cfg4"); h = capi.BatchOcp(p0, device=0)
d_in = torch.from_numpy(h.pack(p0)).cuda()[None, :].repeat(n, 1)
xi = torch.from_numpy(problems.instance_xi(n)[:, 2].copy()).cuda()
for s in range(p0.N + 1):
    nux = p0.nx[s] + p0.nu[s]
    o = h.off[s]["RSQ"] + nux * (nux + 1) // 2
    d_in[:, o:o + nux] *= (1.0 + 0.3 * xi[:, None])
z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
res = {}
for team in ("1", "0"):
    os.environ["HPMPC_B200_TEAM"] = team
    ux, pi, lam, t, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * 40)
    L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, d_in.data_ptr(), 40, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(), pi.data_ptr(), lam.data_ptr(), t.data_ptr(), info.data_ptr(), None)
    torch.cuda.synchronize()
    res[team] = (ux.cpu().numpy(), info.cpu().numpy())
a, b = res["1"][1], res["0"][1]
diff = np.nonzero(a[:, 0] != b[:, 0])[0]
print("instances with different iteration counts:", len(diff), "of", n)
for i in diff[:8]:
    ka, kb = int(a[i, 0]), int(b[i, 0])
    print(f"  instance {i}: team kk {ka}, one-warp kk {kb}; info head team {a[i, :6]}, one-warp {b[i, :6]}")
    for k in range(max(ka, kb)):
        print(f"     it {k}: team stat {a[i, 6 + 5 * k:11 + 5 * k]}   one-warp stat {b[i, 6 + 5 * k:11 + 5 * k]}")
same = a[:, 0] == b[:, 0]
ua, ub = res["1"][0][same], res["0"][0][same]
print("max rel diff of ux over instances with equal counts:", float(np.max(np.abs(ua - ub) / np.maximum(1.0, np.abs(ub)))))
