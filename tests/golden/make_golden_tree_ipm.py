"""Generate tests/golden/golden_tree_ipm_v1.npz from the REAL reference: each bounded tree problem is turned into the stacked
chain problem of the reference's own tree tests (test_problems/test_d_tree_ric_libstr.c:797-1018; bounds stacked the same
way) and solved by the reference's lib4 IPM fortran_order_d_ip_ocp_hard_tv (oracle/_ref/libhpmpc_ref_c99.so, built from
/root/reference by oracle/Makefile; the tree IPM of the reference itself needs BLASFEO, which is absent).  The chain
solution is mapped back to node-indexed u, x, pi; lam is kept in the stacked order together with the map.
Run in the build container:  python tests/golden/make_golden_tree_ipm.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hpmpc_b200 import problems, tree as T  # noqa: E402
from oracle import api  # noqa: E402

CASES = {   # name: (nx, nu, md, Nr, Nh, instance id)
    "ipm_path": (4, 2, 1, 0, 6, 0),
    "ipm_2x2": (4, 2, 2, 2, 5, 1),
    "ipm_3x2": (6, 2, 3, 2, 5, 2),
    "ipm_2x1": (8, 3, 2, 1, 4, 3),
    "ipm_cfg5_small": (12, 5, 4, 2, 6, 4),      # config-5 sizes, robust horizon 2
}
K_MAX, MU0, MU_TOL = 40, 2.0, 1e-8


def build(case):
    nx, nu, md, Nr, Nh, inst = CASES[case]
    xi = tuple(problems.instance_xi(1, first=inst)[0])
    return T.mass_spring_tree(nx, nu, md, Nr, Nh, xi=xi, bounds=True)


def node_lam(t, maps, p, lam_chain):
    """lam of the stacked chain (per level [lower ; upper], bounds sorted by stacked index) -> per node [lower ; upper]."""
    out = []
    for n in range(t.topo["Nn"]):
        s = t.topo["stage"][n]
        nb2 = p.nb[s]
        lo, up = [], []
        for j in range(t.nb[n]):
            i = int(t.idxb[n][j])
            k = maps["uo"][s][n] + i if i < t.nu[n] else p.nu[s] + maps["xo"][s][n] + i - t.nu[n]
            pos = int(np.searchsorted(p.idxb[s], k))
            lo.append(lam_chain[s][pos]); up.append(lam_chain[s][nb2 + pos])
        out.append(np.asarray(lo + up))
    return out


if __name__ == "__main__":
    assert api.have_reference(), "needs oracle/_ref/libhpmpc_ref_c99.so (make -C oracle with /root/reference present)"
    ref = api.reference("c99")
    out = {}
    for case in CASES:
        t = build(case)
        p, maps = T.stacked_chain(t)
        sol = ref.ip_ocp_hard_tv(p, k_max=K_MAX, mu0=MU0, mu_tol=MU_TOL)
        u, x, pi = T.unstack(t, maps, sol)
        lam = node_lam(t, maps, p, sol["lam"])
        for f, v in (("u", u), ("x", x), ("pi", pi), ("lam", lam)):
            out[f"{case}/{f}"] = np.concatenate([np.asarray(a).ravel() for a in v])
        out[f"{case}/kk"] = np.asarray([sol["kk"], sol["status"]])
        print(case, "kk", sol["kk"], "status", sol["status"])
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "golden_tree_ipm_v1.npz"), **out)
    print("wrote", len(out), "arrays")
