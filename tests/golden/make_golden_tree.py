"""Generate tests/golden/golden_tree_v1.npz from the REAL reference: each tree problem is turned into the stacked chain
problem of the reference's own tree test (test_problems/test_d_tree_ric_libstr.c:797-1018) and solved by the reference's
lib4 d_back_ric_rec_sv_tv_res (oracle/_ref/libhpmpc_ref_c99.so, built from /root/reference by oracle/Makefile); the chain
solution is mapped back to node-indexed u, x, pi.  Run in the build container:  python tests/golden/make_golden_tree.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hpmpc_b200 import problems, tree as T  # noqa: E402
from oracle import api  # noqa: E402

CASES = {   # name: (nx, nu, md, Nr, Nh, instance id)
    "tree_path": (4, 2, 1, 0, 6, 0),
    "tree_2x2": (4, 2, 2, 2, 5, 1),
    "tree_3x2": (6, 2, 3, 2, 5, 2),
    "tree_2x1": (8, 3, 2, 1, 4, 3),
    "tree_cfg5_small": (12, 5, 4, 2, 6, 4),      # config-5 sizes, robust horizon 2 (stacks to nx2 = 192)
}


def build(case):
    nx, nu, md, Nr, Nh, inst = CASES[case]
    xi = tuple(problems.instance_xi(1, first=inst)[0])
    return T.mass_spring_tree(nx, nu, md, Nr, Nh, xi=xi)


if __name__ == "__main__":
    assert api.have_reference(), "needs oracle/_ref/libhpmpc_ref_c99.so (make -C oracle with /root/reference present)"
    ref = api.reference("c99")
    out = {}
    for case in CASES:
        t = build(case)
        p, maps = T.stacked_chain(t)
        u, x, pi = T.unstack(t, maps, ref.ric(p, "sv"))
        for f, v in (("u", u), ("x", x), ("pi", pi)):
            out[f"{case}/{f}"] = np.concatenate([np.asarray(a).ravel() for a in v])
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "golden_tree_v1.npz"), **out)
    print("wrote", len(out), "arrays")
