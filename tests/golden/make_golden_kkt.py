"""Golden vectors for the re-solve with a new right-hand side, produced by the REAL reference.

Run in the build container (needs /root/reference compiled into oracle/_ref by `make -C oracle`):
    python tests/golden/make_golden_kkt.py
Each case runs the reference's d_ip2_res_mpc_hard_tv on a mass-spring problem p and then its
d_kkt_solve_new_rhs_res_mpc_hard_tv (mpc_solvers/d_ip2_res_hard.c:1922) on the same work memory with the vectors
(b, q, r, lb, ub) of `perturbed(p)`, exactly as test_problems/test_d_ip_hard.c:1040 chains the two calls (C99_4X4 build).
The .npz travels to the GPU box; /root/reference does not.
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from hpmpc_b200 import problems  # noqa: E402

# name -> (nx, nu, N, keyword arguments of mass_spring_ocp, xi)
CASES = {
    "ms_8_3_10": (8, 3, 10, {}, (0.3, -0.2, 0.1, 0.4)),
    "ms_12_5_30": (12, 5, 30, {}, (-0.5, 0.7, 0.2, -0.3)),
    "ms_24_11_50": (24, 11, 50, {}, (0.1, 0.1, -0.6, 0.9)),
    "ms_8_3_10_free_x0": (8, 3, 10, dict(free_x0=True), (0.3, -0.2, 0.1, 0.4)),
    "ms_4_2_5": (4, 2, 5, {}, (0.9, -0.9, 0.0, 0.0)),
}
K_MAX, MU0, MU_TOL = 40, 2.0, 1e-8


def build(name):
    nx, nu, N, kw, xi = CASES[name]
    return problems.mass_spring_ocp(nx, nu, N, bounds=True, xi=xi, **kw)


def perturbed(p, seed=5, scale=0.05):
    """Same sizes and matrices, new b, q, r and slightly moved bounds: the 'new right-hand side'."""
    rng = np.random.default_rng(seed)
    p2 = copy.deepcopy(p)
    p2.b = [v + scale * rng.standard_normal(v.shape) for v in p.b]
    p2.q = [v + scale * rng.standard_normal(v.shape) for v in p.q]
    p2.r = [v + scale * rng.standard_normal(v.shape) for v in p.r]
    p2.lb = [v - 0.01 for v in p.lb]
    p2.ub = [v + 0.02 for v in p.ub]
    return p2


def cat(v):
    v = [np.asarray(a, dtype=np.float64).ravel() for a in v]
    return np.concatenate(v) if v else np.zeros(0)


def main():
    from oracle import api
    assert api.have_reference(), "compile the reference first: make -C oracle"
    ref = api.reference("c99")
    out = {}
    for name in CASES:
        p = build(name)
        r = ref.ip2_then_kkt_new_rhs(p, perturbed(p), k_max=K_MAX, mu0=MU0, mu_tol=MU_TOL)
        out[f"{name}/kk"] = np.array([r["kk"], r["status"]])
        for f in ("u", "x", "pi", "lam", "t"):
            out[f"{name}/{f}"] = cat(r[f])
        print(name, "kk", r["kk"], "status", r["status"], "|u|max", float(np.max(np.abs(out[f"{name}/u"]))))
    np.savez_compressed(os.path.join(HERE, "golden_kkt_v1.npz"), **out)


if __name__ == "__main__":
    main()
