"""Generate tests/golden/golden_newton_v1.npz: the reference's single Newton step (fortran_order_d_ip_ocp_hard_tv_single_newton_step,
interfaces/c/fortran_order_interface.c:695 -> d_ip2_res_mpc_hard_tv_single_newton_step, mpc_solvers/d_ip2_res_hard.c:1348) from an
interior iterate, by the compiled reference (oracle/_ref/libhpmpc_ref_c99.so).  The starting iterate is the reference IPM's own
iterate after a few iterations (d_ip2_res_mpc_hard_tv with a small k_max) and is stored with the results.
    python tests/golden/make_golden_newton.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hpmpc_b200 import problems  # noqa: E402
from oracle import api  # noqa: E402

CASES = {
    "nt_8_3_10_one": dict(shape=(8, 3, 10), xi=(0.3, -0.2, 0.1, 0.4), k0=3, k_max=1, mu0=1e-3),
    "nt_8_3_10_three": dict(shape=(8, 3, 10), xi=(-0.5, 0.6, 0.2, -0.3), k0=4, k_max=3, mu0=1e-4),
    "nt_12_5_8": dict(shape=(12, 5, 8), xi=(0.1, 0.1, -0.4, 0.2), k0=2, k_max=2, mu0=1e-2),
    "nt_cfg3": dict(shape=(24, 11, 50), xi=(0.2, -0.3, 0.1, 0.0), k0=4, k_max=1, mu0=1e-4),
}


def build_problem(name):
    c = CASES[name]
    return problems.mass_spring_ocp(*c["shape"], bounds=True, xi=c["xi"])


def cat(L):
    return np.concatenate([np.asarray(v, dtype=np.float64).ravel() for v in L] + [np.zeros(0)])


def main():
    ref = api.reference("c99")
    out = {}
    for name, c in CASES.items():
        p = build_problem(name)
        s = ref.ip2_res_mpc_hard_tv(p, k_max=c["k0"])
        ux0 = [np.concatenate([s["u"][n] if n < p.N else np.zeros(0), s["x"][n]]) for n in range(p.N + 1)]
        r = ref.single_newton_step(p, ux0, s["pi"], s["lam"], s["t"], k_max=c["k_max"], mu0=c["mu0"])
        out[name + "/ux0"], out[name + "/pi0"], out[name + "/lam0"], out[name + "/t0"] = cat(ux0), cat(s["pi"]), cat(s["lam"]), cat(s["t"])
        for f in ("u", "x", "pi", "lam", "t"):
            out[name + "/" + f] = cat(r[f])
        out[name + "/kk"] = np.array(r["kk"]); out[name + "/status"] = np.array(r["status"])
        out[name + "/stat"] = r["stat"]; out[name + "/inf_norm_res"] = r["inf_norm_res"]
        print(name, "kk", r["kk"], "status", r["status"], r["inf_norm_res"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_newton_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
