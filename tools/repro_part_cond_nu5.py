"""Repro: the reference's lib4 partial condensing (lqcp_solvers/d_part_cond.c through fortran_order_d_ip_ocp_hard_tv with N2 < N) returns a
point that is NOT stationary when nu > 4 and a block holds three or more stages: its own exit norm inf_norm_res[0] (stationarity of the
FULL problem at the expanded solution) is O(1), while the same call is right for nu <= 4 or blocks of two stages.  The oracle's
restatement of the same algorithm (oracle/ric_oracle.c: orc_part_cond) gives a KKT point in every case and agrees with the reference
to 1e-14 wherever the reference is right.  Needs the reference build (oracle/_ref, `make -C oracle`); CPU only.
usage: python tools/repro_part_cond_nu5.py > profiles/rNN_repro_part_cond_nu5.txt"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from hpmpc_b200 import problems
from oracle import api

ref = api.reference("c99")
cat = lambda L: np.concatenate([np.asarray(v).ravel() for v in L])
print("shape (nx,nu,N)   N2  block   | reference: kk  |rq|inf      | oracle: kk  |rq|inf      | max |u_ref - u_oracle|")
for (nx, nu, N) in ((8, 3, 10), (12, 4, 10), (12, 5, 10), (10, 5, 10), (12, 6, 10), (24, 11, 10), (12, 5, 30)):
    p = problems.mass_spring_ocp(nx, nu, N, bounds=True, xi=(0.1, 0.2, -0.5, 0.7))
    for N2 in (N // 2, N // 3, 2, 1):
        r, o = ref.ip_ocp_hard_tv(p, N2=N2), api.ipm(p, N2=N2)
        T = -(-N // N2)
        flag = "   <-- reference not stationary" if r["inf_norm_res"][0] > 1e-6 else ""
        print(f"({nx:2d},{nu:2d},{N:2d})        {N2:3d}  {T:3d}     |  {r['kk']:3d}  {r['inf_norm_res'][0]:10.2e}     |  {o['kk']:3d}  {o['inf_norm_res'][0]:10.2e}     |"
              f"  {np.abs(cat(r['u']) - cat(o['u'])).max():9.2e}{flag}")
