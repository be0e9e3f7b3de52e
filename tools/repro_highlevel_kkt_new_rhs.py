"""Repro (build container only, needs oracle/_ref/libhpmpc_ref_c99.so): the REFERENCE's high-level re-solve pair
fortran_order_d_ip_ocp_hard_tv + fortran_order_d_solve_kkt_new_rhs_ocp_hard_tv (include/c_interface.h:65,67) on the same work0
against its own low-level pair d_ip2_res_mpc_hard_tv + d_kkt_solve_new_rhs_res_mpc_hard_tv (include/mpc_solvers.h:42,46).

The second high-level routine lays work0 out differently from the first: interfaces/c/fortran_order_interface.c:1193 places the IPM
work space right behind the packed matrices, :459 (the IPM wrapper) places it last, behind the partial-condensing and
residual arrays -- so the re-solve reads its factor, t_inv and backup iterate from memory the IPM never wrote them to.
usage: python tools/repro_highlevel_kkt_new_rhs.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hpmpc_b200 import problems
from oracle import api as oracle

ref = oracle.reference("c99")
p = problems.mass_spring_ocp(8, 3, 10, bounds=True, xi=(0.3, -0.1, 0.2, 0.1))
p2 = problems.mass_spring_ocp(8, 3, 10, bounds=True, xi=(0.35, -0.05, 0.2, 0.1))
for n in range(p.N + 1):
    p2.Q[n], p2.R[n], p2.S[n] = p.Q[n], p.R[n], p.S[n]
lo = ref.ip2_then_kkt_new_rhs(p, p2)
hi = ref.ip_then_solve_kkt_new_rhs_high_level(p, p2)
cat = lambda L: np.concatenate([np.asarray(v).ravel() for v in L])
print("low-level pair : kk =", lo["kk"], " max|u| = %.6f" % np.max(np.abs(cat(lo["u"]))), " u0 =", lo["u"][0])
print("high-level pair: kk =", hi["kk"], " max|u| = %.6e" % np.max(np.abs(cat(hi["u"]))), " u0 =", hi["u"][0])
print("max |u_high - u_low| = %.3e   (bounds are |u| <= 0.5)" % np.max(np.abs(cat(hi["u"]) - cat(lo["u"]))))
