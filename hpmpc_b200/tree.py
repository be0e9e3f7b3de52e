"""Scenario-tree OCPs: topology, synthetic mass-spring tree problems, the stacked-chain reformulation used as oracle,
and the ctypes view of the tree handle of libhpmpc_b200.so (include/hpmpc_b200_tree.h).

Reference recipes restated (paths relative to the HPMPC tree):
  setup_tree / node count   test_problems/test_d_tree_ip_hard_libstr.c:61-176 (BFS numbering; md kids while stage < Nr,
                            one kid while stage < Nh, none at stage Nh)
  cost scaling              :776-800 (stage s < Nr carries md^(Nr-s) times the nominal cost)
  stacked chain problem     test_problems/test_d_tree_ric_libstr.c:797-1018 (stage s = all nodes of stage s, [all u ; all x])
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np

from .problems import Ocp, mass_spring_AB


class Node(C.Structure):          # reference include/tree.h:34-44
    _fields_ = [("kids", C.POINTER(C.c_int)), ("idx", C.c_int), ("dad", C.c_int), ("nkids", C.c_int), ("stage", C.c_int),
                ("real", C.c_int), ("idxkid", C.c_int)]


class TreeSizes(C.Structure):
    _fields_ = [("in_stride", C.c_longlong), ("ux_stride", C.c_longlong), ("pi_stride", C.c_longlong), ("L_stride", C.c_longlong),
                ("Nn", C.c_int), ("nzM", C.c_int), ("nxM", C.c_int), ("n_tails", C.c_int), ("n_top_nodes", C.c_int), ("cut_stage", C.c_int),
                ("n_shard_nodes", C.c_int)]


def number_of_nodes(md: int, Nr: int, Nh: int) -> int:
    if md == 1:
        return Nh + 1
    return (Nh - Nr) * md ** Nr + (md ** (Nr + 1) - 1) // (md - 1)


def setup_tree(md: int, Nr: int, Nh: int):
    """BFS tree: arrays dad, first_kid, nkids, stage, real (one entry per node)."""
    Nn = number_of_nodes(md, Nr, Nh)
    dad, stage, real = [-1], [0], [-1]
    first_kid, nkids = [], []
    nxt = 1
    n = 0
    while n < len(dad):
        s = stage[n]
        k = md if s < Nr else (1 if s < Nh else 0)
        nkids.append(k)
        first_kid.append(nxt if k else -1)
        for i in range(k):
            dad.append(n); stage.append(s + 1); real.append(i if k > 1 else max(real[n], 0))
        nxt += k
        n += 1
    assert len(dad) == Nn, (len(dad), Nn)
    return dict(Nn=Nn, dad=dad, first_kid=first_kid, nkids=nkids, stage=stage, real=real, md=md, Nr=Nr, Nh=Nh)


@dataclass
class TreeOcp:
    """Unconstrained LQ problem on a tree; node/edge-indexed dense arrays (edge k = the edge into node k, k >= 1)."""
    topo: dict
    nx: List[int]
    nu: List[int]
    A: List[np.ndarray] = field(default_factory=list)    # [Nn] (entry 0 unused) nx_k x nx_dad
    B: List[np.ndarray] = field(default_factory=list)    # nx_k x nu_dad
    b: List[np.ndarray] = field(default_factory=list)
    Q: List[np.ndarray] = field(default_factory=list)
    S: List[np.ndarray] = field(default_factory=list)    # nu x nx
    R: List[np.ndarray] = field(default_factory=list)
    q: List[np.ndarray] = field(default_factory=list)
    r: List[np.ndarray] = field(default_factory=list)
    # box bounds per node (empty lists = unconstrained): idxb[n] indexes [u_n ; x_n], lb/ub have nb[n] entries
    nb: List[int] = field(default_factory=list)
    idxb: List[np.ndarray] = field(default_factory=list)
    lb: List[np.ndarray] = field(default_factory=list)
    ub: List[np.ndarray] = field(default_factory=list)


def mass_spring_tree(nx: int, nu: int, md: int, Nr: int, Nh: int, xi=(0.0, 0.0, 0.0, 0.0), bounds: bool = False) -> TreeOcp:
    """per-branch dynamics A, B (1 + 0.05 real), cost scaled md^(Nr - stage) on the robust
    stages, x0 eliminated at the root (nx[0] = 0, b of the root's edges = A x0 + b); bounds: u in [-0.5, 0.5] on every
    input of every node (test_d_tree_ip_hard_libstr.c: "u-bounds")."""
    topo = setup_tree(md, Nr, Nh)
    Nn = topo["Nn"]
    A0, B0 = mass_spring_AB(nx, nu)
    x0 = np.zeros(nx); x0[0] = 2.5 * (1 + 0.2 * xi[0]); x0[1] = 2.5 * (1 + 0.2 * xi[1])
    qs, rs = 1.0 + 0.1 * xi[2], 2.0 + 0.2 * xi[3]
    nxs = [0] + [nx] * (Nn - 1)
    nus = [nu if topo["nkids"][n] > 0 else 0 for n in range(Nn)]
    t = TreeOcp(topo=topo, nx=nxs, nu=nus)
    for n in range(Nn):
        s, d = topo["stage"][n], topo["dad"][n]
        if n == 0:
            t.A.append(np.zeros((0, 0))); t.B.append(np.zeros((0, 0))); t.b.append(np.zeros(0))
        else:
            Bn = B0 * (1.0 + 0.05 * max(topo["real"][n], 0))
            if d == 0:
                t.A.append(np.zeros((nx, 0))); t.b.append(A0 @ x0 + 0.1)
            else:
                t.A.append(A0.copy()); t.b.append(0.1 * np.ones(nx))
            t.B.append(np.ascontiguousarray(Bn))
        w = float(md ** (Nr - s)) if s < Nr else 1.0
        t.Q.append(w * qs * np.eye(nxs[n])); t.S.append(np.zeros((nus[n], nxs[n]))); t.R.append(w * rs * np.eye(nus[n]))
        t.q.append(w * 0.1 * np.ones(nxs[n])); t.r.append(w * 0.2 * np.ones(nus[n]))
        if bounds:
            t.nb.append(nus[n]); t.idxb.append(np.arange(nus[n], dtype=np.int32))
            t.lb.append(-0.5 * np.ones(nus[n])); t.ub.append(0.5 * np.ones(nus[n]))
    return t


def stacked_chain(t: TreeOcp):
    """The chain OCP whose stage s stacks all tree nodes of stage s (variables [all u ; all x]); returns (Ocp, index maps).
    Solving it with the ordinary chain Riccati gives the tree solution (reference test_d_tree_ric_libstr.c:797-1018)."""
    topo = t.topo
    Nh = max(topo["stage"])
    levels = [[n for n in range(topo["Nn"]) if topo["stage"][n] == s] for s in range(Nh + 1)]
    nx2 = [sum(t.nx[n] for n in lv) for lv in levels]
    nu2 = [sum(t.nu[n] for n in lv) for lv in levels]
    uo = [{n: sum(t.nu[m] for m in lv[:i]) for i, n in enumerate(lv)} for lv in levels]
    xo = [{n: sum(t.nx[m] for m in lv[:i]) for i, n in enumerate(lv)} for lv in levels]
    has_b = len(t.nb) > 0
    idx2, lb2, ub2 = [], [], []
    for s, lv in enumerate(levels):
        ii, ll, uu = [], [], []
        for n in (lv if has_b else []):
            for j in range(t.nb[n]):
                i = int(t.idxb[n][j])
                ii.append(uo[s][n] + i if i < t.nu[n] else nu2[s] + xo[s][n] + i - t.nu[n]); ll.append(t.lb[n][j]); uu.append(t.ub[n][j])
        o = np.argsort(np.asarray(ii, dtype=np.int64), kind="stable") if ii else np.zeros(0, dtype=np.int64)
        idx2.append(np.asarray(ii, dtype=np.int32)[o] if ii else np.zeros(0, dtype=np.int32))
        lb2.append(np.asarray(ll, dtype=np.float64)[o] if ii else np.zeros(0)); ub2.append(np.asarray(uu, dtype=np.float64)[o] if ii else np.zeros(0))
    p = Ocp(N=Nh, nx=nx2, nu=nu2, nb=[len(v) for v in idx2], idxb=idx2)
    for s in range(Nh + 1):
        lv = levels[s]
        Q = np.zeros((nx2[s], nx2[s])); R = np.zeros((nu2[s], nu2[s])); S = np.zeros((nu2[s], nx2[s]))
        q = np.zeros(nx2[s]); r = np.zeros(nu2[s])
        for n in lv:
            a, e = xo[s][n], uo[s][n]
            Q[a:a + t.nx[n], a:a + t.nx[n]] = t.Q[n]; R[e:e + t.nu[n], e:e + t.nu[n]] = t.R[n]
            S[e:e + t.nu[n], a:a + t.nx[n]] = t.S[n]; q[a:a + t.nx[n]] = t.q[n]; r[e:e + t.nu[n]] = t.r[n]
        p.Q.append(Q); p.R.append(R); p.S.append(S); p.q.append(q); p.r.append(r)
        p.lb.append(lb2[s]); p.ub.append(ub2[s])
        if s < Nh:
            A = np.zeros((nx2[s + 1], nx2[s])); B = np.zeros((nx2[s + 1], nu2[s])); b = np.zeros(nx2[s + 1])
            for k in levels[s + 1]:
                d = topo["dad"][k]
                a1 = xo[s + 1][k]
                A[a1:a1 + t.nx[k], xo[s][d]:xo[s][d] + t.nx[d]] = t.A[k]
                B[a1:a1 + t.nx[k], uo[s][d]:uo[s][d] + t.nu[d]] = t.B[k]
                b[a1:a1 + t.nx[k]] = t.b[k]
            p.A.append(A); p.B.append(B); p.b.append(b)
    return p, dict(levels=levels, uo=uo, xo=xo)


def unstack(t: TreeOcp, maps, sol):
    """Chain solution of the stacked problem -> node-indexed (u, x, pi)."""
    topo = t.topo
    u = [None] * topo["Nn"]; x = [None] * topo["Nn"]; pi = [None] * topo["Nn"]
    for s, lv in enumerate(maps["levels"]):
        for n in lv:
            u[n] = sol["u"][s][maps["uo"][s][n]:maps["uo"][s][n] + t.nu[n]] if s < len(sol["u"]) else np.zeros(0)
            x[n] = sol["x"][s][maps["xo"][s][n]:maps["xo"][s][n] + t.nx[n]]
            pi[n] = sol["pi"][s - 1][maps["xo"][s][n]:maps["xo"][s][n] + t.nx[n]] if s > 0 else np.zeros(0)
    return u, x, pi


class TreeBatch:
    """hpmpc_b200_tree handle + numpy-side packing / splitting."""

    def __init__(self, t: TreeOcp, device: int = 0):
        from . import capi
        L = capi.product()
        self.L, self.t, self.device = L, t, device
        topo = t.topo
        Nn = topo["Nn"]
        self._kids = [(C.c_int * max(topo["nkids"][n], 1))(*[topo["first_kid"][n] + i for i in range(topo["nkids"][n])]) for n in range(Nn)]
        self._nodes = (Node * Nn)()
        for n in range(Nn):
            nd = self._nodes[n]
            nd.kids = C.cast(self._kids[n], C.POINTER(C.c_int)); nd.idx = n; nd.dad = topo["dad"][n]; nd.nkids = topo["nkids"][n]
            nd.stage = topo["stage"][n]; nd.real = topo["real"][n]; nd.idxkid = 0
        L.hpmpc_b200_tree_create.restype = C.c_int
        L.hpmpc_b200_tree_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hpmpc_b200_tree_destroy.argtypes = [C.c_void_p]
        L.hpmpc_b200_tree_sizes_get.argtypes = [C.c_void_p, C.POINTER(TreeSizes)]
        L.hpmpc_b200_tree_node_offsets.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 5
        L.hpmpc_b200_tree_tail_root.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3
        L.hpmpc_b200_tree_pack_instance.restype = C.c_int
        L.hpmpc_b200_tree_pack_instance.argtypes = [C.c_void_p] + [C.c_void_p] * 9
        L.hpmpc_b200_d_tree_back_ric_rec_sv_batch.restype = C.c_int
        L.hpmpc_b200_d_tree_back_ric_rec_sv_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 5
        L.hpmpc_b200_d_tree_back_ric_rec_sv_phase.restype = C.c_int
        L.hpmpc_b200_d_tree_back_ric_rec_sv_phase.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
        L.hpmpc_b200_tree_create_box.restype = C.c_int
        L.hpmpc_b200_tree_create_box.argtypes = [C.POINTER(C.c_void_p), C.c_int] + [C.c_void_p] * 5 + [C.c_int]
        L.hpmpc_b200_tree_pack_bounds.restype = C.c_int
        L.hpmpc_b200_tree_pack_bounds.argtypes = [C.c_void_p] * 4
        L.hpmpc_b200_tree_bound_offsets.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3
        L.hpmpc_b200_d_tree_ip2_res_mpc_hard_batch.restype = C.c_int
        L.hpmpc_b200_d_tree_ip2_res_mpc_hard_batch.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                                               C.c_int] + [C.c_void_p] * 6
        self.h = C.c_void_p()
        self.nb = list(t.nb) if t.nb else [0] * Nn
        if t.nb:
            self._idxb = [np.ascontiguousarray(t.idxb[n], dtype=np.int32) if self.nb[n] else np.zeros(1, dtype=np.int32) for n in range(Nn)]
            rc = L.hpmpc_b200_tree_create_box(C.byref(self.h), Nn, C.cast(self._nodes, C.c_void_p), capi.int_array(t.nx), capi.int_array(t.nu),
                                              capi.int_array(self.nb), capi.ptr_array(self._idxb), device)
        else:
            rc = L.hpmpc_b200_tree_create(C.byref(self.h), Nn, C.cast(self._nodes, C.c_void_p), capi.int_array(t.nx), capi.int_array(t.nu), device)
        if rc != 0:
            raise RuntimeError(f"hpmpc_b200_tree_create failed ({rc})")
        self.nbtot = sum(self.nb)
        self.sz = TreeSizes()
        L.hpmpc_b200_tree_sizes_get(self.h, C.byref(self.sz))
        self.off = []
        for n in range(Nn):
            v = [C.c_int() for _ in range(5)]
            L.hpmpc_b200_tree_node_offsets(self.h, n, *[C.byref(x) for x in v])
            self.off.append(dict(zip(("BAbt", "RSQ", "ux", "pi", "L"), [x.value for x in v])))
        L.hpmpc_b200_tree_shard_node.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 5
        self.subtrees = []
        for k in range(self.sz.n_shard_nodes):
            v = [C.c_int() for _ in range(5)]
            L.hpmpc_b200_tree_shard_node(self.h, k, *[C.byref(x) for x in v])
            self.subtrees.append(dict(node=v[0].value, off_L=v[1].value, len_L=v[2].value, tail_lo=v[3].value, tail_hi=v[4].value))
        self.tails = []
        for j in range(self.sz.n_tails):
            v = [C.c_int() for _ in range(3)]
            L.hpmpc_b200_tree_tail_root(self.h, j, *[C.byref(x) for x in v])
            self.tails.append(dict(node=v[0].value, off_L=v[1].value, len_L=v[2].value))

    def close(self):
        if self.h:
            self.L.hpmpc_b200_tree_destroy(self.h)
            self.h = C.c_void_p()

    def pack(self, t: TreeOcp) -> np.ndarray:
        from .capi import ptr_array
        blk = np.zeros(self.sz.in_stride)
        f = lambda M: np.asfortranarray(M, dtype=np.float64)
        arrs = [[f(M) for M in L] for L in (t.A, t.B)] + [[np.ascontiguousarray(v) for v in t.b]] + \
               [[f(M) for M in L] for L in (t.Q, t.S, t.R)] + [[np.ascontiguousarray(v) for v in L] for L in (t.q, t.r)]
        self._keep = arrs
        rc = self.L.hpmpc_b200_tree_pack_instance(self.h, *[ptr_array(a) for a in arrs], blk.ctypes.data)
        assert rc == 0
        if t.nb:
            lb = [np.ascontiguousarray(v, dtype=np.float64) if len(v) else np.zeros(1) for v in t.lb]
            ub = [np.ascontiguousarray(v, dtype=np.float64) if len(v) else np.zeros(1) for v in t.ub]
            rc = self.L.hpmpc_b200_tree_pack_bounds(self.h, ptr_array(lb), ptr_array(ub), blk.ctypes.data)
            assert rc == 0
        return blk

    def split_lam(self, lam: np.ndarray):
        """per tree [node: lower(nb) upper(nb)] -> list per node"""
        out, o = [], 0
        for n in range(self.t.topo["Nn"]):
            out.append(lam[o:o + 2 * self.nb[n]].copy()); o += 2 * self.nb[n]
        return out

    def split(self, ux: np.ndarray, pi: np.ndarray):
        t = self.t
        Nn = t.topo["Nn"]
        u = [ux[self.off[n]["ux"]:self.off[n]["ux"] + t.nu[n]].copy() for n in range(Nn)]
        x = [ux[self.off[n]["ux"] + t.nu[n]:self.off[n]["ux"] + t.nu[n] + t.nx[n]].copy() for n in range(Nn)]
        p = [pi[self.off[n]["pi"]:self.off[n]["pi"] + t.nx[n]].copy() if n > 0 else np.zeros(0) for n in range(Nn)]
        return u, x, p
