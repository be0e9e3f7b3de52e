"""Synthetic OCP inputs: the reference's mass-spring chain, restated (test / bench harness, not the solver).

Follows the recipe of the reference's test programs (paths relative to the HPMPC tree):
  mass_spring_system      test_problems/test_d_ric_mpc.c:59-145 (continuous chain -> expm -> ZOH B)
  cost / bounds / x0      test_problems/test_d_ip_hard.c:165-185,345-410 ; test_d_ric_libstr.c:258-343
  x0 elimination          nx[0] = 0, b0 = A x0 + b  (test_d_ric_libstr.c:209,297)
Per-instance variation (not in the reference): SplitMix64(seed 20260101, stream i).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

SEED = 20260101
_M64 = (1 << 64) - 1


def splitmix64_uniform(seed: int, stream: np.ndarray, count: int) -> np.ndarray:
    """`count` U(-1,1) numbers for every stream id (vectorised SplitMix64)."""
    stream = np.asarray(stream, dtype=np.uint64)
    with np.errstate(over="ignore"):
        state = np.uint64(seed) + (stream + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        out = np.empty((stream.size, count), dtype=np.float64)
        for k in range(count):
            state = state + np.uint64(0x9E3779B97F4A7C15)
            z = state.copy()
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[:, k] = (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -53) * 2.0 - 1.0
    return out


def _expm(M: np.ndarray) -> np.ndarray:
    """Scaling-and-squaring Pade(6) matrix exponential (own implementation, no scipy dependency)."""
    n = M.shape[0]
    nrm = np.linalg.norm(M, 1)
    s = max(0, int(np.ceil(np.log2(max(nrm, 1e-300)))) + 1)
    A = M / (2.0 ** s)
    q = 6
    c = 0.5
    X = A.copy()
    I = np.eye(n)
    E = I + c * A
    D = I - c * A
    p = True
    for k in range(2, q + 1):
        c = c * (q - k + 1) / (k * (2 * q - k + 1))
        X = A @ X
        E = E + c * X
        D = D + (c * X if p else -c * X)
        p = not p
    E = np.linalg.solve(D, E)
    for _ in range(s):
        E = E @ E
    return E


def mass_spring_AB(nx: int, nu: int, Ts: float = 0.5):
    pp = nx // 2
    T = -2.0 * np.eye(pp) + np.eye(pp, k=1) + np.eye(pp, k=-1)
    Ac = np.zeros((nx, nx))
    Ac[:pp, pp:] = np.eye(pp)
    Ac[pp:, :pp] = T
    Bc = np.zeros((nx, nu))
    Bc[pp:pp + nu, :] = np.eye(nu)
    A = _expm(Ts * Ac)
    B = np.linalg.solve(Ac, (A - np.eye(nx)) @ Bc)
    return A, B


@dataclass
class Ocp:
    """One OCP in the reference's stage-wise form (dense, C-contiguous 2-D arrays = row-major)."""
    N: int
    nx: List[int]
    nu: List[int]          # length N+1, nu[N] = 0
    nb: List[int]
    idxb: List[np.ndarray]
    A: List[np.ndarray] = field(default_factory=list)   # [N]  nx1 x nx
    B: List[np.ndarray] = field(default_factory=list)   # [N]  nx1 x nu
    b: List[np.ndarray] = field(default_factory=list)   # [N]  nx1
    Q: List[np.ndarray] = field(default_factory=list)   # [N+1] nx x nx
    S: List[np.ndarray] = field(default_factory=list)   # [N+1] nu x nx
    R: List[np.ndarray] = field(default_factory=list)   # [N+1] nu x nu
    q: List[np.ndarray] = field(default_factory=list)
    r: List[np.ndarray] = field(default_factory=list)
    lb: List[np.ndarray] = field(default_factory=list)
    ub: List[np.ndarray] = field(default_factory=list)
    # general (polytopic) constraints lg <= D u + C x <= ug; empty lists = none (reference include/c_interface.h:62: C, D, lg, ug)
    ng: List[int] = field(default_factory=list)
    C: List[np.ndarray] = field(default_factory=list)   # [N+1] ng x nx
    D: List[np.ndarray] = field(default_factory=list)   # [N+1] ng x nu
    lg: List[np.ndarray] = field(default_factory=list)
    ug: List[np.ndarray] = field(default_factory=list)

    def ng_list(self):
        return list(self.ng) if self.ng else [0] * (self.N + 1)

    def general_arrays(self):
        """C, D, lg, ug with (ng x n) zero-size placeholders where a stage has no general constraints."""
        ng = self.ng_list()
        if not self.ng:
            z = [np.zeros((0, 1))] * (self.N + 1)
            return [np.zeros((0, self.nx[n])) for n in range(self.N + 1)], [np.zeros((0, self.nu[n])) for n in range(self.N + 1)], \
                [np.zeros(0)] * (self.N + 1), [np.zeros(0)] * (self.N + 1)
        return self.C, self.D, self.lg, self.ug


def mass_spring_ocp(nx: int, nu: int, N: int, *, bounds: bool = False, xi=(0.0, 0.0, 0.0, 0.0),
                    nx_profile: Optional[List[int]] = None, free_x0: bool = False) -> Ocp:
    """Mass-spring OCP.  xi = (xi1..xi4) in [-1,1] perturbs x0, Q, R as SURVEY.md section 8d describes.

    nx_profile (config 4): per-stage state sizes; A_n, B_n are the leading blocks of the nx_profile[0]-state system,
    x0 stays a free variable, only inputs are bounded.
    """
    if nx_profile is None:
        nxs = [0 if not free_x0 else nx] + [nx] * N
        nfull = nx
    else:
        nxs = list(nx_profile)
        nfull = max(nxs)
        assert len(nxs) == N + 1
    nus = [nu] * N + [0]
    A0, B0 = mass_spring_AB(nfull, nu)
    x0 = np.zeros(nfull)
    x0[0] = 2.5 * (1.0 + 0.2 * xi[0])
    x0[1] = 2.5 * (1.0 + 0.2 * xi[1])
    qs, rs = 1.0 + 0.1 * xi[2], 2.0 + 0.2 * xi[3]
    p = Ocp(N=N, nx=nxs, nu=nus, nb=[0] * (N + 1), idxb=[np.zeros(0, dtype=np.int32) for _ in range(N + 1)])
    for n in range(N + 1):
        nxn, nun = nxs[n], nus[n]
        if n < N:
            nx1 = nxs[n + 1]
            bn = 0.1 * np.ones(nx1)
            if nx_profile is None and not free_x0 and n == 0:
                An = np.zeros((nx1, 0))
                bn = A0 @ x0 + 0.1
            else:
                An = A0[:nx1, :nxn].copy()
            p.A.append(np.ascontiguousarray(An))
            p.B.append(np.ascontiguousarray(B0[:nx1, :nun]))
            p.b.append(bn)
        p.Q.append(qs * np.eye(nxn))
        p.S.append(np.zeros((nun, nxn)))
        p.R.append(rs * np.eye(nun))
        p.q.append(0.1 * np.ones(nxn))
        p.r.append(0.2 * np.ones(nun))
        if bounds:
            nbx = 0
            if nx_profile is None and not free_x0:
                nbx = nxn // 2 if n >= 1 else 0
            idx = list(range(nun)) + [nun + i for i in range(nbx)]
            lbn = [-0.5] * nun + [-4.0] * nbx
            ubn = [0.5] * nun + [4.0] * nbx
            p.nb[n] = len(idx)
            p.idxb[n] = np.asarray(idx, dtype=np.int32)
            p.lb.append(np.asarray(lbn, dtype=np.float64))
            p.ub.append(np.asarray(ubn, dtype=np.float64))
        else:
            p.lb.append(np.zeros(0))
            p.ub.append(np.zeros(0))
    return p


def add_general(p: Ocp, stage_rows) -> Ocp:
    """Attach general constraints: stage_rows maps stage -> (C, D, lg, ug); other stages get ng = 0."""
    N = p.N
    p.ng = [0] * (N + 1)
    p.C = [np.zeros((0, p.nx[n])) for n in range(N + 1)]
    p.D = [np.zeros((0, p.nu[n])) for n in range(N + 1)]
    p.lg = [np.zeros(0) for _ in range(N + 1)]
    p.ug = [np.zeros(0) for _ in range(N + 1)]
    for n, (C, D, lg, ug) in stage_rows.items():
        g = len(lg)
        C = np.ascontiguousarray(C, dtype=np.float64).reshape(g, p.nx[n])
        D = np.ascontiguousarray(D, dtype=np.float64).reshape(g, p.nu[n])
        p.ng[n] = g
        p.C[n], p.D[n] = C, D
        p.lg[n], p.ug[n] = np.asarray(lg, dtype=np.float64), np.asarray(ug, dtype=np.float64)
    return p


def guide_problem(N: int = 30, nx: int = 8, nu: int = 3, terminal_band: float = 0.0) -> Ocp:
    """The problem of the reference's user guide (doc/guide.tex:333-346): mass-spring chain, forces in [-0.5, 0.5], positions in
    [-4, 4], Q = I, R = 2 I, q = r = 0, and a terminal constraint x_N = 0 imposed as ngN = nx general constraints (C_N = I,
    lg = ug = 0).  x0 = (2.5, 2.5, 0, ...) is eliminated into b_0 as in test_problems/test_d_ip_hard.c:303-314."""
    p = mass_spring_ocp(nx, nu, N, bounds=True)
    for n in range(N + 1):
        p.q[n] = np.zeros(p.nx[n]); p.r[n] = np.zeros(p.nu[n])
    return add_general(p, {N: (np.eye(nx), np.zeros((nx, 0)), -terminal_band * np.ones(nx), terminal_band * np.ones(nx))})


def general_test_problem(nx: int = 8, nu: int = 3, N: int = 10, xi=(0.0, 0.0, 0.0, 0.0)) -> Ocp:
    """Box bounds plus general constraints at EVERY stage (inputs only at stage 0, mixed C / D rows in the middle, states only
    at stage N): exercises every ng > 0 code path with a different ng per stage."""
    p = mass_spring_ocp(nx, nu, N, bounds=True, xi=xi)
    rows = {}
    for n in range(N + 1):
        nxn, nun = p.nx[n], p.nu[n]
        if n == 0:
            rows[n] = (np.zeros((1, 0)), np.ones((1, nun)), [-1.2], [1.2])
        elif n < N:
            C = np.zeros((2, nxn)); C[0, 0] = 1.0; C[0, 1] = -1.0; C[1, nxn // 2] = 1.0
            D = np.zeros((2, nun)); D[1, :] = 0.5; D[0, 0] = 0.3
            rows[n] = (C, D, [-2.0, -1.5], [2.0, 1.5])
        else:
            C = np.zeros((3, nxn)); C[0, 0] = 1.0; C[0, 1] = 1.0; C[1, 2] = 1.0; C[2, nxn // 2 + 1] = 1.0
            rows[n] = (C, np.zeros((3, 0)), [-1.5, -1.0, -1.0], [1.5, 1.0, 1.0])
    return add_general(p, rows)


def config(name: str):
    """The BASELINE.json configurations: (nx, nu, N, bounds, nx_profile, n_inst)."""
    if name == "cfg1":
        return dict(nx=12, nu=5, N=10, bounds=False, n_inst=1)
    if name == "cfg2":
        return dict(nx=12, nu=5, N=30, bounds=False, n_inst=65536)
    if name == "cfg3":
        return dict(nx=24, nu=11, N=50, bounds=True, n_inst=16384)
    if name == "cfg4":
        N = 20
        return dict(nx=40, nu=8, N=N, bounds=True, nx_profile=[40 - (9 * n) // 5 for n in range(N + 1)], n_inst=8192)
    raise KeyError(name)


def make(name_or_cfg, xi=(0.0, 0.0, 0.0, 0.0), **over) -> Ocp:
    cfg = dict(config(name_or_cfg)) if isinstance(name_or_cfg, str) else dict(name_or_cfg)
    cfg.update(over)
    cfg.pop("n_inst", None)
    return mass_spring_ocp(cfg["nx"], cfg["nu"], cfg["N"], bounds=cfg.get("bounds", False), xi=xi,
                           nx_profile=cfg.get("nx_profile"))


def instance_xi(n_inst: int, first: int = 0) -> np.ndarray:
    return splitmix64_uniform(SEED, np.arange(first, first + n_inst, dtype=np.uint64), 4)
