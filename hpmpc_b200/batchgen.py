"""Synthetic instance batches in the native packed layout (harness code for tests / bench).

Instance i of a batch is problems.mass_spring_ocp(..., xi = instance_xi(i)): only x0 (-> b_0), the diagonal of Q and the
diagonal of R differ between instances, so a batch is the packed base block with those entries overwritten.  Every
instance still owns a full private copy of all its matrices (fully time-varying storage).
"""
from __future__ import annotations

import numpy as np

from . import problems
from .capi import BatchOcp


class BatchSpec:
    def __init__(self, cfg, device: int = 0):
        self.cfg = dict(problems.config(cfg)) if isinstance(cfg, str) else dict(cfg)
        self.base = problems.make(self.cfg)
        self.h = BatchOcp(self.base, device=device)
        p, h = self.base, self.h
        self.base_block = h.pack(p)
        tri = lambda i: i * (i + 1) // 2
        iq, ir = [], []
        for n in range(p.N + 1):
            nu, nx = p.nu[n], p.nx[n]
            o = h.off[n]["RSQ"]
            ir += [o + tri(i) + i for i in range(nu)]
            iq += [o + tri(nu + i) + nu + i for i in range(nx)]
        self.idx_Q, self.idx_R = np.asarray(iq, dtype=np.int64), np.asarray(ir, dtype=np.int64)
        # b_0 = A x0 + 0.1 when x0 is eliminated (nx[0] == 0): columns 0 and 1 of the full A
        self.x0_elim = p.nx[0] == 0
        if self.x0_elim:
            nfull = p.nx[1]
            A0, _ = problems.mass_spring_AB(nfull, p.nu[0])
            self.A_cols = A0[:, :2].copy()
            nux0 = p.nu[0] + p.nx[0]
            self.idx_b0 = h.off[0]["BAbt"] + nux0 * p.nx[1] + np.arange(p.nx[1], dtype=np.int64)
        self.A0, self.B0 = problems.mass_spring_AB(max(p.nx), p.nu[0])

    def scalars(self, n_inst: int, first: int = 0):
        xi = problems.instance_xi(n_inst, first)
        x01 = 2.5 * (1.0 + 0.2 * xi[:, 0]); x02 = 2.5 * (1.0 + 0.2 * xi[:, 1])
        return x01, x02, 1.0 + 0.1 * xi[:, 2], 2.0 + 0.2 * xi[:, 3]

    def numpy_batch(self, n_inst: int, first: int = 0) -> np.ndarray:
        x01, x02, qs, rs = self.scalars(n_inst, first)
        blk = np.repeat(self.base_block[None, :], n_inst, axis=0)
        blk[:, self.idx_Q] = qs[:, None]
        blk[:, self.idx_R] = rs[:, None]
        if self.x0_elim:
            blk[:, self.idx_b0] = x01[:, None] * self.A_cols[None, :, 0] + x02[:, None] * self.A_cols[None, :, 1] + 0.1
        else:
            raise NotImplementedError("free-x0 batches vary only Q and R")
        return blk

    def torch_batch(self, n_inst: int, first: int = 0, device="cuda"):
        """Same as numpy_batch but assembled on the device (6.4 GB at config 2 -- never staged through the host)."""
        import torch
        x01, x02, qs, rs = (torch.from_numpy(v).to(device) for v in self.scalars(n_inst, first))
        base = torch.from_numpy(self.base_block).to(device)
        blk = base.unsqueeze(0).repeat(n_inst, 1)
        blk[:, torch.from_numpy(self.idx_Q).to(device)] = qs[:, None]
        blk[:, torch.from_numpy(self.idx_R).to(device)] = rs[:, None]
        if self.x0_elim:
            Ac = torch.from_numpy(self.A_cols).to(device)
            blk[:, torch.from_numpy(self.idx_b0).to(device)] = x01[:, None] * Ac[None, :, 0] + x02[:, None] * Ac[None, :, 1] + 0.1
        return blk

    def problem(self, inst: int) -> problems.Ocp:
        xi = tuple(problems.instance_xi(1, first=inst)[0])
        return problems.make(self.cfg, xi=xi)
