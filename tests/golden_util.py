"""Shared access to tests/golden/golden_v1.npz (vectors produced by the real reference, see make_golden.py)."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)

GOLD = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
CASES = make_golden.CASES


def case_ids(kind=None):
    return [(name, inst) for name, (k, spec, insts) in CASES.items() for inst in insts if kind is None or k == kind]


def problem(name, inst):
    kind, spec, _ = CASES[name]
    p = make_golden.build_problem(spec, inst)
    assert abs(make_golden.checksum(p) - float(GOLD[f"{name}/{inst}/checksum"])) < 1e-9, "input generator drifted"
    return kind, p


def gold(name, inst, field):
    return GOLD[f"{name}/{inst}/{field}"]


def cat(L):
    return make_golden.cat(L)
