"""Scenario-tree Riccati (BASELINE config 5; SURVEY.md section 8 rows a9, a10, e).

CPU: the tree oracle (oracle/ric_oracle.c: orc_tree_ric_sv) against golden vectors made by the real reference on the
stacked chain problem, against the stacked chain solved by the chain oracle, topology / layout / sharding host logic
(world_size-2 gloo).  GPU: the batched tree kernels through the C ABI against the oracle, and the phase-split path
(what two GPUs would each compute, with the tail-root factor blocks exchanged in between) against the one-GPU path."""
import os
import sys

import numpy as np
import pytest

from conftest import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from hpmpc_b200 import problems, tree as T  # noqa: E402
from oracle import api as oracle  # noqa: E402
import make_golden_tree as G  # noqa: E402

TOL = 1e-9
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "golden_tree_v1.npz"))


def cat(v):
    return np.concatenate([np.asarray(a).ravel() for a in v]) if len(v) else np.zeros(0)


@pytest.mark.parametrize("case", list(G.CASES))
def test_tree_oracle_matches_reference_golden(case):
    t = G.build(case)
    o = oracle.tree_ric(t)
    for f in ("u", "x", "pi"):
        assert rel_err([cat(o[f])], [GOLD[f"{case}/{f}"]]) < 1e-11, (case, f)


def test_tree_oracle_matches_stacked_chain_oracle_and_node_count():
    for md, Nr, Nh in ((1, 0, 5), (2, 2, 5), (3, 2, 4), (4, 3, 20)):
        topo = T.setup_tree(md, Nr, Nh)
        assert topo["Nn"] == T.number_of_nodes(md, Nr, Nh)
        for n in range(1, topo["Nn"]):
            d = topo["dad"][n]
            assert topo["first_kid"][d] <= n < topo["first_kid"][d] + topo["nkids"][d] and topo["stage"][n] == topo["stage"][d] + 1
    assert T.number_of_nodes(4, 3, 20) == 1173
    t = T.mass_spring_tree(6, 2, 2, 2, 4, xi=(0.1, 0.2, -0.3, 0.4))
    p, maps = T.stacked_chain(t)
    u, x, pi = T.unstack(t, maps, oracle.ric(p, "sv"))
    o = oracle.tree_ric(t)
    assert rel_err(o["u"], u) < 1e-12 and rel_err(o["x"], x) < 1e-12 and rel_err(o["pi"], pi) < 1e-11


def test_tree_handle_layout_host_only():
    """device = -1: topology analysis, offsets and packing work without a GPU; compute entry points refuse to run."""
    t = T.mass_spring_tree(12, 5, 4, 3, 20)
    h = T.TreeBatch(t, device=-1)
    assert (h.sz.Nn, h.sz.n_tails, h.sz.n_top_nodes, h.sz.cut_stage) == (1173, 64, 21, 3)
    assert [tl["node"] for tl in h.tails] == list(range(21, 85))
    assert h.sz.n_shard_nodes == 16 and [st["node"] for st in h.subtrees] == list(range(5, 21))
    assert [(st["tail_lo"], st["tail_hi"]) for st in h.subtrees] == [(4 * k, 4 * k + 4) for k in range(16)]
    blk = h.pack(t)
    n = 100
    d = t.topo["dad"][n]
    M = blk[h.off[n]["BAbt"]:h.off[n]["BAbt"] + (t.nu[d] + t.nx[d] + 1) * t.nx[n]].reshape(-1, t.nx[n])
    np.testing.assert_array_equal(M[:t.nu[d]], t.B[n].T)
    np.testing.assert_array_equal(M[t.nu[d]:t.nu[d] + t.nx[d]], t.A[n].T)
    np.testing.assert_array_equal(M[-1], t.b[n])
    rc = h.L.hpmpc_b200_d_tree_back_ric_rec_sv_batch(h.h, 1, None, None, None, None, None)
    assert rc != 0
    h.close()


def tail_range(n_tails, rank, world):
    return n_tails * rank // world, n_tails * (rank + 1) // world


def _worker(rank, world, port, q):
    """The exchange of the multi-GPU tree path on CPU tensors: every rank owns a contiguous range of tails, fills the factor
    blocks of its tail roots, and all ranks end up with all blocks after one all_gather."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_trees, n_tails, blk = 3, 64, 204
    lo, hi = tail_range(n_tails, rank, world)
    mine = torch.stack([torch.full((hi - lo, blk), float(t * 1000 + rank)) + torch.arange(lo, hi)[:, None] for t in range(n_trees)])
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    full = torch.cat(parts, dim=1)
    ok = full.shape == (n_trees, n_tails, blk) and all(float(full[t, j, 0]) == t * 1000 + (j * world // n_tails) + j for t in range(n_trees) for j in range(n_tails))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_tail_block_exchange_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out == [(0, True), (1, True)]
    for world in (1, 2, 4, 8):
        r = [tail_range(64, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == 64 and all(r[i][1] == r[i + 1][0] for i in range(world - 1))


# ---------------------------------------------------------------------------------------------------- GPU
def _solve(h, blocks, phases=None):
    import torch
    n = blocks.shape[0]
    d_in = torch.from_numpy(blocks).cuda()
    ux = torch.zeros((n, h.sz.ux_stride), dtype=torch.float64, device="cuda")
    pi = torch.zeros((n, h.sz.pi_stride), dtype=torch.float64, device="cuda")
    Lst = torch.zeros((n, h.sz.L_stride), dtype=torch.float64, device="cuda")
    rc = h.L.hpmpc_b200_d_tree_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), Lst.data_ptr(), None)
    assert rc == 0
    torch.cuda.synchronize()
    return d_in, ux, pi, Lst


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 2, 1, 0, 6), (4, 2, 2, 2, 5), (6, 2, 3, 2, 5), (8, 3, 2, 1, 4), (12, 5, 4, 3, 20)])
def test_tree_gpu_vs_oracle(shape):
    nx, nu, md, Nr, Nh = shape
    xis = problems.instance_xi(5, first=70)
    trees = [T.mass_spring_tree(nx, nu, md, Nr, Nh, xi=tuple(x)) for x in xis]
    h = T.TreeBatch(trees[0], device=0)
    _, ux, pi, _ = _solve(h, np.stack([h.pack(t) for t in trees]))
    uxh, pih = ux.cpu().numpy(), pi.cpu().numpy()
    for i, t in enumerate(trees):
        o = oracle.tree_ric(t)
        u, x, p = h.split(uxh[i], pih[i])
        assert rel_err(u, o["u"]) < TOL and rel_err(x, o["x"]) < TOL and rel_err(p, o["pi"]) < TOL, (shape, i)
    h.close()


@pytest.mark.gpu
def test_tree_gpu_golden_and_phase_split_equals_single_pass():
    """Golden vectors (reference on the stacked chain) through the C ABI, then config 5 solved the way two GPUs would: each
    'rank' runs phase 0 on its half of the tails into its own stash, the tail-root blocks are exchanged, both run phase 1 and
    phase 2 on their tails; the union must equal the single-pass result bit for bit."""
    import torch
    for case in G.CASES:
        t = G.build(case)
        h = T.TreeBatch(t, device=0)
        _, ux, pi, _ = _solve(h, h.pack(t)[None, :])
        u, x, p = h.split(ux[0].cpu().numpy(), pi[0].cpu().numpy())
        for f, v in (("u", u), ("x", x), ("pi", p)):
            assert rel_err([cat(v)], [GOLD[f"{case}/{f}"]]) < TOL, (case, f)
        h.close()
    xis = problems.instance_xi(6, first=200)
    trees = [T.mass_spring_tree(12, 5, 4, 3, 20, xi=tuple(x)) for x in xis]
    h = T.TreeBatch(trees[0], device=0)
    blocks = np.stack([h.pack(t) for t in trees])
    d_in, ux0, pi0, _ = _solve(h, blocks)
    n, world = len(trees), 2
    st = [dict(ux=torch.zeros_like(ux0), pi=torch.zeros_like(pi0), L=torch.zeros((n, h.sz.L_stride), dtype=torch.float64, device="cuda")) for _ in range(world)]
    ph = h.L.hpmpc_b200_d_tree_back_ric_rec_sv_phase
    rng = [tail_range(h.sz.n_tails, r, world) for r in range(world)]
    for r in range(world):
        assert ph(h.h, n, 0, rng[r][0], rng[r][1], d_in.data_ptr(), st[r]["ux"].data_ptr(), st[r]["pi"].data_ptr(), st[r]["L"].data_ptr(), None) == 0
    torch.cuda.synchronize()
    for r in range(world):                       # the exchange: every rank receives the blocks of the tails it does not own
        for s in range(world):
            if s != r:
                for j in range(*rng[s]):
                    o, ln = h.tails[j]["off_L"], h.tails[j]["len_L"]
                    st[r]["L"][:, o:o + ln] = st[s]["L"][:, o:o + ln]
    for r in range(world):
        assert ph(h.h, n, 1, 0, 0, d_in.data_ptr(), st[r]["ux"].data_ptr(), st[r]["pi"].data_ptr(), st[r]["L"].data_ptr(), None) == 0
        assert ph(h.h, n, 2, rng[r][0], rng[r][1], d_in.data_ptr(), st[r]["ux"].data_ptr(), st[r]["pi"].data_ptr(), st[r]["L"].data_ptr(), None) == 0
    torch.cuda.synchronize()
    topo = trees[0].topo
    owner = {}
    for j, tl in enumerate(h.tails):
        m = tl["node"]
        while True:
            owner[m] = next(r for r in range(world) if rng[r][0] <= j < rng[r][1])
            if topo["nkids"][m] == 0:
                break
            m = topo["first_kid"][m]
    for node in range(topo["Nn"]):
        for r in ([owner[node]] if node in owner else range(world)):
            a, ln = h.off[node]["ux"], trees[0].nu[node] + trees[0].nx[node]
            assert torch.equal(st[r]["ux"][:, a:a + ln], ux0[:, a:a + ln]), node
            a, ln = h.off[node]["pi"], trees[0].nx[node]
            assert torch.equal(st[r]["pi"][:, a:a + ln], pi0[:, a:a + ln]), node
    h.close()


@pytest.mark.gpu
def test_tree_gpu_subtree_sharding_equals_single_pass():
    """Config 5 solved the way four GPUs would with SUBTREE sharding: every 'rank' owns four of the 16 depth-2 subtrees (their
    tails and their roots), the subtree-root factor blocks are exchanged, the five nodes above are solved redundantly, and the
    union of the results must equal the single-pass result bit for bit."""
    import torch
    xis = problems.instance_xi(5, first=300)
    trees = [T.mass_spring_tree(12, 5, 4, 3, 20, xi=tuple(x)) for x in xis]
    h = T.TreeBatch(trees[0], device=0)
    blocks = np.stack([h.pack(t) for t in trees])
    d_in, ux0, pi0, _ = _solve(h, blocks)
    n, world = len(trees), 4
    st = [dict(ux=torch.zeros_like(ux0), pi=torch.zeros_like(pi0), L=torch.zeros((n, h.sz.L_stride), dtype=torch.float64, device="cuda")) for _ in range(world)]
    ph = h.L.hpmpc_b200_d_tree_back_ric_rec_sv_phase
    rng = [tail_range(h.sz.n_shard_nodes, r, world) for r in range(world)]
    trng = [(h.subtrees[a]["tail_lo"], h.subtrees[b - 1]["tail_hi"]) for a, b in rng]
    args = lambda r: (d_in.data_ptr(), st[r]["ux"].data_ptr(), st[r]["pi"].data_ptr(), st[r]["L"].data_ptr(), None)
    for r in range(world):
        assert ph(h.h, n, 0, trng[r][0], trng[r][1], *args(r)) == 0
        assert ph(h.h, n, 3, rng[r][0], rng[r][1], *args(r)) == 0
    torch.cuda.synchronize()
    for r in range(world):                       # the exchange: factor blocks of the subtree roots owned by the others
        for s in range(world):
            if s != r:
                for k in range(*rng[s]):
                    o, ln = h.subtrees[k]["off_L"], h.subtrees[k]["len_L"]
                    st[r]["L"][:, o:o + ln] = st[s]["L"][:, o:o + ln]
    for r in range(world):
        assert ph(h.h, n, 4, 0, 0, *args(r)) == 0
        assert ph(h.h, n, 5, rng[r][0], rng[r][1], *args(r)) == 0
        assert ph(h.h, n, 2, trng[r][0], trng[r][1], *args(r)) == 0
    torch.cuda.synchronize()
    topo = trees[0].topo
    owner = {}
    for r in range(world):
        for k in range(*rng[r]):
            stack = [h.subtrees[k]["node"]]
            while stack:
                m = stack.pop()
                owner[m] = r
                stack += [topo["first_kid"][m] + i for i in range(topo["nkids"][m])]
    assert len(owner) == topo["Nn"] - 5
    for node in range(topo["Nn"]):
        for r in ([owner[node]] if node in owner else range(world)):
            a, ln = h.off[node]["ux"], trees[0].nu[node] + trees[0].nx[node]
            assert torch.equal(st[r]["ux"][:, a:a + ln], ux0[:, a:a + ln]), node
            a, ln = h.off[node]["pi"], trees[0].nx[node]
            assert torch.equal(st[r]["pi"][:, a:a + ln], pi0[:, a:a + ln]), node
    h.close()


def test_tree_oracle_trf_trs_equals_sv():
    for shape in [(4, 2, 2, 2, 5), (6, 2, 3, 2, 5), (12, 5, 4, 2, 6)]:
        t = T.mass_spring_tree(*shape, xi=(0.1, -0.3, 0.2, 0.5))
        a, b = oracle.tree_ric(t), oracle.tree_ric_trf_trs(t)
        assert max(np.max(np.abs(cat(a[f]) - cat(b[f]))) for f in ("u", "x", "pi")) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 2, 2, 2, 5), (6, 2, 3, 2, 5), (12, 5, 4, 3, 20)])
def test_tree_gpu_trf_then_trs_with_new_right_hand_sides(shape):
    """d_tree_back_ric_rec_trf / _trs: factorize once, then solve for the original and for new b, q, r (same matrices)"""
    import ctypes as C
    import copy
    import torch
    t = T.mass_spring_tree(*shape, xi=(0.2, 0.1, -0.4, 0.3))
    t2 = copy.deepcopy(t)
    rng = np.random.default_rng(5)
    for n in range(t.topo["Nn"]):
        t2.q[n] = t.q[n] + 0.3 * rng.standard_normal(t.nx[n]); t2.r[n] = t.r[n] - 0.2 * rng.standard_normal(t.nu[n])
        if n > 0:
            t2.b[n] = t.b[n] + 0.1 * rng.standard_normal(t.nx[n])
    tb = T.TreeBatch(t)
    try:
        L = tb.L
        L.hpmpc_b200_d_tree_back_ric_rec_trf_batch.restype = C.c_int
        L.hpmpc_b200_d_tree_back_ric_rec_trf_batch.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hpmpc_b200_d_tree_back_ric_rec_trs_batch.restype = C.c_int
        L.hpmpc_b200_d_tree_back_ric_rec_trs_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 5
        d1 = torch.from_numpy(np.stack([tb.pack(t), tb.pack(t)])).cuda()
        d2 = torch.from_numpy(np.stack([tb.pack(t2), tb.pack(t)])).cuda()          # tree 0: new right-hand sides, tree 1: unchanged
        z = lambda m: torch.zeros((2, int(m)), dtype=torch.float64, device="cuda")
        Lst, ux, pi = z(tb.sz.L_stride), z(tb.sz.ux_stride), z(tb.sz.pi_stride)
        assert L.hpmpc_b200_d_tree_back_ric_rec_trf_batch(tb.h, 2, d1.data_ptr(), Lst.data_ptr(), None) == 0
        assert L.hpmpc_b200_d_tree_back_ric_rec_trs_batch(tb.h, 2, d2.data_ptr(), Lst.data_ptr(), ux.data_ptr(), pi.data_ptr(), None) == 0
        torch.cuda.synchronize()
        for i, prob in enumerate((t2, t)):
            u, x, p = tb.split(ux[i].cpu().numpy(), pi[i].cpu().numpy())
            r = oracle.tree_ric(prob)
            for f, v in (("u", u), ("x", x), ("pi", p)):
                a, b = cat(v), cat(r[f])
                assert np.max(np.abs(a - b)) <= 1e-9 * max(1.0, np.max(np.abs(b))), (i, f)
    finally:
        tb.close()


@pytest.mark.gpu
def test_tree_multi_gpu_exchange_inside_the_library():
    """hpmpc_b200_d_tree_back_ric_rec_sv_batch_mg: subtrees sharded over 2 ranks, the NCCL all-gather issued by the C library;
    every rank's nodes bit-identical to the single-GPU solve (tools/tree_mg_check.py under torchrun).  Needs 2 GPUs."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "tree_mg_check.py"), "256"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mismatching node vectors (max over ranks): 0" in r.stdout
