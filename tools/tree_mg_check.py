"""Scenario-tree Riccati on several GPUs through the C library's own NCCL exchange (hpmpc_b200_d_tree_back_ric_rec_sv_batch_mg):
every rank solves its subtrees, results are compared bit for bit with the single-GPU solve of the same trees, and timed.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/tree_mg_check.py [n_trees]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from hpmpc_b200 import tree as T
from hpmpc_b200.problems import instance_xi

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("gloo")                       # only to hand the NCCL unique id around; the data path is the library's own
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
t0 = T.mass_spring_tree(12, 5, 4, 3, 20)
h = T.TreeBatch(t0, device=local)
L = h.L
L.hpmpc_b200_comm_unique_id.argtypes = [C.c_void_p, C.c_int]
L.hpmpc_b200_comm_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_int]
L.hpmpc_b200_comm_destroy.argtypes = [C.c_void_p]
L.hpmpc_b200_d_tree_back_ric_rec_sv_batch_mg.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong] + [C.c_void_p] * 5
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    buf = (C.c_char * 128)()
    assert L.hpmpc_b200_comm_unique_id(buf, 128) == 0
    uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
dist.broadcast(uid, 0)
comm = C.c_void_p()
idb = (C.c_char * 128).from_buffer_copy(uid.numpy().tobytes())
assert L.hpmpc_b200_comm_create(C.byref(comm), world, rank, idb, local) == 0
base = torch.from_numpy(h.pack(t0)).to(dev)
mask = torch.zeros_like(base)
for nd in range(h.sz.Nn):
    nux = t0.nu[nd] + t0.nx[nd]
    mask[h.off[nd]["RSQ"]:h.off[nd]["RSQ"] + nux * (nux + 1) // 2 + nux] = 1.0
xi = torch.from_numpy(instance_xi(n)[:, 2].copy()).to(dev)
d_in = base[None, :] * (1.0 + 0.1 * xi[:, None] * mask[None, :])
z = lambda m: torch.zeros((n, m), dtype=torch.float64, device=dev)
ux, pi, Lst = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.L_stride)
ux1, pi1, Lst1 = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.L_stride)
st = torch.cuda.current_stream().cuda_stream
assert L.hpmpc_b200_d_tree_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux1.data_ptr(), pi1.data_ptr(), Lst1.data_ptr(), st) == 0
run = lambda: L.hpmpc_b200_d_tree_back_ric_rec_sv_batch_mg(h.h, comm, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), Lst.data_ptr(), st)
assert run() == 0
torch.cuda.synchronize()
# nodes this rank owns: the levels above the subtree roots (everybody), its subtrees' roots and their tails
ns = h.sz.n_shard_nodes
lo, hi = ns * rank // world, ns * (rank + 1) // world
own = [nd for nd in range(h.sz.Nn) if t0.topo["stage"][nd] < h.sz.cut_stage - 1]
def subtree(nd):
    out = [nd]
    for k in range(t0.topo["first_kid"][nd], t0.topo["first_kid"][nd] + t0.topo["nkids"][nd]):
        out += subtree(k)
    return out
for k in range(lo, hi):
    own += subtree(h.subtrees[k]["node"])
bad = 0
for nd in own:
    o, nux = h.off[nd]["ux"], t0.nu[nd] + t0.nx[nd]
    bad += int(not torch.equal(ux[:, o:o + nux], ux1[:, o:o + nux]))
    if nd > 0:
        o = h.off[nd]["pi"]
        bad += int(not torch.equal(pi[:, o:o + t0.nx[nd]], pi1[:, o:o + t0.nx[nd]]))
dist.barrier()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
reps = 5
ev[0].record()
for _ in range(reps):
    run()
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / reps
tm = torch.tensor([ms, float(bad)], dtype=torch.float64)
dist.all_reduce(tm, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"tree_mg: {world} rank(s), {n} trees, {len(own)} of {h.sz.Nn} nodes owned by rank 0, mismatching node vectors (max over ranks): {int(tm[1])}, "
          f"{tm[0]:.3f} ms per solve -> {n / tm[0] * 1e3:.0f} trees/s")
assert bad == 0
L.hpmpc_b200_comm_destroy(comm)
dist.destroy_process_group()
