"""Small driver for ncu / timing sweeps: runs the batched Riccati sv kernel a few times.
usage: python tools/prof_sv.py [n_inst] [ctas_per_sm] [warps] [reps] [cfg]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hpmpc_b200 import capi
from hpmpc_b200.batchgen import BatchSpec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
warps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
cfg = sys.argv[5] if len(sys.argv) > 5 else "cfg2"
L = capi.product()
spec = BatchSpec(cfg)
h = spec.h
if cps or warps:
    assert h.set_launch(cps, warps) == 0
d_in = spec.torch_batch(n)
ux = torch.zeros((n, h.sz.ux_stride), dtype=torch.float64, device="cuda")
pi = torch.zeros((n, h.sz.pi_stride), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
if os.environ.get("SV_PROBE"):          # the traffic-equivalent probe instead of the solver (same arguments)
    import ctypes as C
    L.hpmpc_b200_sv_traffic_probe.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 4
    _probe = L.hpmpc_b200_sv_traffic_probe
    L = type("P", (), {"hpmpc_b200_d_back_ric_rec_sv_batch": staticmethod(lambda h_, n_, a, b, c, _none, s_: _probe(h_, n_, a, b, c, s_))})
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
for i in range(2):
    L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), None, st)
torch.cuda.synchronize()
ev[0].record()
for i in range(reps):
    assert L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), None, st) == 0
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
best = min(ms)
print(f"n={n} grid={h.sz.grid} warps={h.sz.warps_per_cta} smem={h.sz.smem_per_cta} fast={h.sz.fast_variant} "
      f"ms={best:.3f} solves/s={n / best * 1e3:.3e} frac_hbm={n * 97968 / (best * 1e-3) / 6454.6e9:.3f}")
