"""Time the generic building blocks on BASELINE config 3 (nx=24, nu=11, N=50): sv, trf, trs and the whole IPM.
usage: python tools/prof_ipm_parts.py [n_inst]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hpmpc_b200 import capi
from hpmpc_b200.batchgen import BatchSpec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
os.environ.setdefault("HPMPC_B200_NO_FAST", "1") if len(sys.argv) > 2 and sys.argv[2] == "nofast" else None
L = capi.product()
spec = BatchSpec("cfg3")
h = spec.h
d_in = spec.torch_batch(n)
z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
ux, pi, Lf, lam, t, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.L_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * 40)
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))


Pb = z(h.sz.pi_stride)
t_sv = timeit(lambda: L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), Pb.data_ptr(), st))
t_svf = timeit(lambda: L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), None, st))
t_trf = timeit(lambda: L.hpmpc_b200_d_back_ric_rec_trf_batch(h.h, n, d_in.data_ptr(), Lf.data_ptr(), st))
t_trs = timeit(lambda: L.hpmpc_b200_d_back_ric_rec_trs_batch(h.h, n, d_in.data_ptr(), Lf.data_ptr(), ux.data_ptr(), pi.data_ptr(), st))
t_ipm = timeit(lambda: L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, d_in.data_ptr(), 40, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(), pi.data_ptr(),
                                                             lam.data_ptr(), t.data_ptr(), info.data_ptr(), st), reps=2)
kk = float(info[:, 0].mean())
print(f"cfg3 n={n} fast_variant={h.sz.fast_variant}: sv(generic, with Pb) {t_sv:.1f} ms | sv(default path) {t_svf:.1f} ms | trf {t_trf:.1f} ms | trs {t_trs:.1f} ms | "
      f"ipm {t_ipm:.1f} ms (mean kk {kk:.2f}, {t_ipm / kk:.1f} ms per iteration, {n / t_ipm * 1e3:.0f} solves/s)")
