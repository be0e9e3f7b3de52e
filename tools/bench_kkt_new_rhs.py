"""Throughput of the re-solve with a new right-hand side at BASELINE config 3 shapes, one GPU.

    python tools/bench_kkt_new_rhs.py [--n-inst 4096] [--steps 5] [--cfg cfg3]

Times, with CUDA events on the launching stream after warm-up: the IPM alone, the IPM leaving its KKT state, and the re-solve
(hpmpc_b200_d_kkt_solve_new_rhs_batch) on that state.  Prints one JSON line; the roofline entry uses the byte model of DESIGN.md 3.5
(factor + [B A] + Hessian read once, six vector passes) against the measured HBM peak of MEASURED_PEAKS.json."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-inst", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cfg", default="cfg3")
    print(json.dumps(measure(ap.parse_args())))


def measure(a):
    """a: namespace with cfg, n_inst, steps, warmup; returns the JSON record (bench.py adds it to its `extra`)"""
    import torch
    from hpmpc_b200 import capi
    from hpmpc_b200.batchgen import BatchSpec
    L = capi.product()
    L.hpmpc_b200_kkt_state_stride.restype = C.c_longlong
    L.hpmpc_b200_kkt_state_stride.argtypes = [C.c_void_p]
    L.hpmpc_b200_d_ip2_res_mpc_hard_kkt_batch.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                                         C.c_double, C.c_int] + [C.c_void_p] * 7
    L.hpmpc_b200_d_kkt_solve_new_rhs_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 8
    spec = BatchSpec(a.cfg)
    h, n, k_max = spec.h, a.n_inst, 40
    ks = L.hpmpc_b200_kkt_state_stride(h.h)
    d_in = spec.torch_batch(n)
    z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
    ux, pi, lam, t, info, info2 = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * k_max), z(6)
    kkt = torch.empty((n, ks), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ipm = lambda: L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, d_in.data_ptr(), k_max, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(), pi.data_ptr(),
                                                        lam.data_ptr(), t.data_ptr(), info.data_ptr(), st)
    ipm_kkt = lambda: L.hpmpc_b200_d_ip2_res_mpc_hard_kkt_batch(h.h, n, d_in.data_ptr(), k_max, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(),
                                                                pi.data_ptr(), lam.data_ptr(), t.data_ptr(), info.data_ptr(), kkt.data_ptr(), st)
    resolve = lambda: L.hpmpc_b200_d_kkt_solve_new_rhs_batch(h.h, n, d_in.data_ptr(), kkt.data_ptr(), ux.data_ptr(), pi.data_ptr(),
                                                             lam.data_ptr(), t.data_ptr(), info2.data_ptr(), st)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            assert fn() == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            assert fn() == 0
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    ms_ipm = timed(ipm, max(1, a.steps // 2), 1)
    ms_ipm_kkt = timed(ipm_kkt, max(1, a.steps // 2), 1)
    assert float(info[:, 1].abs().max()) == 0.0
    ms_re = timed(resolve, a.steps, a.warmup)
    assert float(info2[:, 1].abs().max()) == 0.0
    p = spec.base
    nux = [p.nu[i] + p.nx[i] for i in range(p.N + 1)]
    D_BAbt = sum((nux[i] + 1) * p.nx[i + 1] for i in range(p.N))
    D_RSQ = sum(nux[i] * (nux[i] + 1) // 2 + nux[i] for i in range(p.N + 1))
    D_vec = sum(nux) + sum(p.nx[1:]) + 6 * sum(p.nb)
    bytes_re = 8 * (D_RSQ + D_BAbt + D_RSQ + 6 * D_vec)          # factor (same count as the Hessian) + [B A b] + Hessian + vectors
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6454.6)) if isinstance(peaks, dict) else 6454.6
    ach = bytes_re * n / (ms_re * 1e-3) / 1e9
    rec = {
        "metric": "kkt_new_rhs_resolves_per_s", "value": n / (ms_re * 1e-3), "unit": "solves/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_re, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"re-solve with a new right-hand side on the IPM's last factor, {a.cfg}, {n} instances", "kkt_state_bytes_per_instance": 8 * ks,
                   "l2": f"state + inputs = {(8 * ks + 8 * h.sz.in_stride) * n / 1e9:.2f} GB, larger than L2"},
        "ipm_ms": ms_ipm, "ipm_with_kkt_state_ms": ms_ipm_kkt, "ipm_solves_per_s": n / (ms_ipm * 1e-3),
        "speedup_vs_full_ipm_solve": ms_ipm / ms_re,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "kernel": "hb_kkt_new_rhs_kernel", "algorithmic_bytes_per_solve": bytes_re}}
    h.close()
    return rec


if __name__ == "__main__":
    main()
