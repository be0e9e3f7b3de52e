"""BASELINE config 4 (time-varying variable-nx OCP: nx 40 -> 4, nu=8, N=20, box bounds, 8 192 instances) on the any-size kernels
(four-warps-per-instance sweeps of hpmpc_b200/csrc/ric_team.cuh): Riccati sv and box IPM, kernels timed with CUDA events on data
resident in HBM.  usage: python tools/bench_cfg4.py [n_inst] [reps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(n=8192, reps=3):
    import torch
    from hpmpc_b200 import capi, problems
    L = capi.product()
    st = torch.cuda.current_stream().cuda_stream

    def timeit(fn, reps):
        fn(); torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            fn(); ev[i + 1].record()
        torch.cuda.synchronize()
        return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))

    z = lambda n, m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device="cuda")
    p0 = problems.make("cfg4"); h = capi.BatchOcp(p0, device=0)
    base = torch.from_numpy(h.pack(p0)).cuda()
    d_in = base[None, :].repeat(n, 1)
    # distinct instances: the gradient rows (q, r) scaled per instance (keeps every problem well posed)
    xi = torch.from_numpy(problems.instance_xi(n)[:, 2].copy()).cuda()
    for s in range(p0.N + 1):
        nux = p0.nx[s] + p0.nu[s]
        o = h.off[s]["RSQ"] + nux * (nux + 1) // 2
        d_in[:, o:o + nux] *= (1.0 + 0.3 * xi[:, None])
    k_max = 40
    ux, pi, lam, t, info = z(n, h.sz.ux_stride), z(n, h.sz.pi_stride), z(n, h.sz.lam_stride), z(n, h.sz.lam_stride), z(n, 6 + 5 * k_max)
    out = {"n_inst": n, "in_bytes_per_instance": int(h.sz.in_stride * 8)}
    ms = timeit(lambda: L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), None, st), reps)
    out["sv_ms"] = ms; out["sv_solves_per_s"] = n / ms * 1e3
    ms = timeit(lambda: L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, d_in.data_ptr(), k_max, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(), pi.data_ptr(),
                                                              lam.data_ptr(), t.data_ptr(), info.data_ptr(), st), reps)
    out["ipm_ms"] = ms; out["ipm_solves_per_s"] = n / ms * 1e3
    out["mean_kk"] = float(info[:, 0].mean()); out["converged"] = int((info[:, 1] == 0).sum())
    out["ux_checksum"] = float(ux.double().abs().sum())
    out["problem"] = p0
    h.close()
    return out


if __name__ == "__main__":
    r = measure(int(sys.argv[1]) if len(sys.argv) > 1 else 8192, int(sys.argv[2]) if len(sys.argv) > 2 else 3)
    r.pop("problem")
    print(json.dumps(r))
