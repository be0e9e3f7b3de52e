#!/usr/bin/env python
"""bench.py -- throughput of the batched Riccati factor+solve (BASELINE.json config 2) and of the fused box-IPM
(config 3) through the C ABI of libhpmpc_b200.so.  Contract: see the task description; one JSON line on stdout.

  python bench.py --gpus 1 --steps 10 --warmup 3                       (our arm, 1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (weak scaling, no collectives)
  python bench.py --impl reference --gpus N --steps K --warmup W       (the reference's own CPU code on the host cores)

A "step" is one pass of the hot path over one batch: 65 536 independent LQCP instances (nx=12, nu=5, N=30) per GPU,
every instance with its own private matrices (6.4 GB of input per GPU, far larger than the 126 MB L2, so no flush
is needed between steps).  `value` has the inputs resident in HBM; `e2e` runs the host-buffer entry point with pinned
host inputs and includes the H2D / D2H copies.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


# ------------------------------------------------------------------------------------------- host logic (CPU-testable)
def shard_range(n: int, rank: int, world: int):
    """Contiguous slice [a, b) of n units owned by `rank` (no data-path collective: instances are independent)."""
    return n * rank // world, n * (rank + 1) // world


def reduce_max_time(t: float, device) -> float:
    """max over ranks (torch.distributed when initialised, identity otherwise)."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        v = torch.tensor([t], dtype=torch.float64, device=device)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())
    return t


def algorithmic_work(p):
    """(flops, bytes) per Riccati factor+solve and per IPM iteration -- SURVEY.md section 8(d) formulas."""
    N = p.N
    nux = [p.nx[n] + p.nu[n] for n in range(N + 1)]
    D_BAbt = sum((nux[n] + 1) * p.nx[n + 1] for n in range(N))
    D_RSQ = sum(nux[n] * (nux[n] + 1) // 2 + nux[n] for n in range(N + 1))
    D_out = sum(nux) + sum(p.nx[1:])
    B_sv = 8 * (D_BAbt + D_RSQ + D_out)
    F_trf = nux[N] ** 3 / 3 + sum(nux[n] * p.nx[n + 1] ** 2 + nux[n] ** 2 * p.nx[n + 1] + nux[n] ** 3 / 3 for n in range(N))
    F_trs = sum(4 * p.nx[n + 1] ** 2 + 4 * nux[n] * p.nx[n + 1] + 2 * p.nu[n] ** 2 + 4 * p.nx[n] * p.nu[n] for n in range(N))
    nx, nu = p.nx[-1], p.nu[0]
    uniform = all(v == nx for v in p.nx[1:]) and all(v == nu for v in p.nu[:N])
    if uniform:
        # the reference's own formula (test_problems/test_d_ric_libstr.c:486) + 2 N nx^2 for pi (test_d_ric_mpc.c:579)
        F_sv = (nx ** 3 / 3 + 1.5 * nx ** 2) + N * (7 / 3 * nx ** 3 + 4 * nx ** 2 * nu + 2 * nx * nu ** 2 + nu ** 3 / 3 + 6.5 * nx ** 2
                                                    + 9 * nx * nu + 2.5 * nu ** 2) - (nx * (nx + nu) + nx ** 3 / 3 + 1.5 * nx ** 2) + 2 * N * nx ** 2
        if p.nx[0] != 0:
            F_sv = F_trf + F_trs
    else:
        F_sv = F_trf + F_trs
    D_vec = sum(nux) + sum(p.nx[1:]) + 6 * sum(p.nb)
    B_it = 8 * (D_BAbt + D_RSQ + 2 * D_RSQ + D_BAbt + 4 * D_vec)
    F_res = sum(2 * nux[n] ** 2 for n in range(N + 1)) + 4 * sum(nux[n] * p.nx[n + 1] for n in range(N))
    F_it = F_trf + 2 * F_trs + F_res
    return dict(F_sv=float(F_sv), B_sv=float(B_sv), F_it=float(F_it), B_it=float(B_it))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


# ------------------------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    """The reference's own CPU implementation (oracle/_ref, built from /root/reference by oracle/Makefile) on all host
    cores.  Riccati: d_back_ric_rec_sv_tv_res on pre-packed panel-major data, exactly like the reference's timing program
    (test_problems/test_d_ric_mpc.c); one private workspace per thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hpmpc_b200.batchgen import BatchSpec
    from oracle import api as oracle
    cores = os.cpu_count()
    if args.workload in ("tree", "tree_ipm"):
        # the reference's tree path needs BLASFEO (absent from /root/reference and from this image): the CPU arm is the oracle's
        # plain-C port, one tree per thread on all host cores (ctypes releases the GIL during the call)
        from concurrent.futures import ThreadPoolExecutor
        from hpmpc_b200 import tree as T
        from hpmpc_b200.problems import instance_xi
        ipm = args.workload == "tree_ipm"
        n_sample = args.ref_sample or 2 * cores
        xis = instance_xi(n_sample)
        trees = [T.mass_spring_tree(12, 5, 4, 3, 20, xi=tuple(xis[i]), bounds=ipm) for i in range(n_sample)]
        packed = [oracle.TreeTimed(t, ipm) for t in trees]        # packing is outside the timed region, like the GPU arm's

        def run_once():
            t0 = time.perf_counter()
            with ThreadPoolExecutor(cores) as ex:
                list(ex.map(lambda p: p.run(), packed))
            return time.perf_counter() - t0
        for _ in range(max(min(args.warmup, 1), 1)):
            run_once()
        steps = min(args.steps, 3)
        tot = sum(run_once() for _ in range(steps))
        value = n_sample * steps / tot
        sample = f"{n_sample} distinct trees per step, oracle/ric_oracle.c port ({'orc_tree_ip2_res_mpc_hard' if ipm else 'orc_tree_ric_sv'}), {cores} threads"
        print(json.dumps({"metric": "tree_box_ipm_solves_per_s" if ipm else "tree_riccati_solves_per_s", "value": value, "unit": "trees/s",
                          "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": 1e3 * tot / steps, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
                          "config": {"workload": "scenario tree md=4 Nr=3 Nh=20 nx=12 nu=5 (1173 nodes)" + (", box IPM, tol 1e-8" if ipm else ", Riccati factor+solve"),
                                     "note": "CPU port (the reference's tree files need BLASFEO), bounded sample"},
                          "cpu_baseline": {"value": value, "unit": "trees/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": value, "unit": "trees/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    cfg = "cfg2" if args.workload == "ric" else "cfg3"
    spec = BatchSpec(cfg, device=-1)
    kind = "avx2" if os.path.exists(oracle.REF_AVX2) else "c99"
    if not os.path.exists(oracle.REF_C99):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libhpmpc_ref_*.so missing (reference not compiled)"}))
        return
    n_sample = args.ref_sample or (max(2048, 16 * cores) if args.workload == "ric" else max(256, 2 * cores))
    rs = oracle.RefSample(spec, n_sample)
    run = (lambda n_pass: rs.time_ric_sv(kind, cores, n_pass)[0]) if args.workload == "ric" else (lambda n_pass: rs.time_ipm(kind, cores, n_pass)[0])
    t1 = run(1)
    n_pass = max(1, int(round(args.ref_step_seconds / max(t1, 1e-6))))
    for _ in range(max(args.warmup, 1) - 1):
        run(n_pass)
    times = [run(n_pass) for _ in range(args.steps)]
    tot = sum(times)
    value = n_sample * n_pass * args.steps / tot
    metric = "lqcp_riccati_solves_per_s" if args.workload == "ric" else "box_ipm_qp_solves_per_s"
    sample = f"{n_sample} distinct instances x {n_pass} passes per step, {kind} lib4 build, {cores} threads, FTZ on"
    out = {"metric": metric, "value": value, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "impl": "reference",
           "config": {"workload": workload_name(args.workload), "note": "reference CPU path, bounded sample of the same workload"},
           "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "reference", "sample": sample},
           "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def workload_name(w):
    return ("batched Riccati factor+solve (d_back_ric_rec_sv), 65536 instances/GPU, mass-spring nx=12 nu=5 N=30, FP64" if w == "ric"
            else "box-constrained Riccati IPM (d_ip2_res_mpc_hard), 16384 instances/GPU, mass-spring nx=24 nu=11 N=50, tol 1e-8, FP64")


# ------------------------------------------------------------------------------------------- our arm (GPU)
def bind_to_gpu_numa_node(local_rank: int):
    """Run this process (and so first-touch its pinned staging buffers) on the host cores of the NUMA node the GPU hangs off:
    with 8 ranks feeding 8 PCIe links, staging memory on one socket was the e2e limiter of round 1 (SCALE_r01: 0.37 efficiency).
    Returns a short description for the JSON line; does nothing when sysfs has no answer."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return "numa: single node"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"numa: rank bound to node {node} ({len(allowed)} cores)"
        return f"numa: node {node} has no allowed cores"
    except Exception as e:                                  # noqa: BLE001
        return f"numa: not bound ({type(e).__name__})"


def time_steps(launch, steps, warmup, stream, barrier):
    import torch
    for _ in range(warmup):
        launch()
    torch.cuda.synchronize()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record(stream)
    for i in range(steps):
        launch()
        ev[i + 1].record(stream)
    torch.cuda.synchronize()
    barrier()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return ev[0].elapsed_time(ev[steps]), per


def run_ours(args):
    import torch
    import torch.distributed as dist
    from hpmpc_b200 import capi
    from hpmpc_b200.batchgen import BatchSpec

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- hpmpc_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_note = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)
    L = capi.product()
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    hbm_peak, peak_src = measured_peaks()

    def bench_ric():
        spec = BatchSpec("cfg2", device=local)
        h = spec.h
        if args.ctas_per_sm or args.warps:
            assert h.set_launch(args.ctas_per_sm, args.warps) == 0
        n = args.n_inst or 65536
        first = rank * n                                   # weak scaling: every GPU gets its own 65 536 instances
        d_in = spec.torch_batch(n, first, device=dev)
        ux = torch.zeros((n, h.sz.ux_stride), dtype=torch.float64, device=dev)
        pi = torch.zeros((n, h.sz.pi_stride), dtype=torch.float64, device=dev)

        def launch():
            rc = L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), None, st)
            assert rc == 0
        clk = ClockSampler(local)
        clk.start()
        tot_ms, per = time_steps(launch, args.steps, args.warmup, stream, barrier)
        clocks = clk.stop()
        tot_ms = reduce_max_time(tot_ms, dev)
        w = algorithmic_work(spec.base)
        launch_ms = float(np.mean(per))
        res = dict(n=n, tot_ms=tot_ms, launch_ms=launch_ms, w=w, clocks=clocks, sizes=h.sz, spec=spec,
                   value=world * n * args.steps / (tot_ms * 1e-3))
        # ---- end to end: pinned host buffers through the host-buffer entry point
        if not args.no_e2e:
            ne = n                          # the same e2e batch at every N: 65 536 instances = 6 GB pinned per rank
            h_in = torch.empty((ne, h.sz.in_stride), dtype=torch.float64, pin_memory=True)
            h_in.copy_(d_in[:ne])
            h_ux = torch.empty((ne, h.sz.ux_stride), dtype=torch.float64, pin_memory=True)
            h_pi = torch.empty((ne, h.sz.pi_stride), dtype=torch.float64, pin_memory=True)
            torch.cuda.synchronize()

            def e2e_step():
                rc = L.hpmpc_b200_d_back_ric_rec_sv_batch_host(h.h, ne, h_in.data_ptr(), h_ux.data_ptr(), h_pi.data_ptr())
                assert rc == 0
            ke = max(1, min(args.steps, args.e2e_steps))
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ke):
                e2e_step()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            barrier()
            te = reduce_max_time(t1 - t0, dev)
            res["e2e"] = {"value": world * ne * ke / te, "unit": "solves/s", "h2d_bytes_per_step": int(ne * h.sz.in_stride * 8),
                          "d2h_bytes_per_step": int(ne * (h.sz.ux_stride + h.sz.pi_stride) * 8), "steps": ke, "instances_per_gpu": ne,
                          "note": "pinned host buffers -> hpmpc_b200_d_back_ric_rec_sv_batch_host (chunked H2D / kernel / D2H), PCIe-bound; " + numa_note}
            assert float((h_ux[:, :8] - ux[:ne].cpu()[:, :8]).abs().max()) == 0.0
            del h_in, h_ux, h_pi
        return res, (d_in, ux, pi)

    def bench_ipm(steps, warmup, e2e=True):
        spec = BatchSpec("cfg3", device=local)
        h = spec.h
        n, k_max = args.n_inst_ipm or 16384, 40
        d_in = spec.torch_batch(n, rank * n, device=dev)
        z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device=dev)
        ux, pi, lam, t, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(h.sz.lam_stride), z(h.sz.lam_stride), z(6 + 5 * k_max)

        def launch():
            rc = L.hpmpc_b200_d_ip2_res_mpc_hard_batch(h.h, n, d_in.data_ptr(), k_max, 2.0, 1e-8, 1e-8, 0, ux.data_ptr(), pi.data_ptr(),
                                                       lam.data_ptr(), t.data_ptr(), info.data_ptr(), st)
            assert rc == 0
        tot_ms, per = time_steps(launch, steps, warmup, stream, barrier)
        tot_ms = reduce_max_time(tot_ms, dev)
        kk = info[:, 0]
        w = algorithmic_work(spec.base)
        mean_kk = float(kk.mean())
        launch_ms = float(np.mean(per))
        wave = max(1, h.sz.ipm_grid * h.sz.ipm_warps_per_cta)
        try:
            ipm_traffic_wave = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("hb_ipm_bytes_per_launch") if h.sz.ipm_fast_variant == 0 else None
        except Exception:
            ipm_traffic_wave = None
        multi = h.sz.ipm_fast_variant >= 0 and n >= 2 * wave and os.environ.get("HPMPC_B200_IPM_FUSED", "0") != "1"
        if multi:
            # multi-kernel driver (cipm_kernels.cu): 3 launches up front, then k_max rounds of factorisation, forward, step, trs, step, res, step
            split = os.environ.get("HPMPC_B200_IPM_SV2", "1") != "0" and os.environ.get("HPMPC_B200_IPM_LIGHT", "1") != "0" and h.sz.ipm_fast_variant == 0
            n_launch, kern = 3 + (7 if split else 6) * k_max, ("multi-kernel driver, k_max rounds enqueued: hb_cipm_sv2_kernel<24,11> (factorisation, two instances per warp; "
                                                               "~50 % of an iteration) + hb_cipm_sweep_kernel<forward> + <trs> + <res> (16 warps/SM) + 3 x hb_cipm_step_kernel")
            try:
                ipm_traffic_wave = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("hb_cipm_sv2_bytes_per_launch") if h.sz.ipm_fast_variant == 0 else None
            except Exception:
                ipm_traffic_wave = None
            tnote = "DRAM bytes of ONE launch of the dominant kernel (hb_cipm_sv2_kernel, all 16384 instances active) from the ncu capture in profiles/"
        else:
            n_launch, kern = (n + wave - 1) // wave, "hb_ipm_kernel<S> (whole IPM in one kernel, one launch per wave)"
            tnote = "DRAM bytes of one wave launch (1184 instances) from the ncu capture in profiles/"
        out = {"metric": "box_ipm_qp_solves_per_s", "value": world * n * steps / (tot_ms * 1e-3), "unit": "solves/s",
               "workload": workload_name("ipm"), "steps": steps, "ms_per_step": tot_ms / steps, "mean_iterations": mean_kk,
               "converged": int((info[:, 1] == 0).sum()), "instances_per_gpu": n,
               "launch": {"grid": h.sz.ipm_grid, "warps_per_cta": h.sz.ipm_warps_per_cta, "wave": wave, "fast_variant": h.sz.ipm_fast_variant},
               "gpu_launches": steps * n_launch,
               "roofline": {"bound": "hbm", "achieved": w["B_it"] * mean_kk * n / (launch_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": w["B_it"] * mean_kk * n / (launch_ms * 1e-3) / 1e9 / hbm_peak, "traffic": ipm_traffic_wave,
                            "traffic_note": tnote, "kernel": kern,
                            "bytes_per_iteration_model": w["B_it"], "flops_per_iteration_model": w["F_it"],
                            "fp64_tflops": w["F_it"] * mean_kk * n / (launch_ms * 1e-3) / 1e12}}
        if e2e:
            pin = lambda m: torch.empty((n, max(int(m), 2)), dtype=torch.float64, pin_memory=True)
            h_in = pin(h.sz.in_stride); h_in.copy_(d_in)
            hux, hpi, hlam, ht, hinfo = pin(h.sz.ux_stride), pin(h.sz.pi_stride), pin(h.sz.lam_stride), pin(h.sz.lam_stride), pin(6 + 5 * k_max)
            torch.cuda.synchronize()

            def e2e_step():
                rc = L.hpmpc_b200_d_ip2_res_mpc_hard_batch_host(h.h, n, h_in.data_ptr(), k_max, 2.0, 1e-8, 1e-8, 0, hux.data_ptr(), hpi.data_ptr(),
                                                                hlam.data_ptr(), ht.data_ptr(), hinfo.data_ptr())
                assert rc == 0
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            e2e_step()
            t1 = time.perf_counter()
            barrier()
            te = reduce_max_time(t1 - t0, dev)
            assert float((hinfo[:, 0] - info.cpu()[:, 0]).abs().max()) == 0.0
            out["e2e"] = {"value": world * n / te, "unit": "solves/s", "h2d_bytes_per_step": int(n * h.sz.in_stride * 8),
                          "d2h_bytes_per_step": int(n * (h.sz.ux_stride + h.sz.pi_stride + 2 * h.sz.lam_stride + 6 + 5 * k_max) * 8), "steps": 1}
        spec.h.close()
        return out

    def bench_shared(steps, warmup):
        """SEPARATELY LABELLED MODE, not the headline: shared dynamics -- one set of config-2 matrices for the whole batch, per-instance
        vectors b, q, r (fleets of identical systems / sweeps over initial states; how the reference's own test programs call the
        solver).  Factor once + batched solve with the stored factor whose matrices sit in shared memory."""
        import ctypes as C
        spec = BatchSpec("cfg2", device=local)
        h, base = spec.h, spec.problem(0)
        n = args.n_inst_shared or 262144
        L.hpmpc_b200_shared_factor_doubles.restype = C.c_longlong; L.hpmpc_b200_shared_factor_doubles.argtypes = [C.c_void_p]
        L.hpmpc_b200_d_back_ric_rec_trf_shared.argtypes = [C.c_void_p] * 4
        L.hpmpc_b200_d_back_ric_rec_trs_shared_batch.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 6
        L.hpmpc_b200_d_back_ric_rec_sv_shared_batch_host.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 4
        n_ux, n_pi = sum(base.nx) + sum(base.nu), sum(base.nx[1:])
        vs = h.sz.ux_stride + h.sz.pi_stride
        blk = h.pack(base)
        v0 = np.zeros(vs)
        v0[:n_ux] = np.concatenate([np.concatenate([base.r[k], base.q[k]]) for k in range(base.N + 1)])
        v0[h.sz.ux_stride:h.sz.ux_stride + n_pi] = np.concatenate(base.b)
        g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
        d_vec = torch.from_numpy(v0).to(dev)[None, :] * (1.0 + 0.2 * (torch.rand((n, vs), generator=g, device=dev, dtype=torch.float64) - 0.5))
        d_blk = torch.from_numpy(blk).to(dev)
        d_L = torch.zeros(L.hpmpc_b200_shared_factor_doubles(h.h) + 8, dtype=torch.float64, device=dev)
        ux = torch.zeros((n, h.sz.ux_stride), dtype=torch.float64, device=dev); pi = torch.zeros((n, h.sz.pi_stride), dtype=torch.float64, device=dev)

        def launch():
            assert L.hpmpc_b200_d_back_ric_rec_trf_shared(h.h, d_blk.data_ptr(), d_L.data_ptr(), st) == 0
            assert L.hpmpc_b200_d_back_ric_rec_trs_shared_batch(h.h, n, d_blk.data_ptr(), d_L.data_ptr(), d_vec.data_ptr(), ux.data_ptr(), pi.data_ptr(), st) == 0
        tot_ms, per = time_steps(launch, steps, warmup, stream, barrier)
        tot_ms = reduce_max_time(tot_ms, dev)
        bytes_inst = 8.0 * (n_ux + n_pi) * 2                     # vectors in, solution out
        out = {"metric": "lqcp_riccati_solves_per_s_shared_dynamics", "value": world * n * steps / (tot_ms * 1e-3), "unit": "solves/s",
               "ms_per_step": tot_ms / steps, "instances_per_gpu": n, "gpu_launches": 2 * steps,
               "workload": "SHARED DYNAMICS (not the headline workload): config-2 matrices stored once per batch, per-instance b, q, r; factor once + batched solve with the stored factor",
               "roofline": {"bound": "hbm", "achieved": bytes_inst * n / (float(np.mean(per)) * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": bytes_inst * n / (float(np.mean(per)) * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                            "kernel": ("hb_ric_trs_shared_kernel" if os.environ.get("HPMPC_B200_SHARED_GENERIC") else "hb_ric_trs_shared_tpi_kernel<12,5>") + " (+ hb_ric_trf_kernel once)", "algorithmic_bytes_per_solve": bytes_inst}}
        # the same solve on STAGE-MAJOR vectors (random data of the same shape: the layout, not the values, is what is measured)
        try:
            L.hpmpc_b200_d_back_ric_rec_trs_shared_batch_stage_major.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 6

            def launch_sm():
                assert L.hpmpc_b200_d_back_ric_rec_trf_shared(h.h, d_blk.data_ptr(), d_L.data_ptr(), st) == 0
                assert L.hpmpc_b200_d_back_ric_rec_trs_shared_batch_stage_major(h.h, n, d_blk.data_ptr(), d_L.data_ptr(), d_vec.data_ptr(), ux.data_ptr(), pi.data_ptr(), st) == 0
            tot_sm, per_sm = time_steps(launch_sm, steps, warmup, stream, barrier)
            tot_sm = reduce_max_time(tot_sm, dev)
            out["stage_major_vectors"] = {"value": world * n * steps / (tot_sm * 1e-3), "unit": "solves/s", "ms_per_step": tot_sm / steps,
                                          "frac": bytes_inst * n / (float(np.mean(per_sm)) * 1e-3) / 1e9 / hbm_peak,
                                          "entry": "hpmpc_b200_d_back_ric_rec_trs_shared_batch_stage_major"}
            launch(); torch.cuda.synchronize()          # ux, pi back in the instance-major layout for the e2e check below
        except Exception as e:      # noqa: BLE001
            out["stage_major_vectors"] = {"error": repr(e)}
        pin = lambda m: torch.empty((n, m), dtype=torch.float64, pin_memory=True)
        h_vec, h_ux, h_pi = pin(vs), pin(h.sz.ux_stride), pin(h.sz.pi_stride)
        h_vec.copy_(d_vec)
        torch.cuda.synchronize()

        def e2e_step():
            assert L.hpmpc_b200_d_back_ric_rec_sv_shared_batch_host(h.h, n, blk.ctypes.data, h_vec.data_ptr(), h_ux.data_ptr(), h_pi.data_ptr()) == 0
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            e2e_step()
        t1 = time.perf_counter()
        barrier()
        te = reduce_max_time(t1 - t0, dev)
        assert float((h_ux[:64, :n_ux] - ux[:64, :n_ux].cpu()).abs().max()) == 0.0
        out["e2e"] = {"value": world * n * 3 / te, "unit": "solves/s", "h2d_bytes_per_step": int(n * vs * 8 + blk.nbytes),
                      "d2h_bytes_per_step": int(n * (h.sz.ux_stride + h.sz.pi_stride) * 8), "steps": 3}
        if rank == 0 and world == 1 and not args.no_cpu:
            try:
                from oracle import api as oracle
                cores = os.cpu_count()
                ns = 4096
                rs = oracle.RefSample(spec, 1, want=("pm",))
                vec_s = np.ascontiguousarray(np.concatenate([h_vec[:ns, :n_ux].numpy(), h_vec[:ns, h.sz.ux_stride:h.sz.ux_stride + n_pi].numpy()], axis=1))
                fn = oracle.lib().ref_harness_ric_trs_shared
                fn.restype = C.c_double
                fn.argtypes = [C.c_char_p, C.c_int, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_long]
                off = np.asarray(rs.off_pm, dtype=np.int64)
                from hpmpc_b200.capi import int_array
                run = lambda n_pass: fn(oracle.REF_AVX2.encode(), cores, ns, n_pass, base.N, int_array(base.nx), int_array(base.nu), rs.pm.ctypes.data,
                                        off.ctypes.data, vec_s.ctypes.data, n_ux + n_pi, n_ux, None, n_ux)
                t1_ = run(1)
                n_pass = max(1, int(round(2.0 / max(t1_, 1e-6))))
                sec = run(n_pass)
                out["cpu_baseline"] = {"value": ns * n_pass / sec, "unit": "solves/s", "cores": cores, "kind": "reference",
                                       "sample": f"{ns} instances x {n_pass} passes, reference X64_AVX2 build: d_back_ric_rec_trf_tv_res once per thread and pass, d_back_ric_rec_trs_tv_res per instance"}
                out["e2e_vs_cpu_baseline"] = out["e2e"]["value"] / out["cpu_baseline"]["value"]
            except Exception as e:      # noqa: BLE001
                out["cpu_baseline"] = {"value": None, "sample": f"failed: {e!r}"}
        h.close()
        return out

    def bench_tree(steps, warmup):
        """BASELINE config 5: scenario trees md=4, Nr=3, Nh=20, nx=12, nu=5 (1173 nodes), subtrees sharded over the ranks (strong scaling)."""
        from hpmpc_b200 import tree as T
        t0 = T.mass_spring_tree(12, 5, 4, 3, 20)
        h = T.TreeBatch(t0, device=local)
        n = args.n_trees or 1024
        base = torch.from_numpy(h.pack(t0)).to(dev)
        mask = torch.zeros_like(base)
        for nd in range(h.sz.Nn):
            nux = t0.nu[nd] + t0.nx[nd]
            mask[h.off[nd]["RSQ"]:h.off[nd]["RSQ"] + nux * (nux + 1) // 2 + nux] = 1.0
        from hpmpc_b200.problems import instance_xi
        xi = torch.from_numpy(instance_xi(n)[:, 2].copy()).to(dev)
        d_in = base[None, :] * (1.0 + 0.1 * xi[:, None] * mask[None, :])       # per-instance cost scaling: distinct, well-posed problems
        ux = torch.zeros((n, h.sz.ux_stride), dtype=torch.float64, device=dev)
        pi = torch.zeros((n, h.sz.pi_stride), dtype=torch.float64, device=dev)
        Lst = torch.zeros((n, h.sz.L_stride), dtype=torch.float64, device=dev)
        # subtree sharding: the 16 depth-2 subtrees (their roots, their 4 tails each) are split over the ranks; the one exchange per
        # solve is an all-gather of the subtree roots' factor blocks (NCCL over NVLink); the 5 nodes above are solved redundantly
        ns = h.sz.n_shard_nodes
        lo, hi = shard_range(ns, rank, world)
        tlo, thi = (h.subtrees[lo]["tail_lo"], h.subtrees[hi - 1]["tail_hi"]) if hi > lo else (0, 0)
        o0, ln = h.subtrees[0]["off_L"], h.subtrees[0]["len_L"]
        assert all(h.subtrees[k]["off_L"] == o0 + k * ln for k in range(ns))             # subtree roots are contiguous in the stash
        assert ns % world == 0, "subtrees must divide evenly over the ranks"
        # the exchange is the LIBRARY's: hpmpc_b200_d_tree_back_ric_rec_sv_batch_mg issues the NCCL all-gather itself on this stream;
        # torch.distributed only hands the communicator's unique id around
        import ctypes as C
        Lt = h.L
        Lt.hpmpc_b200_comm_unique_id.argtypes = [C.c_void_p, C.c_int]
        Lt.hpmpc_b200_comm_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_int]
        Lt.hpmpc_b200_comm_destroy.argtypes = [C.c_void_p]
        Lt.hpmpc_b200_d_tree_back_ric_rec_sv_batch_mg.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong] + [C.c_void_p] * 5
        comm = C.c_void_p()
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                buf = (C.c_char * 128)()
                assert Lt.hpmpc_b200_comm_unique_id(buf, 128) == 0
                uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone().to(dev)
            dist.broadcast(uid, 0)
            idb = (C.c_char * 128).from_buffer_copy(uid.cpu().numpy().tobytes())
            assert Lt.hpmpc_b200_comm_create(C.byref(comm), world, rank, idb, local) == 0

        def launch():
            assert Lt.hpmpc_b200_d_tree_back_ric_rec_sv_batch_mg(h.h, comm, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), Lst.data_ptr(), st) == 0
        tot_ms, per = time_steps(launch, steps, warmup, stream, barrier)
        tot_ms = reduce_max_time(tot_ms, dev)
        # flops: edge-wise sums of SURVEY.md section 8d; bytes: inputs once + factors written and read + outputs
        F = 0.0
        for nd in range(h.sz.Nn):
            nux = t0.nu[nd] + t0.nx[nd]
            F += nux ** 3 / 3
            if nd > 0:
                dd = t0.topo["dad"][nd]
                nuxd, nxk = t0.nu[dd] + t0.nx[dd], t0.nx[nd]
                F += nuxd * nxk ** 2 + nuxd ** 2 * nxk + 4 * nxk ** 2 + 4 * nuxd * nxk
        Bt = 8.0 * (h.sz.in_stride + h.sz.ux_stride + h.sz.pi_stride)
        out = {"metric": "tree_riccati_solves_per_s", "value": n * steps / (tot_ms * 1e-3), "unit": "trees/s", "n_gpus": world, "steps": steps,
               "warmup": warmup, "ms_per_step": tot_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": {"workload": f"scenario-tree Riccati factor+solve (d_tree_back_ric_rec_sv), {n} trees, md=4 Nr=3 Nh=20 nx=12 nu=5 (1173 nodes), FP64",
                          "parallelism": f"16 depth-2 subtrees sharded over {world} GPU(s), 5 nodes above replicated, one NCCL all-gather of subtree-root factor blocks per solve, issued inside the C library (hpmpc_b200_d_tree_back_ric_rec_sv_batch_mg)"},
               "roofline": {"bound": "hbm", "achieved": Bt * n / (float(np.mean(per)) * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": Bt * n / (float(np.mean(per)) * 1e-3) / 1e9 / hbm_peak, "traffic": 9.63e6 * n / world,
                            "traffic_source": "static: profiles/r02_tree_traffic.txt (ncu dram__bytes_read.sum + dram__bytes_write.sum of the 8 launches of one solve, "
                                              "1024 trees on one GPU: 9.63 MB per tree; the tail kernels run at 3.8 / 6.3 TB/s of DRAM traffic)",
                            "kernel": "hbk_tail_kernel (backward, forward) + 6 x hbk_top_kernel",
                            "algorithmic_bytes_per_tree": Bt, "algorithmic_flops_per_tree": F},
               "gpu_launches": 8 * steps}
        if world > 1:
            Lt.hpmpc_b200_comm_destroy(comm)
        h.close()
        return out

    def bench_tree_ipm(steps, warmup):
        """BASELINE config 5 as named (d_tree_ip2_res_hard): box IPM over scenario trees md=4, Nr=3, Nh=20, nx=12, nu=5 (1173 nodes,
        u in [-0.5, 0.5] at every node); the batch of trees is split over the ranks (independent units, no collective)."""
        import ctypes as C
        from hpmpc_b200 import tree as T
        from hpmpc_b200.problems import instance_xi
        t0 = T.mass_spring_tree(12, 5, 4, 3, 20, bounds=True)
        h = T.TreeBatch(t0, device=local)
        n = args.n_trees or 2048
        k_max = 40
        base = torch.from_numpy(h.pack(t0)).to(dev)
        mask = torch.zeros_like(base)
        for nd in range(h.sz.Nn):
            nux = t0.nu[nd] + t0.nx[nd]
            mask[h.off[nd]["RSQ"]:h.off[nd]["RSQ"] + nux * (nux + 1) // 2 + nux] = 1.0
        xi = torch.from_numpy(instance_xi(n, first=rank * n)[:, 2].copy()).to(dev)
        d_in = base[None, :] * (1.0 + 0.1 * xi[:, None] * mask[None, :])       # per-tree cost scaling: distinct problems, same bounds
        z = lambda m: torch.zeros((n, max(int(m), 2)), dtype=torch.float64, device=dev)
        ux, pi, lam, tt, info = z(h.sz.ux_stride), z(h.sz.pi_stride), z(2 * h.nbtot), z(2 * h.nbtot), z(6 + 5 * k_max)
        fn = h.L.hpmpc_b200_d_tree_ip2_res_mpc_hard_batch

        def launch():
            assert fn(h.h, n, d_in.data_ptr(), k_max, C.c_double(2.0), C.c_double(1e-8), C.c_double(1e-8), 0, ux.data_ptr(), pi.data_ptr(),
                      lam.data_ptr(), tt.data_ptr(), info.data_ptr(), st) == 0
        h.L.hpmpc_b200_tree_launch_count.restype = C.c_longlong
        h.L.hpmpc_b200_tree_launch_count.argtypes = [C.c_void_p]
        launch(); torch.cuda.synchronize()
        l0 = h.L.hpmpc_b200_tree_launch_count(h.h); launch(); torch.cuda.synchronize(); per_solve = h.L.hpmpc_b200_tree_launch_count(h.h) - l0
        tot_ms, per = time_steps(launch, steps, warmup, stream, barrier)
        tot_ms = reduce_max_time(tot_ms, dev)
        kk = info[:, 0]
        conv = int((info[:, 1] == 0).sum().item())
        out = {"metric": "tree_box_ipm_solves_per_s", "value": world * n * steps / (tot_ms * 1e-3), "unit": "trees/s", "n_gpus": world, "steps": steps,
               "warmup": warmup, "ms_per_step": tot_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": {"workload": f"scenario-tree box IPM (d_tree_ip2_res_mpc_hard), {n} trees/GPU, md=4 Nr=3 Nh=20 nx=12 nu=5 (1173 nodes, "
                                      f"{h.nbtot} bounded inputs), tol 1e-8, FP64",
                          "parallelism": f"trees sharded over {world} GPU(s), no collective (one warp per tree, whole IPM in one kernel)"},
               "mean_iterations": float(kk.mean().item()), "converged": conv, "trees_per_gpu": n, "gpu_launches": int(per_solve) * steps}
        # per-iteration byte model of SURVEY.md section 8d applied edge-wise: factor pass reads the data and writes L, the corrector solve
        # re-reads L and [B A b]', vectors four times:  B_it = 8 (2 D_BAbt + D_RSQ + 2 D_L + 4 D_vec)
        D_B = sum((t0.nu[t0.topo["dad"][k]] + t0.nx[t0.topo["dad"][k]] + 1) * t0.nx[k] for k in range(1, h.sz.Nn))
        D_Q = sum((t0.nu[k] + t0.nx[k]) * (t0.nu[k] + t0.nx[k] + 1) // 2 + t0.nu[k] + t0.nx[k] for k in range(h.sz.Nn))
        D_v = sum(t0.nu) + 2 * sum(t0.nx) + 6 * h.nbtot
        B_it = 8.0 * (2 * D_B + D_Q + 2 * D_Q + 4 * D_v)
        ach = B_it * out["mean_iterations"] * n / (float(np.mean(per)) * 1e-3) / 1e9
        out["roofline"] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                           "kernel": "whole solve (hb_tipm_step_kernel + hbk_tail_kernel + hbk_top_kernel + hb_tipm_res_kernel)",
                           "bytes_per_iteration_model": B_it}
        if rank == 0 and not args.no_cpu:
            from oracle import api
            tt_ = api.TreeTimed(t0, True)
            t1 = time.perf_counter(); tt_.run(); dt = time.perf_counter() - t1
            out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "trees/s", "cores": 1, "kind": "port",
                                   "sample": f"1 tree, oracle/ric_oracle.c orc_tree_ip2_res_mpc_hard (kk={tt_.kk.value}); the reference's own tree IPM needs BLASFEO (absent)"}
        h.close()
        return out

    if args.workload == "tree_ipm":
        tr = bench_tree_ipm(min(args.steps, 3), min(args.warmup, 1))
        if rank == 0:
            print(json.dumps(tr))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    if args.workload == "tree":
        tr = bench_tree(args.steps, args.warmup)
        if rank == 0:
            print(json.dumps(tr))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    if args.workload == "shared":
        sh = bench_shared(max(args.steps, 3), max(args.warmup, 3))
        if rank == 0:
            print(json.dumps(sh))
        return

    if args.workload == "ipm":
        ipm = bench_ipm(args.steps, args.warmup)
        if rank == 0:
            ipm.update({"n_gpus": world, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                        "dtype": "f64", "data": "synthetic", "config": {"workload": ipm.pop("workload"), "launch": ipm.pop("launch")}})
            print(json.dumps(ipm))
        return

    res, keep = bench_ric()
    w, n, sz = res["w"], res["n"], res["sizes"]
    launch_s = res["launch_ms"] * 1e-3
    achieved = w["B_sv"] * n / launch_s / 1e9
    fp64_peak = L.hpmpc_b200_fp64_peak_tflops(local)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and sz.fast_variant == 0 and n == 65536:
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("hb_ric_sv_bytes_per_launch")
            traffic_src = "static: " + tj.get("hb_ric_sv_source", "profiles/ncu_traffic.json") + " (ncu cannot run inside the timed bench)"
        except Exception:
            traffic = None
    out = {"metric": "lqcp_riccati_solves_per_s", "value": res["value"], "unit": "solves/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": res["tot_ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name("ric"), "instances_per_gpu": n, "l2": "inputs (6.4 GB per GPU) exceed the 126 MB L2; no flush between steps",
                      "launch": {"grid": sz.grid, "warps_per_cta": sz.warps_per_cta, "smem_per_cta": sz.smem_per_cta}, "parallelism": f"instances sharded, {world} GPU(s), no collective"},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": peak_src, "kernel": ("hbk_ric_sv_kernel<12,5,8,2>" if sz.fast_variant == 0 else "hb_ric_sv_kernel"), "launch_ms": res["launch_ms"],
                        "algorithmic_bytes_per_solve": w["B_sv"], "algorithmic_flops_per_solve": w["F_sv"],
                        "fp64": {"achieved_tflops": w["F_sv"] * n / launch_s / 1e12, "peak_tflops": fp64_peak, "frac": w["F_sv"] * n / launch_s / 1e12 / fp64_peak if fp64_peak > 0 else None,
                                 "peak_source": "measured DFMA probe (hb_fp64_probe)"}},
           "clocks": res["clocks"], "gpu_launches": args.steps}
    if "e2e" in res:
        out["e2e"] = res["e2e"]
    del keep
    torch.cuda.empty_cache()
    if not args.no_ipm:
        try:
            out["extra"] = {"ipm": bench_ipm(max(1, min(args.steps, args.ipm_steps)), 1, e2e=(world == 1))}
        except Exception as e:      # the secondary workload must not hide the headline number
            out["extra"] = {"ipm_error": repr(e)}
        # BASELINE config 5 (scenario trees) at EVERY N, so that the driver's scaling run records it: the Riccati solve of a fixed
        # batch of trees with the subtrees sharded over the ranks (STRONG scaling, one exchange of subtree-root factor blocks per
        # solve) and the box IPM over the trees (whole trees per rank, weak).  Short runs; --workload tree / tree_ipm are the full lines.
        torch.cuda.empty_cache()
        try:
            saved = args.n_trees, args.no_cpu
            args.n_trees, args.no_cpu = args.n_trees or 4096, True          # the same 4096 trees at every N: strong scaling
            tr = bench_tree(3, 2)
            if rank == 0:
                out["extra"]["tree"] = {k: tr[k] for k in ("metric", "value", "unit", "ms_per_step", "gpu_launches", "scaling", "n_gpus", "roofline")} | \
                    {"workload": tr["config"]["workload"], "parallelism": tr["config"]["parallelism"]}
            args.n_trees = saved[0] or 1024                # per GPU (weak)
            ti = bench_tree_ipm(2, 1)
            if rank == 0:
                out["extra"]["tree_ipm"] = {k: ti[k] for k in ("metric", "value", "unit", "ms_per_step", "mean_iterations", "converged", "scaling", "n_gpus")} | \
                    {"workload": ti["config"]["workload"]}
            args.n_trees, args.no_cpu = saved
        except Exception as e:
            out["extra"]["tree_error"] = repr(e)
        torch.cuda.empty_cache()
        try:
            sh = bench_shared(3, 2)
            if rank == 0:
                out["extra"]["shared_dynamics"] = sh
        except Exception as e:      # noqa: BLE001
            out["extra"]["shared_dynamics_error"] = repr(e)
        if world == 1:
            # a shape WITHOUT kernels of its own, nx=10 nu=4 N=25: embedded in the (12,5) frame (hpmpc_b200_ocp_create_padded) against
            # the any-size kernels on the same batch and against the compiled (12,5) shape at the same N
            torch.cuda.empty_cache()
            try:
                from hpmpc_b200 import problems as _pr
                n_u = 32768
                res_u = {}
                for tag, shp, padded in (("padded_into_12_5", (10, 4, 25), True), ("any_size_kernels", (10, 4, 25), False), ("compiled_12_5", (12, 5, 25), False)):
                    ps = [_pr.mass_spring_ocp(*shp, xi=tuple(xi)) for xi in _pr.instance_xi(64, first=0)]
                    hu = capi.BatchOcp(ps[0], device=local, padded=padded)
                    d_u = torch.from_numpy(np.tile(np.stack([hu.pack(q) for q in ps]), (n_u // 64, 1))).to(dev)
                    uxu = torch.zeros((n_u, hu.sz.ux_stride), dtype=torch.float64, device=dev); piu = torch.zeros((n_u, hu.sz.pi_stride), dtype=torch.float64, device=dev)
                    run_u = lambda: L.hpmpc_b200_d_back_ric_rec_sv_batch(hu.h, n_u, d_u.data_ptr(), uxu.data_ptr(), piu.data_ptr(), None, st)
                    tot_u, _ = time_steps(run_u, 3, 2, stream, barrier)
                    res_u[tag] = {"solves_per_s": n_u * 3 / (tot_u * 1e-3), "fast_variant": hu.sz.fast_variant, "in_bytes_per_solve": int(hu.sz.in_stride * 8)}
                    hu.close()
                    del d_u, uxu, piu
                out["extra"]["unlisted_shape"] = {"workload": "batched Riccati factor+solve, nx=10 nu=4 N=25 (no size-specialised kernel), 32768 instances", **res_u,
                                                  "padded_vs_compiled": res_u["padded_into_12_5"]["solves_per_s"] / res_u["compiled_12_5"]["solves_per_s"],
                                                  "padded_vs_any_size": res_u["padded_into_12_5"]["solves_per_s"] / res_u["any_size_kernels"]["solves_per_s"]}
            except Exception as e:      # noqa: BLE001
                out["extra"]["unlisted_shape_error"] = repr(e)
            # BASELINE config 4: time-varying variable-nx OCP on the any-size kernels (four warps per instance, ric_team.cuh)
            torch.cuda.empty_cache()
            try:
                import importlib.util
                sp4 = importlib.util.spec_from_file_location("bench_cfg4", os.path.join(ROOT, "tools", "bench_cfg4.py"))
                mod4 = importlib.util.module_from_spec(sp4)
                sp4.loader.exec_module(mod4)
                r4 = mod4.measure(8192, 3)
                w4 = algorithmic_work(r4.pop("problem"))
                ach_sv = w4["B_sv"] * r4["n_inst"] / (r4["sv_ms"] * 1e-3) / 1e9
                ach_ipm = w4["B_it"] * r4["mean_kk"] * r4["n_inst"] / (r4["ipm_ms"] * 1e-3) / 1e9
                out["extra"]["cfg4"] = {
                    "workload": "time-varying variable-nx OCP (nx 40 -> 4, nu=8, N=20, box bounds), 8192 instances, any-size kernels, FP64",
                    "sv": {"metric": "lqcp_riccati_solves_per_s", "value": r4["sv_solves_per_s"], "unit": "solves/s", "ms_per_step": r4["sv_ms"],
                           "roofline": {"bound": "hbm", "achieved": ach_sv, "peak": hbm_peak, "unit": "GB/s", "frac": ach_sv / hbm_peak, "traffic": None,
                                        "kernel": "hbt_ric_sv_kernel", "algorithmic_bytes_per_solve": w4["B_sv"],
                                        "fp64_tflops": w4["F_sv"] * r4["sv_solves_per_s"] / 1e12}},
                    "ipm": {"metric": "box_ipm_qp_solves_per_s", "value": r4["ipm_solves_per_s"], "unit": "solves/s", "ms_per_step": r4["ipm_ms"],
                            "mean_iterations": r4["mean_kk"], "converged": r4["converged"],
                            "roofline": {"bound": "hbm", "achieved": ach_ipm, "peak": hbm_peak, "unit": "GB/s", "frac": ach_ipm / hbm_peak, "traffic": None,
                                         "kernel": "multi-kernel driver: hb_cipm_team_kernel<0> (factor + solve) + hb_cipm_team_kernel<1> (solve) + "
                                                   "hb_cipm_sweep_kernel<generic,2> (residuals) + hb_cipm_step_kernel",
                                         "bytes_per_iteration_model": w4["B_it"], "fp64_tflops": w4["F_it"] * r4["mean_kk"] * r4["ipm_solves_per_s"] / 1e12}}}
            except Exception as e:      # noqa: BLE001
                out["extra"]["cfg4_error"] = repr(e)
            # the IPM's last KKT system solved again for a new right-hand side (cfg 3 shapes, 4096 instances)
            torch.cuda.empty_cache()
            try:
                import importlib.util
                import types
                sp = importlib.util.spec_from_file_location("bench_kkt_new_rhs", os.path.join(ROOT, "tools", "bench_kkt_new_rhs.py"))
                mod = importlib.util.module_from_spec(sp)
                sp.loader.exec_module(mod)
                kr = mod.measure(types.SimpleNamespace(cfg="cfg3", n_inst=4096, steps=3, warmup=3))
                out["extra"]["kkt_new_rhs"] = {k: kr[k] for k in ("metric", "value", "unit", "ms_per_step", "ipm_ms", "ipm_with_kkt_state_ms",
                                                                  "speedup_vs_full_ipm_solve", "roofline")} | {"workload": kr["config"]["workload"]}
            except Exception as e:
                out["extra"]["kkt_new_rhs_error"] = repr(e)
    if rank == 0 and world == 1 and not args.no_cpu and isinstance(out.get("extra", {}).get("ipm"), dict):
        # the reference's own box IPM (fortran_order_d_ip_ocp_hard_tv, AVX2 lib4 build) on the host cores, bounded sample of config 3
        try:
            from oracle import api as oracle
            cores = os.cpu_count()
            spec3 = BatchSpec("cfg3", device=-1)
            n_s = max(128, 4 * cores)
            rs3 = oracle.RefSample(spec3, n_s, want=("cm",))
            t1 = rs3.time_ipm("avx2", cores, 1)[0]
            n_pass = max(1, int(round(3.0 / max(t1, 1e-6))))
            sec = rs3.time_ipm("avx2", cores, n_pass)[0]
            v = n_s * n_pass / sec
            out["extra"]["ipm"]["cpu_baseline"] = {"value": v, "unit": "solves/s", "cores": cores, "kind": "reference",
                                                   "sample": f"{n_s} distinct config-3 instances x {n_pass} passes, reference X64_AVX2 lib4 build, fortran_order_d_ip_ocp_hard_tv"}
            if "e2e" in out["extra"]["ipm"]:
                out["extra"]["ipm"]["e2e_vs_cpu_baseline"] = out["extra"]["ipm"]["e2e"]["value"] / v
            spec3.h.close()
        except Exception as e:      # noqa: BLE001
            out["extra"]["ipm"]["cpu_baseline"] = {"value": None, "sample": f"failed: {e!r}"}
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            from oracle import api as oracle
            cores = os.cpu_count()
            kind = "avx2" if os.path.exists(oracle.REF_AVX2) else None
            if kind is None:
                raise RuntimeError("oracle/_ref reference build missing")
            n_sample = max(2048, 16 * cores)
            rs = oracle.RefSample(res["spec"], n_sample)
            t1 = rs.time_ric_sv(kind, cores, 1)[0]
            n_pass = max(1, int(round(2.0 / max(t1, 1e-6))))
            sec = rs.time_ric_sv(kind, cores, n_pass)[0]
            out["cpu_baseline"] = {"value": n_sample * n_pass / sec, "unit": "solves/s", "cores": cores, "kind": "reference",
                                   "sample": f"{n_sample} distinct instances x {n_pass} passes, reference X64_AVX2 lib4 build (oracle/_ref), "
                                             f"d_back_ric_rec_sv_tv_res on pre-packed panel-major data, one workspace per thread, FTZ on"}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "unit": "solves/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e!r}"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ric", choices=["ric", "ipm", "tree", "tree_ipm", "shared"])
    ap.add_argument("--n-trees", type=int, default=0)
    ap.add_argument("--n-inst", type=int, default=0)
    ap.add_argument("--n-inst-ipm", type=int, default=0)
    ap.add_argument("--n-inst-shared", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ipm", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ipm-steps", type=int, default=2)
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--ref-step-seconds", type=float, default=2.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
