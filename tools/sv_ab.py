"""A/B runs of the headline kernel (batched Riccati factor+solve, BASELINE config 2): library variants x L2 flags x launch shapes.
Every configuration runs in its own process (the flags are read once per process); the results of every run are compared bit for
bit with the first configuration's.   usage: python tools/sv_ab.py [n_inst] [reps] ; worker: python tools/sv_ab.py --worker ..."""
import hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "--worker":
    import torch
    from hpmpc_b200 import capi
    from hpmpc_b200.batchgen import BatchSpec
    n, reps, cps, warps = (int(v) for v in sys.argv[2:6])
    L = capi.product()
    spec = BatchSpec("cfg2"); h = spec.h
    if cps or warps:
        assert h.set_launch(cps, warps) == 0
    d_in = spec.torch_batch(n)
    ux = torch.zeros((n, h.sz.ux_stride), dtype=torch.float64, device="cuda"); pi = torch.zeros((n, h.sz.pi_stride), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: L.hpmpc_b200_d_back_ric_rec_sv_batch(h.h, n, d_in.data_ptr(), ux.data_ptr(), pi.data_ptr(), None, st)
    for _ in range(3):
        assert run() == 0
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        run(); ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    n_ux = sum(spec.base.nx) + sum(spec.base.nu)
    dig = hashlib.sha1(ux[:, :n_ux].contiguous().cpu().numpy().tobytes() + pi.cpu().numpy().tobytes()).hexdigest()[:16]
    print(json.dumps(dict(ms_best=ms[0], ms_med=ms[len(ms) // 2], grid=h.sz.grid, warps=h.sz.warps_per_cta, digest=dig,
                          finite=bool(torch.isfinite(ux[:, :n_ux]).all()))))
    sys.exit(0)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
V = os.path.join(ROOT, "hpmpc_b200", "lib", "variants")
configs = json.loads(os.environ["SV_AB_CONFIGS"]) if "SV_AB_CONFIGS" in os.environ else [
    dict(name="base", lib="", flags=0, keep=0),
    dict(name="base+in_first", lib="", flags=1, keep=0),
    dict(name="base+in_first+re_first", lib="", flags=3, keep=0),
    dict(name="bulk", lib="bulk", flags=0, keep=0),
    dict(name="bulk+discard", lib="bulk", flags=8, keep=0),
    dict(name="bulk+first+discard", lib="bulk", flags=11, keep=0),
    dict(name="bulk+first+keep8+discard", lib="bulk", flags=15, keep=8),
    dict(name="bulk+first+keep16+discard", lib="bulk", flags=15, keep=16),
    dict(name="bulk+first+keep24+discard", lib="bulk", flags=15, keep=24),
    dict(name="bulk+first+keep31+discard", lib="bulk", flags=15, keep=31),
    dict(name="bulk+first+keep31+discard w6", lib="bulk", flags=15, keep=31, warps=6),
    dict(name="bulk+first+keep31+discard w4", lib="bulk", flags=15, keep=31, warps=4),
    dict(name="bulk+first+keep16+discard w6", lib="bulk", flags=15, keep=16, warps=6),
]
ref = None
for c in configs:
    env = dict(os.environ)
    if c.get("lib"):
        env["HPMPC_B200_LIB"] = os.path.join(V, f"libhpmpc_b200_{c['lib']}.so")
    env["HPMPC_B200_SV_FLAGS"] = str(c.get("flags", 0)); env["HPMPC_B200_SV_KEEP"] = str(c.get("keep", 0))
    for k, v in c.get("env", {}).items():
        env[k] = str(v)
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", str(n), str(reps), str(c.get("cps", 0)), str(c.get("warps", 0))],
                         env=env, capture_output=True, text=True)
    if out.returncode != 0:
        print(f"{c['name']:40s} FAILED: {out.stderr[-300:]}")
        continue
    r = json.loads(out.stdout.strip().splitlines()[-1])
    ref = ref or r["digest"]
    same = "bit-identical" if r["digest"] == ref else "DIFFERENT RESULTS"
    print(f"{c['name']:40s} grid {r['grid']:4d} x {r['warps']} warps  best {r['ms_best']:7.3f} ms  median {r['ms_med']:7.3f} ms  "
          f"{n / r['ms_best'] / 1e3:6.2f} M solves/s  frac {n * 97968 / (r['ms_best'] * 1e-3) / 6454.6e9:.3f}  {same}", flush=True)
