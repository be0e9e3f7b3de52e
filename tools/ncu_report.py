"""Write a compact, committable text summary of an ncu --set full report: headline metrics, stall reasons,
shared-memory wavefronts per LDS/STS kind, top source lines.  usage: python tools/ncu_report.py <file.ncu-rep> <kernel substring> <out.txt>"""
import collections, csv, io, os, re, subprocess, sys
rep, key, out = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit = rows[0], rows[1]
vals = next(r for r in rows[2:] if key in r[hdr.index("Kernel Name")])
d, u = dict(zip(hdr, vals)), dict(zip(hdr, unit))
o = []
o.append(f"# ncu --set full --clock-control none, report {os.path.basename(rep)}")
o.append(f"kernel: {d['Kernel Name']}")
for k in ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
          "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
          "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed.ratio"):
    if k in d:
        o.append(f"  {k:78s} {d[k]:>18s} {u.get(k, '')}")
o.append("stall reasons (pc samples):")
st = sorted(((float(d[k].replace(",", "")), k) for k in d if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and d[k]), reverse=True)
tot = sum(v for v, _ in st) or 1
for v, k in st[:12]:
    o.append(f"  {k.replace('smsp__pcsamp_warps_issue_stalled_', ''):28s} {100 * v / tot:5.1f} %")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; ix = {h: i for i, h in enumerate(h2)}
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
need = max(ix["Source"], ix["Instructions Executed"], ix["L1 Wavefronts Shared"])
for r in rows[2:]:
    if len(r) <= need or r[ix["Source"]] == "Source":      # a report with several launches repeats the header block per launch
        continue
    s = r[ix["Source"]].strip()
    if not s:
        continue
    t = s.split()
    op = t[1] if s.startswith("@") and len(t) > 1 else t[0]
    a = agg[op]
    try:
        a[0] += float(r[ix["Instructions Executed"]] or 0); a[1] += float(r[ix["L1 Wavefronts Shared"]] or 0); a[2] += 1
    except ValueError:
        pass
ti = sum(a[0] for a in agg.values()) or 1
o.append("instruction mix (warp instructions executed; shared-memory wavefronts per instruction):")
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
    o.append(f"  {op:18s} sass {a[2]:5d}  {100 * a[0] / ti:5.1f} %" + (f"   wavefronts/inst {a[1] / a[0]:.2f}" if a[1] > 0 and a[0] > 0 else ""))
open(out, "w").write("\n".join(o) + "\n")
print("\n".join(o))
