"""Join an ncu SASS source page (CSV) with nvdisasm -g line info and aggregate per CUDA source line.
usage: python tools/ncu_by_line.py <ncu_source.csv> <nvdisasm -g -c output> <kernel mangled substring> [top]"""
import csv, re, sys
src_csv, dis, key = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; body = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
# disassembly: collect (line, text) for instructions in the kernel's section
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("\t.section\t.text.") and key in l)
cur = None; ins = []
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith("//-----"):
        if ins: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((cur, m.group(2)))
print("sass instrs:", len(ins), "ncu rows:", len(body))
n = min(len(ins), len(body))
agg = {}
tot_i = tot_s = 0
for (loc, txt), r in zip(ins[:n], body[:n]):
    ie = float(r[ix["Instructions Executed"]] or 0); ss = float(r[ix["# Samples"]] or 0)
    a = agg.setdefault(loc, [0, 0, 0]); a[0] += ie; a[1] += ss; a[2] += 1
    tot_i += ie; tot_s += ss
srcs = {}
def srcline(loc):
    if loc is None: return ""
    f = "/root/repo/hpmpc_b200/csrc/" + loc[0]
    if f not in srcs:
        try: srcs[f] = open(f).read().split("\n")
        except Exception: srcs[f] = []
    L = srcs[f]
    return L[loc[1] - 1].strip()[:90] if 0 < loc[1] <= len(L) else ""
print(f"total warp-instr {tot_i:.3e}  samples {tot_s:.0f}")
print("---- by samples")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{str(loc):28s} inst {100*a[0]/tot_i:5.1f}%  samp {100*a[1]/tot_s:5.1f}%  sass {a[2]:4d} | {srcline(loc)}")
