"""Per-kernel SASS evidence of the built library: registers / spills / stack (cuobjdump -res-usage) and counts of the mnemonics that
matter on this path -- DFMA/DMUL/DADD (FP64 pipe), DMMA (FP64 tensor-core tiles of the any-size team kernels), LDGSTS (cp.async staging), MUFU.RSQ64H (pivot), UBLKCP (1-D bulk async copies, both directions), SYNCS
(mbarrier), LDS/STS, LDG/STG, SHFL, ATOM/RED, CCTL/discard, LDL/STL (local memory = spills).  No GPU needed.
usage: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "hpmpc_b200", "lib", "libhpmpc_b200.so")
CUOBJ = "/usr/local/cuda/bin/cuobjdump"

res = subprocess.run([CUOBJ, "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1); continue
    if cur and "REG:" in line:
        usage[cur] = dict((k, int(v)) for k, v in re.findall(r"(\w+):(\d+)", line)); cur = None

sass = subprocess.run([CUOBJ, "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
KEYS = ["DFMA", "DMMA", "LDGSTS", "DMUL", "DADD", "MUFU.RSQ64H", "MUFU.RCP64H", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "SHFL", "ATOM", "RED", "CCTL", "LDL", "STL", "BAR"]
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1); counts[cur]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k.startswith("MUFU") and op == k):
                counts[cur][k] += 1

def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))

dm = demangle(list(counts))
print(f"# SASS summary of `{os.path.relpath(lib, ROOT)}` (sm_100a; `cuobjdump -res-usage` and `-sass`, mnemonic counts are static)\n")
print("| kernel | regs | stack B | smem B (static) | instr | DFMA | DMMA | LDGSTS | DMUL+DADD | RSQ64H | UBLKCP | SYNCS | LDS | STS | LDG | STG | SHFL | ATOM/RED | CCTL | LDL/STL |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for k, c in counts.items():
    u = usage.get(k, {})
    name = dm.get(k, k)
    name = re.sub(r"^void ", "", name); name = re.sub(r"\(.*$", "", name)
    print(f"| `{name}` | {u.get('REG', '?')} | {u.get('STACK', '?')} | {u.get('SHARED', '?')} | {c['_total']} | {c['DFMA']} | {c['DMMA']} | {c['LDGSTS']} | {c['DMUL'] + c['DADD']} | "
          f"{c['MUFU.RSQ64H']} | {c['UBLKCP']} | {c['SYNCS']} | {c['LDS']} | {c['STS']} | {c['LDG']} | {c['STG']} | {c['SHFL']} | {c['ATOM'] + c['RED']} | {c['CCTL']} | {c['LDL'] + c['STL']} |")
print("\nDMMA = `mma.sync.m8n8k4.f64` (FP64 tensor-core tiles; the counts of a kernel include the device functions it calls, which cuobjdump lists inside the kernel's section); LDGSTS = `cp.async`. UBLKCP = `cp.async.bulk` (global<->shared 1-D bulk copies issued by one lane, completion on an mbarrier = SYNCS); a non-zero "
      "LDL/STL column is local-memory traffic (spills or indexed private arrays).")
