"""Summarise an ncu --csv metrics log of tools/microbench.bin: wavefronts per LDS and cycles per LDS for each launch."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
d = collections.OrderedDict()
for r in rows:
    d.setdefault(int(r[0]), {"k": r[4][:24], "blk": r[7]})[r[12]] = float(r[14].replace(",", ""))
for i, m in d.items():
    nw = int(m["blk"].strip("()").split(",")[0]) // 32
    n_lds = 148 * nw * 2000 * 16
    print(i, m["k"], "nw", nw, "wf/LDS %.2f" % (m["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"] / n_lds),
          "cyc/LDS/SM %.2f" % (m["sm__cycles_elapsed.max"] * 148 / n_lds), "wf%% %.1f" % m["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"])
