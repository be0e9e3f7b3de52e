"""Print the handful of ncu raw metrics we track.  usage: python tools/ncu_summary.py <file.ncu-rep>"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, unit = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, unit))
    print("kernel:", d.get("Kernel Name", "")[:90])
    for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
              "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
              "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"):
        if k in d:
            print(f"  {k:70s} {d[k]:>16s} {u.get(k,'')}")
