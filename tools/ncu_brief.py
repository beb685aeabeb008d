"""Print a brief per-kernel table from an .ncu-rep (run where ncu is installed): python tools/ncu_brief.py file.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = [("time_us", "gpu__time_duration.sum"), ("inst_M", "smsp__inst_executed.sum"), ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("heavy%", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"), ("alu%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed"), ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("rdMB", "dram__bytes_read.sum"), ("wrMB", "dram__bytes_write.sum"), ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("smemKB", "launch__shared_mem_per_block_dynamic"),
        ("st_long", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("st_short", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("st_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("st_bar", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("st_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
        ("st_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
        ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("st_nosel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
        ("st_noinst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
        ("st_disp", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"),
        ("l2hit%", "lts__t_sector_hit_rate.pct"),
        ("bankconf_M", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), ("smem_wf_M", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")]
print("%-34s" % "kernel" + "".join("%10s" % w[0] for w in want))
seen = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200::", "")[:33]
    key = (name, r[idx["Grid Size"]])
    if key in seen:
        continue
    seen[key] = 1
    vals = []
    for short, m in want:
        v = r[idx[m]] if m in idx else ""
        try:
            f = float(v.replace(",", ""))
            u = rows[1][idx[m]]
            if short == "time_us":
                f = f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
            if short in ("inst_M", "bankconf_M", "smem_wf_M"):
                f /= 1e6
            if short in ("rdMB", "wrMB"):
                f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
            vals.append("%10.2f" % f)
        except ValueError:
            vals.append("%10s" % v[:9])
    print("%-34s" % name + "".join(vals))
