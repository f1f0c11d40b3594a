"""Summaries of ncu output for profiles/ (run here, on the CPU box, over files brought back in gpurun_out/):

    python tools/ncu_summary.py rep  gpurun_out/prof_x.ncu-rep [more.ncu-rep ...]   # key raw metrics per captured launch
    python tools/ncu_summary.py list gpurun_out/launches_x.csv                      # per-kernel share of a launch list
"""
import csv
import re
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA pipe % (inst, active)"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA pipe cycles active %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio throttle / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg throttle / issue"),
]


def rep(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        print("# %s: no launches" % path)
        return
    hdr, units = rows[0], rows[1]
    print("# %s  (ncu --set full --clock-control none; %d launches)" % (path, len(rows) - 2))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("%s  grid %s block %s" % (name[:110], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for key, label in KEYS:
            for i, h in enumerate(hdr):
                if h == key or h.endswith("." + key):
                    print("    %-34s %s %s" % (label, r[i], units[i]))
                    break
        # any other dmma metric the section files carry
        for i, h in enumerate(hdr):
            if "dmma" in h and not any(h == k or h.endswith("." + k) for k, _ in KEYS):
                print("    %-34s %s %s" % (h.split(".")[-3] if h.count(".") > 2 else h, r[i], units[i]))


def launch_list(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r and r[0].isdigit()]
    tot, by = 0.0, {}
    hdr = None
    for r in csv.reader(open(path, errors="replace")):
        if r and r[0] == "ID":
            hdr = r
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rows:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)     # -> microseconds
        m = re.search(r"(gett_\w+|\w+_kernel\w*|nccl\w+)", r[ki])
        name = m.group(1) if m else r[ki][:40]
        tot += v
        d = by.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += v
    print("# %s: %d launches, %.1f ms of kernel time under ncu (serialised, cold cache: shares only)" % (path, len(rows), tot / 1e3))
    for name, (n, v) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print("%-40s %6d launches %10.1f ms  %5.1f %%  %8.1f us/launch" % (name, n, v / 1e3, 100 * v / tot, v / n))


if __name__ == "__main__":
    if sys.argv[1] == "rep":
        for p in sys.argv[2:]:
            rep(p)
    else:
        for p in sys.argv[2:]:
            launch_list(p)
