#!/usr/bin/env python
"""One block of key metrics per kernel launch of an ncu report:  python tools/ncu_summary.py report.ncu-rep [--json out.json]
(the --json file maps kernel base names to dram bytes per launch; bench.py reads profiles/dominant_kernel_traffic.json)."""
import csv
import json
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "CTAs"),
    ("launch__block_size", "threads/CTA"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"),
    ("launch__shared_mem_per_block_static", "static smem/CTA"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_bytes.sum", "L1 bytes"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory pipes throughput"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (elapsed)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (active)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("sm__cycles_active.max", "SM active cycles (max)"),
    ("sm__cycles_elapsed.avg", "SM elapsed cycles"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    traffic = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        base = re.sub(r"^void\s+", "", name).split("<")[0].split("(")[0].split("::")[-1]
        print("=== %s" % name[:150])
        for key, label in WANT:
            if key in col:
                print("    %-34s %14s %s" % (label, r[col[key]], units[col[key]]))
        try:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
            traffic.setdefault(base, []).append(rd + wr)
        except (KeyError, ValueError):
            pass
    if "--json" in sys.argv:
        path = sys.argv[sys.argv.index("--json") + 1]
        json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
