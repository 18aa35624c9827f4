"""Compact text summary of an `ncu --set full --import-source on` report: headline metrics per kernel launch plus the
executed-instruction mix (opcode shares) from the SASS source page.
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/r01_ncu_x_summary.txt"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic"]
for li, r in enumerate(rows[2:]):
    print(f"== launch {li}: {r[h.index('Kernel Name')]}  grid {r[h.index('Grid Size')]} block {r[h.index('Block Size')]}")
    for n in want:
        if n in h:
            print(f"   {n:86s} {r[h.index(n)]:>14s} {units[h.index(n)]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    hi = [i for i, x in enumerate(srows) if x and x[0] == "Address"]
    if not hi:
        continue
    sh = srows[hi[0]]
    isrc, iex = sh.index("Source"), sh.index("Instructions Executed")
    seen, op, tot = set(), collections.Counter(), 0
    for x in srows[hi[0] + 1:]:
        if len(x) <= iex or not x[iex].isdigit() or x[0] in seen:
            continue
        seen.add(x[0])
        t = x[isrc].split()
        o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        op[o] += int(x[iex])
        tot += int(x[iex])
    print(f"   executed warp instructions: {tot}")
    print("   " + ", ".join(f"{o} {100 * c / tot:.1f}%" for o, c in op.most_common(16)))
