"""Summarise an .ncu-rep: headline metrics + executed instructions / stall samples per SASS region.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [region_size]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 150
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
units = rows[1] if len(rows) > 2 else None
for h, v in zip(hdr, vals):
    if h in want:
        print(f"{h} = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci["Instructions Executed"]]) for r in data)
tots = sum(int(r[ci["# Samples"]]) for r in data)
print(f"\nSASS instructions: {len(data)}; executed warp-instructions: {tot}; stall samples: {tots}")
print("region(first SASS idx)  %inst  %samples  avg-threads")
for k in range(0, len(data), step):
    ch = data[k:k + step]
    i = sum(int(r[ci["Instructions Executed"]]) for r in ch)
    s = sum(int(r[ci["# Samples"]]) for r in ch)
    th = sum(int(r[ci["Thread Instructions Executed"]]) for r in ch)
    if i * 200 > tot or s * 200 > tots:
        print(f"{k:6d}  {100*i/tot:5.1f}  {100*s/tots:5.1f}  {th/max(i,1):5.1f}")
