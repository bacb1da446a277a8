"""Summarises an .ncu-rep (raw + source pages) into a short text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.txt]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fma.sum",
        "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_lsu.sum",
        "smsp__inst_executed_pipe_uniform.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sass__inst_executed_shared_loads", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("== kernel:", name, file=out)
    for h, u, v in zip(hdr, units, r):
        if h in KEYS:
            print("%-85s %-12s %s" % (h, u, v), file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hi:
    h = rows[hi[0]]
    data = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else None)] if len(r) == len(h)]
    ia, isrc = h.index("Instructions Executed"), h.index("Source")
    ops = collections.Counter()
    tot = 0
    for r in data:
        toks = r[isrc].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ops[op.split(".")[0]] += int(r[ia])
        tot += int(r[ia])
    print("== executed warp-instructions by opcode (total %d)" % tot, file=out)
    for k, v in ops.most_common(22):
        print("%-10s %14d %5.1f%%" % (k, v, 100.0 * v / tot), file=out)
    fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print("FP64-pipe warp-instructions: %d (%.1f%% of issued); non-FP64 per FP64: %.2f" % (fp64, 100.0 * fp64 / tot, (tot - fp64) / fp64), file=out)
