"""Summarises an .ncu-rep (read with `ncu -i`, no GPU needed): headline metrics + per-opcode instruction mix.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [units_per_launch unit_name]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
uname = sys.argv[3] if len(sys.argv) > 3 else "unit"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit_row = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_read.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, unit_row))
    print("== launch ==")
    for k in KEYS:
        if k in d:
            print("%-86s %s %s" % (k, d[k], u.get(k, "")))
    if units and "smsp__inst_executed.sum" in d:
        print("warp instructions per %s: %.1f" % (uname, float(d["smsp__inst_executed.sum"]) / units))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for i, r in enumerate(rows):
    if "Address" in r and "Source" in r:
        h = r; data = [x for x in rows[i + 1:] if len(x) == len(h)]
        break
else:
    sys.exit(0)
isrc, iex, ism = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
op, smp = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc]); o = m.group(2) if m else "?"
    op[o] += int(r[iex]); smp[o] += int(r[ism])
tot, ts = sum(op.values()), max(1, sum(smp.values()))
print("== instruction mix (executed warp instructions%s; %% of stall samples) ==" % (" per " + uname if units else ""))
for o, c in op.most_common(24):
    print("%-10s %14s  %5.1f%%  samples %5.1f%%" % (o, ("%.2f" % (c / units)) if units else str(c), 100.0 * c / tot, 100.0 * smp[o] / ts))
