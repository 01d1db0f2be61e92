#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the small JSON summary kept under profiles/.
    python tools/profile_summary.py gpurun_out/X.ncu-rep profiles/NAME.json "command that was profiled" """
import csv, json, subprocess, sys
rep, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__block_size", "launch__grid_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
kernels = []
for r in rows[2:]:
    m = {k: {"value": r[hdr.index(k)], "unit": units[hdr.index(k)]} for k in KEYS if k in hdr}
    kernels.append({"kernel": r[hdr.index("Kernel Name")], "metrics": m})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
regions = None
if len(srows) > 2:
    h = srows[1]; iI, iT = h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    iW = h.index("L1 Wavefronts Shared"); iS = h.index("Source")
    body = [r for r in srows[2:] if len(r) == len(h)]
    ti = sum(int(r[iI] or 0) for r in body); tt = sum(int(r[iT] or 0) for r in body); tw = sum(int(r[iW] or 0) for r in body)
    regions = {"sass_warp_instructions": ti, "sass_thread_instructions": tt, "avg_active_lanes": tt / max(ti, 1), "shared_wavefronts": tw,
               "hot_64_instruction_buckets": []}
    for s in range(0, len(body), 64):
        seg = body[s:s + 64]
        i = sum(int(r[iI] or 0) for r in seg); t = sum(int(r[iT] or 0) for r in seg); w = sum(int(r[iW] or 0) for r in seg)
        if i > 0.02 * ti:
            regions["hot_64_instruction_buckets"].append({"first_sass_index": s, "first_instruction": seg[0][iS].strip(), "share_of_warp_instructions": round(i / ti, 4),
                                                          "avg_active_lanes": round(t / max(i, 1), 2), "share_of_shared_wavefronts": round(w / max(tw, 1), 4)})
json.dump({"command": cmd, "report": rep, "kernels": kernels, "source_page": regions}, open(out, "w"), indent=1)
print("wrote", out)
