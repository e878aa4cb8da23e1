"""Extracts the judged numbers from ncu reports brought back in gpurun_out/ into profiles/ (tracked).

    python profiles/summarize_ncu.py r1 gpurun_out/prof_c4.ncu-rep:c4 gpurun_out/prof_c1.ncu-rep:c1[:stale] ...

Writes profiles/<round>_ncu_summary.json (read by bench.py for roofline.traffic) and profiles/<round>_ncu_summary.md.
"""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    # instruction supply: these kernels are up to 55 KB of straight-line code
    "sm__icc_request_hit_rate.pct",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}


def read_report(path):
    text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text[text.index('"ID"'):])))
    header, units, values = rows[0], rows[1], rows[2]
    out = {"kernel": values[header.index("Kernel Name")]}
    for name in METRICS:
        if name in header:
            i = header.index(name)
            v = float(values[i].replace(",", ""))
            out[name] = v * UNIT_SCALE.get(units[i], 1)
    return out


def packed_thread_instructions(path):
    """Thread-instructions of the packed FP32 opcodes (FFMA2 / FMUL2 / FADD2): each does two lane-operations."""
    text = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text[text.index('"Address"'):])))
    header = rows[0]
    isrc, ithread = header.index("Source"), header.index("Thread Instructions Executed")
    total = 0
    for r in rows[1:]:
        if len(r) <= ithread:
            continue
        src = r[isrc].strip()
        if src.startswith("@"):
            src = src.split(" ", 1)[1].strip()
        if src.split()[0].split(".")[0] in ("FFMA2", "FMUL2", "FADD2"):
            total += int(r[ithread])
    return total


def main():
    tag = sys.argv[1]
    here = os.path.dirname(os.path.abspath(__file__))
    summary = {}
    sys.path.insert(0, os.path.dirname(here))
    import bench
    for arg in sys.argv[2:]:
        stale = arg.endswith(":stale")  # the report predates the current kernel source: kept for the pipe / DRAM picture only
        path, workload = (arg[:-len(":stale")] if stale else arg).rsplit(":", 1)
        summary[workload] = read_report(path)
        m = summary[workload]
        m["report"] = os.path.basename(path)
        if stale:
            m["stale"] = "captured before the kernel source last changed for this workload; bench.py does not use its instruction count"
        m["dram_bytes_per_launch"] = m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
        points = 1
        for n in bench.WORKLOADS[workload]["n"]:
            points *= n
        m["points_per_launch"] = points
        m["thread_inst_per_point"] = m["smsp__inst_executed.sum"] * m["smsp__thread_inst_executed_per_inst_executed.ratio"] / points
        m["packed_thread_inst_per_point"] = packed_thread_instructions(path) / points
        m["lane_ops_per_point"] = m["thread_inst_per_point"] + m["packed_thread_inst_per_point"]  # a packed instruction = 2 lane-operations
    json.dump(summary, open(os.path.join(here, "%s_ncu_summary.json" % tag), "w"), indent=1)
    lines = ["# ncu --set full --clock-control none, kernel fn2_entry (one launch per workload), round %s" % tag, "",
             "| workload | time | grid | regs | warps active % | issue slots busy % | alu pipe % | fma pipe % | xu pipe % | lsu pipe % | tensor % | thread-inst/point (of which packed) | ICC hit % | stalled no-instruction / not-selected / math-pipe / dispatch (warps per issue) | DRAM read | DRAM write | local ld/st |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for w, m in summary.items():
        lines.append("| %s%s | %.3f ms | %d | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f (%.1f) | %.1f | %.2f / %.2f / %.2f / %.2f | %.3g B | %.3g B | %d/%d |" % (
            w, " (stale)" if "stale" in m else "", m["gpu__time_duration.sum"] * 1e3, m.get("launch__grid_size", 0), m["launch__registers_per_thread"], m["sm__warps_active.avg.pct_of_peak_sustained_active"],
            m["smsp__issue_active.avg.pct_of_peak_sustained_active"], m["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"],
            m["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"], m["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"],
            m.get("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 0),
            m.get("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", 0), m["thread_inst_per_point"], m["packed_thread_inst_per_point"],
            m.get("sm__icc_request_hit_rate.pct", 0), m.get("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", 0),
            m.get("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", 0), m.get("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 0),
            m.get("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", 0), m.get("dram__bytes_read.sum", 0),
            m.get("dram__bytes_write.sum", 0), m.get("sass__inst_executed_local_loads", 0), m.get("sass__inst_executed_local_stores", 0)))
    if any("stale" in m for m in summary.values()):
        lines += ["", "(stale): captured before the last change of the kernel source that affects this workload; pipe shares and DRAM bytes still describe",
                  "the kernel, the instruction count does not."]
    open(os.path.join(here, "%s_ncu_summary.md" % tag), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
