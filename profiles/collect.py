"""Turns the files one `profiles/capture.sh <tag>` run left under gpurun_out/ into the tracked evidence under profiles/.

    python profiles/collect.py r2z

Nothing is edited: bench lines, probe outputs and test summaries are copied; ncu reports are summarised (summarize_ncu.py),
attributed to device functions (attribute_sass.py) and checked against the SASS the committed sources build (tools/sass_identity.py).
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
HERE = os.path.join(ROOT, "profiles")
WORKLOADS = ["c4", "c1", "c2a", "c2b", "c3a", "c3b", "c5"]


def last_line(src, dst):
    lines = [l for l in open(src).read().splitlines() if l.strip()]
    open(dst, "w").write(lines[-1] + "\n")


def main():
    tag = sys.argv[1]
    py = sys.executable
    reports = ["%s/prof_%s_%s.ncu-rep:%s" % (OUT, w, tag, w) for w in WORKLOADS]
    subprocess.check_call([py, os.path.join(HERE, "summarize_ncu.py"), "r2"] + reports, cwd=ROOT)
    identity = []
    for w in WORKLOADS:
        report = "%s/prof_%s_%s.ncu-rep" % (OUT, w, tag)
        text = subprocess.run([py, os.path.join(HERE, "attribute_sass.py"), report, w, "--top", "25"], capture_output=True, text=True, cwd=ROOT).stdout
        open(os.path.join(HERE, "r2_attribution_%s.txt" % w), "w").write(text[text.index("# "):] if "# " in text else text)
        text = subprocess.run([py, os.path.join(ROOT, "tools", "sass_identity.py"), report, w], capture_output=True, text=True, cwd=ROOT).stdout
        identity += [l for l in text.splitlines() if l.startswith(w + ":") or l.startswith("   #")]
    open(os.path.join(HERE, "r2_sass_identity.txt"), "w").write(
        "# tools/sass_identity.py: the fn2_entry SASS stored in each ncu report of this capture against the SASS the committed sources build\n"
        + "\n".join(identity) + "\n")
    os.makedirs(os.path.join(HERE, "r2_bench_lines"), exist_ok=True)
    for w in WORKLOADS:
        last_line("%s/bench_%s_%s.json" % (OUT, w, tag), os.path.join(HERE, "r2_bench_lines", w + ".json"))
    last_line("%s/bench_reference_c4_%s.json" % (OUT, tag), os.path.join(HERE, "r2_bench_lines", "reference_c4.json"))
    last_line("%s/bench_default_%s.json" % (OUT, tag), os.path.join(HERE, "r2_bench_lines", "default_no_flags.json"))
    launches = [l for l in open("%s/launches_c4_%s.csv" % (OUT, tag)).read().splitlines() if not l.startswith("==")]
    open(os.path.join(HERE, "r2_launches_bench_c4.csv"), "w").write("\n".join(launches) + "\n")
    for src, dst in (("gputest", "r2_gputest_summary.txt"), ("alu_probe", "r2_alu_probe.txt"), ("call_latency", "r2_call_latency.txt"),
                     ("odd_width", "r2_odd_width.txt"), ("node_benchmark", "r2_node_benchmark_b200.txt"),
                     ("node_benchmark_trees", "r2_node_benchmark_trees_b200.txt"), ("smoke", "r2_smoke.txt"), ("d2h_bandwidth", "r2_d2h_bandwidth_1gpu.txt")):
        shutil.copyfile("%s/%s_%s.txt" % (OUT, src, tag), os.path.join(HERE, dst))
    shutil.copyfile(os.path.join(OUT, "fullsize_parity.txt"), os.path.join(HERE, "r2_fullsize_parity.txt"))  # written by tests/test_gpu_fullsize.py
    # (r2_pageable_probe.txt keeps the thread-count x streaming-store sweep; the final library's own line goes beside it)
    shutil.copyfile("%s/pageable_probe_%s.txt" % (OUT, tag), os.path.join(HERE, "r2_pageable_probe_final.txt"))
    print(open(os.path.join(HERE, "r2_ncu_summary.md")).read())
    print(open(os.path.join(HERE, "r2_sass_identity.txt")).read())


if __name__ == "__main__":
    main()
