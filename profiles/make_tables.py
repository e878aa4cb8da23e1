"""Prints the result tables of DESIGN.md section 7 and BASELINE.md section 4 from the evidence files under profiles/
(r2_bench_lines/*.json, r2_ncu_summary.json, r2_scaling/*.json), so the documents quote the committed lines, not retyped numbers.

    python profiles/make_tables.py
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = [("c1", "C1 Simplex 256³"), ("c2a", "C2a FBm5∘Perlin 16384²"), ("c2b", "C2b tileable 4096²"), ("c3a", "C3a CellularDistance 512³"),
         ("c3b", "C3b CellularLookup(White) 512³"), ("c4", "**C4 terrain 1024³**"), ("c5", "C5 SuperSimplex 4D 256 Mi positions")]


def sp(value):
    text = str(int(round(value)))
    out = ""
    while len(text) > 3:
        out = " " + text[-3:] + out
        text = text[:-3]
    return text + out


def line(name):
    return json.loads(open(os.path.join(HERE, "r2_bench_lines", name + ".json")).read())


def main():
    summary = json.load(open(os.path.join(HERE, "r2_ncu_summary.json")))
    print("| config | GPU Mpoints/s (device) | step time | e2e Mpoints/s (pinned host buffer) | thread-inst/point (ncu; of which packed) | dispatch roofline frac | issue slots busy (ncu) | regs / warps active | CPU Mpoints/s (threads) |")
    print("|---|---|---|---|---|---|---|---|---|")
    for key, name in NAMES:
        d, s = line(key), summary[key]
        r = d["roofline"]
        e2e = sp(d["e2e"]["value"]) + (" (pageable destination: %s)" % sp(d["e2e_pageable"]["value"]) if "e2e_pageable" in d else "")
        print("| %s | %s | %.3f ms | %s | %s (%s) | %.3f | %.1f %% | %d / %.0f %% | %s (%d) |" % (
            name, sp(d["value"]), d["ms_per_step"], e2e, "%.1f" % s["thread_inst_per_point"], "%.1f" % s.get("packed_thread_inst_per_point", 0.0),
            r["frac"], s["smsp__issue_active.avg.pct_of_peak_sustained_active"], s["launch__registers_per_thread"],
            s["sm__warps_active.avg.pct_of_peak_sustained_active"], sp(d["cpu_baseline"]["value"]), d["cpu_baseline"]["cores"]))
    ref = line("reference_c4")
    print("\n`--impl reference` (C4, %d threads): %s Mpoints/s, %.0f ms per 1024³ step" % (ref["cpu_baseline"]["cores"], sp(ref["value"]), ref["ms_per_step"]))
    print("\n| N | strong: device (`value`) | ms/step | efficiency vs N=1 | weak: device | e2e pinned (N processes) | e2e pageable | generate / + NCCL gather / fused generate+gather (ms) |")
    print("|---|---|---|---|---|---|---|---|")
    one = line("c4")
    print("| 1 | %s | %.2f | — | %s | %s | %s | — |" % (sp(one["value"]), one["ms_per_step"], sp(one["value"]), sp(one["e2e"]["value"]), sp(one["e2e_pageable"]["value"])))
    for n in (2, 4, 8):
        d = json.loads(open(os.path.join(HERE, "r2_scaling", "n%d.json" % n)).read())
        g = d["gather"]
        print("| %d | %s | %.2f | %.3f | %s | %s | %s | %.2f / +%.2f / %.2f |" % (
            n, sp(d["value"]), d["ms_per_step"], d["value"] / (n * one["value"]), sp(d["weak"]["value"]), sp(d["e2e"]["value"]), sp(d["e2e_pageable"]["value"]),
            g["generate_ms"], g["gather_ms"], g["fused_generate_and_gather_ms"]))
    fan = json.loads(open(os.path.join(HERE, "r2_scaling", "n1_one_process_fanout_8_visible.json")).read())
    print("\none process, one call, %d GPUs visible: e2e %s Mpoints/s (one device: %s)" % (fan["e2e_fanout"]["devices"], sp(fan["e2e_fanout"]["value"]), sp(fan["e2e"]["value"])))


if __name__ == "__main__":
    main()
