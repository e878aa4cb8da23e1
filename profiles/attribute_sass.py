"""Attributes the executed warp-instructions of one ncu capture to device functions and tree nodes.

    python profiles/attribute_sass.py <report.ncu-rep> <workload> [--math strict] [--top 25]

ncu cannot show the CUDA source of an NVRTC kernel (the file names are virtual), but the cubin carries -lineinfo.
This script (1) re-creates the workload's cubin locally through the product's own lowering + NVRTC path (no GPU
needed; it is byte-identical to what ran because the in-tree cubin cache travels with the snapshot), (2) reads the
per-instruction line table and inline chains with `nvdisasm -gi`, (3) joins them by instruction offset with the
per-instruction "Instructions Executed" column of the report, and (4) sums per innermost device function, per
outermost generated node function, and per opcode/pipe.
"""
import argparse
import collections
import csv
import glob
import io
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PIPE = {  # sm_100 issue pipes of the opcodes these kernels use (ncu pipe metrics: alu / fma / xu / lsu)
    "FFMA": "fma", "FMUL": "fma", "FADD": "fma", "IMAD": "fma", "HFMA2": "fma",
    "LOP3": "alu", "SHF": "alu", "ISETP": "alu", "FSETP": "alu", "FSEL": "alu", "SEL": "alu", "FMNMX": "alu", "FMNMX3": "alu", "IADD3": "alu",
    "VIADD": "alu", "I2FP": "alu", "MOV": "alu", "LEA": "alu", "PLOP3": "alu", "IABS": "alu", "FCHK": "alu", "PRMT": "alu", "VIMNMX": "alu",
    "MUFU": "xu", "F2I": "xu", "FRND": "xu", "I2F": "xu", "F2F": "xu", "F2FP": "xu",
    "LDS": "lsu", "STS": "lsu", "LDG": "lsu", "STG": "lsu", "ATOMS": "lsu", "RED": "lsu", "ATOMG": "lsu",
}


PRIMITIVES = {"fma_", "fnma_", "fms_", "fxor", "shl", "lsr", "imul", "iadd", "isub", "cvt_f", "cvt_i", "floor_", "round_", "abs_", "min_", "max_",
              "lerp", "hermite", "quintic", "hash_finish", "sel_hi", "pow4", "inv_sqrt", "reciprocal", "fast_div", "bits", "from_bits", "and_f", "or_f"}


def build_cubin(workload, math):
    import bench
    cache = tempfile.mkdtemp(prefix="fn2_attr_")
    os.environ["FASTNOISE2_B200_CACHE"] = cache
    # fnb200Prepare compiles every points-per-thread form the launcher can choose; large calls (what ncu captured) run the 4-point one
    os.environ["FASTNOISE2_B200_PREPARE_ONLY"] = "4"
    import fastnoise2_b200 as fn
    from fastnoise2_b200 import api
    fn.set_math_mode(bench.MATH_MODES[math])
    w = bench.WORKLOADS[workload]
    kind = {"grid2d": api.KIND_GRID2D, "grid3d": api.KIND_GRID3D, "grid4d": api.KIND_GRID4D, "tileable2d": api.KIND_TILEABLE2D,
            "pos2d": api.KIND_POS2D, "pos3d": api.KIND_POS3D, "pos4d": api.KIND_POS4D}[w["kind"]]
    g = fn.Generator(fn.TREES[w["tree"]])
    g.prepare(kind)
    source = g.kernel_source(kind)
    cubins = glob.glob(os.path.join(cache, "*.cubin"))
    assert len(cubins) == 1, cubins
    return cubins[0], source, cache


def function_table(text):
    """line number -> name of the function whose definition most recently started at or before it."""
    starts = []
    pattern = re.compile(r"^\s*(?:template<[^>]*>\s*)?(?:static\s+)?(?:FN2_DI|__device__|extern \"C\" __global__)(?:\s+__launch_bounds__\([^)]*\))?[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    for i, line in enumerate(text.split("\n"), 1):
        m = pattern.match(line)
        if m and not line.rstrip().endswith(";"):
            starts.append((i, m.group(1)))
    return starts


def lookup(starts, line):
    name = "?"
    for start, fn_name in starts:
        if start > line:
            break
        name = fn_name
    return name


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("workload")
    ap.add_argument("--math", default="strict")
    ap.add_argument("--top", type=int, default=25)
    args = ap.parse_args()

    cubin, tree_source, cache = build_cubin(args.workload, args.math)
    tables = {"fn2_tree.cu": function_table(tree_source)}
    for path in glob.glob(os.path.join(ROOT, "fastnoise2_b200", "csrc", "device", "*.cuh")):
        tables[os.path.basename(path)] = function_table(open(path).read())

    # offset -> (innermost function, outermost tree function)
    where = {}
    chain = []
    fresh = True
    for line in subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.split("\n"):
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            if fresh:
                chain = []
                fresh = False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", line)
        if m:
            fresh = True
            inner, outer = "?", "?"
            for f, ln in chain:  # innermost frame first; skip one-line arithmetic helpers so the cost lands on their caller
                name = lookup(tables.get(f, []), ln)
                if name not in PRIMITIVES:
                    inner = name
                    break
            for f, ln in chain:  # the innermost frame that lies in the generated tree source = the node function
                if f == "fn2_tree.cu":
                    outer = lookup(tables["fn2_tree.cu"], ln)
                    break
            where[int(m.group(1), 16)] = (inner, outer)
    shutil.rmtree(cache, ignore_errors=True)

    text = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text[text.index('"Address"'):])))
    header = rows[0]
    ia, isrc, iexec = header.index("Address"), header.index("Source"), header.index("Instructions Executed")
    base = None
    by_inner, by_outer, by_op, by_pipe = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    total = 0
    warps = None
    mismatched = 0
    for r in rows[1:]:
        if len(r) <= iexec or not r[ia].startswith("0x"):
            continue
        addr = int(r[ia], 16)
        base = addr if base is None else base
        n = int(r[iexec])
        warps = n if warps is None else warps  # the first instruction runs once per warp
        src = r[isrc].strip()
        if src.startswith("@"):
            src = src.split(" ", 1)[1].strip()
        op = src.split()[0].split(".")[0].rstrip(";")
        inner, outer = where.get(addr - base, ("?", "?"))
        if (addr - base) not in where:
            mismatched += 1
        by_inner[inner] += n
        by_outer[outer] += n
        by_op[op] += n
        by_pipe[PIPE.get(op, "other")] += n
        total += n
    import bench
    # printed numbers are warp-instructions x 32 per point of the workload = thread-instructions per point with full warps
    # (a thread walks several 4-point groups, so the warp count alone does not give the point count)
    import bench
    points = 1
    for extent in bench.WORKLOADS[args.workload]["n"]:
        points *= extent
    per_point = lambda n: n * 32.0 / points
    print("# %s (%s math): %.1f warp-instructions per point, %d warps, %d unmatched offsets" % (args.workload, args.math, per_point(total), warps, mismatched))
    print("\n## by pipe (instructions per point, share)")
    for k, n in by_pipe.most_common():
        print("  %-6s %8.1f  %5.1f %%" % (k, per_point(n), 100.0 * n / total))
    print("\n## by opcode")
    for k, n in by_op.most_common(args.top):
        print("  %-8s %8.1f  %5.1f %%" % (k, per_point(n), 100.0 * n / total))
    print("\n## by innermost device function")
    for k, n in by_inner.most_common(args.top):
        print("  %-28s %8.1f  %5.1f %%" % (k, per_point(n), 100.0 * n / total))
    print("\n## by generated node function (outermost frame in the tree source)")
    for k, n in by_outer.most_common(args.top):
        print("  %-28s %8.1f  %5.1f %%" % (k, per_point(n), 100.0 * n / total))


if __name__ == "__main__":
    main()
