"""Offline instruction budget of a generated kernel (no GPU needed).

    python tools/sass_count.py c4_terrain --kind 3 [--ppt 4] [--env FASTNOISE2_B200_PACK=1] [--top 30]

Lowers the tree through the product's own path, compiles the text with NVRTC into a scratch cubin cache, disassembles
fn2_entry with cuobjdump and prints registers, code size and the static opcode histogram. The generated kernels are
straight-line code apart from the octave loops; `--unroll` compiles with FN2_OCTAVE_LOOP = unroll so that the static count IS
the executed count per thread (divide by points per thread) -- the number ncu reports as thread-instructions per point.
"""
import argparse
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tree")
    ap.add_argument("--kind", type=int, default=3)
    ap.add_argument("--ppt", type=int, default=0)
    ap.add_argument("--env", action="append", default=[])
    ap.add_argument("--unroll", action="store_true")
    ap.add_argument("--math", type=int, default=0)
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--keep", default="")
    args = ap.parse_args()

    cache = tempfile.mkdtemp(prefix="fn2_sass_")
    os.environ["FASTNOISE2_B200_CACHE"] = cache
    for kv in args.env:
        k, v = kv.split("=", 1)
        os.environ[k] = v
    if args.unroll:
        os.environ["FASTNOISE2_B200_NVRTC_FLAGS"] = (os.environ.get("FASTNOISE2_B200_NVRTC_FLAGS", "") + ' -DFN2_OCTAVE_LOOP=_Pragma("unroll")').strip()
    if args.ppt:
        os.environ["FASTNOISE2_B200_PPT"] = str(args.ppt)
    import fastnoise2_b200 as fn
    fn.set_math_mode(args.math)
    text = fn.TREES.get(args.tree, args.tree)
    g = fn.Generator(text)
    g.prepare(args.kind)  # compiles the 4- and the 1-point-per-thread variant (only the latter with --ppt 1)
    cubins = sorted(glob.glob(os.path.join(cache, "*.cubin")), key=os.path.getmtime)
    assert cubins, fn.last_error()
    for cubin, ppt in zip(cubins, [4, 1] if len(cubins) == 2 else [args.ppt or 4]):
        open(cubin + ".txt", "w").write("%d point(s) per thread" % ppt)
        if args.ppt and ppt != args.ppt:
            continue
        report(cubin, args)
        if args.keep:
            subprocess.check_call(["cp", cubin, args.keep])


def report(cubin, args):
    res = subprocess.run(["cuobjdump", "-res-usage", cubin], capture_output=True, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    regs = re.search(r"REG:(\d+)", res)
    ops = collections.Counter()
    n = 0
    for line in sass.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1).split(".")[0]
        if op == "NOP":
            continue
        ops[op] += 1
        n += 1
    ppt = int(re.search(r"(\d+) point\(s\) per thread", open(cubin + ".txt").read()).group(1)) if os.path.exists(cubin + ".txt") else (args.ppt or 4)
    # the persistent grid-stride loop: the LAST backward branch of fn2_entry; with the octave loops unrolled its body is what a
    # thread executes per group of `ppt` points (divide slow paths, which are CALLed, are outside and almost never run)
    rows = []
    for line in sass.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*)", line)
        if m and m.group(2).split(".")[0] != "NOP":
            rows.append((int(m.group(1), 16), m.group(2).split(".")[0], m.group(3)))
    loops = []
    for addr, op, rest in rows:
        t = re.search(r"BRA(?:\.U)?\s.*?(0x[0-9a-f]+)", op + rest)
        if op == "BRA" and t and int(t.group(1), 16) < addr:
            loops.append((int(t.group(1), 16), addr))
    print("%s: ppt %d, regs %s, %d instructions (%.1f KB), %d backward branches" % (os.path.basename(cubin), ppt, regs.group(1) if regs else "?", n, n * 16 / 1024.0, len(loops)))
    body = ops
    if loops:
        lo, hi = min(l[0] for l in loops), max(l[1] for l in loops)
        body = collections.Counter(op for addr, op, _ in rows if lo <= addr <= hi)
        nb = sum(body.values())
        print("  outermost loop body: %d instructions = %.1f per point; outside the loop: %d" % (nb, nb / ppt, n - nb))
    for op, c in body.most_common(args.top):
        print("  %-10s %6d  %7.1f /pt" % (op, c, c / ppt))
    print("  %-10s %6d  %7.1f /pt" % ("total", sum(body.values()), sum(body.values()) / ppt))


if __name__ == "__main__":
    main()
