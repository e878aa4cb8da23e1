"""Per-call latency of the reference-shaped entry points at the chunk sizes FastNoise2 consumers typically use (16^3 .. 128^3).

    python tools/call_latency.py [--tree c4_terrain] [--calls 300]

Host rows: fnGenUniformGrid3D with a pageable numpy output (what an unmodified consumer passes) and with a pinned one;
device row: the same call writing into a device buffer (no copy, returns after the kernel is enqueued + min/max read back).
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tree", default="c4_terrain")
    ap.add_argument("--calls", type=int, default=300)
    args = ap.parse_args()
    import numpy as np
    import torch
    import fastnoise2_b200 as fn
    assert fn.device_available(), fn.last_error()
    g = fn.Generator(fn.TREES[args.tree])
    print("tree %s" % args.tree)
    for n in (16, 32, 64, 128):
        pts = n ** 3
        host = np.empty(pts, np.float32)
        pinned = torch.empty(pts, dtype=torch.float32).pin_memory()
        dev = torch.empty(pts, dtype=torch.float32, device="cuda")
        row = []
        for label, buf in (("pageable", host), ("pinned", pinned), ("device", dev)):
            for _ in range(5):
                g.grid3d(buf, 0.0, 0.0, 0.0, n, n, n, 1.0, 1.0, 1.0, 1337)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(args.calls):
                g.grid3d(buf, float(i), 0.0, 0.0, n, n, n, 1.0, 1.0, 1.0, 1337)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / args.calls
            row.append("%s %8.1f us (%7.1f Mpts/s)" % (label, dt * 1e6, pts / dt / 1e6))
        print("%4d^3  " % n + "   ".join(row), flush=True)

    # A float edit between calls (the Node Editor's slider pattern, tools/NodeEditor/NoiseTexture.cpp:515-535, preview 224 x 224):
    # float parameters are kernel arguments, so the edit costs a re-lowering of the text (same text -> same loaded kernel), no compile.
    lib = fn.lib
    root_id = lib.fnGetMetadataID(g.handle)
    floats = [i for i in range(lib.fnGetMetadataVariableCount(root_id)) if lib.fnGetMetadataVariableType(root_id, i) == 0]
    n = 224
    pinned = torch.empty(n * n, dtype=torch.float32).pin_memory()
    for _ in range(5):
        g.grid2d(pinned, 0.0, 0.0, n, n, 1.0, 1.0, 1337)

    def loop(edit):
        t0 = time.perf_counter()
        for i in range(args.calls):
            if edit and floats:
                assert lib.fnSetVariableFloat(g.handle, floats[0], 2.0 + (i % 7) * 0.015625)
            g.grid2d(pinned, 0.0, 0.0, n, n, 1.0, 1.0, 1337)
        return (time.perf_counter() - t0) / args.calls
    same, edited = loop(False), loop(True)
    print("224 x 224 preview, pinned output: %8.1f us per call with unchanged parameters, %8.1f us with one float (%s) changed before every call: +%.1f us for the edit (no recompile)" % (
        same * 1e6, edited * 1e6, lib.fnGetMetadataVariableName(root_id, floats[0]).decode() if floats else "-", (edited - same) * 1e6))


if __name__ == "__main__":
    main()
