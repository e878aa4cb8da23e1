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


if __name__ == "__main__":
    main()
