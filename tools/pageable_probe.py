"""End-to-end rate of one fnGenUniformGrid3D call into PAGEABLE and PINNED host memory (terrain tree, 1024 x 1024 x nz).

    python tools/pageable_probe.py [--nz 256] [--calls 4]

The pageable path is device -> pinned bounce buffer -> worker-thread copy; FASTNOISE2_B200_COPY_THREADS and
FASTNOISE2_B200_STREAM_COPY (non-temporal stores, default on) are read once per process, so A/B runs are separate processes.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nz", type=int, default=256)
    ap.add_argument("--calls", type=int, default=4)
    args = ap.parse_args()
    import torch
    import fastnoise2_b200 as fn
    g = fn.Generator(fn.TREES["c4_terrain"])
    n = (1024, 1024, args.nz)
    points = n[0] * n[1] * n[2]
    out = {"points": points, "copy_threads": os.environ.get("FASTNOISE2_B200_COPY_THREADS", "default"),
           "stream_copy": os.environ.get("FASTNOISE2_B200_STREAM_COPY", "1"), "cpus": os.cpu_count()}
    pageable = np.empty(points, np.float32)
    pageable[:] = 0
    pinned = torch.empty(points, dtype=torch.float32, pin_memory=True)
    for label, buf in (("pageable", pageable), ("pinned", pinned)):
        g.grid3d(buf, 0.0, 0.0, 0.0, *n, 1.0, 1.0, 1.0, 1337)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(args.calls):
            g.grid3d(buf, 0.0, 0.0, 0.0, *n, 1.0, 1.0, 1.0, 1337)
        sec = (time.time() - t0) / args.calls
        out[label + "_Mpoints_s"] = round(points / sec / 1e6)
        out[label + "_GBps"] = round(points * 4 / sec / 1e9, 1)
    assert np.array_equal(pageable.view(np.uint32), pinned.numpy().view(np.uint32))
    out["pinned_over_pageable"] = round(out["pinned_Mpoints_s"] / out["pageable_Mpoints_s"], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
