"""Bare device->host copy microbenchmark: what the host side of this box can absorb from 1, 2, 4, 8 GPUs at once.

    python tools/d2h_bandwidth.py [--mib 1024] [--repeats 5]

One process; for every GPU a stream and a `--mib` MiB device buffer, copied with cudaMemcpyAsync (torch's non_blocking copy_)
into its own pinned host buffer -- nothing else runs. The reference-shaped fnGen* calls end in exactly these copies, so this
is the ceiling of their end-to-end rate (bench.py `e2e`, `e2e_fanout`): 4 bytes per point. Also times the same copy into
PAGEABLE memory (what cudaMemcpyAsync degrades to without the library's pinned bounce buffers).
"""
import argparse
import json
import time

import torch


def measure(devices, mib, repeats, pinned=True):
    n = mib * (1 << 20) // 4
    src, dst, streams = [], [], []
    for d in devices:
        with torch.cuda.device(d):
            src.append(torch.empty(n, dtype=torch.float32, device="cuda:%d" % d).fill_(1.0))
            dst.append(torch.empty(n, dtype=torch.float32, pin_memory=pinned))
            if not pinned:
                dst[-1].zero_()
            streams.append(torch.cuda.Stream(device=d))
    best = 0.0
    for _ in range(repeats + 1):
        for d in devices:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for i, d in enumerate(devices):
            with torch.cuda.stream(streams[i]):
                dst[i].copy_(src[i], non_blocking=True)
        for s in streams:
            s.synchronize()
        dt = time.perf_counter() - t0
        best = max(best, len(devices) * n * 4 / dt / 1e9)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--repeats", type=int, default=5)
    args = ap.parse_args()
    count = torch.cuda.device_count()
    rows = []
    for k in (1, 2, 4, 8):
        if k > count:
            break
        devices = list(range(k))
        row = {"gpus": k, "pinned_GBps": measure(devices, args.mib, args.repeats, True), "mib_per_gpu": args.mib}
        if k == 1:
            row["pageable_GBps"] = measure(devices, min(args.mib, 512), 2, False)
        row["points_per_second_ceiling_M"] = row["pinned_GBps"] * 1e9 / 4 / 1e6
        rows.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
