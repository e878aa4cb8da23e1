"""Device-resident A/B of kernel build / launch knobs in ONE process (no per-variant interpreter start-up).

    python tools/ab_sweep.py --workloads c4,c2a --variants "base:;iters8:FASTNOISE2_B200_ITERS=8" [--steps 5] [--prepare]

Every variant is a set of environment settings (read by the lowering at program-build time and by the launcher at launch
time); each (workload, variant) gets a fresh Generator, so the text is lowered and compiled under those settings. Times are
CUDA events on the launching stream around `steps` launches after 2 warm-up launches, exactly as bench.py's `value`.
--prepare only compiles every variant into the in-tree cubin cache (no GPU needed), so the GPU run does not pay NVRTC.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse_variants(text):
    out = []
    for item in text.split(";"):
        item = item.strip()
        if not item:
            continue
        name, _, settings = item.partition(":")
        env = {}
        for kv in settings.split(","):
            kv = kv.strip()
            if kv:
                k, _, v = kv.partition("=")
                env[k] = v
        out.append((name, env))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c4")
    ap.add_argument("--variants", required=True)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--prepare", action="store_true")
    args = ap.parse_args()
    variants = parse_variants(args.variants)
    knobs = sorted({k for _, env in variants for k in env})

    import bench
    import fastnoise2_b200 as fn
    from fastnoise2_b200 import api
    kinds = {"grid2d": api.KIND_GRID2D, "grid3d": api.KIND_GRID3D, "grid4d": api.KIND_GRID4D, "tileable2d": api.KIND_TILEABLE2D,
             "pos2d": api.KIND_POS2D, "pos3d": api.KIND_POS3D, "pos4d": api.KIND_POS4D}
    if not args.prepare:
        import torch
        stream = torch.cuda.Stream()
        sptr = stream.cuda_stream

    class Args:
        pass

    for wl in args.workloads.split(","):
        w = bench.WORKLOADS[wl]
        for name, env in variants:
            for k in knobs:
                os.environ.pop(k, None)
            os.environ.update(env)
            gen = fn.Generator(fn.TREES[w["tree"]])
            if args.prepare:
                gen.prepare(kinds[w["kind"]])
                print(wl, name, "prepared", flush=True)
                continue
            work = bench.Work(Args(), w, gen, fn.lib, torch, 0, 1, "strong", sptr, None)
            with torch.cuda.stream(stream):
                for _ in range(2):
                    work.step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(args.steps):
                    work.step()
                e1.record(stream)
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            mm = work.read_minmax()
            print("%s %-28s %9.0f Mpts/s  %8.3f ms  mm %r" % (wl, name + ":", work.total_points / ms / 1e3, ms, mm), flush=True)
            del work
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
