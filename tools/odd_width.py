"""Device-resident throughput on grids whose row width 4 does not divide: four points per thread over any width (forced,
FASTNOISE2_B200_RAGGED=2), the library's own choice (1: that form unless the kernel is above 48 KB of code), one point per
thread (0, what such grids ran before), and the nearest multiple-of-4 width.

    python tools/odd_width.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import fastnoise2_b200 as fn
    lib = fn.lib
    stream = torch.cuda.Stream()
    sp = stream.cuda_stream
    cases = [("c4_terrain", 3, (1030, 1030, 250)), ("c4_terrain", 3, (130, 130, 130)), ("c1_simplex", 3, (1030, 1030, 250)),
             ("c3a_cellular_distance", 3, (1030, 1030, 250)), ("c2_fbm5_perlin", 2, (16390, 16390))]
    for tree, dims, n in cases:
        g = fn.Generator(fn.TREES[tree])
        for label, shape, ragged in (("any-width x4", n, "2"), ("library's choice", n, "1"), ("1 per thread", n, "0"), ("width % 4 == 0", (n[0] - n[0] % 4,) + n[1:], "1")):
            os.environ["FASTNOISE2_B200_RAGGED"] = ragged
            points = 1
            for e in shape:
                points *= e
            out = torch.empty(points, dtype=torch.float32, device="cuda")
            mm = torch.zeros(2, dtype=torch.int32, device="cuda")

            def call():
                if dims == 3:
                    rc = lib.fnb200GenUniformGrid3D(g.handle, out.data_ptr(), 0.0, 0.0, 0.0, shape[0], shape[1], shape[2], 1.0, 1.0, 1.0, 1337, 0, shape[2], mm.data_ptr(), sp)
                else:
                    rc = lib.fnb200GenUniformGrid2D(g.handle, out.data_ptr(), 0.0, 0.0, shape[0], shape[1], 0.25, 0.25, 1337, 0, shape[1], mm.data_ptr(), sp)
                assert rc == 0, fn.last_error()
            with torch.cuda.stream(stream):
                for _ in range(2):
                    call()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5 if points > 1 << 24 else 50
                e0.record(stream)
                for _ in range(reps):
                    call()
                e1.record(stream)
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print("%-22s %-18s %-16s %9.0f Mpts/s  %8.3f ms" % (tree, "x".join(str(e) for e in shape), label, points / ms / 1e3, ms), flush=True)
            del out
    os.environ.pop("FASTNOISE2_B200_RAGGED", None)


if __name__ == "__main__":
    main()
