"""Per-node throughput harness: the B200 counterpart of the reference's tests/FastNoiseBenchmark.cpp.

Like the reference harness (FastNoiseBenchmark.cpp:14-46, :137-175) it walks the reflection tables, builds every node type
with `Constant` sources (falling back to the first node type a lookup accepts, its own lookups fed by `Constant`), and
times GenPositionArray{2,3,4}D over positions `idx * 0.6` with a seed that increments every iteration. Names follow the
reference's `<D>D/<feature set>/<group>/<node>` scheme so the two tables line up.

    python tools/node_benchmark.py                       # product on cuda:0, 4 Mi positions per call (throughput)
    python tools/node_benchmark.py --positions 8192      # the reference harness's own size (launch-latency bound on a GPU)
    python tools/node_benchmark.py --impl reference      # the same loop on oracle/_ref (CPU, one thread, like the reference harness)
    python tools/node_benchmark.py --filter Cellular --dims 3

Positions and outputs stay on the device for the product arm (the reference harness likewise keeps them in host memory
for the CPU): this measures the fused tree kernel, not PCIe. Times are CUDA-event times on the launching stream.
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_generator(lib, node_id, handles):
    """FastNoiseBenchmark.cpp:14-46 through the C ABI. Returns a node handle or None."""
    def new(i):
        h = lib.fnNewFromMetadata(i, 0xFFFFFFFF)
        handles.append(h)
        return h

    generator = new(node_id)
    constant = new(0)
    for lookup in range(lib.fnGetMetadataNodeLookupCount(node_id)):
        if lib.fnSetNodeLookup(generator, lookup, constant):
            continue
        for try_id in range(lib.fnGetMetadataCount()):  # a warp-fractal wants a DomainWarp source: try every type in order
            candidate = new(try_id)
            if lib.fnSetNodeLookup(generator, lookup, candidate):
                for inner in range(lib.fnGetMetadataNodeLookupCount(try_id)):
                    if not lib.fnSetNodeLookup(candidate, inner, constant):
                        return None
                break
        else:
            return None
    return generator


# the "Node Trees" group of the reference harness (FastNoiseBenchmark.cpp:152-166): the two trees shipped with its Node Editor
# (tools/NodeEditor/util/DemoNodeTrees.inl:5-6, written by the reference's own encoder)
NODE_TREES = [
    ("Mountain Terrain", "E@BBZEG@BD8JFgIECArXIzwECiQIw/UoPwkuAAE@BJDQAH@BC@AIEAJBw@ABZEED0KV78YZmZmPwQDmpkZPwsAAIA/HAMAAHBCBA=="),
    ("Cellular Caves", "FgIcCS4AAQ@BklCQs@BlRBDNzMw9G@AIMAgAw@ADgC@BCiQIzczMPgkJ@BPkIQH4XrPhjNzEw/DBIkCM3MzD4JCQ@ADBCCAE@BQzczMvhg@B/JAL/BAAL7FE4PgQKFwkNCQg@CQQQDuB4FPwt7FC4/BAOPwnU8DA=="),
]


def subjects(lib, name_filter):
    """(group, name, builder) for every node type, then the Node Trees group -- the registration order of the reference harness."""
    for node_id in range(lib.fnGetMetadataCount()):
        name = lib.fnGetMetadataName(node_id).decode()
        if name_filter and name_filter.lower() not in name.lower():
            continue
        yield group_name(lib, node_id), name, (lambda handles, i=node_id: build_generator(lib, i, handles))
    for name, text in NODE_TREES:
        if name_filter and name_filter.lower() not in name.lower() and name_filter.lower() != "node trees":
            continue

        def decode(handles, text=text):
            h = lib.fnNewFromEncodedNodeTree(text.encode(), 0xFFFFFFFF)
            if h:
                handles.append(h)
            return h or None
        yield "Node Trees", name, decode


def group_name(lib, node_id):
    count = lib.fnGetMetadataGroupCount(node_id)
    return lib.fnGetMetadataGroupName(node_id, count - 1).decode() if count else "Misc"


def bench_product(args):
    import torch
    import fastnoise2_b200 as fn
    if not fn.device_available():
        raise SystemExit("no CUDA device / NVRTC: " + fn.last_error())
    lib = fn.lib
    n = args.positions
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        pos = torch.arange(n, dtype=torch.float32, device="cuda") * 0.6
        out = torch.empty(n, dtype=torch.float32, device="cuda")
    stream.synchronize()
    sp = ctypes.c_void_p(stream.cuda_stream)
    rows = []
    for group, name, build in subjects(lib, args.filter):
        handles = []
        gen = build(handles)
        for dim in args.dims:
            label = "%dD/B200_SM100A/%s/%s" % (dim, group, name)
            if gen is None:
                rows.append({"name": label, "error": "could not set valid sources"})
                continue
            ptrs = [pos.data_ptr()] * dim + [None] * (4 - dim)
            seed = 0

            def call(seed):
                rc = lib.fnb200GenPositionArray(gen, dim, out.data_ptr(), n, ptrs[0], ptrs[1], ptrs[2], ptrs[3], 0.0, 0.0, 0.0, 0.0, seed, None, sp)
                if rc != 0:
                    raise RuntimeError(label + ": " + fn.last_error())

            for _ in range(3):
                call(seed)
                seed += 1
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            stream.synchronize()
            start.record(stream)
            for _ in range(args.iterations):
                call(seed)
                seed += 1
            stop.record(stream)
            stop.synchronize()
            ms = start.elapsed_time(stop) / args.iterations
            rows.append({"name": label, "us_per_call": ms * 1e3, "items_per_second": n / (ms * 1e-3)})
            print("%-64s %10.2f us  %10.1f M items/s" % (label, ms * 1e3, n / (ms * 1e-3) / 1e6), flush=True)
        for h in handles:
            lib.fnDeleteNodeRef(h)
    return rows


def bench_reference(args):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests import helpers
    variant = "relaxed_avx512" if helpers.load_ref("relaxed_avx512") is not None else "relaxed_avx2"
    lib = helpers.load_ref(variant)
    if lib is None:
        raise SystemExit("oracle/_ref is not built (make -C oracle)")
    n = args.positions
    pos = (np.arange(n, dtype=np.float32) * np.float32(0.6)).astype(np.float32)
    out = np.empty(n, np.float32)
    p = pos.ctypes.data
    rows = []
    for group, name, build in subjects(lib, args.filter):
        handles = []
        gen = build(handles)
        for dim in args.dims:
            label = "%dD/%s/%s/%s" % (dim, variant, group, name)
            if gen is None:
                rows.append({"name": label, "error": "could not set valid sources"})
                continue
            call = {2: lambda s: lib.fnGenPositionArray2D(gen, out.ctypes.data, n, p, p, 0.0, 0.0, s, None),
                    3: lambda s: lib.fnGenPositionArray3D(gen, out.ctypes.data, n, p, p, p, 0.0, 0.0, 0.0, s, None),
                    4: lambda s: lib.fnGenPositionArray4D(gen, out.ctypes.data, n, p, p, p, p, 0.0, 0.0, 0.0, 0.0, s, None)}[dim]
            call(0)
            iterations, seed, t0 = 0, 1, time.perf_counter()
            while iterations < args.iterations and time.perf_counter() - t0 < args.cpu_seconds:
                call(seed)
                seed += 1
                iterations += 1
            dt = (time.perf_counter() - t0) / iterations
            rows.append({"name": label, "us_per_call": dt * 1e6, "items_per_second": n / dt})
            print("%-64s %10.2f us  %10.1f M items/s" % (label, dt * 1e6, n / dt / 1e6), flush=True)
        for h in handles:
            lib.fnDeleteNodeRef(h)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--positions", type=int, default=4 << 20)
    ap.add_argument("--iterations", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=0.5, help="reference arm: stop a node after this much time")
    ap.add_argument("--dims", type=int, nargs="+", default=[2, 3, 4])
    ap.add_argument("--filter", default="")
    ap.add_argument("--json", default="", help="also write the rows to this file")
    args = ap.parse_args()
    rows = bench_product(args) if args.impl == "b200" else bench_reference(args)
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"impl": args.impl, "positions": args.positions, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
