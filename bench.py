#!/usr/bin/env python
"""Benchmark of the fused node-graph evaluation path (BASELINE.json metric: Mpoints/s of a 3D uniform grid).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c4] [--scaling strong|weak]

One "step" = one pass of the hot path over one batch: the whole grid of the workload generated into HBM.
Default workload = BASELINE.json configs[3], the config the metric is quoted on: the domain-warped terrain tree
(DomainWarpFractalProgressive + FBm + Ridged + MaxSmooth + Remap + Terrace, SURVEY.md 8d "C4"),
GenUniformGrid3D 1024^3, seed 1337, "sharded over 8 GPUs". It fits one GPU (4 GiB of output), so N=1 runs it whole.
With N ranks (one process per GPU) the default is STRONG scaling, as the config names it: the same 1024^3 volume, rank r
generates z-slab r (1024/N planes) through the slab entry point with GLOBAL integer indices, so the N slabs are bit-identical
to the one-GPU volume. The only cross-rank traffic is the OutputMinMax all-reduce: NCCL max over the accumulator's two
unsigned keys, enqueued on the generating stream (fnb200CommAllReduceMinMax) -- inside the timed region, no host round trip.
The same run also times the weak-scaling variant (every rank 1024^3 points of a 1024 x 1024 x 1024N volume) and reports it
under "weak"; --scaling weak makes that the headline instead.

JSON keys: see the task contract. `value` is device-resident throughput (CUDA events around the K steps on the
launching stream, max over ranks). `e2e` goes through the reference-shaped C ABI call fnGenUniformGrid3D with a
pinned HOST output buffer: device->host copies are inside the timed region (all K steps). `roofline` is the instruction-issue
roofline (this path has no tensor work and writes 4 B/point): achieved = thread-instructions the kernel executes per point
(ncu, profiles/; a packed FP32 instruction counts as the two dispatch cycles it takes) x points / kernel time, peak = the theoretical
dispatch rate SMs x 4 x 32 x sampled SM clock (the live FFMA probe of fnb200MeasureAluPeak is reported beside it). `cpu_baseline` = the
reference's own node sources (oracle/_ref, relaxed build = the reference's default flags) on all host cores over a bounded
sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workload table: tree, call shape, reference lane-ops per point (SURVEY.md 8d / BASELINE.md section 3)
WORKLOADS = {
    "c1": dict(tree="c1_simplex", kind="grid3d", n=(256, 256, 256), step=1.0, ops=177, desc="Simplex 3D GenUniformGrid3D 256^3"),
    "c2a": dict(tree="c2_fbm5_perlin", kind="grid2d", n=(16384, 16384), step=0.25, ops=525, desc="FBm5(Perlin) GenUniformGrid2D 16384^2"),
    "c2b": dict(tree="c2_fbm5_perlin", kind="tileable2d", n=(4096, 4096), step=1.0, ops=2401, desc="FBm5(Perlin) GenTileable2D 4096^2"),
    "c3a": dict(tree="c3a_cellular_distance", kind="grid3d", n=(512, 512, 512), step=1.0, ops=1016, desc="CellularDistance GenUniformGrid3D 512^3"),
    "c3b": dict(tree="c3b_cellular_lookup_white", kind="grid3d", n=(512, 512, 512), step=1.0, ops=1024, desc="CellularLookup(White) GenUniformGrid3D 512^3"),
    "c4": dict(tree="c4_terrain", kind="grid3d", n=(1024, 1024, 1024), step=1.0, ops=2310, desc="terrain tree (DomainWarpFractalProgressive+FBm+Ridged+MaxSmooth+Remap+Terrace) GenUniformGrid3D 1024^3"),
    "c5": dict(tree="c5_supersimplex", kind="pos4d", n=(256 * 1024 * 1024,), step=0.0, ops=843, desc="SuperSimplex GenPositionArray4D 256Mi positions U[-1000,1000)^4 (Philox key 1337)"),
}
SEED = 1337
MATH_MODES = {"strict": 0, "relaxed": 1, "strict_generic": 2, "relaxed_generic": 3}
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r2_ncu_summary.json")


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    p.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="N > 1: split the workload's volume over the ranks (strong) or give every rank a whole one (weak)")
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="target wall time of one CPU baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--extras", action="store_true", help="also time e2e into a PAGEABLE destination, the one-process multi-GPU fan-out (when several GPUs are visible) and, for N > 1, the optional all-gather of the volume")
    p.add_argument("--math", default="strict", choices=sorted(MATH_MODES), help="fnb200SetMathMode: strict/relaxed arithmetic x table/generic gradient flavour")
    return p.parse_args()


def metric_name(args, w):
    return "Mpoints/sec 3D grid (terrain tree)" if args.workload == "c4" else "Mpoints/sec " + w["desc"]


def config_of(args, w, world):
    """The workload description: identical in both arms (the driver compares the two lines' `config`)."""
    from fastnoise2_b200.trees import TREES
    points = 1
    for n in w["n"]:
        points *= n
    out_mib = points * 4 >> 20
    return {"workload": w["desc"] + ", seed %d, step %g" % (SEED, w["step"]), "tree": w["tree"], "encoded_tree": TREES[w["tree"]],
            "points": points, "scaling": args.scaling if world > 1 else "single GPU",
            "sharding": "outermost-axis slab per rank, global integer indices" if world > 1 else "none",
            "l2": "the output of one step (%d MiB) exceeds L2; no inputs are re-read between steps" % out_mib}


def slab_range(total, rank, world):
    base, extra = divmod(int(total), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for (_, l) in self.lines]
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in rows:
            parts = [x.strip() for x in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag == "Active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU reference arm
def pick_ref_variant():
    from tests import helpers
    for variant in ("relaxed_avx512", "relaxed_avx2", "strict_avx2"):
        lib = helpers.load_ref(variant)
        if lib is not None:
            return variant, lib
    return None, None


def cpu_sample_shape(w, planes):
    """A bounded sample of the workload: the first `planes` z-planes (rows for 2D) of the grid / first positions."""
    if w["kind"] == "grid3d":
        return (w["n"][0], w["n"][1], max(1, min(w["n"][2], planes)))
    if w["kind"] in ("grid2d", "tileable2d"):
        return (w["n"][0], max(1, min(w["n"][1], planes)))
    return (max(1, min(w["n"][0], planes * 1024 * 1024)),)


def run_cpu(lib, w, shape, threads):
    """One pass of the reference CPU implementation over `shape`; returns (seconds, points, threads used)."""
    import numpy as np
    from fastnoise2_b200.trees import TREES
    from tests import helpers
    node = helpers.RefNode(lib, TREES[w["tree"]])
    if w["kind"] == "grid3d":
        nx, ny, nz = shape
        out = np.empty(nx * ny * nz, np.float32)
        sec = lib.fn2refTimedGridMT(node.handle, out.ctypes.data, 3, nx, ny, nz, 0.0, 0.0, 0.0, w["step"], SEED, threads, None)
        return sec, out.size, min(threads, nz)
    if w["kind"] == "grid2d":
        nx, ny = shape
        out = np.empty(nx * ny, np.float32)
        sec = lib.fn2refTimedGridMT(node.handle, out.ctypes.data, 2, nx, ny, 1, 0.0, 0.0, 0.0, w["step"], SEED, threads, None)
        return sec, out.size, min(threads, ny)
    if w["kind"] == "tileable2d":
        # GenTileable2D has no slab form in the reference API (it walks the whole image): the threaded driver splits the rows and
        # evaluates each slab through the reference's GenPositionArray4D over the torus positions (oracle/ref/ref_extra.cpp)
        nx, ny = shape
        out = np.empty(nx * ny, np.float32)
        if hasattr(lib, "fn2refTimedTileableMT") and threads > 1:
            sec = lib.fn2refTimedTileableMT(node.handle, out.ctypes.data, nx, ny, w["step"], w["step"], SEED, threads, None)
            return sec, out.size, min(threads, ny)
        t0 = time.time()
        lib.fnGenTileable2D(node.handle, out.ctypes.data, nx, ny, w["step"], w["step"], SEED, None)
        return time.time() - t0, out.size, 1
    n = shape[0]
    rng = np.random.default_rng(SEED)
    pos = [rng.uniform(-1000, 1000, n).astype(np.float32) for _ in range(4)]
    out = np.empty(n, np.float32)
    sec = lib.fn2refTimedPositionsMT(node.handle, out.ctypes.data, 4, n, pos[0].ctypes.data, pos[1].ctypes.data, pos[2].ctypes.data, pos[3].ctypes.data, SEED, threads, None)
    return sec, n, min(threads, n)


def cpu_baseline(w, target_seconds):
    variant, lib = pick_ref_variant()
    if lib is None:
        return None
    threads = os.cpu_count() or 1
    planes = 2
    shape = cpu_sample_shape(w, planes)
    for _ in range(4):  # grow the sample until one pass takes about target_seconds (first pass also warms the tree)
        sec, pts, used = run_cpu(lib, w, shape, threads)
        if sec >= target_seconds * 0.5:
            break
        grown = int(min(planes * max(2.0, target_seconds / max(sec, 1e-3)), 1 << 20))
        bigger = cpu_sample_shape(w, grown)
        if bigger == shape:
            break
        planes, shape = grown, bigger
    best = None
    for _ in range(2):
        sec, pts, used = run_cpu(lib, w, shape, threads)
        if best is None or pts / sec > best[1] / best[0]:
            best = (sec, pts)
    sec, pts = best
    return {"value": pts / sec / 1e6, "unit": "Mpoints/s", "cores": used, "kind": "reference",
            "sample": "oracle/_ref %s (reference node sources over the FastSIMD stand-in, reference default flags), %s of the workload grid = %d points in %.2f s, %d thread(s)%s" % (
                variant, "x".join(str(s) for s in shape), pts, sec, used, (", rows split over the threads through the reference's GenPositionArray4D on the torus positions (GenTileable2D has no slab form)" if w["kind"] == "tileable2d" else ", one slab per thread") if used > 1 else " (the reference API has no slab form of this call)"),
            "_lib": lib, "_shape": shape, "_variant": variant}


def reference_arm(args, w, rank, world):
    """--impl reference: the reference's CPU implementation timed through the same step/warmup protocol (rank 0 only)."""
    if rank != 0:
        return
    base = cpu_baseline(w, min(args.cpu_seconds, 6.0))
    if base is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built and /root/reference is not on this box"}))
        return
    lib, shape, threads = base["_lib"], base["_shape"], os.cpu_count() or 1
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        run_cpu(lib, w, shape, threads)
    total_sec, total_pts = 0.0, 0
    for _ in range(args.steps):
        sec, pts, used = run_cpu(lib, w, shape, threads)
        total_sec += sec
        total_pts += pts
    value = total_pts / total_sec / 1e6
    line = {
        "impl": "reference", "metric": metric_name(args, w),
        "value": value, "unit": "Mpoints/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup,
        "ms_per_step": total_sec / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_of(args, w, world),
        "cpu_baseline": {"value": value, "unit": "Mpoints/s", "cores": base["cores"], "kind": "reference",
                         "sample": base["sample"] + "; each of the %d timed steps = that sample (%s points)" % (args.steps, "x".join(str(s) for s in shape))},
        "e2e": {"value": value, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- B200 arm
class Work:
    """This rank's share of one step: output buffer, accumulator and the closure that enqueues the kernels."""

    def __init__(self, args, w, gen, lib, torch, rank, world, scaling, sptr, comm):
        kind = w["kind"]
        self.lib, self.sptr, self.comm = lib, sptr, comm
        self.mm = torch.zeros(2, dtype=torch.int32, device="cuda")
        self.in_bytes = 0
        outer = w["n"][-1]
        if kind == "pos4d":
            total = w["n"][0]
            if scaling == "weak":
                begin, end = rank * total, (rank + 1) * total
            else:  # 16-byte aligned shares
                begin, end = slab_range(total, rank, world)
                begin, end = (begin + 3) & ~3, total if rank == world - 1 else (end + 3) & ~3
            self.points = end - begin
            self.total_points = total * world if scaling == "weak" else total
            self.out = torch.empty(self.points, dtype=torch.float32, device="cuda")
            self.pos = [torch.empty(self.points, dtype=torch.float32, device="cuda") for _ in range(4)]
            for axis in range(4):  # synthetic positions generated on the device, identical on any rank layout
                lib.fnb200FillUniformPhilox(self.pos[axis].data_ptr(), self.points, begin, SEED, axis, -1000.0, 1000.0, sptr)
            self.in_bytes = self.points * 16

            def gen_call():
                return lib.fnb200GenPositionArray(gen.handle, 4, self.out.data_ptr(), self.points, self.pos[0].data_ptr(), self.pos[1].data_ptr(), self.pos[2].data_ptr(),
                                                  self.pos[3].data_ptr(), 0.0, 0.0, 0.0, 0.0, SEED, self.mm.data_ptr(), sptr)
        else:
            full = outer * world if scaling == "weak" else outer
            begin, end = (rank * outer, (rank + 1) * outer) if scaling == "weak" else slab_range(outer, rank, world)
            per_outer = 1
            for n in w["n"][:-1]:
                per_outer *= n
            self.points = per_outer * (end - begin)
            self.total_points = per_outer * full
            self.out = torch.empty(self.points, dtype=torch.float32, device="cuda")
            s = w["step"]
            if kind == "grid3d":
                nx, ny, _ = w["n"]

                def gen_call():
                    return lib.fnb200GenUniformGrid3D(gen.handle, self.out.data_ptr(), 0.0, 0.0, 0.0, nx, ny, full, s, s, s, SEED, begin, end, self.mm.data_ptr(), sptr)
            elif kind == "grid2d":
                nx, _ = w["n"]

                def gen_call():
                    return lib.fnb200GenUniformGrid2D(gen.handle, self.out.data_ptr(), 0.0, 0.0, nx, full, s, s, SEED, begin, end, self.mm.data_ptr(), sptr)
            else:
                nx, _ = w["n"]

                def gen_call():
                    return lib.fnb200GenTileable2D(gen.handle, self.out.data_ptr(), nx, full, s, s, SEED, begin, end, self.mm.data_ptr(), sptr)
        self.begin, self.end = begin, end
        self.out_bytes = self.points * 4
        self._gen_call = gen_call

    def step(self):
        """Kernels of one step on the stream: reset the accumulator, generate the slab, reduce min/max across ranks on the device."""
        import fastnoise2_b200 as fn
        self.lib.fnb200MinMaxReset(self.mm.data_ptr(), self.sptr)
        rc = self._gen_call()
        assert rc == 0, fn.last_error()
        if self.comm:
            rc = self.lib.fnb200CommAllReduceMinMax(self.comm, self.mm.data_ptr(), self.sptr)
            assert rc == 0, fn.last_error()

    def read_minmax(self):
        import numpy as np
        mm_host = np.zeros(2, np.float32)
        self.lib.fnb200MinMaxRead(self.mm.data_ptr(), self.sptr, mm_host.ctypes.data)
        return float(mm_host[0]), float(mm_host[1])


def timed_steps(work, steps, stream, torch, dist, world):
    """Exactly `steps` steps between a barrier + synchronize on both sides; CUDA events on the launching stream; max over ranks."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0 = time.time()
    ev0.record(stream)
    for k in range(steps):
        kev[k][0].record(stream)
        work.step()
        kev[k][1].record(stream)
    ev1.record(stream)
    lo, hi = work.read_minmax()  # the steps' result: one 8-byte read behind the last step
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.time()
    elapsed_ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    kernel_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in kev) / steps], device="cuda")
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kernel_ms, op=dist.ReduceOp.MAX)
    return elapsed_ms.item(), kernel_ms.item(), (lo, hi), (t0, t1)


def main():
    args = parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, w, rank, world)
        return

    # the headline e2e number is one GPU per process; the one-process multi-GPU fan-out is measured separately (--extras)
    os.environ.setdefault("FASTNOISE2_B200_FANOUT", "0")

    import numpy as np
    import torch
    import torch.distributed as dist
    import fastnoise2_b200 as fn

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not fn.device_available():
        raise SystemExit("fastnoise2_b200: no usable CUDA device: " + fn.last_error())
    fn.set_math_mode(MATH_MODES[args.math])
    lib = fn.lib
    gen = fn.Generator(fn.TREES[w["tree"]])
    stream = torch.cuda.Stream()  # the stream every kernel of the timed region is launched on, and the events are recorded on
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream

    # ---- the library's own communicator for the device-side min/max all-reduce (and the optional gather): id from rank 0
    comm = None
    if world > 1:
        import ctypes
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            assert lib.fnb200CommGetUniqueId(buf) == 0, fn.last_error()
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        uid_buf = ctypes.create_string_buffer(uid.cpu().numpy().tobytes(), 128)
        comm = lib.fnb200CommInitRank(world, rank, uid_buf)
        assert comm, fn.last_error()

    scaling = args.scaling if world > 1 else "weak"
    work = Work(args, w, gen, lib, torch, rank, world, scaling, sptr, comm)

    # ---- warm-up (includes the NVRTC compile / cubin-cache load on the first step)
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        work.step()
    work.read_minmax()
    torch.cuda.synchronize()

    # ---- ALU roofline denominator, measured live on this GPU (rank 0 is enough)
    alu_peak = lib.fnb200MeasureAluPeak(0, 4000) if rank == 0 else 0.0
    torch.cuda.synchronize()

    # ---- timed region
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = lib.fnb200LaunchCount()
    elapsed_ms, kernel_ms, (lo, hi), (t_wall0, t_wall1) = timed_steps(work, args.steps, stream, torch, dist, world)
    launches = lib.fnb200LaunchCount() - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = work.total_points * args.steps / (elapsed_ms * 1e-3) / 1e6
    points, out_bytes, in_bytes = work.points, work.out_bytes, work.in_bytes

    # ---- the other scaling mode, same protocol (N > 1 only), reported next to the headline
    other = None
    if world > 1:
        other_mode = "weak" if scaling == "strong" else "strong"
        work2 = Work(args, w, gen, lib, torch, rank, world, other_mode, sptr, comm)
        for _ in range(2):
            work2.step()
        o_ms, o_kernel_ms, _, _ = timed_steps(work2, args.steps, stream, torch, dist, world)
        other = {"scaling": other_mode, "value": work2.total_points * args.steps / (o_ms * 1e-3) / 1e6, "unit": "Mpoints/s", "ms_per_step": o_ms / args.steps,
                 "points_per_gpu": work2.points, "total_points": work2.total_points}
        del work2

    # ---- optional all-gather of the volume (only when a caller wants one contiguous buffer), timed on its own
    gather = None
    if args.extras and world > 1 and w["kind"] != "pos4d":
        gather = measure_gather(args, w, work, gen, lib, torch, dist, rank, world, sptr, comm, stream)

    # ---- e2e: the reference-shaped call with HOST buffers (pinned), copies inside the timed region
    e2e, extras = None, {}
    if not args.no_e2e:
        e2e = measure_e2e(args, w, gen, lib, rank, world, work, torch, dist, np, pinned=True)
        if args.extras:
            extras["e2e_pageable"] = measure_e2e(args, w, gen, lib, rank, world, work, torch, dist, np, pinned=False)
            if world == 1 and torch.cuda.device_count() > 1:
                os.environ["FASTNOISE2_B200_FANOUT"] = "1024"
                fan = measure_e2e(args, w, gen, lib, rank, world, work, torch, dist, np, pinned=True)
                fan["devices"] = torch.cuda.device_count()
                fan["note"] = "ONE process, one fnGen* call: rows split over every visible GPU, each draining its share over its own PCIe link"
                extras["e2e_fanout"] = fan
                os.environ["FASTNOISE2_B200_FANOUT"] = "0"

    if comm:
        lib.fnb200CommDestroy(comm)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None if args.no_cpu_baseline else cpu_baseline(w, args.cpu_seconds)
    if cpu:
        cpu = {k: v for k, v in cpu.items() if not k.startswith("_")}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    # From the committed `ncu --set full` capture of this workload's kernel (never measured under a timer here): DRAM bytes per
    # launch and the thread-instructions the kernel executes per point -- the work this implementation actually issues.
    traffic, inst_per_point, packed_per_point = None, None, 0.0
    try:
        prof = json.load(open(NCU_SUMMARY)).get(args.workload, {})
        traffic = prof.get("dram_bytes_per_launch")
        inst_per_point = prof.get("thread_inst_per_point") if args.math == "strict" and "stale" not in prof else None
        packed_per_point = prof.get("packed_thread_inst_per_point", 0.0) if inst_per_point else 0.0
    except (OSError, ValueError):
        pass
    ref_ops = w["ops"] * points / (kernel_ms * 1e-3) / 1e12
    # issue-slot utilisation = thread-instructions / time / issue peak (an FFMA2 is one slot); lane-operations counts the two
    # operations of a packed FP32 instruction, the unit the FFMA probe measures
    issue_rate = inst_per_point * points / (kernel_ms * 1e-3) / 1e12 if inst_per_point else None
    lane_ops = (inst_per_point + packed_per_point) * points / (kernel_ms * 1e-3) / 1e12 if inst_per_point else None
    # The bound: an SM sub-partition dispatches one warp-instruction (32 lane-operations) per clock; a packed FP32 instruction
    # holds the dispatch port for two clocks (profiles/r2_alu_probe.txt). Peak = SMs x 4 x 32 x the SM clock sampled under load --
    # the theoretical dispatch rate, which no instruction mix can exceed; the live FFMA probe (97 % of it) is reported beside it.
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz")
    peak = sms * 4 * 32 * sm_mhz * 1e6 / 1e12 if sm_mhz else alu_peak / 1e12
    store_gbs = out_bytes / (kernel_ms * 1e-3) / 1e9
    name = os.path.relpath(NCU_SUMMARY, ROOT)
    line = {
        "metric": metric_name(args, w),
        "value": value, "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_of(args, w, world),
        "run": {"points_per_gpu": points, "total_points": work.total_points, "math": args.math, "output_minmax": [lo, hi],
                "minmax_reduction": "ncclAllReduce(max) over the accumulator's two unsigned keys, on the launching stream, every step" if world > 1 else "single GPU: none"},
        "roofline": {"bound": "alu", "achieved": lane_ops, "peak": peak, "unit": "Tlane-op/s", "frac": lane_ops / peak if peak and lane_ops else None,
                     "issue_slot_frac": issue_rate / peak if peak and issue_rate else None,
                     "achieved_is": ("lane-operations per second = (%.1f thread-inst/point, of which %.1f packed FP32 counted twice -- an FFMA2 holds the dispatch port for two cycles, profiles/r2_alu_probe.txt; %s, ncu smsp__inst_executed x threads/inst) x points / kernel time" % (inst_per_point, packed_per_point, name)) if inst_per_point
                     else "no ncu instruction count of the current kernel is committed for this workload + math mode (%s): see reference_ops" % name,
                     "traffic": traffic, "traffic_source": (name + " (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture)") if traffic else None,
                     "algorithmic_bytes": out_bytes + in_bytes, "kernel": "fn2_entry (NVRTC-compiled tree kernel, persistent grid-stride launch)", "kernel_ms": kernel_ms,
                     "inst_per_point": inst_per_point,
                     "reference_ops": {"ops_per_point": w["ops"], "tlane_ops_per_s": ref_ops, "note": "SURVEY 8d count of the reference's generic-path SIMD lane-ops; the GPU needs fewer instructions"},
                     "peak_source": ("theoretical dispatch rate: %d SMs x 4 sub-partitions x 32 lanes x %.0f MHz (SM clock sampled during the timed region); MEASURED_PEAKS.json has no ALU entry" % (sms, sm_mhz)) if sm_mhz
                     else "FFMA issue rate measured live by fnb200MeasureAluPeak (no clock sample available)",
                     "ffma_probe_tlane_ops_per_s": alu_peak / 1e12,
                     "hbm_store": {"achieved_gbs": store_gbs, "peak_gbs": peaks.get("hbm_gbs"), "frac": store_gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                                   "algorithmic_bytes_per_point": 4 + (16 if w["kind"] == "pos4d" else 0)}},
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if other:
        line[other["scaling"]] = other
    if gather:
        line["gather"] = gather
    line.update(extras)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_gather(args, w, work, gen, lib, torch, dist, rank, world, sptr, comm, stream):
    """Strong-scaling slabs generated IN PLACE inside a whole-volume buffer, then all-gathered over NVLink (fnb200CommGatherSlabs)."""
    import ctypes
    import fastnoise2_b200 as fn
    total = 1
    for n in w["n"]:
        total *= n
    per_outer = total // w["n"][-1]
    counts = [(slab_range(w["n"][-1], r, world)[1] - slab_range(w["n"][-1], r, world)[0]) * per_outer for r in range(world)]
    begin, end = slab_range(w["n"][-1], rank, world)
    volume = torch.empty(total, dtype=torch.float32, device="cuda")
    mine = volume[begin * per_outer:end * per_outer]
    carr = (ctypes.c_longlong * world)(*counts)
    s = w["step"]

    def generate():
        if w["kind"] == "grid3d":
            return lib.fnb200GenUniformGrid3D(gen.handle, mine.data_ptr(), 0.0, 0.0, 0.0, w["n"][0], w["n"][1], w["n"][2], s, s, s, SEED, begin, end, None, sptr)
        if w["kind"] == "grid2d":
            return lib.fnb200GenUniformGrid2D(gen.handle, mine.data_ptr(), 0.0, 0.0, w["n"][0], w["n"][1], s, s, SEED, begin, end, None, sptr)
        return lib.fnb200GenTileable2D(gen.handle, mine.data_ptr(), w["n"][0], w["n"][1], s, s, SEED, begin, end, None, sptr)

    def gather_call():
        return lib.fnb200CommGatherSlabs(comm, None, volume.data_ptr(), carr, sptr)

    assert generate() == 0 and gather_call() == 0, fn.last_error()
    torch.cuda.synchronize()
    dist.barrier()
    steps = max(1, min(args.steps, 5))
    g0, g1, g2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    gen_ms = gat_ms = 0.0
    for _ in range(steps):
        g0.record(stream)
        generate()
        g1.record(stream)
        gather_call()
        g2.record(stream)
        torch.cuda.synchronize()
        gen_ms += g0.elapsed_time(g1)
        gat_ms += g1.elapsed_time(g2)
    t = torch.tensor([gen_ms / steps, gat_ms / steps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gen_ms, gat_ms = t.tolist()
    inbound = (total - counts[rank]) * 4
    result = {"generate_ms": gen_ms, "gather_ms": gat_ms, "volume_bytes": total * 4, "inbound_bytes_per_gpu": inbound,
              "inbound_GBps_per_gpu": inbound / (gat_ms * 1e-3) / 1e9, "steps": steps,
              "api": "fnb200CommGatherSlabs (NCCL all-gather over NVLink), slabs generated in place inside the volume; timed separately from compute"}
    del volume, mine
    # the fused form: the generating kernel stores every value into all ranks' copies (peer memory, CUDA IPC), no gather pass
    if w["kind"] == "grid3d" and world <= 8:
        shared = lib.fnb200CommAllocVolume(comm, total, sptr)
        if shared:
            def fused():
                return lib.fnb200GenUniformGrid3DShared(gen.handle, shared, 0.0, 0.0, 0.0, w["n"][0], w["n"][1], w["n"][2], s, s, s, SEED, begin, end, work.mm.data_ptr(), sptr)
            assert fused() == 0, fn.last_error()
            lib.fnb200CommAllReduceMinMax(comm, work.mm.data_ptr(), sptr)
            torch.cuda.synchronize()
            dist.barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            for _ in range(steps):
                fused()
                lib.fnb200CommAllReduceMinMax(comm, work.mm.data_ptr(), sptr)  # the cross-rank completion barrier of a step
            f1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([f0.elapsed_time(f1) / steps], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            result["fused_generate_and_gather_ms"] = t.item()
            result["fused_outbound_GBps_per_gpu"] = (world - 1) * counts[rank] * 4 / (t.item() * 1e-3) / 1e9
            result["fused_api"] = "fnb200GenUniformGrid3DShared: peer stores from the generating kernel into every rank's copy (CUDA IPC over NVLink) + min/max all-reduce as the step barrier"
            lib.fnb200CommFreeVolume(comm, shared, sptr)
        else:
            result["fused_error"] = fn.last_error()
    return result


def measure_e2e(args, w, gen, lib, rank, world, work, torch, dist, np, pinned):
    """Same metric through the reference-facing call (fnGenUniformGrid3D & co) with host buffers: pinned, or pageable numpy memory."""
    import fastnoise2_b200 as fn
    kind = w["kind"]
    steps = args.steps
    points = work.points
    if pinned:
        host_t = torch.empty(points, dtype=torch.float32, pin_memory=True)
        host_ptr = host_t.data_ptr()
    else:
        host_t = np.empty(points, np.float32)
        host_t[:] = 0  # touch the pages: first-touch faults belong to the allocator, not to the call
        host_ptr = host_t.ctypes.data
    mm = np.zeros(2, np.float32)
    h2d = 0
    s = w["step"]
    if kind == "grid3d":
        nx, ny, _ = w["n"]
        nz = work.end - work.begin
        # a rank's slab through the unmodified reference signature: float z offset (bit-exact slabs need the fnb200 entry; this is the drop-in path)
        z0 = float(work.begin) * s

        def call():
            lib.fnGenUniformGrid3D(gen.handle, host_ptr, 0.0, 0.0, z0, nx, ny, nz, s, s, s, SEED, mm.ctypes.data)
    elif kind == "grid2d":
        nx, _ = w["n"]
        ny = work.end - work.begin
        y0 = float(work.begin) * s

        def call():
            lib.fnGenUniformGrid2D(gen.handle, host_ptr, 0.0, y0, nx, ny, s, s, SEED, mm.ctypes.data)
    elif kind == "tileable2d":
        nx, ny = w["n"]  # the reference call has no slab form: every rank generates the whole tile
        points = nx * ny
        if pinned:
            host_t = torch.empty(points, dtype=torch.float32, pin_memory=True)
            host_ptr = host_t.data_ptr()
        else:
            host_t = np.zeros(points, np.float32)
            host_ptr = host_t.ctypes.data

        def call():
            lib.fnGenTileable2D(gen.handle, host_ptr, nx, ny, s, s, SEED, mm.ctypes.data)
    else:
        n = min(points, 0x7fffffff)
        points = n
        rng = np.random.default_rng(SEED + rank)
        if pinned:
            host_pos = [torch.empty(n, dtype=torch.float32, pin_memory=True) for _ in range(4)]
            for t in host_pos:
                t.numpy()[:] = rng.uniform(-1000, 1000, n).astype(np.float32)
            pos_ptr = [t.data_ptr() for t in host_pos]
        else:
            host_pos = [rng.uniform(-1000, 1000, n).astype(np.float32) for _ in range(4)]
            pos_ptr = [t.ctypes.data for t in host_pos]
        h2d = n * 16

        def call():
            lib.fnGenPositionArray4D(gen.handle, host_ptr, n, pos_ptr[0], pos_ptr[1], pos_ptr[2], pos_ptr[3], 0.0, 0.0, 0.0, 0.0, SEED, mm.ctypes.data)
    call()  # warm-up (allocates the staging buffers)
    assert mm[0] <= mm[1], fn.last_error()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(steps):
        call()
    sec = torch.tensor([time.time() - t0], device="cuda")
    total = torch.tensor([float(points)], device="cuda")
    if world > 1:
        dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return {"value": total.item() * steps / sec.item() / 1e6, "unit": "Mpoints/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": points * 4 + 8,
            "steps": steps, "host_memory": "pinned" if pinned else "pageable (numpy)",
            "api": "fnGen* C ABI (include/FastNoise/FastNoise_C.h) with host buffers; the call returns after the last byte is on the host; wall clock around %d calls, max over ranks" % steps}


if __name__ == "__main__":
    main()
