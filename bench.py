#!/usr/bin/env python
"""Benchmark of the fused node-graph evaluation path (BASELINE.json metric: Mpoints/s of a 3D uniform grid).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c4]

One "step" = one pass of the hot path over one batch: the whole grid of the workload generated into HBM.
Default workload = BASELINE.json configs[3], the config the metric is quoted on: the domain-warped terrain tree
(DomainWarpFractalProgressive + FBm + Ridged + MaxSmooth + Remap + Terrace, SURVEY.md 8d "C4"),
GenUniformGrid3D 1024^3, seed 1337. It fits one GPU (4 GiB of output), so N=1 runs it whole. With N ranks the
volume is 1024 x 1024 x (1024*N) and rank r generates z-slab r through the slab entry point with GLOBAL integer
indices (weak scaling: every GPU always does 1024^3 points); the only cross-rank traffic is the min/max
all-reduce of 2 floats (NCCL), which is inside the timed region.

JSON keys: see the task contract. `value` is device-resident throughput (CUDA events around the K steps on the
launching stream, max over ranks). `e2e` goes through the reference-shaped C ABI call fnGenUniformGrid3D with a
pinned HOST output buffer: device->host copies are inside the timed region. `roofline` is the ALU issue roofline
(this path has no tensor work and writes 4 B/point): achieved = reference lane-ops/point (measured
with an op-counting build of the reference sources) x points / kernel time, peak = FFMA issue rate measured live
by fnb200MeasureAluPeak. `cpu_baseline` = the reference's own node sources (oracle/_ref, relaxed build = the
reference's default flags) on all host cores over a bounded sample of the same workload.
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


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="target wall time of one CPU baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--math", default="strict", choices=sorted(MATH_MODES), help="fnb200SetMathMode: strict/relaxed arithmetic x table/generic gradient flavour")
    return p.parse_args()


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
    """One pass of the reference CPU implementation over `shape`; returns (seconds, points)."""
    import numpy as np
    import fastnoise2_b200 as fn
    from tests import helpers
    node = helpers.RefNode(lib, fn.TREES[w["tree"]])
    if w["kind"] == "grid3d":
        nx, ny, nz = shape
        out = np.empty(nx * ny * nz, np.float32)
        sec = lib.fn2refTimedGridMT(node.handle, out.ctypes.data, 3, nx, ny, nz, 0.0, 0.0, 0.0, w["step"], SEED, threads, None)
        return sec, out.size
    if w["kind"] == "grid2d":
        nx, ny = shape
        out = np.empty(nx * ny, np.float32)
        sec = lib.fn2refTimedGridMT(node.handle, out.ctypes.data, 2, nx, ny, 1, 0.0, 0.0, 0.0, w["step"], SEED, threads, None)
        return sec, out.size
    if w["kind"] == "tileable2d":
        # the tileable driver has no slab form in the reference API: time the FULL-width rows of a smaller tile
        nx, ny = shape
        out = np.empty(nx * ny, np.float32)
        t0 = time.time()
        lib.fnGenTileable2D(node.handle, out.ctypes.data, nx, ny, w["step"], w["step"], SEED, None)
        return time.time() - t0, out.size
    n = shape[0]
    rng = np.random.default_rng(SEED)
    pos = [rng.uniform(-1000, 1000, n).astype(np.float32) for _ in range(4)]
    out = np.empty(n, np.float32)
    sec = lib.fn2refTimedPositionsMT(node.handle, out.ctypes.data, 4, n, pos[0].ctypes.data, pos[1].ctypes.data, pos[2].ctypes.data, pos[3].ctypes.data, SEED, threads, None)
    return sec, n


def cpu_baseline(w, target_seconds):
    variant, lib = pick_ref_variant()
    if lib is None:
        return None
    threads = os.cpu_count() or 1
    planes = 2
    shape = cpu_sample_shape(w, planes)
    for _ in range(4):  # grow the sample until one pass takes about target_seconds (first pass also warms the tree)
        sec, pts = run_cpu(lib, w, shape, threads)
        if sec >= target_seconds * 0.5:
            break
        grown = int(min(planes * max(2.0, target_seconds / max(sec, 1e-3)), 1 << 20))
        bigger = cpu_sample_shape(w, grown)
        if bigger == shape:
            break
        planes, shape = grown, bigger
    best = None
    for _ in range(2):
        sec, pts = run_cpu(lib, w, shape, threads)
        if best is None or pts / sec > best[1] / best[0]:
            best = (sec, pts)
    sec, pts = best
    return {"value": pts / sec / 1e6, "unit": "Mpoints/s", "cores": threads, "kind": "reference",
            "sample": "oracle/_ref %s (reference node sources over the FastSIMD stand-in, reference default flags), %s of the workload grid = %d points in %.2f s, %d threads, one slab per thread" % (
                variant, "x".join(str(s) for s in shape), pts, sec, threads),
            "_lib": lib, "_shape": shape, "_variant": variant}


def reference_arm(args, w, rank):
    """--impl reference: the reference's CPU implementation timed through the same step/warmup protocol (rank 0 only)."""
    if rank != 0:
        return
    base = cpu_baseline(w, min(args.cpu_seconds, 6.0))
    if base is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built and /root/reference is not on this box"}))
        return
    lib, shape, threads = base["_lib"], base["_shape"], base["cores"]
    for _ in range(min(args.warmup, 1)):
        run_cpu(lib, w, shape, threads)
    total_sec, total_pts = 0.0, 0
    for _ in range(args.steps):
        sec, pts = run_cpu(lib, w, shape, threads)
        total_sec += sec
        total_pts += pts
    value = total_pts / total_sec / 1e6
    line = {
        "impl": "reference", "metric": "Mpoints/sec 3D grid (terrain tree)" if args.workload == "c4" else "Mpoints/sec " + w["desc"],
        "value": value, "unit": "Mpoints/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"] + ", seed %d; each step = a bounded sample of it: %s points" % (SEED, "x".join(str(s) for s in shape)),
                   "tree": w["tree"], "threads": threads},
        "cpu_baseline": {"value": value, "unit": "Mpoints/s", "cores": threads, "kind": "reference", "sample": base["sample"]},
        "e2e": {"value": value, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- B200 arm
def main():
    args = parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, w, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import fastnoise2_b200 as fn
    from fastnoise2_b200 import api

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

    # ---- per-rank work: a slab of the N-times-larger volume (weak scaling), global integer indices
    kind = w["kind"]
    if kind == "grid3d":
        nx, ny, nz = w["n"]
        points = nx * ny * nz
        out = torch.empty(points, dtype=torch.float32, device="cuda")
        mm = torch.zeros(2, dtype=torch.int32, device="cuda")

        def step():
            lib.fnb200MinMaxReset(mm.data_ptr(), sptr)
            rc = lib.fnb200GenUniformGrid3D(gen.handle, out.data_ptr(), 0.0, 0.0, 0.0, nx, ny, nz * world, w["step"], w["step"], w["step"], SEED,
                                            rank * nz, (rank + 1) * nz, mm.data_ptr(), sptr)
            assert rc == 0, fn.last_error()
        out_bytes = points * 4
        in_bytes = 0
    elif kind in ("grid2d", "tileable2d"):
        nx, ny = w["n"]
        points = nx * ny
        out = torch.empty(points, dtype=torch.float32, device="cuda")
        mm = torch.zeros(2, dtype=torch.int32, device="cuda")

        def step():
            lib.fnb200MinMaxReset(mm.data_ptr(), sptr)
            if kind == "grid2d":
                rc = lib.fnb200GenUniformGrid2D(gen.handle, out.data_ptr(), 0.0, 0.0, nx, ny * world, w["step"], w["step"], SEED, rank * ny, (rank + 1) * ny, mm.data_ptr(), sptr)
            else:
                rc = lib.fnb200GenTileable2D(gen.handle, out.data_ptr(), nx, ny * world, w["step"], w["step"], SEED, rank * ny, (rank + 1) * ny, mm.data_ptr(), sptr)
            assert rc == 0, fn.last_error()
        out_bytes = points * 4
        in_bytes = 0
    else:
        points = w["n"][0]
        out = torch.empty(points, dtype=torch.float32, device="cuda")
        mm = torch.zeros(2, dtype=torch.int32, device="cuda")
        pos = [torch.empty(points, dtype=torch.float32, device="cuda") for _ in range(4)]
        for axis in range(4):  # synthetic positions generated on the device, identical on any rank layout
            lib.fnb200FillUniformPhilox(pos[axis].data_ptr(), points, rank * points, SEED, axis, -1000.0, 1000.0, sptr)

        def step():
            lib.fnb200MinMaxReset(mm.data_ptr(), sptr)
            rc = lib.fnb200GenPositionArray(gen.handle, 4, out.data_ptr(), points, pos[0].data_ptr(), pos[1].data_ptr(), pos[2].data_ptr(), pos[3].data_ptr(),
                                            0.0, 0.0, 0.0, 0.0, SEED, mm.data_ptr(), sptr)
            assert rc == 0, fn.last_error()
        out_bytes = points * 4
        in_bytes = points * 16

    mm_host = np.zeros(2, np.float32)
    mm_dev = torch.zeros(2, dtype=torch.float32, device="cuda")

    def reduce_minmax():
        """OutputMinMax of the whole volume: decode this rank's accumulator, then min/max all-reduce of 2 floats."""
        lib.fnb200MinMaxRead(mm.data_ptr(), sptr, mm_host.ctypes.data)
        if world > 1:
            mm_dev.copy_(torch.tensor([mm_host[0], -mm_host[1]]))
            dist.all_reduce(mm_dev, op=dist.ReduceOp.MIN)
            lo, hi = mm_dev.tolist()
            return lo, -hi
        return float(mm_host[0]), float(mm_host[1])

    # ---- warm-up (includes the NVRTC compile / cubin-cache load on the first step)
    for _ in range(max(args.warmup, 3)):
        step()
    reduce_minmax()
    torch.cuda.synchronize()

    # ---- ALU roofline denominator, measured live on this GPU (rank 0 is enough)
    alu_peak = lib.fnb200MeasureAluPeak(0, 4000) if rank == 0 else 0.0
    torch.cuda.synchronize()

    # ---- timed region: exactly K steps, device timing with CUDA events on the launching stream
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = lib.fnb200LaunchCount()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.time()
    ev0.record(stream)
    for k in range(args.steps):
        kev[k][0].record(stream)
        step()
        kev[k][1].record(stream)
        lo, hi = reduce_minmax()
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    launches = lib.fnb200LaunchCount() - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    elapsed_ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    kernel_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in kev) / args.steps], device="cuda")
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kernel_ms, op=dist.ReduceOp.MAX)
    elapsed_ms, kernel_ms = elapsed_ms.item(), kernel_ms.item()
    total_points = points * world
    value = total_points * args.steps / (elapsed_ms * 1e-3) / 1e6

    # ---- e2e: the reference-shaped call with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(args, w, gen, lib, rank, world, points, torch, dist, np)

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
        prof = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_summary.json"))).get(args.workload, {})
        traffic = prof.get("dram_bytes_per_launch")
        inst_per_point = prof.get("thread_inst_per_point") if args.math == "strict" and "stale" not in prof else None
        packed_per_point = prof.get("packed_thread_inst_per_point", 0.0) if inst_per_point else 0.0
    except (OSError, ValueError):
        pass
    ref_ops = w["ops"] * points / (kernel_ms * 1e-3) / 1e12
    # lane-operations = thread-instructions, with a packed FP32 instruction (FFMA2 / FMUL2 / FADD2) counted as the two it performs --
    # the unit the peak is measured in (FFMA lane-ops/s; FFMA2 reaches the same lane-op rate with half the instructions)
    achieved = (inst_per_point + packed_per_point) * points / (kernel_ms * 1e-3) / 1e12 if inst_per_point else None
    issue_frac = inst_per_point * points / (kernel_ms * 1e-3) / alu_peak if inst_per_point and alu_peak else None
    peak = alu_peak / 1e12
    store_gbs = out_bytes / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": "Mpoints/sec 3D grid (terrain tree)" if args.workload == "c4" else "Mpoints/sec " + w["desc"],
        "value": value, "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"] + ", seed %d, step %g" % (SEED, w["step"]), "tree": w["tree"], "encoded_tree": fn.TREES[w["tree"]],
                   "points_per_gpu": points, "sharding": "z-slab (outermost axis) per rank, global integer indices" if world > 1 else "none",
                   "math": args.math, "l2": "output per step (%d MiB) exceeds L2; no inputs are re-read" % (out_bytes >> 20),
                   "output_minmax": [lo, hi]},
        "roofline": {"bound": "alu", "achieved": achieved, "peak": peak, "unit": "Tlane-op/s", "frac": achieved / peak if peak and achieved else None, "issue_slot_frac": issue_frac,
                     "achieved_is": ("lane-operations per second = (%.1f thread-inst/point, of which %.1f packed FP32 counted twice; profiles/r1_ncu_summary.json, ncu) x points / kernel time" % (inst_per_point, packed_per_point)) if inst_per_point
                     else "no ncu instruction count of the current kernel is committed for this workload + math mode (profiles/r1_ncu_summary.json): see reference_ops",
                     "traffic": traffic, "traffic_source": "profiles/r1_ncu_summary.json (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture)" if traffic else None,
                     "algorithmic_bytes": out_bytes + in_bytes, "kernel": "fn2_entry (NVRTC-specialised tree kernel)", "kernel_ms": kernel_ms,
                     "inst_per_point": inst_per_point,
                     "reference_ops": {"ops_per_point": w["ops"], "tlane_ops_per_s": ref_ops, "note": "SURVEY 8d count of the reference's generic-path SIMD lane-ops; the GPU needs fewer instructions"},
                     "peak_source": "FFMA issue rate measured live by fnb200MeasureAluPeak (148 SM x 128 lanes x clock = the SM issue limit); MEASURED_PEAKS.json has no ALU entry",
                     "hbm_store": {"achieved_gbs": store_gbs, "peak_gbs": peaks.get("hbm_gbs"), "frac": store_gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                                   "algorithmic_bytes_per_point": 4 + (16 if kind == "pos4d" else 0)}},
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_e2e(args, w, gen, lib, rank, world, points, torch, dist, np):
    """Same metric through the reference-facing call (fnGenUniformGrid3D & co) with pinned host buffers."""
    import fastnoise2_b200 as fn
    kind = w["kind"]
    steps = max(1, min(args.steps, 3))
    host_out = torch.empty(points, dtype=torch.float32, pin_memory=True)
    mm = np.zeros(2, np.float32)
    h2d = 0
    if kind == "grid3d":
        nx, ny, nz = w["n"]
        # a rank's slab through the unmodified reference signature: float z offset (bit-exact slabs need the fnb200 entry; this is the drop-in path)
        z0 = float(rank * nz) * w["step"]

        def call():
            lib.fnGenUniformGrid3D(gen.handle, host_out.data_ptr(), 0.0, 0.0, z0, nx, ny, nz, w["step"], w["step"], w["step"], SEED, mm.ctypes.data)
    elif kind == "grid2d":
        nx, ny = w["n"]
        y0 = float(rank * ny) * w["step"]

        def call():
            lib.fnGenUniformGrid2D(gen.handle, host_out.data_ptr(), 0.0, y0, nx, ny, w["step"], w["step"], SEED, mm.ctypes.data)
    elif kind == "tileable2d":
        nx, ny = w["n"]

        def call():
            lib.fnGenTileable2D(gen.handle, host_out.data_ptr(), nx, ny, w["step"], w["step"], SEED, mm.ctypes.data)
    else:
        n = min(points, 0x7fffffff)
        host_pos = [torch.empty(n, dtype=torch.float32, pin_memory=True) for _ in range(4)]
        rng = np.random.default_rng(SEED + rank)
        for t in host_pos:
            t.numpy()[:] = rng.uniform(-1000, 1000, n).astype(np.float32)
        h2d = n * 16

        def call():
            lib.fnGenPositionArray4D(gen.handle, host_out.data_ptr(), n, host_pos[0].data_ptr(), host_pos[1].data_ptr(), host_pos[2].data_ptr(), host_pos[3].data_ptr(),
                                     0.0, 0.0, 0.0, 0.0, SEED, mm.ctypes.data)
    call()  # warm-up (allocates the staging buffers)
    assert mm[0] <= mm[1], fn.last_error()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(steps):
        call()
    sec = torch.tensor([time.time() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(sec, op=dist.ReduceOp.MAX)
    return {"value": points * world * steps / sec.item() / 1e6, "unit": "Mpoints/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": points * 4 + 8,
            "steps": steps, "api": "fnGen* C ABI (include/FastNoise/FastNoise_C.h) with pinned host buffers; the call returns after the last byte is on the host"}


if __name__ == "__main__":
    main()
