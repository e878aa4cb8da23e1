"""Several GPUs (run on a multi-GPU box: gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu; skipped with one GPU).

  * one process, one fnGen* call, rows fanned out over every visible GPU (each draining its share over its own PCIe link):
    bit-identical to the one-GPU result, pinned and pageable destinations;
  * one process per GPU: the library's NCCL communicator -- device-side min/max all-reduce of the accumulator keys and the
    optional all-gather of ragged slabs -- against the whole volume generated on one GPU.
"""
import ctypes
import os
import sys

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs_two = pytest.mark.skipif(gpu_count() < 2, reason="needs at least two GPUs")


@needs_two
def test_one_call_fans_out_over_every_gpu_bit_for_bit(fn):
    import subprocess
    code = r'''
import os, sys
sys.path.insert(0, %r)
os.environ["FASTNOISE2_B200_FANOUT_MIN_MPOINTS"] = "1"
os.environ["FASTNOISE2_B200_STAGING_MB"] = "4"
import numpy as np, torch
import fastnoise2_b200 as fn
g = fn.Generator(fn.TREES["c4_terrain"])
n = (256, 200, 99)
total = n[0] * n[1] * n[2]
os.environ["FASTNOISE2_B200_FANOUT"] = "0"
one = np.empty(total, np.float32)
mm_one = g.grid3d(one, 0.5, -3.0, 7.0, *n, 1.3, 1.3, 1.3, 1337)
os.environ["FASTNOISE2_B200_FANOUT"] = "1024"
launches = fn.lib.fnb200LaunchCount()
fan = np.empty(total, np.float32)                      # pageable: bounce buffers + copy workers, on every device
mm_fan = g.grid3d(fan, 0.5, -3.0, 7.0, *n, 1.3, 1.3, 1.3, 1337)
assert fn.lib.fnb200LaunchCount() - launches >= torch.cuda.device_count()  # every device took a share (one 4 MB chunk each on 8 GPUs)
pin = torch.empty(total, dtype=torch.float32, pin_memory=True)
mm_pin = g.grid3d(pin, 0.5, -3.0, 7.0, *n, 1.3, 1.3, 1.3, 1337)
assert mm_one == mm_fan == mm_pin, (mm_one, mm_fan, mm_pin)
assert np.array_equal(one.view(np.uint32), fan.view(np.uint32)) and np.array_equal(one.view(np.uint32), pin.numpy().view(np.uint32))
assert torch.cuda.current_device() == 0
s = fn.Generator(fn.TREES["c5_supersimplex"])
rng = np.random.default_rng(1)
cnt = 3000001
pos = [rng.uniform(-1000, 1000, cnt).astype(np.float32) for _ in range(4)]
os.environ["FASTNOISE2_B200_FANOUT"] = "0"
a = np.empty(cnt, np.float32); mm_a = s.positions(a, pos, [0, 0, 0, 0], 1337)
os.environ["FASTNOISE2_B200_FANOUT"] = "1024"
b = np.empty(cnt, np.float32); mm_b = s.positions(b, pos, [0, 0, 0, 0], 1337)
assert mm_a == mm_b and np.array_equal(a.view(np.uint32), b.view(np.uint32))
print("ok", torch.cuda.device_count())
''' % helpers.ROOT
    result = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert result.returncode == 0 and result.stdout.strip().startswith("ok"), result.stderr[-3000:]


def _rank_main(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    import fastnoise2_b200 as fn
    lib = fn.lib
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        assert lib.fnb200CommGetUniqueId(buf) == 0, fn.last_error()
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    comm = lib.fnb200CommInitRank(world, rank, ctypes.create_string_buffer(uid.cpu().numpy().tobytes(), 128))
    assert comm, fn.last_error()

    g = fn.Generator(fn.TREES["c4_terrain"])
    n = (128, 96, 61)  # 61 planes: ragged slabs
    per = n[0] * n[1]
    from fastnoise2_b200.sharding import slab_range
    counts = [(slab_range(n[2], r, world)[1] - slab_range(n[2], r, world)[0]) * per for r in range(world)]
    begin, end = slab_range(n[2], rank, world)
    stream = torch.cuda.Stream()
    sp = stream.cuda_stream
    volume = torch.zeros(per * n[2], dtype=torch.float32, device="cuda")
    acc = torch.zeros(2, dtype=torch.int32, device="cuda")
    with torch.cuda.stream(stream):
        assert lib.fnb200MinMaxReset(acc.data_ptr(), sp) == 0
        mine = volume[begin * per:end * per]  # generated in place
        assert lib.fnb200GenUniformGrid3D(g.handle, mine.data_ptr(), 0.0, 0.0, 0.0, n[0], n[1], n[2], 2.7, 2.7, 2.7, 1337, begin, end, acc.data_ptr(), sp) == 0, fn.last_error()
        assert lib.fnb200CommAllReduceMinMax(comm, acc.data_ptr(), sp) == 0, fn.last_error()
        assert lib.fnb200CommGatherSlabs(comm, None, volume.data_ptr(), (ctypes.c_longlong * world)(*counts), sp) == 0, fn.last_error()
        mm = np.zeros(2, np.float32)
        assert lib.fnb200MinMaxRead(acc.data_ptr(), sp, mm.ctypes.data) == 0
    stream.synchronize()
    # the same volume, whole, on this rank's GPU
    whole = torch.empty(per * n[2], dtype=torch.float32, device="cuda")
    mm_whole = g.grid3d(whole, 0.0, 0.0, 0.0, *n, 2.7, 2.7, 2.7, 1337)
    assert torch.equal(volume, whole), "gathered volume differs on rank %d" % rank
    assert (float(mm[0]), float(mm[1])) == mm_whole, (mm, mm_whole)
    # a separate send buffer (slab not in place) gathers to the same volume
    volume2 = torch.zeros_like(volume)
    with torch.cuda.stream(stream):
        slab = whole[begin * per:end * per].clone()
        assert lib.fnb200CommGatherSlabs(comm, slab.data_ptr(), volume2.data_ptr(), (ctypes.c_longlong * world)(*counts), sp) == 0, fn.last_error()
    stream.synchronize()
    assert torch.equal(volume2, whole)
    # fused generate + all-gather: every rank's kernel stores its slab into ALL ranks' copies of a shared volume (peer memory)
    shared = lib.fnb200CommAllocVolume(comm, per * n[2], sp)
    assert shared, fn.last_error()
    with torch.cuda.stream(stream):
        assert lib.fnb200MinMaxReset(acc.data_ptr(), sp) == 0
        assert lib.fnb200GenUniformGrid3DShared(g.handle, shared, 0.0, 0.0, 0.0, n[0], n[1], n[2], 2.7, 2.7, 2.7, 1337, begin, end, acc.data_ptr(), sp) == 0, fn.last_error()
        assert lib.fnb200CommAllReduceMinMax(comm, acc.data_ptr(), sp) == 0  # also the cross-rank barrier: every slab has landed
        mm3 = np.zeros(2, np.float32)
        assert lib.fnb200MinMaxRead(acc.data_ptr(), sp, mm3.ctypes.data) == 0
        got = torch.empty_like(whole)
        import ctypes as ct
        ct.cdll.LoadLibrary("libcudart.so.12").cudaMemcpyAsync(ct.c_void_p(got.data_ptr()), ct.c_void_p(lib.fnb200VolumeData(shared)), ct.c_size_t(per * n[2] * 4), 3, ct.c_void_p(sp))
    stream.synchronize()
    assert torch.equal(got, whole), "shared volume differs on rank %d" % rank
    assert (float(mm3[0]), float(mm3[1])) == mm_whole
    assert lib.fnb200CommFreeVolume(comm, shared, sp) == 0
    # the same fused path on rows 4 does not divide (the any-width four-point entry: groups straddle rows, so a group's peer
    # stores may be scalar at the slab's end) -- forced onto the large-call path, the volume is small
    os.environ["FASTNOISE2_B200_FILL"] = "0"
    os.environ["FASTNOISE2_B200_RAGGED"] = "2"
    m = (130, 7, 2 * world + 1)
    per2 = m[0] * m[1]
    b2, e2 = slab_range(m[2], rank, world)
    shared2 = lib.fnb200CommAllocVolume(comm, per2 * m[2], sp)
    assert shared2, fn.last_error()
    with torch.cuda.stream(stream):
        assert lib.fnb200MinMaxReset(acc.data_ptr(), sp) == 0
        assert lib.fnb200GenUniformGrid3DShared(g.handle, shared2, 0.0, 0.0, 0.0, m[0], m[1], m[2], 2.7, 2.7, 2.7, 1337, b2, e2, acc.data_ptr(), sp) == 0, fn.last_error()
        assert lib.fnb200CommAllReduceMinMax(comm, acc.data_ptr(), sp) == 0
        mm4 = np.zeros(2, np.float32)
        assert lib.fnb200MinMaxRead(acc.data_ptr(), sp, mm4.ctypes.data) == 0
        got2 = torch.empty(per2 * m[2], dtype=torch.float32, device="cuda")
        ct.cdll.LoadLibrary("libcudart.so.12").cudaMemcpyAsync(ct.c_void_p(got2.data_ptr()), ct.c_void_p(lib.fnb200VolumeData(shared2)), ct.c_size_t(per2 * m[2] * 4), 3, ct.c_void_p(sp))
    stream.synchronize()
    whole2 = torch.empty(per2 * m[2], dtype=torch.float32, device="cuda")
    mm_whole2 = g.grid3d(whole2, 0.0, 0.0, 0.0, *m, 2.7, 2.7, 2.7, 1337)
    assert torch.equal(got2, whole2), "any-width shared volume differs on rank %d" % rank
    assert (float(mm4[0]), float(mm4[1])) == mm_whole2
    assert lib.fnb200CommFreeVolume(comm, shared2, sp) == 0
    os.environ.pop("FASTNOISE2_B200_FILL", None)
    os.environ.pop("FASTNOISE2_B200_RAGGED", None)
    assert lib.fnb200CommDestroy(comm) == 0
    dist.destroy_process_group()
    open(os.path.join(tmp, "rank%d.ok" % rank), "w").write("ok")


@needs_two
def test_nccl_minmax_all_reduce_and_slab_gather(fn, tmp_path):
    import torch.multiprocessing as mp
    world = min(gpu_count(), 4)
    port = 29500 + os.getpid() % 1000
    mp.spawn(_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(str(tmp_path / ("rank%d.ok" % r))) for r in range(world))
