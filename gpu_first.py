"""First GPU contact: every config tree, small sizes, GPU vs strict oracle."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import fastnoise2_b200 as fn
from tests import helpers
print("device available:", fn.device_available(), torch.cuda.get_device_name(0))
ref = helpers.load_ref("strict_avx2")
print("ref strict:", ref is not None)
def report(name, got, mm, want, wmm, dt):
    c = helpers.compare(got, want)
    print(f"{name:34s} ident {c['identical']*100:7.3f}% maxabs {c['max_abs']:.3g} over2e-5 {c['over_tol']} n {c['n']}  mm gpu {mm} ref {tuple(float(x) for x in wmm)}  {dt*1e3:.1f} ms", flush=True)
cases = [("c1_simplex", 3, 1.0), ("c1_simplex", 3, 8.0), ("c3a_cellular_distance", 3, 1.0), ("c3b_cellular_lookup_white", 3, 1.0),
         ("c3c_cellular_lookup_gradient", 3, 1.0), ("c4_terrain", 3, 1.0), ("c4_terrain", 3, 4.0)]
for name, d, step in cases:
    g = fn.Generator(fn.TREES[name]); r = helpers.RefNode(ref, fn.TREES[name])
    n = 96
    out = np.empty(n**3, np.float32)
    t = time.time(); mm = g.grid3d(out, 0,0,0, n,n,n, step,step,step, 1337); dt = time.time()-t
    t = time.time(); mm = g.grid3d(out, 0,0,0, n,n,n, step,step,step, 1337); dt = time.time()-t
    want, wmm = r.grid3d(0,0,0, n,n,n, step,step,step, 1337)
    report(f"{name} step {step}", out, mm, want, wmm, dt)
for name in ["fbm_simplex_readme", "c2_fbm5_perlin"]:
    g = fn.Generator(fn.TREES[name]); r = helpers.RefNode(ref, fn.TREES[name])
    n = 1024; out = np.empty(n*n, np.float32)
    mm = g.grid2d(out, 0,0, n,n, 0.25,0.25, 1337); t=time.time(); mm = g.grid2d(out, 0,0, n,n, 0.25,0.25, 1337); dt=time.time()-t
    want, wmm = r.grid2d(0,0, n,n, 0.25,0.25, 1337)
    report(name+" grid2d", out, mm, want, wmm, dt)
g = fn.Generator(fn.TREES["c2_fbm5_perlin"]); r = helpers.RefNode(ref, fn.TREES["c2_fbm5_perlin"])
n = 512; out = np.empty(n*n, np.float32)
mm = g.tileable2d(out, n,n, 1,1, 1337); t=time.time(); mm = g.tileable2d(out, n,n, 1,1, 1337); dt=time.time()-t
want, wmm = r.tileable2d(n,n,1,1,1337)
report("c2 tileable", out, mm, want, wmm, dt)
rng = np.random.default_rng(1)
pos = [rng.uniform(-1000,1000, 1000003).astype(np.float32) for _ in range(4)]
g = fn.Generator(fn.TREES["c5_supersimplex"]); r = helpers.RefNode(ref, fn.TREES["c5_supersimplex"])
out = np.empty(len(pos[0]), np.float32)
mm = g.positions(out, pos, [0,0,0,0], 1337); t=time.time(); mm = g.positions(out, pos, [0,0,0,0], 1337); dt=time.time()-t
want, wmm = r.positions(pos, [0,0,0,0], 1337)
report("c5 positions4d", out, mm, want, wmm, dt)
# device-pointer throughput, quick look
for name, kind in [("c1_simplex",3),("c3a_cellular_distance",3),("c4_terrain",3)]:
    g = fn.Generator(fn.TREES[name]); n=256
    o = torch.empty(n**3, dtype=torch.float32, device='cuda')
    g.grid3d(o, 0,0,0,n,n,n,1,1,1,1337); torch.cuda.synchronize()
    t=time.time()
    for _ in range(5): g.grid3d(o, 0,0,0,n,n,n,1,1,1,1337)
    torch.cuda.synchronize(); dt=(time.time()-t)/5
    print(f"{name} 256^3 device out: {dt*1e3:.2f} ms = {n**3/dt/1e6:.0f} Mpts/s")
for mode in range(4):
    print("alu peak mode", mode, fn.lib.fnb200MeasureAluPeak(mode, 2000)/1e12, "T lane-ops/s")
print("launches", fn.lib.fnb200LaunchCount())
