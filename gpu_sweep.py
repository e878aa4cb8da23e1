import sys; sys.path.insert(0,'/root/repo')
import numpy as np, fastnoise2_b200 as fn
from tests import helpers, trees
ref = helpers.load_ref("strict_avx2"); lib = fn.lib
F=lambda a: a.ctypes.data
def gen(l, h, dim, n, step, seed):
    out = np.empty(n**dim, np.float32); mm = np.zeros(2, np.float32)
    if dim==2: l.fnGenUniformGrid2D(h, F(out), -3.3, 2.1, n,n, step,step, seed, F(mm))
    elif dim==3: l.fnGenUniformGrid3D(h, F(out), -3.3, 2.1, 0.7, n,n,n, step,step,step, seed, F(mm))
    else: l.fnGenUniformGrid4D(h, F(out), -3.3,2.1,0.7,-1.9, n,n,n,n, step,step,step,step, seed, F(mm))
    return out, mm
bad = 0
for hybrid in (False, True):
  for node_id in range(47):
    if node_id in trees.UNSUPPORTED_IDS: continue
    for dim, n in ((2, 96), (3, 24), (4, 10)):
        a = trees.Tree(lib); b = trees.Tree(ref)
        ha = a.wire(node_id, hybrid); hb = b.wire(node_id, hybrid)
        got, mg = gen(lib, ha, dim, n, 1.9, 1337); want, mw = gen(ref, hb, dim, n, 1.9, 1337)
        if not (mg[0] <= mg[1]):
            print("FAIL", node_id, lib.fnGetMetadataName(node_id).decode(), dim, hybrid, fn.last_error()[:300]); bad += 1; continue
        c = helpers.compare(got, want)
        nanmatch = np.array_equal(np.isnan(got), np.isnan(want))
        flag = "" if (c["identical"] == 1.0) else ("  <-- DIFF" if c["over_tol"] or not nanmatch else "  (ulp)")
        if flag: bad += 1
        print(f"{node_id:2d} {lib.fnGetMetadataName(node_id).decode():30s} {dim}D hyb={int(hybrid)} ident {c['identical']*100:7.3f}% maxabs {c['max_abs']:.3g} nan {int(np.isnan(got).sum())}/{int(np.isnan(want).sum())}{flag}", flush=True)
        a.close(); b.close()
print("non-identical cases:", bad)
