import numpy as np, sys
sys.path.insert(0,'.')
from collisiondetection_b200 import api
from oracle import bind
g = np.load('tests/golden/testctcd.npz'); r = np.load('tests/golden/prims_random.npz')
ctx = api.Context(0); port = bind.Port()
for name, fn in (("ee", ctx.edgeEdgeCTCD), ("vf", ctx.vertexFaceCTCD), ("ve", ctx.vertexEdgeCTCD), ("vv", ctx.vertexVertexCTCD)):
    hit, t = fn(g[name + "_pts"], 1e-6)
    print(name, hit, t, g["ref_"+name+"_t"])
    hit, t = fn(r[name + "_pts"], r[name+"_eta"]); ph, pt = getattr(port, name+"_batch")(r[name + "_pts"], r[name+"_eta"])
    bad = np.nonzero(hit != ph)[0]
    print("  random: hits gpu %d port %d mismatches %d first %s"%(hit.sum(), ph.sum(), len(bad), bad[:5]))
for deg in (2,3,4,6):
    rng = np.random.default_rng(deg); co = rng.uniform(-1,1,(2000,deg+1))
    for pos in (True, False):
        cnt, lo, hi = ctx.findIntervals(co, deg, pos)
        import ctypes as C
        dp = C.POINTER(C.c_double); nb=0
        for i in range(2000):
            op = co[i].copy(); l=np.zeros(8); u=np.zeros(8)
            k = port.lib.orc_find_intervals(op.ctypes.data_as(dp), deg, int(pos), l.ctypes.data_as(dp), u.ctypes.data_as(dp))
            if k != cnt[i] or not np.array_equal(l[:k], lo[i,:k]): nb+=1
        print("find_intervals deg", deg, pos, "bad", nb)
