import numpy as np, sys
sys.path.insert(0,'.')
from collisiondetection_b200 import api
g = np.load('tests/golden/alec_prob17_30957.npz')
ctx = api.Context(0)
vf, ee = ctx.findCollisionCandidatesStep(13, g["faces"], g["q0"], g["q1"], 1e-8)
print(len(vf), len(ee))
