"""One BSDF probe on the device and in the oracle for the inputs of a traced sample (debugging aid)."""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from arendur_b200 import api, _lib as L
import oracle_lib as O
import test_gpu_round2 as T
hs, cam, film, smp, prm = T._random_scene(7)
d = hs.desc()
m = C.cast(d.materials, C.POINTER(L.Material))[1]
wo = np.float32([[0.462853849, -0.0552956164, 0.884708285]])
fr = np.float32([[0.901513219, -0.816214204, -0.499537945, 0.707514226, 0.691110313, 0.147615314, 0.707514167, 0.691110253, 0.147615269]])
u = np.float32([[0.364191055, 0.073315382]])
wi = np.float32([[0.391029, 0.50922662, -0.766671062]])
ctx = api.Context(0)
g = np.zeros((1, 12), np.float32); o = np.zeros((1, 12), np.float32)
rc = ctx.lib.arn_selftest_bsdf(ctx.c, C.byref(m), 1, wo.ctypes.data, u.ctypes.data, wi.ctypes.data, fr.ctypes.data, g.ctypes.data)
O.load().arn_oracle_bsdf_probe2(C.byref(m), wo.ctypes.data, u.ctypes.data, wi.ctypes.data, fr.ctypes.data, o.ctypes.data)
np.set_printoptions(precision=9, linewidth=200)
print("gpu   ", g[0]); print("oracle", o[0]); print("same bits", (g.view(np.uint32) == o.view(np.uint32))[0])
