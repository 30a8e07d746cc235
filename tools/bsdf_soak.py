"""Soak run of the kernel-level BSDF probe: python tools/bsdf_soak.py N_MATERIALS_PER_KIND  (device arn_selftest_bsdf vs the oracle's probe)."""
import os, sys, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from arendur_b200 import api, _lib as L
import oracle_lib as O
import test_gpu_round2 as T
nm = int(sys.argv[1]); n = 2048
ctx = api.Context(0); olib = O.load()
rng = np.random.default_rng(777)
bad = 0; t0 = time.time()
for kind in range(4):
    for rep in range(nm):
        m = T._probe_material(rng, kind)
        def dirs():
            v = rng.normal(size=(n, 3)); v[: n // 8, 2] *= 1e-3
            return np.ascontiguousarray(v / np.linalg.norm(v, axis=1, keepdims=True), np.float32)
        wo, wi = dirs(), dirs()
        wi[n // 2: n // 2 + 256] = wo[n // 2: n // 2 + 256] * np.float32([-1, -1, 1])
        u = np.ascontiguousarray(rng.uniform(0, 1, (n, 2)), np.float32)
        ns = rng.normal(size=(n, 3)); ns /= np.linalg.norm(ns, axis=1, keepdims=True)
        ng = ns + rng.normal(scale=0.2, size=(n, 3)); ng /= np.linalg.norm(ng, axis=1, keepdims=True)
        dpdu = np.cross(ng, rng.normal(size=(n, 3))) * rng.uniform(0.1, 5, (n, 1))
        fr = np.ascontiguousarray(np.concatenate([dpdu, ns, ng], 1), np.float32)
        g = np.zeros((n, 12), np.float32); o = np.zeros((n, 12), np.float32)
        assert ctx.lib.arn_selftest_bsdf(ctx.c, C.byref(m), n, wo.ctypes.data, u.ctypes.data, wi.ctypes.data, fr.ctypes.data, g.ctypes.data) == 0
        for i in range(n):
            olib.arn_oracle_bsdf_probe2(C.byref(m), wo[i].ctypes.data, u[i].ctypes.data, wi[i].ctypes.data, fr[i].ctypes.data, o[i].ctypes.data)
        same = (g.view(np.uint32) == o.view(np.uint32)) | (np.isnan(g) & np.isnan(o))
        k = int((~same.all(axis=1)).sum())
        if k:
            bad += 1; i = int(np.argwhere(~same.all(axis=1))[0, 0])
            print(f"kind {kind} material {rep} (rough {m.roughness} eta {m.eta} dissolve {m.dissolve} sigma {m.sigma}): {k} probes differ, first wo {wo[i]} u {u[i]} wi {wi[i]}\n gpu {g[i]}\n ora {o[i]}", flush=True)
print(f"bsdf soak: {4 * nm} materials x {n} probes, {bad} materials with differences, {time.time() - t0:.1f} s")
