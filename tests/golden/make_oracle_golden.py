"""Regression vectors of the ORACLE itself (not of the reference): the film and the closest hits it produces for the
committed Cornell fixture.  They pin the restatement across refactors — a change of the oracle would silently move
the target every GPU parity test compares against.  Run from the repository root: python tests/golden/make_oracle_golden.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from arendur_b200 import scenes


def compute():
    hs, cam, film, smp, prm = scenes.cornell_scene(64, 48, 2, 2)
    osc = O.OracleScene(hs.desc())
    f, st, _ = osc.render_pt(cam, film, smp, prm, nthreads=1)           # one thread: a fixed order of the film additions
    _, rad = osc.render_pt_samples(cam, film, smp, prm, nthreads=1)
    xs, ys = np.meshgrid(np.arange(0, 64, 2) + 0.5, np.arange(0, 48, 2) + 0.5, indexing="xy")
    pf = np.zeros((xs.size, 4), np.float32); pf[:, 0], pf[:, 1] = xs.reshape(-1), ys.reshape(-1)
    hits = osc.intersect_closest(O.camera_rays(cam, pf))
    osc.close()
    return dict(film=f, radiance=rad, prim_id=hits["prim_id"].copy(), t=hits["t"].copy(),
                rays=np.array([st.camera_rays, st.extend_rays, st.shadow_rays, st.mis_rays, st.invalid_samples], np.int64))


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_cornell_64x48.npz")
    np.savez_compressed(out, **compute())
    print("wrote", out, os.path.getsize(out), "bytes")
