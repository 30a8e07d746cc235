"""Full cb.json configuration (1024 x 768, 1024 spp) on the GPU against the reference's cornellbox.png fixture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arendur_b200 import api, scenes
ref = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cornellbox_ref_64x48.npy")).astype(np.float64)
hs, cam, film, smp, prm = scenes.cornell_scene(1024, 768, 32, 32)
ctx = api.Context(0); sc = ctx.upload(hs.desc())
f, st = sc.render_pt(cam, film, smp, prm)
_, rgb8 = api.film_finalize(f)
mine = rgb8.astype(np.float64).reshape(48, 16, 64, 16, 3).mean(axis=(1, 3))
rel = np.abs(mine.mean((0, 1)) - ref.mean((0, 1))) / ref.mean((0, 1))
err = np.abs(mine - ref) / np.maximum(ref, 50.0)
print(f"render {st.gpu_ms:.0f} ms, invalid {st.invalid_samples}; mean colour rel diff {rel}; block err median {np.median(err):.4f} p99 {np.quantile(err, 0.99):.4f} max {err.max():.4f}; "
      f"abs diff (8-bit units) mean {np.abs(mine-ref).mean():.3f} max {np.abs(mine-ref).max():.2f}")
