"""Generates tests/golden/cornellbox_ref_64x48.npy from the reference's only result artefact,
/root/reference/cornellbox.png (1024x768, 8-bit RGB, "1024 spp", unseeded RNG, no gamma):
box-filtered 16x down to 64x48, float32 in [0, 255].  Run in the build container:
    python tests/golden/make_png_fixture.py
"""
import os
import numpy as np
from PIL import Image

src = "/root/reference/cornellbox.png"
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cornellbox_ref_64x48.npy")
a = np.asarray(Image.open(src).convert("RGB"), dtype=np.float64)
assert a.shape == (768, 1024, 3)
small = a.reshape(48, 16, 64, 16, 3).mean(axis=(1, 3)).astype(np.float32)
np.save(out, small)
print("wrote", out, small.shape, small.mean(axis=(0, 1)))
