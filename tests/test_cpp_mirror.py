"""C++ mirror API (arendur_b200/csrc/host/arendur.hpp) and the arencli equivalent: compile on CPU, run on GPU."""
import os
import subprocess

import pytest

from arendur_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.dirname(L.LIB_PATH)


def _compile(src, out):
    cmd = ["g++", "-std=c++17", "-O1", "-o", out, src, f"-L{LIBDIR}", "-larn_b200", f"-Wl,-rpath,{LIBDIR}", "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)


def test_mirror_header_and_arencli_compile(tmp_path):
    _compile(os.path.join(ROOT, "tests", "cpp", "test_mirror_api.cpp"), str(tmp_path / "t"))
    _compile(os.path.join(ROOT, "examples", "arencli.cpp"), str(tmp_path / "arencli"))
    # without a scene file arencli prints its usage and exits 2 (no GPU needed)
    assert subprocess.run([str(tmp_path / "arencli")], capture_output=True).returncode == 2


@pytest.mark.gpu
def test_mirror_api_on_gpu(tmp_path):
    exe = str(tmp_path / "t")
    _compile(os.path.join(ROOT, "tests", "cpp", "test_mirror_api.cpp"), exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "mirror API tests: OK" in r.stdout, r.stdout + r.stderr


MINI = os.path.join(ROOT, "tests", "golden", "mini_scene")


def test_mini_scene_json_loads():
    """arencli-format JSON + OBJ/MTL fixture through the host loader (no GPU)."""
    from arendur_b200 import api
    hs = api.HostScene()
    cam, film, smp, prm, out = hs.load_json(os.path.join(MINI, "scene.json"), base_dir=MINI)
    d = hs.build()
    assert (d.n_triangles, d.n_spheres, d.n_lights, d.n_meshes) == (6, 1, 1, 3)
    assert (film.res_x, film.res_y, smp.sampledx, prm.max_depth, prm.min_depth) == (128, 96, 2, 6, 3)
    assert [d.materials[i].type for i in range(d.n_materials)] == [L.ARN_MAT_MATTE, L.ARN_MAT_PLASTIC, L.ARN_MAT_GLASS, L.ARN_MAT_MATTE, L.ARN_MAT_MATTE]
    assert out == "mini_scene.png"


@pytest.mark.gpu
def test_arencli_renders_json_scene_like_the_oracle(tmp_path):
    """The CLI renders the fixture on the GPU; its PNG equals the oracle's render of the same scene
    (8-bit, at most 1 level off on a handful of pixels: summation order)."""
    import numpy as np
    from PIL import Image
    import oracle_lib as O
    from arendur_b200 import api
    exe = str(tmp_path / "arencli")
    _compile(os.path.join(ROOT, "examples", "arencli.cpp"), exe)
    out = str(tmp_path / "o.png")
    r = subprocess.run([exe, os.path.join(MINI, "scene.json"), "-t", "4", "-o", out], capture_output=True, text=True, cwd=MINI)
    assert r.returncode == 0 and "Done! Time used:" in r.stdout, r.stdout + r.stderr
    img = np.asarray(Image.open(out)).astype(int)
    hs = api.HostScene()
    cam, film, smp, prm, _ = hs.load_json(os.path.join(MINI, "scene.json"), base_dir=MINI)
    osc = O.OracleScene(hs.build())
    of, st, _ = osc.render_pt(cam, film, smp, prm)
    _, o8 = O.film_finalize(of)
    assert img.shape == (96, 128, 3) and img.max() > 30
    diff = np.abs(img - o8.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.01
