"""Point / Spot / Distant lights: oracle known answers from the reference's formulas
(lighting/pointlights.rs, lighting/distantlight.rs — the reference has no tests or fixtures for them, so
these values are derived by hand from the cited lines), host construction, JSON ingest."""
import ctypes as C
import json
import os
import math

import numpy as np
import pytest

import oracle_lib as O
from arendur_b200 import api, scenes, _lib as L

f32 = np.float32


def _sample(light, pos):
    out = np.zeros(7, f32)
    p = np.asarray(pos, f32)
    O.load().arn_oracle_analytic_sample(C.byref(light), p.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out[:3], out[3], out[4:7]


def test_point_light_known_answers():
    l = api.point_light((0, 2, 0), (8, 4, 2))
    rad, pdf, pfrom = _sample(l, (0, 0, 0))
    assert pdf == 1.0 and np.array_equal(pfrom, f32([0, 2, 0]))
    assert np.array_equal(rad, f32([2, 1, 0.5]))                      # I / |pto - pfrom|^2, pointlights.rs:53
    rad, _, _ = _sample(l, (3, 2, 4))
    assert np.array_equal(rad, f32([8, 4, 2]) / f32(25))
    # power = I * (pi * 4) -> luminance (pointlights.rs:79-81, spectrum to_xyz().y)
    p = f32([8, 4, 2]) * (f32(math.pi) * f32(4))
    y = f32(0.212671) * p[0] + f32(0.715160) * p[1] + f32(0.072169) * p[2]
    assert O.load().arn_oracle_analytic_power_y(C.byref(l)) == y


def test_spot_light_known_answers():
    total, start = math.radians(40), math.radians(20)
    l = api.spot_light((0, 3, 0), (0, -1, 0), (9, 9, 9), total, start)
    assert l.cost == f32(math.cos(f32(total))) and l.cosf == f32(math.cos(f32(start)))
    # parent_local = rotation(towards -> +z) * translation(+pos) (pointlights.rs:108-112; sic: +pos, so the light
    # itself does not land on the local origin — only transform_vector, which ignores it, is ever used)
    m = np.array(l.parent_local, f32).reshape(4, 4).T                 # column-major -> rows
    assert np.allclose(m[:3, :3] @ f32([0, -1, 0]), [0, 0, 1], atol=1e-6)
    assert np.allclose(m[:3, 3], m[:3, :3] @ f32([0, 3, 0]), atol=1e-6)
    assert np.allclose(m[:3, :3] @ m[:3, :3].T, np.eye(3), atol=1e-6)
    # on the axis: falloff 1, radiance I / d^2
    rad, pdf, pfrom = _sample(l, (0, 0, 0))
    assert pdf == 1.0 and np.allclose(rad, 1.0, rtol=1e-6)
    # outside the cone (60 degrees off axis): black
    rad, _, _ = _sample(l, (3 * math.tan(math.radians(60)), 0, 0))
    assert np.array_equal(rad, f32([0, 0, 0]))
    # inside the falloff band (30 degrees): ((cos30 - cost) / (cosf - cost))^4 * I / d^2   (:147-158,186-187)
    x = 3 * math.tan(math.radians(30))
    rad, _, _ = _sample(l, (x, 0, 0))
    d = ((math.cos(math.radians(30)) - math.cos(total)) / (math.cos(start) - math.cos(total))) ** 4
    assert np.allclose(rad, 9 * d / (x * x + 9), rtol=2e-5)
    # power = I * 2 pi * (1 - 0.5 (cosf - cost))   (:222-226)
    y = 9 * 2 * math.pi * (1 - 0.5 * (math.cos(start) - math.cos(total)))
    assert abs(O.load().arn_oracle_analytic_power_y(C.byref(l)) / y - 1) < 1e-6
    with pytest.raises(api.ArnError):
        api.spot_light((0, 0, 0), (0, 0, 1), (1, 1, 1), 0.2, 0.3)     # assert!(total_angle > start_falloff_angle)
    # degenerate arcs of Quaternion::from_arc: already aligned, and opposite
    for towards in ((0, 0, 1), (0, 0, -1), (0.3, -0.2, 0.9)):
        s = api.spot_light((1, 2, 3), towards, (1, 1, 1), 1.0, 0.5)
        m = np.array(s.parent_local, f32).reshape(4, 4).T
        t = f32(towards) / np.linalg.norm(f32(towards))
        assert np.allclose(m[:3, :3] @ t, [0, 0, 1], atol=1e-6), towards


def test_distant_light_known_answers():
    l = api.distant_light((3, 2, 1), (0, -2, 0), 10.0)
    assert np.array_equal(np.array(l.dir, f32), f32([0, -1, 0]))       # normalised by ::new
    rad, pdf, pfrom = _sample(l, (1, 1, 1))
    assert pdf == 1.0 and np.array_equal(rad, f32([3, 2, 1]))
    assert np.array_equal(pfrom, f32([1, 21, 1]))                     # pos + (-2 r) dir   (distantlight.rs:70)
    p = f32([3, 2, 1]) * (f32(10) * f32(10) * f32(math.pi))
    y = f32(0.212671) * p[0] + f32(0.715160) * p[1] + f32(0.072169) * p[2]
    assert O.load().arn_oracle_analytic_power_y(C.byref(l)) == y


def _floor_scene(lights, res=65):
    """A 20 x 20 matte floor at y = 0 seen from straight above."""
    hs = api.HostScene()
    for l in lights:
        hs.add_light(l)
    mat = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.6, 0.7)))
    pos = f32([[-10, 0, -10], [10, 0, -10], [10, 0, 10], [-10, 0, 10]])
    hs.add_mesh(pos, np.uint32([0, 2, 1, 0, 3, 2]), mat)
    hs.build()
    # camera at (0, 5, 0) looking down -y: parent_view maps world to view space (view looks along +z)
    view_parent = np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, -1, 0, 0], [0, 5, 0, 1]], f32)    # columns: x, y, z axes of the camera, origin
    parent_view = np.linalg.inv(view_parent.T).T.astype(f32)
    cam = api.make_camera(parent_view.reshape(-1), (-1, -1, 1, 1), 0.1, 100.0, 0.2, res, res)
    return hs, cam, api.make_film(res, res), api.make_sampler(1, 1, 8, 0), api.make_pt_params(max_depth=1)


def test_scene_light_order_and_distribution():
    lights = [api.point_light((0, 2, 0), (8, 4, 2)), api.spot_light((1, 3, 0), (0, -1, 0), (5, 5, 5), 0.8, 0.4), api.distant_light((1, 1, 1), (0, -1, 0), 20.0)]
    hs, cam, film, smp, prm = scenes.cornell_scene(32, 24, 1, 1, lights=lights)
    d = hs.desc()
    assert d.n_analytic_lights == 3 and d.n_lights == 5
    lp = np.ctypeslib.as_array(d.light_prims, (5,))
    assert list(lp[:3]) == [L.ARN_LIGHT_ANALYTIC | k for k in range(3)]          # the file's lights first (arencli.rs:95-98)
    assert all(p < d.n_prims for p in lp[3:])
    func = np.ctypeslib.as_array(d.light_func, (5,))
    for k in range(3):
        assert func[k] == O.load().arn_oracle_analytic_power_y(C.byref(lights[k]))
    cdf = np.ctypeslib.as_array(d.light_cdf, (6,))
    assert cdf[0] == 0 and abs(cdf[-1] - 1) < 1e-6 and np.all(np.diff(cdf) >= 0)


def test_oracle_point_light_on_a_floor_known_answer():
    """max_depth 1, one delta light: L = kd/pi * I/d^2 * cos(theta) (scene.rs:95-115, light pdf and selection pdf = 1)."""
    I = f32([8, 4, 2])
    hs, cam, film, smp, prm = _floor_scene([api.point_light((0, 2, 0), I)])
    osc = O.OracleScene(hs.desc())
    _, rad = osc.render_pt_samples(cam, film, smp, prm)
    _, st, _ = osc.render_pt(cam, film, smp, prm)
    c = rad[32, 32, 0, :3]                                               # centre pixel: hit ~ (0, 0, 0), d = 2, cos = 1
    want = f32([0.5, 0.6, 0.7]) / math.pi * I / 4.0
    assert np.allclose(c, want, rtol=2e-3), (c, want)
    assert st.shadow_rays == 65 * 65 and st.mis_rays == 0               # delta light: no BSDF-sampled light ray
    # a corner pixel obeys the same formula with its own distance and cosine
    x = rad[0, 0, 0, :3]
    assert 0 < x[0] < c[0]
    osc.close()


def test_oracle_distant_light_uses_the_mis_weight():
    """DistantLight is not `is_delta()` (LIGHT_INFINITE), so the light sample is weighted by
    power_heuristic(1, spdf) = 1 / (1 + spdf^2) (scene.rs:116-124) and nothing comes back from the BSDF half."""
    hs, cam, film, smp, prm = _floor_scene([api.distant_light((3, 3, 3), (0, -1, 0), 50.0)])
    osc = O.OracleScene(hs.desc())
    _, rad = osc.render_pt_samples(cam, film, smp, prm)
    _, st, _ = osc.render_pt(cam, film, smp, prm)
    spdf = 1 / math.pi                                                   # Lambert pdf at normal incidence
    want = f32([0.5, 0.6, 0.7]) / math.pi * 3.0 * (1 / (1 + spdf * spdf))
    assert np.allclose(rad[32, 32, 0, :3], want, rtol=2e-3)
    assert st.mis_rays == 0                                              # Lambert is not specular: stops at lpdf == 0
    osc.close()


def test_json_lights(tmp_path):
    src = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mini_scene", "scene.json")))
    ident = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    spot = api.spot_light((0.5, 2.5, 1.0), (0, -1, 0.2), (6, 6, 5), 0.9, 0.5)
    pl = np.array(spot.parent_local, f32).reshape(4, 4)
    src["lights"] = [
        {"Point": {"posw": {"x": 0.0, "y": 1.5, "z": 0.5}, "intensity": {"inner": {"x": 3.0, "y": 2.0, "z": 1.0}}}},
        {"Spot": {"posw": [0.5, 2.5, 1.0], "intensity": {"inner": [6.0, 6.0, 5.0]}, "cost": float(spot.cost), "cosf": float(spot.cosf),
                  "local_parent": ident, "parent_local": [[float(v) for v in col] for col in pl]}},
        {"Distant": {"intensity": {"inner": [0.5, 0.5, 0.5]}, "dir": {"x": 0.0, "y": -1.0, "z": 0.0}, "world_center": [0, 0, 0], "world_radius": 12.0}},
    ]
    import shutil
    for f in ("room.obj", "room.mtl"):
        shutil.copy(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mini_scene", f), tmp_path / f)
    p = tmp_path / "scene.json"
    p.write_text(json.dumps(src))
    hs = api.HostScene()
    cam, film, smp, prm, out = hs.load_json(p, base_dir=tmp_path)
    hs.build()
    d = hs.desc()
    assert d.n_analytic_lights == 3 and d.n_lights == 4
    a = d.analytic_lights
    assert (a[0].type, a[1].type, a[2].type) == (L.ARN_LIGHT_POINT, L.ARN_LIGHT_SPOT, L.ARN_LIGHT_DISTANT)
    assert list(a[0].pos) == [0.0, 1.5, 0.5] and list(a[0].intensity) == [3.0, 2.0, 1.0]
    assert bytes(a[1])[28:] == bytes(spot)[28:] and list(a[1].pos) == [0.5, 2.5, 1.0]
    assert a[2].world_radius == 12.0 and list(a[2].dir) == [0.0, -1.0, 0.0]
    src["lights"] = [{"Area": "x"}]
    p.write_text(json.dumps(src))
    with pytest.raises(api.ArnError):
        api.HostScene().load_json(p, base_dir=tmp_path)
