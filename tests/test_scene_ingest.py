"""Host layer (product): scene ingest and flattening vs the committed fixture and the oracle.
Covers TriangleMesh::from_model_transformed, load_obj's tobj semantics and material choice,
arencli's JSON reader, PerspecCam::new, Scene::new's light distribution."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import oracle_lib as O
from arendur_b200 import api, scenes, _lib as L

REF = "/root/reference/examples/cornellbox"
have_ref = os.path.exists(os.path.join(REF, "cb.json"))
MINI_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mini_scene")


def _desc_arrays(d):
    return dict(
        positions=np.ctypeslib.as_array(d.positions, shape=(d.n_vertices, 3)).copy(),
        normals=np.ctypeslib.as_array(d.normals, shape=(d.n_vertices, 3)).copy() if d.normals else None,
        uvs=np.ctypeslib.as_array(d.uvs, shape=(d.n_vertices, 2)).copy() if d.uvs else None,
        indices=np.ctypeslib.as_array(d.indices, shape=(d.n_triangles, 3)).copy(),
        tri_mesh=np.ctypeslib.as_array(d.tri_mesh, shape=(d.n_triangles,)).copy(),
        prims=np.ctypeslib.as_array(d.prims, shape=(d.n_prims,)).copy(),
        materials=bytes(C.string_at(d.materials, d.n_materials * C.sizeof(L.Material))),
        spheres=bytes(C.string_at(d.spheres, d.n_spheres * 176)),
        meshes=bytes(C.string_at(d.meshes, d.n_meshes * 16)),
    )


def test_cornell_fixture_scene_shape():
    hs, cam, film, smp, prm = scenes.cornell_scene(256, 256, 4, 4)
    d = hs.desc()
    assert (d.n_triangles, d.n_spheres, d.n_prims, d.n_meshes, d.n_lights) == (1112, 2, 1114, 8, 2)
    a = _desc_arrays(d)
    # component order of BASELINE.md C1: OBJ triangles, light_sphere, blue_sphere
    assert np.array_equal(a["prims"][:1112], np.arange(1112)) and a["prims"][1112] == L.ARN_PRIM_SPHERE and a["prims"][1113] == L.ARN_PRIM_SPHERE | 1
    mats = np.frombuffer(a["materials"], dtype=np.uint32).reshape(-1, 16)
    # sphere Plastic, shortBox Glass, floor Plastic, ceiling/backWall/leftWall/rightWall/light Matte, fallback Matte, sphere light Matte
    assert mats[:, 0].tolist() == [1, 2, 1, 0, 0, 0, 0, 0, 0, 0]
    # Scene::new power distribution (Y-power 994.6 / 685.1 -> pdf 0.592 / 0.408)
    assert abs(d.light_func[0] - 994.6) < 0.1 and abs(d.light_func[1] - 685.1) < 0.1
    assert abs(d.light_cdf[1] - 0.592) < 1e-3 and d.light_cdf[2] == 1.0
    assert prm.max_depth == 8 and prm.min_depth == 4 and abs(prm.rr_threshold - 0.05) < 1e-9


def test_mesh_transform_matches_oracle():
    """from_model_transformed with the projective cb.json matrix (quirk A-12): positions through
    the homogeneous divide, normals through the normalised inverse-transpose — vs the oracle."""
    z = np.load(scenes.CORNELL_FIXTURE)
    t = z["mesh_transform"]
    pos, nrm, idx = z["m0_positions"], z["m0_normals"], z["m0_indices"]
    hs = api.HostScene()
    m = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.5, 0.5)))
    hs.add_mesh(pos, idx, m, normals=nrm, transform=t)
    d = hs.build()
    a = _desc_arrays(d)
    op, on = np.zeros_like(pos), np.zeros_like(nrm)
    O.load().arn_oracle_mesh_transform(O._p(np.ascontiguousarray(t, np.float32)), pos.shape[0], O._p(pos), O._p(nrm), O._p(op), O._p(on))
    assert np.array_equal(a["positions"].view(np.uint32), op.view(np.uint32))
    assert np.array_equal(a["normals"].view(np.uint32), on.view(np.uint32))
    assert abs(np.linalg.norm(a["normals"], axis=1) - 1).max() < 1e-6


def test_camera_matches_oracle():
    for (w, h, screen, fov) in ((1024, 768, (-1.0, -0.75, 1.0, 0.7), 1.2707964), (1920, 1080, (-16 / 9, -1.0, 16 / 9, 1.0), math.pi / 2), (3840, 2160, (-1.0, -0.75, 1.0, 0.7), 1.2707964)):
        view = np.eye(4, dtype=np.float32); view[3, :3] = (0.5, -0.25, 3.0)
        for pv in (np.eye(4, dtype=np.float32), view):
            assert bytes(api.make_camera(pv, screen, 0.1, 1000.0, fov, w, h)) == bytes(O.camera_make(pv, screen, 0.1, 1000.0, fov, w, h))
    with pytest.raises(api.ArnError):
        api.make_camera(np.zeros((4, 4), np.float32), (-1, -1, 1, 1), 0.1, 1000.0, 1.0, 64, 64)     # "matrix inversion failure"
    with pytest.raises(api.ArnError):
        api.make_camera(np.eye(4), (-1, -1, 1, 1), 10.0, 1.0, 1.0, 64, 64)                            # assert!(znear < zfar)


def test_sphere_construction_matches_oracle():
    hs = api.HostScene()
    m = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.5, 0.5), sigma=3.0))
    hs.add_sphere(1.5, -2.0, 2.0, 6.28, m, emission=(15.5, 10.5, 5.5), transform=np.float32([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [-3, 0, -4.5, 1]]))
    hs.add_sphere(1.0, -0.5, 0.25, 9.0, m)            # clamps phimax to 2 pi, keeps z range
    d = hs.build()
    s0, s1 = d.spheres[0], d.spheres[1]
    o = L.Sphere(); O.load().arn_oracle_sphere_new(1.5, -2.0, 2.0, 6.28, C.byref(o))
    assert (s0.radius, s0.zmin, s0.zmax, s0.phimax, s0.thetamin, s0.thetamax) == (o.radius, o.zmin, o.zmax, o.phimax, o.thetamin, o.thetamax)
    assert s0.zmin == -1.5 and s0.zmax == 1.5 and s0.has_transform == 1 and s0.emissive == 1
    inv = np.zeros(16, np.float32)
    assert O.load().arn_oracle_m4_invert(O._p(np.ascontiguousarray(s0.local_parent, np.float32)), O._p(inv)) == 0
    assert np.array_equal(inv, np.array(s0.parent_local, np.float32))
    assert s1.phimax == np.float32(2 * np.float32(math.pi)) and s1.has_transform == 0 and s1.emissive == 0
    assert abs(O.load().arn_oracle_light_power_y(C.byref(s0)) - d.light_func[0]) == 0.0
    # sphere bounds through BBox3::apply_transform vs the oracle's ComponentInfo::new
    ob, oc = O.prim_bounds(d)
    assert np.array_equal(ob[0], np.float32([-4.5, -1.5, -6.0, -1.5, 1.5, -3.0])) and oc.tolist() == [2.0, 1.0]
    with pytest.raises(api.ArnError):
        hs.add_sphere(-1.0, -1, 1, 1, m)              # assert!(radius > 0)
    with pytest.raises(api.ArnError):
        hs.add_sphere(1.0, 0.5, 0.5, 1, m)            # assert!(zmin < zmax)


def test_obj_loader_small(tmp_path):
    """tobj semantics on a hand-written file: groups, usemtl split, (v,vt,vn) re-indexing, quad fan,
    negative indices, material choice rules of load_obj (component/mod.rs:118-164)."""
    (tmp_path / "t.mtl").write_text(
        "newmtl glass\nNs 10\nNi 1.5\nd 0.5\nillum 4\nKd 0.7 0.7 0.7\nKs 1 1 1\n"
        "newmtl shiny\nNs 200\nKd 0.1 0.2 0.3\nKs 0.5 0.5 0.5\n"
        "newmtl dull\nKd 0.4 0.4 0.4\nKs 0 0 0\n"
        "newmtl ghost\nd 0.25\nKd 1 1 1\nKs 0 0 0\n")
    (tmp_path / "t.obj").write_text(
        "mtllib t.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n"
        "g quad\nusemtl glass\nf 1/1/1 2/2/1 3/3/1 4/4/1\n"
        "g two\nusemtl shiny\nf 1 2 3\nusemtl dull\nf -4 -3 -2\n"
        "g nomtl\nusemtl missing\nf 1//1 3//1 4//1\n"
        "g t\nusemtl ghost\nf 2 3 4\n")
    hs = api.HostScene()
    assert hs.load_obj(tmp_path / "t.obj") == 6
    d = hs.build()
    a = _desc_arrays(d)
    assert d.n_meshes == 5 and d.n_triangles == 6
    mats = np.frombuffer(a["materials"], dtype=np.uint32).reshape(-1, 16)
    fm = np.frombuffer(a["materials"], dtype=np.float32).reshape(-1, 16)
    assert mats[:, 0].tolist() == [L.ARN_MAT_GLASS, L.ARN_MAT_PLASTIC, L.ARN_MAT_MATTE, L.ARN_MAT_TRANSLUCENT, L.ARN_MAT_MATTE]
    assert fm[0, 10] == np.float32(1.5) and fm[3, 11] == np.float32(0.25) and fm[1, 8] == np.float32(0.8)      # eta, dissolve, roughness (1000-200)/1000
    assert np.allclose(fm[4, 1:4], (0.5, 0.6, 0.7))                       # fallback material
    meshes = np.frombuffer(a["meshes"], dtype=np.uint32).reshape(-1, 4)
    assert meshes[:, 0].tolist() == [0, 1, 2, 4, 3] and meshes[:, 1].tolist() == [1, 0, 0, 1, 0] and meshes[:, 2].tolist() == [1, 0, 0, 0, 0]
    assert a["indices"][:2].tolist() == [[0, 1, 2], [0, 2, 3]]              # quad fanned from its first vertex
    with pytest.raises(api.ArnError):
        api.HostScene().load_obj(tmp_path / "missing.obj")


@pytest.mark.skipif(not have_ref, reason="reference tree not present (GPU box)")
def test_json_and_obj_loader_match_fixture():
    """arencli parse_input on the reference's cb.json + OBJ/MTL == the scene assembled from the
    committed fixture (which an independent Python parser produced)."""
    hs = api.HostScene()
    cam, film, smp, prm, out = hs.load_json(os.path.join(REF, "cb.json"), base_dir="/root/reference")
    hs.build()
    a = _desc_arrays(hs.desc())
    fs, fcam, ffilm, fsmp, fprm = scenes.cornell_scene(1024, 768, 32, 32)
    b = _desc_arrays(fs.desc())
    for k in ("positions", "normals", "uvs", "indices", "tri_mesh", "prims"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    assert a["meshes"] == b["meshes"] and a["spheres"] == b["spheres"]
    # material tables: same entries; the fixture scene appends the sphere material after the OBJ ones as well
    ma, mb = np.frombuffer(a["materials"], np.uint32).reshape(-1, 16), np.frombuffer(b["materials"], np.uint32).reshape(-1, 16)
    assert ma.shape == mb.shape
    assert np.array_equal(ma[:, :10], mb[:, :10])                      # type, kd, ks, sigma, roughness, alpha
    glass = ma[:, 0] == L.ARN_MAT_GLASS
    assert np.array_equal(ma[glass, 10], mb[glass, 10])                # eta only matters for Glass
    assert np.array_equal(hs.nodes(), fs.nodes()) and np.array_equal(hs.order(), fs.order())
    assert bytes(cam) == bytes(fcam)
    assert (film.res_x, film.res_y, film.crop_max_x, film.crop_max_y, film.filter_radius_x) == (1024, 768, 1024, 768, 4.0)
    assert (smp.sampledx, smp.sampledy, smp.ndim) == (32, 32, 8) and prm.max_depth == 8
    assert out.endswith("CornellBox-Glossy44.png")


def test_json_errors(tmp_path):
    p = tmp_path / "bad.json"
    p.write_text("{ not json")
    with pytest.raises(api.ArnError) as e:
        api.HostScene().load_json(p)
    assert e.value.code == L.ARN_E_IO
    p.write_text('{"lights": [{"Point": {}}], "components": []}')
    with pytest.raises(api.ArnError) as e:
        api.HostScene().load_json(p)
    assert e.value.code == L.ARN_E_INVALID          # malformed Point light


def test_json_transformed_components(tmp_path):
    """ComponentDesc::Transformed (examples/arencli.rs:162-181): an instance of a bare named sphere becomes a transformed
    sphere that is NOT in `lights` even when emissive; a missing original or a singular matrix skips the component (the
    reference prints and continues); an instance of an already transformed primitive is rejected (nested round trip)."""
    import json, shutil
    src = json.load(open(os.path.join(MINI_DIR, "scene.json")))
    for f in ("room.obj", "room.mtl"):
        shutil.copy(os.path.join(MINI_DIR, f), tmp_path / f)
    bulb = {"name": "bulb", "value": {"Shaped": {
        "shape": {"Sphere": {"radius": 0.3, "zmin": -1.0, "zmax": 1.0, "phimax": 6.28}},
        "material": {"name": "lamp_matte", "value": None},
        "light": {"name": "bulb_light", "value": {"Constant": {"value": {"inner": [4.0, 4.0, 4.0]}}}},
        "transform": None}}}
    move = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [-0.7, 1.0, 0.5, 1]]
    singular = [[1, 0, 0, 0], [0, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    src["components"] += [bulb,
                          {"name": "bulb2", "value": {"Transformed": {"transform": move, "original": "bulb"}}},
                          {"name": "ghost", "value": {"Transformed": {"transform": move, "original": "nobody"}}},
                          {"name": "flat", "value": {"Transformed": {"transform": singular, "original": "bulb"}}}]
    p = tmp_path / "scene.json"
    p.write_text(json.dumps(src))
    hs = api.HostScene()
    hs.load_json(p, base_dir=tmp_path)
    d = hs.build()
    assert d.n_spheres == 3                                   # lamp, bulb, bulb2 (ghost and flat are skipped)
    bulb_s, inst = d.spheres[1], d.spheres[2]
    assert bulb_s.has_transform == 0 and inst.has_transform == 1 and inst.emissive == 1
    assert list(inst.local_parent)[12:15] == [np.float32(-0.7), 1.0, 0.5] and inst.radius == bulb_s.radius
    assert d.n_lights == 2                                    # lamp and bulb: the instance is not a light
    lp = np.ctypeslib.as_array(d.light_prims, (2,))
    assert (d.prims[lp[0]] & 0x7fffffff, d.prims[lp[1]] & 0x7fffffff) == (0, 1)
    src["components"].append({"name": "lamp2", "value": {"Transformed": {"transform": move, "original": "lamp"}}})
    p.write_text(json.dumps(src))
    with pytest.raises(api.ArnError) as e:
        api.HostScene().load_json(p, base_dir=tmp_path)
    assert e.value.code == L.ARN_E_UNSUPPORTED


def test_oracle_only_flattening_equals_the_product_host_layer():
    """bench.py --impl reference assembles the Cornell scene with oracle functions only (tests/oracle_lib.OracleFlatScene);
    it must hand the oracle the very bytes the product's host layer (FlatScene, arn_bvh_build) hands the GPU."""
    import ctypes as C
    from arendur_b200 import _lib as L
    o = O.OracleFlatScene()
    d = o.desc
    hs, cam, film, smp, prm = scenes.cornell_scene(64, 48, 2, 2)
    p = hs.desc()

    def raw(ptr, nbytes):
        return C.string_at(C.cast(ptr, C.c_void_p), nbytes) if nbytes else b""
    for name in ("n_vertices", "n_triangles", "n_meshes", "n_spheres", "n_materials", "n_prims", "n_nodes", "n_lights", "n_analytic_lights"):
        assert getattr(d, name) == getattr(p, name), name
    for name, nbytes in (("positions", d.n_vertices * 12), ("normals", d.n_vertices * 12), ("uvs", d.n_vertices * 8), ("indices", d.n_triangles * 12),
                         ("tri_mesh", d.n_triangles * 4), ("prims", d.n_prims * 4), ("order", d.n_prims * 4), ("nodes", d.n_nodes * 32),
                         ("meshes", d.n_meshes * 16), ("spheres", d.n_spheres * C.sizeof(L.Sphere)), ("materials", d.n_materials * C.sizeof(L.Material)),
                         ("light_prims", d.n_lights * 4), ("light_func", d.n_lights * 4), ("light_cdf", (d.n_lights + 1) * 4)):
        assert raw(getattr(d, name), nbytes) == raw(getattr(p, name), nbytes), name
    assert d.light_func_integral == p.light_func_integral
    assert bytes(o.camera(64, 48)) == bytes(cam)
    assert o.max_depth() == prm.max_depth


def test_random_transforms_and_cameras_match_the_oracle():
    """Host-layer fuzz: `from_model_transformed` under random affine and projective matrices (rotation x non-uniform scale x shear,
    translation, a non-trivial w row) and `PerspecCam::new` / `OrthoCam::new` under random view matrices, screens, fovs, lenses —
    byte-identical to the oracle's restatement."""
    rng = np.random.default_rng(2024)
    for k in range(24):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        t = np.eye(4, dtype=np.float32)
        t[:3, :3] = (q * rng.uniform(0.2, 4.0, 3)) @ (np.eye(3) + np.triu(rng.normal(scale=0.3, size=(3, 3)), 1))
        t[3, :3] = rng.uniform(-5, 5, 3)
        if k % 3 == 0:
            t[:3, 3] = rng.uniform(-0.05, 0.05, 3); t[3, 3] = rng.uniform(0.5, 2.0)          # projective (cb.json's own matrix is, quirk A-12)
        n = 64
        pos = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
        nrm = rng.normal(size=(n, 3)); nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
        idx = rng.integers(0, n, 3 * 20).astype(np.uint32)
        hs = api.HostScene()
        m = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.5, 0.5)))
        hs.add_mesh(pos, idx, m, normals=nrm, transform=t)
        a = _desc_arrays(hs.build())
        op, on = np.zeros_like(pos), np.zeros_like(nrm)
        O.load().arn_oracle_mesh_transform(O._p(np.ascontiguousarray(t, np.float32)), n, O._p(pos), O._p(nrm), O._p(op), O._p(on))
        assert np.array_equal(a["positions"].view(np.uint32), op.view(np.uint32)), k
        assert np.array_equal(a["normals"].view(np.uint32), on.view(np.uint32)), k
        # cameras
        view = np.eye(4, dtype=np.float32); view[:3, :3] = q.astype(np.float32); view[3, :3] = rng.uniform(-3, 3, 3)
        w, h = int(rng.integers(8, 4000)), int(rng.integers(8, 2200))
        x0, y0 = float(rng.uniform(-2, -0.1)), float(rng.uniform(-2, -0.1))
        screen = (x0, y0, float(rng.uniform(0.1, 2)), float(rng.uniform(0.1, 2)))
        fov = float(rng.uniform(0.2, 2.8)); znear = float(rng.uniform(0.01, 1)); zfar = znear + float(rng.uniform(1, 1000))
        lens = (float(rng.uniform(0.01, 0.5)), float(rng.uniform(0.5, 20))) if k % 2 else None
        assert bytes(api.make_camera(view, screen, znear, zfar, fov, w, h, lens=lens)) == bytes(O.camera_make(view, screen, znear, zfar, fov, w, h, lens=lens)), k
        assert bytes(api.make_ortho_camera(view, screen, znear, zfar, w, h, lens=lens)) == bytes(O.ortho_camera_make(view, screen, znear, zfar, w, h, lens=lens)), k


def test_random_spheres_bounds_and_light_powers_match_the_oracle():
    """Host-layer fuzz: `Sphere::new` (z clamps, theta range, phimax clamp), the inverse of random affine instance transforms,
    `ComponentInfo::new` bounds / costs of every primitive (triangles and transformed, clipped spheres) and the light power that
    feeds the light-selection CDF — the product's flattened scene against the oracle's own computation, bit for bit."""
    rng = np.random.default_rng(31337)
    lib = O.load()
    for rep in range(6):
        hs = api.HostScene()
        m = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.5, 0.5)))
        pos = rng.uniform(-3, 3, (30, 3)).astype(np.float32)
        hs.add_mesh(pos, np.arange(30, dtype=np.uint32), m)
        params = []
        for k in range(10):
            rad = float(rng.uniform(0.05, 3.0))
            zmin, zmax = sorted(rng.uniform(-1.5 * rad, 1.5 * rad, 2).tolist())
            if zmax - zmin < 1e-3: zmax = zmin + 0.1
            # Sphere::new clamps zmin from below and zmax from above only (sphere.rs:137-138): a z range wholly outside [-r, r] comes out
            # inverted, its area and light power negative, and Distribution1D::new asserts — the loader refuses such a scene as well
            zmin, zmax = min(zmin, 0.9 * rad), max(zmax, -0.9 * rad)
            phimax = float(rng.uniform(0.1, 8.0))
            t = None
            if k % 3:
                q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
                t = np.eye(4, dtype=np.float32); t[:3, :3] = (q * rng.uniform(0.3, 3.0, 3)).astype(np.float32); t[3, :3] = rng.uniform(-4, 4, 3)
            em = tuple(float(x) for x in rng.uniform(0.5, 40, 3)) if k % 2 else None
            hs.add_sphere(rad, zmin, zmax, phimax, m, emission=em, transform=t)
            params.append((rad, zmin, zmax, phimax, t, em))
        d = hs.build()
        for k, (rad, zmin, zmax, phimax, t, em) in enumerate(params):
            s = d.spheres[k]
            o = L.Sphere(); assert lib.arn_oracle_sphere_new(rad, zmin, zmax, phimax, C.byref(o)) == 0
            got = np.float32([s.radius, s.zmin, s.zmax, s.phimax, s.thetamin, s.thetamax]); want = np.float32([o.radius, o.zmin, o.zmax, o.phimax, o.thetamin, o.thetamax])
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (rep, k, got, want)
            if t is not None:
                inv = np.zeros(16, np.float32)
                assert lib.arn_oracle_m4_invert(O._p(np.ascontiguousarray(s.local_parent, np.float32)), O._p(inv)) == 0
                assert np.array_equal(inv.view(np.uint32), np.array(s.parent_local, np.float32).view(np.uint32)), (rep, k)
        # light table: emissive spheres in component order, power luminance as the CDF's function values
        lights = [k for k, p in enumerate(params) if p[5] is not None]
        assert d.n_lights == len(lights)
        for i, k in enumerate(lights):
            assert np.float32(lib.arn_oracle_light_power_y(C.byref(d.spheres[k]))) == np.float32(d.light_func[i]), (rep, k)
        cdf = np.zeros(d.n_lights + 1, np.float32); integral = C.c_float()
        lf = np.ctypeslib.as_array(d.light_func, (d.n_lights,)).astype(np.float32)
        lib.arn_oracle_light_distribution(d.n_lights, O._p(lf), O._p(cdf), C.byref(integral))
        assert np.array_equal(cdf.view(np.uint32), np.ctypeslib.as_array(d.light_cdf, (d.n_lights + 1,)).astype(np.float32).view(np.uint32))
        assert np.float32(integral.value) == np.float32(d.light_func_integral)
        if rep == 0:       # the inverted z range: refused like the reference's assert
            bad = api.HostScene(); bm = bad.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.5, 0.5)))
            bad.add_sphere(0.35, -0.46, -0.43, 5.6, bm, emission=(1.0, 1.0, 1.0))
            with pytest.raises(api.ArnError):
                bad.build()
        # bounds and costs the BVH was built from: recomputed by the oracle from the flattened scene
        ob, oc = O.prim_bounds(d)
        assert np.isfinite(ob).all() and ob.shape[0] == d.n_prims
        nodes = np.ctypeslib.as_array(C.cast(d.nodes, C.POINTER(C.c_float)), (d.n_nodes, 8))
        root = nodes[0, :6]
        assert np.array_equal(root[:3], ob[:, :3].min(0)) and np.array_equal(root[3:6], ob[:, 3:].max(0))
