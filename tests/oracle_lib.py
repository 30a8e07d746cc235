"""ctypes bindings of oracle/liboracle.so — the CPU restatement of arendur's hot path.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, by __graft_entry__.smoke() and by bench.py's
cpu_baseline / --impl reference legs; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from arendur_b200 import _lib as L

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(_ROOT, "oracle")
ORACLE_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

_lib = None


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ORACLE_PATH):
        build()
    lib = C.CDLL(ORACLE_PATH)
    vp = C.c_void_p
    lib.arn_oracle_bvh_build.restype = C.c_int
    lib.arn_oracle_bvh_build.argtypes = [C.c_uint32, vp, vp, C.c_int, vp, vp, C.POINTER(C.c_uint32)]
    lib.arn_oracle_prim_bounds.restype = C.c_int
    lib.arn_oracle_prim_bounds.argtypes = [C.POINTER(L.SceneDesc), vp, vp]
    lib.arn_oracle_light_distribution.argtypes = [C.c_uint32, vp, vp, vp]
    lib.arn_oracle_light_power_y.restype = C.c_float
    lib.arn_oracle_light_power_y.argtypes = [C.POINTER(L.Sphere)]
    lib.arn_oracle_analytic_power_y.restype = C.c_float
    lib.arn_oracle_analytic_power_y.argtypes = [C.POINTER(L.AnalyticLight)]
    lib.arn_oracle_analytic_sample.restype = None
    lib.arn_oracle_analytic_sample.argtypes = [C.POINTER(L.AnalyticLight), vp, vp]
    lib.arn_oracle_scene_create.restype = C.c_int
    lib.arn_oracle_scene_create.argtypes = [C.POINTER(L.SceneDesc), C.POINTER(vp)]
    lib.arn_oracle_scene_destroy.argtypes = [vp]
    lib.arn_oracle_intersect_closest.restype = C.c_int
    lib.arn_oracle_intersect_closest.argtypes = [vp, vp, C.c_size_t, vp, vp, C.c_int]
    lib.arn_oracle_intersect_any.restype = C.c_int
    lib.arn_oracle_intersect_any.argtypes = [vp, vp, C.c_size_t, vp]
    lib.arn_oracle_camera_make.restype = C.c_int
    lib.arn_oracle_camera_make.argtypes = [vp, vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(L.Camera)]
    lib.arn_oracle_camera_rays.argtypes = [C.POINTER(L.Camera), vp, C.c_size_t, vp]
    lib.arn_oracle_render_pt.restype = C.c_int
    lib.arn_oracle_render_pt.argtypes = [vp, C.POINTER(L.Camera), C.POINTER(L.Film), C.POINTER(L.Sampler), C.POINTER(L.PTParams), vp, C.POINTER(L.Stats), vp, C.c_int]
    lib.arn_oracle_render_pt_samples.restype = C.c_int
    lib.arn_oracle_render_pt_samples.argtypes = [vp, C.POINTER(L.Camera), C.POINTER(L.Film), C.POINTER(L.Sampler), C.POINTER(L.PTParams), vp, vp, C.c_int]
    lib.arn_oracle_film_finalize.argtypes = [vp, C.c_size_t, vp, vp]
    lib.arn_oracle_mesh_transform.argtypes = [vp, C.c_uint32, vp, vp, vp, vp]
    lib.arn_oracle_m4_invert.restype = C.c_int
    lib.arn_oracle_m4_invert.argtypes = [vp, vp]
    lib.arn_oracle_sphere_new.restype = C.c_int
    lib.arn_oracle_sphere_new.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(L.Sphere)]
    lib.arn_oracle_sphere_intersect.restype = C.c_int
    lib.arn_oracle_sphere_intersect.argtypes = [C.POINTER(L.Sphere), C.POINTER(L.Ray), vp, vp, vp, vp]
    lib.arn_oracle_bbox2i.restype = C.c_int
    lib.arn_oracle_bbox2i.argtypes = [C.c_int, vp, vp, vp]
    lib.arn_oracle_bbox2f_lerp.argtypes = [vp, C.c_float, C.c_float, vp]
    lib.arn_oracle_sampler_draws.argtypes = [C.c_uint32] * 6 + [vp]
    lib.arn_oracle_sampler_draws2.argtypes = [C.POINTER(L.Sampler)] + [C.c_uint32] * 5 + [vp]
    lib.arn_oracle_lanczos.restype = C.c_float
    lib.arn_oracle_lanczos.argtypes = [C.c_float, C.c_float]
    lib.arn_oracle_roughness_to_alpha.restype = C.c_float
    lib.arn_oracle_roughness_to_alpha.argtypes = [C.c_float]
    lib.arn_oracle_bsdf_probe.argtypes = [C.POINTER(L.Material), vp, vp, vp, vp]
    lib.arn_oracle_bsdf_probe2.argtypes = [C.POINTER(L.Material), vp, vp, vp, vp, vp]
    _lib = lib
    return lib


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def bvh_build(bounds6, costs, strategy=0):
    lib = load()
    b = np.ascontiguousarray(bounds6, np.float32).reshape(-1, 6)
    c = np.ascontiguousarray(costs, np.float32)
    n = b.shape[0]
    nodes = np.zeros((2 * n, 8), np.uint32)
    order = np.zeros(n, np.uint32)
    nn = C.c_uint32(0)
    rc = lib.arn_oracle_bvh_build(n, _p(b), _p(c), strategy, _p(nodes), _p(order), C.byref(nn))
    assert rc == 0
    return nodes[:nn.value].copy(), order


def prim_bounds(desc):
    lib = load()
    b = np.zeros((desc.n_prims, 6), np.float32)
    c = np.zeros(desc.n_prims, np.float32)
    assert lib.arn_oracle_prim_bounds(C.byref(desc), _p(b), _p(c)) == 0
    return b, c


class OracleScene:
    def __init__(self, desc):
        self.lib = load()
        self.h = C.c_void_p()
        assert self.lib.arn_oracle_scene_create(C.byref(desc), C.byref(self.h)) == 0

    def close(self):
        if self.h:
            self.lib.arn_oracle_scene_destroy(self.h)
            self.h = C.c_void_p()

    def intersect_closest(self, rays, nthreads=None, counters=False):
        from arendur_b200.api import RAY_DTYPE, HIT_DTYPE
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.empty(rays.shape[0], HIT_DTYPE)
        ctr = np.zeros(3, np.uint64)
        nt = nthreads or (os.cpu_count() or 1)
        assert self.lib.arn_oracle_intersect_closest(self.h, _p(rays), rays.shape[0], _p(hits), _p(ctr), nt) == 0
        return (hits, ctr) if counters else hits

    def intersect_any(self, rays):
        from arendur_b200.api import RAY_DTYPE
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        out = np.empty(rays.shape[0], np.uint8)
        assert self.lib.arn_oracle_intersect_any(self.h, _p(rays), rays.shape[0], _p(out)) == 0
        return out

    def render_pt(self, cam, film, sampler, params, nthreads=None):
        w = film.crop_max_x - film.crop_min_x
        h = film.crop_max_y - film.crop_min_y
        out = np.zeros((h, w, 4), np.float32)
        st = L.Stats()
        trav = np.zeros(3, np.uint64)
        nt = nthreads or (os.cpu_count() or 1)
        rc = self.lib.arn_oracle_render_pt(self.h, C.byref(cam), C.byref(film), C.byref(sampler), C.byref(params), _p(out), C.byref(st), _p(trav), nt)
        assert rc == 0, rc
        return out, st, trav


def _render_pt_samples(self, cam, film, sampler, params, nthreads=None):
    w = film.crop_max_x - film.crop_min_x
    h = film.crop_max_y - film.crop_min_y
    n = (params.spp_end or sampler.sampledx * sampler.sampledy) - params.spp_begin
    out = np.zeros((h, w, 4), np.float32)
    rad = np.zeros((h, w, n, 4), np.float32)
    rc = self.lib.arn_oracle_render_pt_samples(self.h, C.byref(cam), C.byref(film), C.byref(sampler), C.byref(params), _p(out), _p(rad), nthreads or (os.cpu_count() or 1))
    assert rc == 0, rc
    return out, rad


OracleScene.render_pt_samples = _render_pt_samples


def camera_make(parent_view, screen, znear, zfar, fov, res_x, res_y, lens=None):
    lib = load()
    cam = L.Camera()
    pv = np.ascontiguousarray(parent_view, np.float32).reshape(16)
    sc = np.ascontiguousarray(screen, np.float32).reshape(4)
    rc = lib.arn_oracle_camera_make(_p(pv), _p(sc), znear, zfar, fov, 1 if lens else 0, lens[0] if lens else 0.0,
                                    lens[1] if lens else 0.0, float(res_x), float(res_y), C.byref(cam))
    assert rc == 0
    return cam


def ortho_camera_make(view_parent, screen, znear, zfar, res_x, res_y, lens=None):
    lib = load()
    cam = L.Camera()
    vpm = np.ascontiguousarray(view_parent, np.float32).reshape(16)
    sc = np.ascontiguousarray(screen, np.float32).reshape(4)
    lib.arn_oracle_ortho_camera_make.restype = C.c_int
    lib.arn_oracle_ortho_camera_make.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(L.Camera)]
    rc = lib.arn_oracle_ortho_camera_make(_p(vpm), _p(sc), znear, zfar, 1 if lens else 0, lens[0] if lens else 0.0,
                                          lens[1] if lens else 0.0, float(res_x), float(res_y), C.byref(cam))
    assert rc == 0
    return cam


def filter_eval(film, dx, dy):
    lib = load()
    lib.arn_oracle_filter.restype = C.c_float
    lib.arn_oracle_filter.argtypes = [C.POINTER(L.Film), C.c_float, C.c_float]
    return lib.arn_oracle_filter(C.byref(film), dx, dy)


def camera_rays(cam, pfilm_plens):
    from arendur_b200.api import RAY_DTYPE
    lib = load()
    a = np.ascontiguousarray(pfilm_plens, np.float32).reshape(-1, 4)
    rays = np.zeros(a.shape[0], RAY_DTYPE)
    lib.arn_oracle_camera_rays(C.byref(cam), _p(a), a.shape[0], _p(rays))
    return rays


def film_finalize(film):
    lib = load()
    f = np.ascontiguousarray(film, np.float32)
    h, w = f.shape[0], f.shape[1]
    rgb = np.zeros((h, w, 3), np.float32)
    rgb8 = np.zeros((h, w, 3), np.uint8)
    lib.arn_oracle_film_finalize(_p(f), h * w, _p(rgb), _p(rgb8))
    return rgb, rgb8


# ---------------------------------------------------------------- oracle-only scene assembly
class OracleFlatScene:
    """The Cornell fixture flattened with the ORACLE's functions only (mesh transform, Sphere::new,
    bounds, BVH::new, Scene::new's light distribution) — no call into libarn_b200.so.  Used by
    `bench.py --impl reference` so that the CPU arm never loads the product library, and by
    tests/test_scene_ingest.py as an independent check of the product's host flattening
    (examples/arencli.rs:113-197 -> arn_scene_desc)."""

    def __init__(self, fixture=None):
        lib = load()
        fixture = fixture or os.path.join(_ROOT, "tests", "golden", "cornell_scene.npz")
        z = np.load(fixture, allow_pickle=False)
        self.z = z
        mats = z["materials"]
        self.materials = (L.Material * len(mats))()
        for i, r in enumerate(mats):
            m = self.materials[i]
            m.type = int(r[0]); m.kd[:] = [float(v) for v in r[1:4]]; m.ks[:] = [float(v) for v in r[4:7]]
            sig = np.float32(r[7])
            if m.type == L.ARN_MAT_MATTE:                                  # material/matte.rs:50-54
                sig = np.float32(0.0) if sig < 0 else (np.float32(90.0) if not (sig < 90) else sig)
            m.sigma, m.roughness, m.eta, m.dissolve = float(sig), float(r[8]), float(r[9]), float(r[10])
            m.alpha = lib.arn_oracle_roughness_to_alpha(C.c_float(float(r[8])))
        t = np.ascontiguousarray(z["mesh_transform"], np.float32)
        pos, nrm, uvs, idx, tri_mesh = [], [], [], [], []
        nm = int(z["n_models"])
        self.meshes = (L.Mesh * nm)()
        base = 0
        any_n = any_uv = False
        for m in range(nm):
            p = np.ascontiguousarray(z[f"m{m}_positions"], np.float32)
            n = np.ascontiguousarray(z[f"m{m}_normals"], np.float32) if f"m{m}_normals" in z.files else None
            uv = np.ascontiguousarray(z[f"m{m}_texcoords"], np.float32) if f"m{m}_texcoords" in z.files else None
            op = np.zeros_like(p); on = np.zeros_like(p)
            lib.arn_oracle_mesh_transform(_p(t), p.shape[0], _p(p), _p(n), _p(op), _p(on) if n is not None else None)
            pos.append(op); nrm.append(on if n is not None else np.zeros_like(p))
            uvs.append(uv if uv is not None else np.zeros((p.shape[0], 2), np.float32))
            ii = np.ascontiguousarray(z[f"m{m}_indices"], np.uint32)
            ntri = ii.shape[0] // 3
            idx.append(ii[:ntri * 3] + np.uint32(base)); tri_mesh.append(np.full(ntri, m, np.uint32))
            me = self.meshes[m]
            me.material, me.has_normals, me.has_uvs, me.reserved = int(z["model_material"][m]), int(n is not None), int(uv is not None), 0
            any_n |= n is not None; any_uv |= uv is not None
            base += p.shape[0]
        self.positions = np.ascontiguousarray(np.concatenate(pos), np.float32)
        self.normals = np.ascontiguousarray(np.concatenate(nrm), np.float32) if any_n else None
        self.uvs = np.ascontiguousarray(np.concatenate(uvs), np.float32) if any_uv else None
        self.indices = np.ascontiguousarray(np.concatenate(idx), np.uint32)
        self.tri_mesh = np.ascontiguousarray(np.concatenate(tri_mesh), np.uint32)
        ntri = self.tri_mesh.shape[0]
        sph = z["spheres"]
        self.spheres = (L.Sphere * len(sph))()
        prims = list(range(ntri)); lights = []
        for k, r in enumerate(sph):
            s = self.spheres[k]
            assert lib.arn_oracle_sphere_new(C.c_float(float(r[0])), C.c_float(float(r[1])), C.c_float(float(r[2])), C.c_float(float(r[3])), C.byref(s)) == 0
            s.material = int(r[4]); s.emissive = 1; s.emission[:] = [float(v) for v in r[5:8]]
            lp = np.ascontiguousarray(r[8:24], np.float32); pl = np.zeros(16, np.float32)
            if lib.arn_oracle_m4_invert(_p(lp), _p(pl)) == 0:
                s.has_transform = 1
            else:                                                           # arencli.rs:133-146
                lp = np.eye(4, dtype=np.float32).reshape(16); pl = lp.copy()
            s.local_parent[:] = [float(v) for v in lp]; s.parent_local[:] = [float(v) for v in pl]
            lights.append(len(prims)); prims.append(L.ARN_PRIM_SPHERE | k)
        self.prims = np.ascontiguousarray(prims, np.uint32)
        d = L.SceneDesc()
        d.n_vertices = self.positions.shape[0]
        d.positions = C.cast(_p(self.positions), L.c_float_p)
        d.normals = C.cast(_p(self.normals), L.c_float_p) if self.normals is not None else None
        d.uvs = C.cast(_p(self.uvs), L.c_float_p) if self.uvs is not None else None
        d.n_triangles = ntri
        d.indices = C.cast(_p(self.indices), L.c_u32_p); d.tri_mesh = C.cast(_p(self.tri_mesh), L.c_u32_p)
        d.n_meshes = nm; d.meshes = self.meshes
        d.n_spheres = len(sph); d.spheres = self.spheres
        d.n_materials = len(mats); d.materials = self.materials
        d.n_prims = self.prims.shape[0]; d.prims = C.cast(_p(self.prims), L.c_u32_p)
        self.desc = d
        # BVH::new(components, SAH) (component/bvh.rs:58-79)
        b6, cost = prim_bounds(d)
        nodes, order = bvh_build(b6, cost, 0)
        self.nodes = np.ascontiguousarray(nodes); self.order = np.ascontiguousarray(order)
        d.n_nodes = self.nodes.shape[0]; d.nodes = C.cast(_p(self.nodes), C.POINTER(L.Node)); d.order = C.cast(_p(self.order), L.c_u32_p)
        # Scene::new (renderer/scene.rs:31-51)
        self.light_prims = np.ascontiguousarray(lights, np.uint32)
        self.light_func = np.ascontiguousarray([lib.arn_oracle_light_power_y(C.byref(self.spheres[int(self.prims[i]) & 0x7fffffff])) for i in lights], np.float32)
        self.light_cdf = np.zeros(len(lights) + 1, np.float32)
        integ = C.c_float(0.0)
        lib.arn_oracle_light_distribution(len(lights), _p(self.light_func), _p(self.light_cdf), C.byref(integ))
        d.n_lights = len(lights); d.light_prims = C.cast(_p(self.light_prims), L.c_u32_p)
        d.light_func = C.cast(_p(self.light_func), L.c_float_p); d.light_cdf = C.cast(_p(self.light_cdf), L.c_float_p)
        d.light_func_integral = integ.value
        d.n_analytic_lights = 0; d.analytic_lights = None

    def camera(self, res_x, res_y):
        c = self.z["camera"]
        return camera_make(c[0:16], c[16:20], float(c[20]), float(c[21]), float(c[22]), res_x, res_y)

    def max_depth(self):
        return int(self.z["max_depth"])


def make_film(res_x, res_y):
    f = L.Film()
    f.res_x, f.res_y = res_x, res_y
    f.crop_min_x, f.crop_min_y, f.crop_max_x, f.crop_max_y = 0, 0, res_x, res_y
    f.filter_radius_x, f.filter_radius_y = 4.0, 4.0
    return f


def make_sampler(sx, sy, ndim=8, seed=0, mode=0):
    s = L.Sampler()
    s.sampledx, s.sampledy, s.ndim, s.seed, s.mode = sx, sy, ndim, seed, mode
    return s


def make_pt_params(max_depth=8, spp_begin=0, spp_end=0, rank=0, world_size=1, tiles=(16, 16), subdiv=0):
    p = L.PTParams()
    p.max_depth, p.min_depth, p.rr_threshold = max_depth, max_depth // 2, 0.05      # renderer/pt.rs:47-48
    p.tiles_x, p.tiles_y = tiles
    p.rank, p.world_size, p.spp_begin, p.spp_end = rank, world_size, spp_begin, spp_end
    p.partition_subdiv = subdiv
    return p
