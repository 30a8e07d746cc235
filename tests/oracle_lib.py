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
    lib.arn_oracle_lanczos.restype = C.c_float
    lib.arn_oracle_lanczos.argtypes = [C.c_float, C.c_float]
    lib.arn_oracle_roughness_to_alpha.restype = C.c_float
    lib.arn_oracle_roughness_to_alpha.argtypes = [C.c_float]
    lib.arn_oracle_bsdf_probe.argtypes = [C.POINTER(L.Material), vp, vp, vp, vp]
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
