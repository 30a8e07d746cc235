"""Thin Python wrappers over the C-ABI (include/arn.h, include/arn_host.h).

Names follow the reference's domain: a `HostScene` collects components (meshes, shaped
primitives), `build()` is `BVH::new` + `Scene::new`; a `Context` is one GPU; `Context.upload`
gives a device `Scene` on which `intersect_closest` (batched `Composable::intersect_ray`) and
`render_pt` (`PTRenderer::render`) run.  numpy is used for host buffers only.
"""
import ctypes as C
import numpy as np

from . import _lib as L


class ArnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"arn error {code}: {msg}")
        self.code = code


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


IDENTITY = np.eye(4, dtype=np.float32)


def make_camera(parent_view, screen, znear, zfar, fov, res_x, res_y, lens=None):
    """PerspecCam::new. `parent_view`: 4x4 given as 4 COLUMNS (cgmath / JSON order); screen = (pmin.x, pmin.y, pmax.x, pmax.y)."""
    lib = L.load()
    cam = L.Camera()
    pv = _f32(parent_view).reshape(16)
    sc = _f32(screen).reshape(4)
    rc = lib.arn_camera_make(_ptr(pv), _ptr(sc), znear, zfar, fov, 1 if lens else 0,
                             lens[0] if lens else 0.0, lens[1] if lens else 0.0, float(res_x), float(res_y), C.byref(cam))
    if rc != 0:
        raise ArnError(rc, lib.arn_hscene_last_error(None).decode())
    return cam


def make_film(res_x, res_y, crop=None, filter_radius=(4.0, 4.0), filter_kind=0, filter_a=0.0, filter_b=0.0):
    """Film (filming/film.rs:38-45).  filter_kind 0 = the deserialised default Lanczos(tau 3); Film::new's other
    filters (sample/filters.rs): L.ARN_FILTER_BOX / _TRIANGLE / _GAUSSIAN (filter_a = alpha) / _MITCHELL (b, c)."""
    f = L.Film()
    f.res_x, f.res_y = res_x, res_y
    c = crop or (0, 0, res_x, res_y)
    f.crop_min_x, f.crop_min_y, f.crop_max_x, f.crop_max_y = c
    f.filter_radius_x, f.filter_radius_y = filter_radius
    f.filter_kind, f.filter_a, f.filter_b = filter_kind, filter_a, filter_b
    return f


def make_ortho_camera(view_parent, screen, znear, zfar, res_x, res_y, lens=None):
    """OrthoCam::new (filming/ortho.rs:30-55). `view_parent`: 4 COLUMNS; note the reference takes view_parent here."""
    lib = L.load()
    cam = L.Camera()
    vpm = _f32(view_parent).reshape(16)
    sc = _f32(screen).reshape(4)
    rc = lib.arn_ortho_camera_make(_ptr(vpm), _ptr(sc), znear, zfar, 1 if lens else 0,
                                   lens[0] if lens else 0.0, lens[1] if lens else 0.0, float(res_x), float(res_y), C.byref(cam))
    if rc != 0:
        raise ArnError(rc, lib.arn_hscene_last_error(None).decode())
    return cam


def make_sampler(sampledx, sampledy, ndim=8, seed=0, mode=0):
    """mode: L.ARN_SAMPLER_PARITY (the reference's effective behaviour) or L.ARN_SAMPLER_STRATIFIED (the sampler as intended)."""
    s = L.Sampler()
    s.sampledx, s.sampledy, s.ndim, s.seed, s.mode = sampledx, sampledy, ndim, seed, mode
    return s


def make_pt_params(max_depth=8, rank=0, world_size=1, spp_begin=0, spp_end=0, tiles=(16, 16), min_depth=None, rr_threshold=0.05, subdiv=0):
    p = L.PTParams()
    p.max_depth = max_depth
    p.min_depth = max_depth // 2 if min_depth is None else min_depth      # renderer/pt.rs:48
    p.rr_threshold = rr_threshold                                         # renderer/pt.rs:47
    p.tiles_x, p.tiles_y = tiles
    p.rank, p.world_size = rank, world_size
    p.spp_begin, p.spp_end = spp_begin, spp_end
    p.partition_subdiv = subdiv
    return p


def material(kind, kd=(0, 0, 0), ks=(0, 0, 0), sigma=0.0, roughness=0.0, eta=1.0, dissolve=1.0, kd_tex=0, ks_tex=0, aux_tex=0, bump_tex=0):
    """*_tex: ids returned by HostScene.add_texture (0 = the constant): RGB textures for kd / ks, Luma for sigma|roughness (aux) and bump."""
    m = L.Material()
    m.kd_tex, m.ks_tex, m.aux_tex, m.bump_tex = kd_tex, ks_tex, aux_tex, bump_tex
    m.type = kind
    m.kd[:] = kd
    m.ks[:] = ks
    m.sigma, m.roughness, m.eta, m.dissolve = sigma, roughness, eta, dissolve
    return m


class HostScene:
    """Component list under construction (examples/arencli.rs:113-197)."""

    def __init__(self):
        self.lib = L.load()
        self.h = C.c_void_p()
        self._check(self.lib.arn_hscene_create(C.byref(self.h)))
        self._keep = []

    def _check(self, rc):
        if rc < 0:
            raise ArnError(rc, self.lib.arn_hscene_last_error(self.h).decode())
        return rc

    def close(self):
        if self.h:
            self.lib.arn_hscene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_material(self, m):
        return self._check(self.lib.arn_hscene_add_material(self.h, C.byref(m)))

    def add_mesh(self, positions, indices, material_id, normals=None, uvs=None, transform=None):
        pos = _f32(positions).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        nrm = _f32(normals).reshape(-1, 3) if normals is not None else None
        uv = _f32(uvs).reshape(-1, 2) if uvs is not None else None
        tr = _f32(transform).reshape(16) if transform is not None else None
        return self._check(self.lib.arn_hscene_add_mesh(self.h, _ptr(pos), pos.shape[0], _ptr(idx), idx.shape[0],
                                                        _ptr(nrm), _ptr(uv), _ptr(tr), material_id))

    def add_sphere(self, radius, zmin, zmax, phimax, material_id, emission=None, transform=None):
        em = _f32(emission).reshape(3) if emission is not None else None
        tr = _f32(transform).reshape(16) if transform is not None else None
        return self._check(self.lib.arn_hscene_add_sphere(self.h, radius, zmin, zmax, phimax, material_id, _ptr(em), _ptr(tr)))

    def add_texture(self, levels, trilinear=True, max_aniso=8.0, wrapping=L.ARN_WRAP_REPEAT, scaling=(1.0, 1.0), shifting=(0.0, 0.0)):
        """An ImageTexture with a ready-made pyramid: `levels` = list of float32 arrays (h, w, 3) or (h, w) (Luma), finest first.
        Returns the id for api.material(..., kd_tex=id)."""
        t = L.Texture()
        lv = [_f32(a) for a in levels]
        t.channels = 3 if lv[0].ndim == 3 else 1
        t.n_levels, t.trilinear, t.wrapping, t.max_aniso = len(lv), int(bool(trilinear)), wrapping, max_aniso
        t.scale_u, t.scale_v = scaling; t.shift_u, t.shift_v = shifting
        off = 0
        for i, a in enumerate(lv):
            t.level_h[i], t.level_w[i], t.level_offset[i] = a.shape[0], a.shape[1], off
            off += a.size
        flat = np.ascontiguousarray(np.concatenate([a.reshape(-1) for a in lv]), dtype=np.float32)
        return self._check(self.lib.arn_hscene_add_texture(self.h, C.byref(t), _ptr(flat), flat.size))

    def add_texture_file(self, path, channels=3, trilinear=False, max_aniso=16.0, wrapping=L.ARN_WRAP_REPEAT, gamma=False, scale=1.0,
                         scaling=(1.0, 1.0), shifting=(0.0, 0.0)):
        """An ImageTexture from a PNG file: MipMap::new restated (decode, Lanczos3 pyramid, convert_in) — defaults are load_obj's.
        Returns (id, mean) with mean = MipMap::mean."""
        t = L.Texture()
        t.channels, t.trilinear, t.wrapping, t.max_aniso = channels, int(bool(trilinear)), wrapping, max_aniso
        t.scale_u, t.scale_v = scaling; t.shift_u, t.shift_v = shifting
        mean = np.zeros(3, np.float32)
        tid = self._check(self.lib.arn_hscene_add_texture_file(self.h, str(path).encode(), C.byref(t), int(bool(gamma)), scale, mean.ctypes.data))
        return tid, mean[:channels].copy()

    def add_light(self, light):
        """`lights.push(light.to_arc())` for a Point / Spot / Distant light (examples/arencli.rs:95-98)."""
        return self._check(self.lib.arn_hscene_add_light(self.h, C.byref(light)))

    def load_obj(self, path, transform=None):
        tr = _f32(transform).reshape(16) if transform is not None else _f32(IDENTITY).reshape(16)
        return self._check(self.lib.arn_hscene_load_obj(self.h, str(path).encode(), _ptr(tr)))

    def load_json(self, path, base_dir=None):
        """arencli's parse_input. Returns (camera, film, sampler, pt_params, outputfilename)."""
        cam, film, smp, prm = L.Camera(), L.Film(), L.Sampler(), L.PTParams()
        out = C.create_string_buffer(1024)
        self._check(self.lib.arn_hscene_load_json(self.h, str(path).encode(), str(base_dir).encode() if base_dir else None,
                                                  C.byref(cam), C.byref(film), C.byref(smp), C.byref(prm), out, 1024))
        return cam, film, smp, prm, out.value.decode()

    def build(self, strategy=L.ARN_BVH_SAH):
        self._check(self.lib.arn_hscene_build(self.h, strategy))
        return self.desc()

    def build_gpu(self, ctx):
        """BVH built on the device (arn_bvh_build_gpu: LBVH, not the reference's topology). Returns (desc, build ms)."""
        ms = C.c_float(0.0)
        self._check(self.lib.arn_hscene_build_gpu(self.h, ctx.c, C.byref(ms)))
        return self.desc(), ms.value

    def desc(self):
        d = self.lib.arn_hscene_desc(self.h)
        if not d:
            raise ArnError(L.ARN_E_INVALID, "scene not built")
        return d.contents

    # numpy views of the flattened description (copies)
    def nodes(self):
        d = self.desc()
        return np.ctypeslib.as_array(C.cast(d.nodes, C.POINTER(C.c_uint32)), shape=(d.n_nodes, 8)).copy()

    def order(self):
        d = self.desc()
        return np.ctypeslib.as_array(d.order, shape=(d.n_prims,)).copy()


def bvh_build(bounds6, costs, strategy=L.ARN_BVH_SAH):
    """arn_bvh_build on raw bounds. Returns (nodes as uint32 (n,8) bit patterns, order)."""
    lib = L.load()
    b = _f32(bounds6).reshape(-1, 6)
    c = _f32(costs).reshape(-1)
    n = b.shape[0]
    nodes = np.zeros((2 * n, 8), dtype=np.uint32)
    order = np.zeros(n, dtype=np.uint32)
    nn = C.c_uint32(0)
    rc = lib.arn_bvh_build(n, C.cast(_ptr(b), L.c_float_p), C.cast(_ptr(c), L.c_float_p), strategy,
                           C.cast(_ptr(nodes), C.POINTER(L.Node)), C.cast(_ptr(order), L.c_u32_p), C.byref(nn))
    if rc != 0:
        raise ArnError(rc, "arn_bvh_build failed")
    return nodes[:nn.value].copy(), order


def film_finalize(film):
    """TilePixel::finalize + u8 quantisation. film: (h, w, 4) float32 -> (rgb float (h,w,3), rgb8 (h,w,3))."""
    lib = L.load()
    f = _f32(film)
    h, w = f.shape[0], f.shape[1]
    rgb = np.zeros((h, w, 3), dtype=np.float32)
    rgb8 = np.zeros((h, w, 3), dtype=np.uint8)
    rc = lib.arn_film_finalize(_ptr(f), h * w, _ptr(rgb), _ptr(rgb8))
    if rc != 0:
        raise ArnError(rc, "arn_film_finalize failed")
    return rgb, rgb8


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3), ("tmax", np.float32)])
HIT_DTYPE = np.dtype([("prim_id", np.int32), ("t", np.float32)])


class Context:
    """One GPU (arn_ctx)."""

    def __init__(self, device=0):
        self.lib = L.load()
        self.c = C.c_void_p()
        rc = self.lib.arn_ctx_create(device, C.byref(self.c))
        if rc != 0:
            raise ArnError(rc, self.lib.arn_last_error(None).decode())
        self.device = device

    def error(self):
        return self.lib.arn_last_error(self.c).decode()

    def close(self):
        if self.c:
            self.lib.arn_ctx_destroy(self.c)
            self.c = C.c_void_p()

    def set_option(self, option, value):
        rc = self.lib.arn_ctx_set_option(self.c, option, int(value))
        if rc != 0:
            raise ArnError(rc, self.error())

    def synchronize(self):
        rc = self.lib.arn_ctx_synchronize(self.c)
        if rc != 0:
            raise ArnError(rc, self.error())

    def stream(self):
        return self.lib.arn_ctx_stream(self.c)

    def upload(self, desc):
        s = C.c_void_p()
        rc = self.lib.arn_scene_upload(self.c, C.byref(desc), C.byref(s))
        if rc != 0:
            raise ArnError(rc, self.error())
        return Scene(self, s)


class FilmComm:
    """The NCCL communicator of the multi-GPU film merge (arn_nccl_comm_create / arn_film_reduce, include/arn.h).
    `exchange(obj_or_None) -> obj` ships rank 0's 128-byte ncclUniqueId to the other ranks (e.g. a
    torch.distributed broadcast over the launcher's rendezvous): plumbing, any transport will do."""

    def __init__(self, ctx, rank, world_size, exchange):
        self.ctx, self.lib, self.rank, self.world = ctx, ctx.lib, rank, world_size
        uid = (C.c_ubyte * 128)()
        if rank == 0:
            rc = self.lib.arn_nccl_unique_id(uid)
            if rc != 0:
                raise ArnError(rc, self.lib.arn_last_error(None).decode())
        data = exchange(bytes(uid) if rank == 0 else None)
        buf = (C.c_ubyte * 128).from_buffer_copy(data)
        self.comm = C.c_void_p()
        rc = self.lib.arn_nccl_comm_create(ctx.c, buf, rank, world_size, C.byref(self.comm))
        if rc != 0:
            raise ArnError(rc, ctx.error())

    def reduce(self, film_dev_ptr, n_pixels, root=0):
        """Film::merge_into across ranks: sum of the per-rank films lands in root's buffer (in place, ctx stream)."""
        rc = self.lib.arn_film_reduce(self.ctx.c, self.comm, C.c_void_p(film_dev_ptr), n_pixels, root)
        if rc != 0:
            raise ArnError(rc, self.ctx.error())

    def close(self):
        if self.comm:
            self.lib.arn_nccl_comm_destroy(self.comm)
            self.comm = C.c_void_p()


def film_merge(ctx, dst_dev_ptr, src_dev_ptr, n_pixels):
    """dst += src on the device (arn_film_merge)."""
    rc = ctx.lib.arn_film_merge(ctx.c, C.c_void_p(dst_dev_ptr), C.c_void_p(src_dev_ptr), n_pixels)
    if rc != 0:
        raise ArnError(rc, ctx.error())


def point_light(pos, intensity):
    """PointLight::new (lighting/pointlights.rs:25-27)."""
    out = L.AnalyticLight()
    p, i = _f32(pos).reshape(3), _f32(intensity).reshape(3)
    rc = L.load().arn_point_light_make(_ptr(p), _ptr(i), C.byref(out))
    if rc != 0:
        raise ArnError(rc, "arn_point_light_make")
    return out


def spot_light(pos, towards, intensity, total_angle, start_falloff_angle):
    """SpotLight::new (lighting/pointlights.rs:103-122)."""
    out = L.AnalyticLight()
    p, t, i = _f32(pos).reshape(3), _f32(towards).reshape(3), _f32(intensity).reshape(3)
    rc = L.load().arn_spot_light_make(_ptr(p), _ptr(t), _ptr(i), total_angle, start_falloff_angle, C.byref(out))
    if rc != 0:
        raise ArnError(rc, L.load().arn_hscene_last_error(None).decode())
    return out


def distant_light(intensity, direction, world_radius):
    """DistantLight::new + the radius set_world_bounds would store (lighting/distantlight.rs:26-50)."""
    out = L.AnalyticLight()
    i, d = _f32(intensity).reshape(3), _f32(direction).reshape(3)
    rc = L.load().arn_distant_light_make(_ptr(i), _ptr(d), world_radius, C.byref(out))
    if rc != 0:
        raise ArnError(rc, "arn_distant_light_make")
    return out


class Scene:
    """Device-resident flattened scene (arn_scene)."""

    def __init__(self, ctx, handle):
        self.ctx, self.s, self.lib = ctx, handle, ctx.lib

    def close(self):
        if self.s:
            self.lib.arn_scene_destroy(self.s)
            self.s = C.c_void_p()

    def _check(self, rc):
        if rc != 0:
            raise ArnError(rc, self.ctx.error())

    def intersect_closest(self, rays):
        """rays: structured array RAY_DTYPE (host). Returns HIT_DTYPE array. H2D/D2H inside."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
        self._check(self.lib.arn_intersect_closest(self.s, _ptr(rays), rays.shape[0], _ptr(hits)))
        return hits

    def intersect_any(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        out = np.empty(rays.shape[0], dtype=np.uint8)
        self._check(self.lib.arn_intersect_any(self.s, _ptr(rays), rays.shape[0], _ptr(out)))
        return out

    def intersect_closest_dev(self, rays_ptr, n, hits_ptr, stats=None):
        self._check(self.lib.arn_intersect_closest_dev(self.s, C.c_void_p(rays_ptr), n, C.c_void_p(hits_ptr),
                                                       C.byref(stats) if stats is not None else None))

    def intersect_any_dev(self, rays_ptr, n, out_ptr, stats=None):
        self._check(self.lib.arn_intersect_any_dev(self.s, C.c_void_p(rays_ptr), n, C.c_void_p(out_ptr),
                                                   C.byref(stats) if stats is not None else None))

    def intersect_closest_counted_dev(self, rays_ptr, n, hits_ptr):
        ctr = (C.c_uint64 * 3)()
        self._check(self.lib.arn_intersect_closest_counted_dev(self.s, C.c_void_p(rays_ptr), n, C.c_void_p(hits_ptr), ctr))
        return int(ctr[0]), int(ctr[1]), int(ctr[2])

    def render_pt(self, cam, film, sampler, params, out=None):
        """PTRenderer::render up to the tile merge. Returns (film (h,w,4) float32 host array, Stats).
        `out`: an existing (h, w, 4) float32 host array to receive the film (e.g. pinned memory)."""
        w = film.crop_max_x - film.crop_min_x
        h = film.crop_max_y - film.crop_min_y
        if out is None:
            out = np.zeros((h, w, 4), dtype=np.float32)
        assert out.shape == (h, w, 4) and out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
        st = L.Stats()
        self._check(self.lib.arn_render_pt(self.s, C.byref(cam), C.byref(film), C.byref(sampler), C.byref(params), _ptr(out), C.byref(st)))
        return out, st

    def render_pt_samples(self, cam, film, sampler, params):
        """Diagnostic: (film, per-sample radiance (h, w, n_spp, 4), Stats)."""
        w = film.crop_max_x - film.crop_min_x
        h = film.crop_max_y - film.crop_min_y
        spp = sampler.sampledx * sampler.sampledy
        n = (params.spp_end or spp) - params.spp_begin
        out = np.zeros((h, w, 4), dtype=np.float32)
        rad = np.zeros((h, w, n, 4), dtype=np.float32)
        st = L.Stats()
        self._check(self.lib.arn_render_pt_samples(self.s, C.byref(cam), C.byref(film), C.byref(sampler), C.byref(params), _ptr(out), _ptr(rad), C.byref(st)))
        return out, rad, st

    def render_pt_dev(self, cam, film, sampler, params, film_dev_ptr, want_stats=True):
        st = L.Stats()
        self._check(self.lib.arn_render_pt_dev(self.s, C.byref(cam), C.byref(film), C.byref(sampler), C.byref(params),
                                               C.c_void_p(film_dev_ptr), C.byref(st) if want_stats else None))
        return st
