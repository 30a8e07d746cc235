"""Benchmark / parity scenes of BASELINE.md §4, assembled through the host layer.

C1/C3/C5: the Cornell box of examples/cornellbox/cb.json (from the committed fixture
tests/golden/cornell_scene.npz, or from the original files when a path is given).
C2: 1 002 528-triangle height field, primary rays.   C4: 20 000 172-triangle closed box.
Synthetic geometry is generated with numpy integer hashing, so it is identical everywhere.
"""
import math
import os

import numpy as np

from . import _lib as L
from . import api

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CORNELL_FIXTURE = os.path.join(_ROOT, "tests", "golden", "cornell_scene.npz")


# ---------------------------------------------------------------- integer-hash noise
def _mix32(h):
    h = h.astype(np.uint64)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    return h


def _lattice(ix, iy, seed):
    h = _mix32((np.uint64(seed) + ix.astype(np.uint64) * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF))
    h = _mix32(h ^ ((iy.astype(np.uint64) * np.uint64(0x85EBCA77)) & np.uint64(0xFFFFFFFF)))
    return (h >> np.uint64(8)).astype(np.float64) * (2.0 / 16777216.0) - 1.0


def hash_noise(i, j, seed):
    """fbm-style value noise in [-1, 1] on integer grid coordinates (4 octaves, periods 64..8)."""
    total = np.zeros(np.broadcast(i, j).shape, dtype=np.float64)
    amp, norm = 1.0, 0.0
    for o, period in enumerate((64, 32, 16, 8)):
        x = i.astype(np.float64) / period
        y = j.astype(np.float64) / period
        x0 = np.floor(x).astype(np.int64)
        y0 = np.floor(y).astype(np.int64)
        fx, fy = x - x0, y - y0
        sx, sy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
        s = seed + 0x101 * o
        v00 = _lattice(x0, y0, s)
        v10 = _lattice(x0 + 1, y0, s)
        v01 = _lattice(x0, y0 + 1, s)
        v11 = _lattice(x0 + 1, y0 + 1, s)
        total += amp * ((v00 * (1 - sx) + v10 * sx) * (1 - sy) + (v01 * (1 - sx) + v11 * sx) * sy)
        norm += amp
        amp *= 0.5
    return total / norm


def heightfield(cells, lo, hi, z0, amplitude, seed):
    """(positions (V,3) f32, indices (T,3) u32) of a cells x cells height field on [lo,hi]^2, z = z0 + amplitude*noise."""
    n = cells + 1
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    xs = lo + (hi - lo) * (ii.astype(np.float64) / cells)
    ys = lo + (hi - lo) * (jj.astype(np.float64) / cells)
    zs = z0 + amplitude * hash_noise(ii, jj, seed)
    pos = np.stack([xs, ys, zs], axis=-1).reshape(-1, 3).astype(np.float32)
    ci, cj = np.meshgrid(np.arange(cells), np.arange(cells), indexing="xy")
    v00 = (cj * n + ci).reshape(-1)
    v10, v01, v11 = v00 + 1, v00 + n, v00 + n + 1
    # each cell split along (i,j)->(i+1,j+1)
    tris = np.stack([np.stack([v00, v11, v10], axis=-1), np.stack([v00, v01, v11], axis=-1)], axis=1).reshape(-1, 3)
    return pos, tris.astype(np.uint32)


# ---------------------------------------------------------------- C2
def c2_heightfield_scene(cells=708, seed=0x5EED):
    """returns (HostScene built, camera, film)."""
    hs = api.HostScene()
    mat = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.7, 0.7, 0.7)))
    pos, idx = heightfield(cells, -2.0, 2.0, 4.0, 0.15, seed)
    hs.add_mesh(pos, idx, mat)
    hs.build()
    w, h = 1920, 1080
    cam = api.make_camera(api.IDENTITY, (-16.0 / 9.0, -1.0, 16.0 / 9.0, 1.0), 0.1, 1000.0, math.pi / 2, w, h)
    return hs, cam, api.make_film(w, h)


def pixel_center_rays(cam, w, h):
    """One ray per pixel centre (pfilm = (x+0.5, y+0.5)), numpy float32 restatement of
    PerspecCam::generate_path for the no-lens case; used for benchmarking only (parity
    tests take their rays from the oracle)."""
    rv = np.array(cam.raster_view, dtype=np.float32).reshape(4, 4)      # rows = columns of the matrix
    vp = np.array(cam.view_parent, dtype=np.float32).reshape(4, 4)
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32) + np.float32(0.5), np.arange(h, dtype=np.float32) + np.float32(0.5), indexing="xy")
    px, py = xs.reshape(-1), ys.reshape(-1)
    one = np.float32(1.0)

    def mulp(m, x, y, z):
        out = [m[0, r] * x + m[1, r] * y + m[2, r] * z + m[3, r] * one for r in range(4)]
        iw = one / out[3]
        return out[0] * iw, out[1] * iw, out[2] * iw

    vx, vy, vz = mulp(rv, px, py, np.zeros_like(px))
    inv = one / np.sqrt(vx * vx + vy * vy + vz * vz)
    dx, dy, dz = vx * inv, vy * inv, vz * inv
    rays = np.zeros(px.shape[0], dtype=api.RAY_DTYPE)
    ox, oy, oz = mulp(vp, np.zeros_like(px), np.zeros_like(px), np.zeros_like(px))
    rays["o"] = np.stack([ox, oy, oz], axis=-1)
    ddx = vp[0, 0] * dx + vp[1, 0] * dy + vp[2, 0] * dz
    ddy = vp[0, 1] * dx + vp[1, 1] * dy + vp[2, 1] * dz
    ddz = vp[0, 2] * dx + vp[1, 2] * dy + vp[2, 2] * dz
    rays["d"] = np.stack([ddx, ddy, ddz], axis=-1)
    rays["tmax"] = np.inf
    return rays


# ---------------------------------------------------------------- Cornell (C1 / C3 / C5)
def cornell_scene(res_x=256, res_y=256, sampledx=4, sampledy=4, seed=0, fixture=CORNELL_FIXTURE, lights=()):
    """The cb.json scene with the film / sampler of the requested config.
    `lights`: extra Point / Spot / Distant lights (api.point_light(...) etc.), placed first in Scene.lights as
    arencli does.  Returns (HostScene built, camera, film, sampler, pt_params)."""
    z = np.load(fixture, allow_pickle=False)
    hs = api.HostScene()
    for l in lights:
        hs.add_light(l)
    mats = z["materials"]          # rows: type, kd3, ks3, sigma, roughness, eta, dissolve
    mat_ids = []
    for r in mats:
        mat_ids.append(hs.add_material(api.material(int(r[0]), kd=r[1:4], ks=r[4:7], sigma=float(r[7]), roughness=float(r[8]),
                                                    eta=float(r[9]), dissolve=float(r[10]))))
    transform = z["mesh_transform"]
    for m in range(int(z["n_models"])):
        nrm = z[f"m{m}_normals"] if f"m{m}_normals" in z.files else None
        uv = z[f"m{m}_texcoords"] if f"m{m}_texcoords" in z.files else None
        hs.add_mesh(z[f"m{m}_positions"], z[f"m{m}_indices"], mat_ids[int(z["model_material"][m])], normals=nrm, uvs=uv, transform=transform)
    sph = z["spheres"]             # rows: radius, zmin, zmax, phimax, material index, emission3, transform16
    for r in sph:
        hs.add_sphere(float(r[0]), float(r[1]), float(r[2]), float(r[3]), mat_ids[int(r[4])], emission=r[5:8], transform=r[8:24])
    hs.build()
    c = z["camera"]                # transform16, screen4, znear, zfar, fov
    cam = api.make_camera(c[0:16], c[16:20], float(c[20]), float(c[21]), float(c[22]), res_x, res_y)
    film = api.make_film(res_x, res_y)
    smp = api.make_sampler(sampledx, sampledy, 8, seed)
    prm = api.make_pt_params(max_depth=int(z["max_depth"]))
    return hs, cam, film, smp, prm


# ---------------------------------------------------------------- C4
def c4_box_scene(cells=1291, seed=0x5EED, res=1024, sampledx=4, sampledy=4, build=True):
    """closed box [-4,4]^3 of six noise-displaced height-field walls
    (6 * 2 * cells^2 triangles), Lambertian kd 0.7, the two emissive spheres of cb.json inside."""
    hs = api.HostScene()
    wall = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.7, 0.7, 0.7)))
    lightm = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.5, 0.5), sigma=3.0))
    for wdx in range(6):
        pos, idx = heightfield(cells, -4.0, 4.0, 4.0, -0.05, seed + wdx)   # inward displacement
        x, y, zc = pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy()
        axis, sign = wdx // 2, 1.0 if wdx % 2 == 0 else -1.0
        p = np.empty_like(pos)
        if axis == 0:
            p[:, 0], p[:, 1], p[:, 2] = sign * zc, x, y
        elif axis == 1:
            p[:, 0], p[:, 1], p[:, 2] = x, sign * zc, y
        else:
            p[:, 0], p[:, 1], p[:, 2] = x, y, sign * zc
        hs.add_mesh(p, idx, wall)
    def tr(x, y, z):
        m = np.eye(4, dtype=np.float32); m[3, 0:3] = (x, y, z); return m     # rows = columns (translation in column w)
    hs.add_sphere(1.5, -2.0, 2.0, 6.28, lightm, emission=(15.5, 10.5, 5.5), transform=tr(-1.5, 0.0, 1.0))
    hs.add_sphere(1.5, -2.0, 2.0, 6.28, lightm, emission=(7.5, 7.5, 10.5), transform=tr(1.5, 1.5, -1.0))
    if build:
        hs.build()          # build=False: the caller builds (HostScene.build_gpu)
    view = np.eye(4, dtype=np.float32); view[3, 0:3] = (0.0, 0.0, 3.5)      # camera at z = -3.5 looking +z
    cam = api.make_camera(view, (-1.0, -1.0, 1.0, 1.0), 0.1, 1000.0, 1.2707964, res, res)
    film = api.make_film(res, res)
    smp = api.make_sampler(sampledx, sampledy, 8, 0)
    prm = api.make_pt_params(max_depth=8)
    return hs, cam, film, smp, prm
