"""GPU tests added in round 2: roofline inputs pinned to the oracle's counters, the in-library film merge
(arn_film_merge / arn_film_reduce), the finer rank partition, and the conservative interior-node walk."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O
from arendur_b200 import api, scenes, _lib as L

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _camera_grid_rays(cam, w, h, step=1):
    xs, ys = np.meshgrid(np.arange(0, w, step) + 0.5, np.arange(0, h, step) + 0.5, indexing="xy")
    pf = np.zeros((xs.size, 4), np.float32)
    pf[:, 0], pf[:, 1] = xs.reshape(-1), ys.reshape(-1)
    return O.camera_rays(cam, pf)


def _dev(rays):
    import torch
    n = rays.shape[0]
    return torch.from_numpy(np.ascontiguousarray(rays).view(np.uint8).reshape(n, 28)).cuda(), torch.empty((n, 8), dtype=torch.uint8, device="cuda")


def test_traversal_counters_equal_the_oracles(ctx, cornell_small):
    """The Nn / Nt of the algorithmic-bytes figure are the REFERENCE traversal's node and primitive tests:
    the instrumented kernel must count exactly what the instrumented oracle counts (Cornell incl. spheres, and a height field)."""
    hs, cam, film, smp, prm = cornell_small
    cases = [(hs.desc(), _camera_grid_rays(cam, film.res_x, film.res_y))]
    h2 = api.HostScene()
    mat = h2.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.7, 0.7, 0.7)))
    pos, idx = scenes.heightfield(96, -2.0, 2.0, 4.0, 0.15, 0x5EED)
    h2.add_mesh(pos, idx, mat)
    cam2 = api.make_camera(api.IDENTITY, (-16.0 / 9.0, -1.0, 16.0 / 9.0, 1.0), 0.1, 1000.0, math.pi / 2, 320, 180)
    cases.append((h2.build(), _camera_grid_rays(cam2, 320, 180)))
    rng = np.random.default_rng(3)
    inc = np.zeros(8000, api.RAY_DTYPE)
    inc["o"] = rng.uniform([-1.8, -1.3, 2.2], [1.8, 2.2, 5.8], (8000, 3)).astype(np.float32)
    v = rng.normal(size=(8000, 3)); inc["d"] = (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32); inc["tmax"] = np.inf
    cases.append((hs.desc(), inc))
    for d, rays in cases:
        sc = ctx.upload(d); osc = O.OracleScene(d)
        rd, hd = _dev(rays)
        g = sc.intersect_closest_counted_dev(rd.data_ptr(), rays.shape[0], hd.data_ptr())
        oh, octr = osc.intersect_closest(rays, counters=True)
        assert tuple(int(x) for x in octr) == g, f"gpu counters {g} != oracle counters {tuple(int(x) for x in octr)}"
        gh = hd.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
        assert np.array_equal(gh["prim_id"], oh["prim_id"]) and np.array_equal(gh["t"], oh["t"])
        sc.close(); osc.close()


def test_render_counters_equal_the_oracles(ctx, cornell_small):
    """Same for the counted render pass bench.py takes bytes_per_ray from: nodes / triangles / spheres tested by ALL traversals."""
    hs, cam, film, smp, prm = cornell_small
    sc = ctx.upload(hs.desc()); osc = O.OracleScene(hs.desc())
    ctx.set_option(L.ARN_OPT_COUNT_TRAVERSAL, 1)
    try:
        _, st = sc.render_pt(cam, film, smp, prm)
    finally:
        ctx.set_option(L.ARN_OPT_COUNT_TRAVERSAL, 0)
    _, ost, trav = osc.render_pt(cam, film, smp, prm)
    assert (st.extend_nodes, st.extend_tris, st.extend_spheres) == tuple(int(x) for x in trav)
    assert (st.extend_rays, st.shadow_rays, st.mis_rays) == (ost.extend_rays, ost.shadow_rays, ost.mis_rays)
    sc.close(); osc.close()


def test_film_merge_on_device(ctx):
    import torch
    a = torch.rand((37, 53, 4), device="cuda"); b = torch.rand((37, 53, 4), device="cuda")
    want = (a + b).cpu()
    api.film_merge(ctx, a.data_ptr(), b.data_ptr(), 37 * 53)
    ctx.synchronize()
    assert torch.equal(a.cpu(), want)


@pytest.mark.parametrize("world,subdiv", [(2, 0), (3, 4), (8, 4), (5, 7)])
def test_subdivided_partition_sums_to_full_frame(ctx, cornell_small, world, subdiv):
    """arn_pt_params.partition_subdiv: the ranks' films add up to the one-rank film; every rank's film equals the oracle's
    film for the same (rank, world, subdiv) — same cells on both sides."""
    hs, cam, film, smp, prm = cornell_small
    sc = ctx.upload(hs.desc()); osc = O.OracleScene(hs.desc())
    full, st_full = sc.render_pt(cam, film, smp, api.make_pt_params(max_depth=4))
    acc = np.zeros_like(full, dtype=np.float64); cams = 0
    for r in range(world):
        p = api.make_pt_params(max_depth=4, rank=r, world_size=world, subdiv=subdiv)
        f, st = sc.render_pt(cam, film, smp, p)
        acc += f; cams += st.camera_rays
        if r in (0, world - 1):
            of, ost, _ = osc.render_pt(cam, film, smp, p)
            assert st.camera_rays == ost.camera_rays
            assert np.allclose(f, of, rtol=2e-5, atol=2e-5)
    assert cams == st_full.camera_rays
    assert np.allclose(acc, full, rtol=2e-5, atol=2e-5)
    sc.close(); osc.close()


def test_film_reduce_over_nccl_two_ranks():
    """arn_film_reduce (ncclReduce inside the C-ABI) + arn_film_merge across two GPUs == the one-GPU film."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 23000 + os.getpid() % 3000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "mp_film_reduce.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "film reduce OK" in r.stdout


def _adversarial_case(cells, seed):
    """Height field + a sphere and the rays of the adversarial-ray tests; returns (host scene, desc, rays)."""
    h = api.HostScene()
    mat = h.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.7, 0.7, 0.7)))
    pos, idx = scenes.heightfield(cells, -2.0, 2.0, 4.0, 0.15, 0x5EED)
    h.add_mesh(pos, idx, mat)
    h.add_sphere(0.4, -0.4, 0.4, 6.28, mat, transform=np.float32([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0.25, -0.5, 3.0, 1]]))
    d = h.build()
    rng = np.random.default_rng(seed)
    grid = (-2.0 + 4.0 * np.arange(cells + 1) / cells).astype(np.float32)      # vertex coordinates = node plane coordinates
    parts = []

    def rays_of(o, dvec, tmax=np.inf):
        r = np.zeros(o.shape[0], api.RAY_DTYPE)
        r["o"], r["d"], r["tmax"] = o.astype(np.float32), dvec.astype(np.float32), tmax
        return r
    n = 6000
    # +z rays from grid-aligned origins (ortho-camera like): inv.x = inv.y = inf, origins on x / y planes of many nodes
    o = np.stack([rng.choice(grid, n), rng.choice(grid, n), np.full(n, 0.5, np.float32)], 1)
    parts.append(rays_of(o, np.tile(np.float32([0, 0, 1]), (n, 1))))
    parts.append(rays_of(o + np.float32([0, 0, 8.0]), np.tile(np.float32([0, 0, -1]), (n, 1))))
    # rays inside the z-range of the surface travelling along x or y, origins on planes of the other axis
    o = np.stack([np.full(n, -3.0, np.float32), rng.choice(grid, n), rng.uniform(3.85, 4.15, n)], 1)
    parts.append(rays_of(o, np.tile(np.float32([1, 0, 0]), (n, 1))))
    o = np.stack([rng.choice(grid, n), np.full(n, 3.0, np.float32), rng.uniform(3.85, 4.15, n)], 1)
    parts.append(rays_of(o, np.tile(np.float32([0, -1, 0]), (n, 1))))
    # tiny but non-zero components (regular / irregular boundary of the fast path), -0.0 components
    for eps in (1e-20, -1e-20, 1e-38, 1e-44, -0.0, 1e-12, -1e-16):
        o = np.stack([rng.choice(grid, n // 4), rng.uniform(-2, 2, n // 4), np.full(n // 4, 0.5, np.float32)], 1)
        dv = np.tile(np.float32([eps, 0.3, 1.0]), (n // 4, 1)); dv[:, 1] = rng.uniform(-0.5, 0.5, n // 4)
        parts.append(rays_of(o, dv))
    # random rays with finite, zero and negative tmax; huge origins; rays aimed at the sphere
    o = rng.uniform([-2.5, -2.5, 0.0], [2.5, 2.5, 6.0], (n, 3)); v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    parts.append(rays_of(o, v, rng.choice(np.float32([np.inf, 3.0, 1.0, 0.0, -1.0]), n)))
    parts.append(rays_of(o * np.float32(1e6), v))
    tgt = np.float32([0.25, -0.5, 3.0]) + rng.normal(scale=0.3, size=(n, 3)); dv = tgt - o
    parts.append(rays_of(o, dv / np.linalg.norm(dv, axis=1, keepdims=True)))
    rays = np.concatenate(parts)
    return h, d, rays


@pytest.mark.parametrize("width", [2, 4, 8])
def test_conservative_interior_walk_is_bit_exact_on_adversarial_rays(ctx, width):
    """Interior nodes are culled with a cheap conservative slab test, leaves with the reference's (traverse.cuh).  Rays that
    stress the argument: axis-parallel rays whose origins sit exactly on node planes (0 * inf = NaN in the reference's slab
    arithmetic: sticky on x, ignored on y / z), direction components of 1e-20 / 1e-38 / denormal, grazing rays along the
    grid, tmax <= 0, huge origins.  Hit ids and t must equal the oracle's bit for bit."""
    h, d, rays = _adversarial_case(64, width)
    osc = O.OracleScene(d)
    ctx.set_option(L.ARN_OPT_BVH_WIDTH, width)          # before the upload: the compressed 8-wide nodes are built there
    try:
        sc = ctx.upload(d)
        gh, ga = sc.intersect_closest(rays), sc.intersect_any(rays)
    finally:
        ctx.set_option(L.ARN_OPT_BVH_WIDTH, 0)
    oh = osc.intersect_closest(rays)
    assert (oh["prim_id"] >= 0).sum() > rays.shape[0] // 10
    assert (oh["prim_id"] == d.n_prims - 1).sum() > 100, "the sphere should be reached"
    mism = np.nonzero((gh["prim_id"] != oh["prim_id"]) | (gh["t"] != oh["t"]))[0]
    assert mism.size == 0, f"{mism.size} mismatches, first rays {mism[:5]}: o {rays['o'][mism[:3]]} d {rays['d'][mism[:3]]} gpu {gh[mism[:3]]} oracle {oh[mism[:3]]}"
    assert np.array_equal(ga != 0, oh["prim_id"] >= 0)
    sc.close(); osc.close()


def _soup_case(rng, scale, n_tri=2500, n=8000):
    """Scene and rays of the traversal fuzz (shared with tools/trav_soak.py)."""
    S = np.float32(scale)
    h = api.HostScene()
    mat = h.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.5, 0.5)))
    c = rng.uniform(-1, 1, (n_tri, 1, 3)); e = rng.normal(scale=0.08, size=(n_tri, 3, 3))
    tri = (c + e).astype(np.float32)
    k = n_tri // 5
    for ax in range(3):                                  # flat in one axis: zero-thickness bounds
        sel = slice(ax * k // 3, (ax + 1) * k // 3)
        tri[sel, :, ax] = np.round(tri[sel, :1, ax] * 8) / 8
    q = n_tri // 40
    tri[k:k + q, 2] = tri[k:k + q, 1]                    # repeated vertex
    tri[k + q:k + 2 * q, 2] = (tri[k + q:k + 2 * q, 0] + tri[k + q:k + 2 * q, 1]) * np.float32(0.5)   # collinear
    tri[k + 2 * q:k + 3 * q] = tri[k + 3 * q:k + 4 * q]  # exact duplicates
    tri *= S
    pos = tri.reshape(-1, 3); idx = np.arange(pos.shape[0], dtype=np.uint32).reshape(-1, 3)
    h.add_mesh(pos, idx, mat)
    for _ in range(6):
        rad = float(rng.uniform(0.05, 0.3)) * float(S)
        t = np.eye(4, dtype=np.float32); t[3, :3] = rng.uniform(-1, 1, 3) * S
        if rng.random() < 0.5:
            a = rng.uniform(0, 6.28); t[0, 0], t[0, 1], t[1, 0], t[1, 1] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
        h.add_sphere(rad, -rad * float(rng.uniform(0.3, 1)), rad * float(rng.uniform(0.3, 1)), float(rng.uniform(2, 6.2832)), mat, transform=t)
    for _ in range(4):                                   # general affine instances (rotation x non-uniform scale): accepted hits rewrite the ray
        q, _r = np.linalg.qr(rng.normal(size=(3, 3)))
        t = np.eye(4, dtype=np.float32); t[:3, :3] = (q * rng.uniform(0.5, 2.0, 3)).astype(np.float32); t[3, :3] = rng.uniform(-1, 1, 3) * S
        rad = float(rng.uniform(0.1, 0.4)) * float(S)
        h.add_sphere(rad, -rad, rad * float(rng.uniform(0.3, 1)), float(rng.uniform(3, 6.2832)), mat, transform=t)
    d = h.build()

    def rays_of(o, dvec, tmax=np.inf):
        r = np.zeros(o.shape[0], api.RAY_DTYPE)
        r["o"], r["d"], r["tmax"] = o.astype(np.float32), dvec.astype(np.float32), tmax
        return r
    parts = []
    o = rng.uniform(-1.5, 1.5, (n, 3)) * S; v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    parts.append(rays_of(o, v))
    parts.append(rays_of(o, v * S))                      # unnormalised directions
    verts = pos[rng.integers(0, pos.shape[0], n)]
    parts.append(rays_of(verts, v))                      # origins on vertices = on box planes
    axes = np.eye(3, dtype=np.float32)[rng.integers(0, 3, n)] * rng.choice(np.float32([-1, 1]), (n, 1))
    o2 = verts.copy(); o2 -= axes * np.float32(3.0) * S
    parts.append(rays_of(o2, axes))                      # axis-parallel rays through vertex coordinates
    tgt = pos[rng.integers(0, pos.shape[0], n)]; dv = tgt - o
    parts.append(rays_of(o, dv, rng.choice(np.float32([np.inf, 1.0, 1.0000001, 0.9999999]), n)))   # aimed at vertices, tmax at the hit
    rays = np.concatenate(parts)
    return h, d, rays                                    # `h` owns the buffers `d` points into


@pytest.mark.parametrize("width", [2, 4, 8])
@pytest.mark.parametrize("scale", [1e-3, 1.0, 3e4])
def test_random_soups_with_degenerate_geometry_are_bit_exact(ctx, width, scale):
    """Fuzz of the walk on geometry the regular fixtures do not have: triangle soups at three coordinate scales with axis-aligned
    (zero-thickness boxes), degenerate (collinear / repeated vertices) and duplicated triangles (exact ties: the first in slot order
    wins on both sides), clipped and transformed spheres; rays from random points, from vertex coordinates and along the axes."""
    h, d, rays = _soup_case(np.random.default_rng(1000 * width + int(math.log10(scale) * 7) + 77), scale)
    osc = O.OracleScene(d)
    ctx.set_option(L.ARN_OPT_BVH_WIDTH, width)
    try:
        sc = ctx.upload(d)
        gh, ga = sc.intersect_closest(rays), sc.intersect_any(rays)
    finally:
        ctx.set_option(L.ARN_OPT_BVH_WIDTH, 0)
    oh = osc.intersect_closest(rays)
    assert (oh["prim_id"] >= 0).sum() > rays.shape[0] // 8
    mism = np.nonzero((gh["prim_id"] != oh["prim_id"]) | (gh["t"] != oh["t"]))[0]
    assert mism.size == 0, f"{mism.size} mismatches, first rays {mism[:5]}: o {rays['o'][mism[:3]]} d {rays['d'][mism[:3]]} gpu {gh[mism[:3]]} oracle {oh[mism[:3]]}"
    assert np.array_equal(ga != 0, oh["prim_id"] >= 0)
    sc.close(); osc.close()


def test_fast_transcendentals_are_the_library_values_for_every_f32(ctx):
    """kernels/cr_math.cuh on the device, all 2^32 arguments: the short f64 kernels + rounding-certainty test return exactly
    (float)libdevice_f64(x) for sin, cos, exp, log (and pow with a pseudo-random exponent per base)."""
    import ctypes as C
    out = (C.c_uint64 * 5)()
    rc = ctx.lib.arn_selftest_math(ctx.c, 0, 32, out)
    assert rc == 0, ctx.error()
    assert list(out) == [0, 0, 0, 0, 0], f"mismatches (sin, cos, exp, log, pow) = {list(out)}"


def test_lane_refilling_trace_is_bit_exact(ctx, cornell_small):
    """ARN_OPT_TRACE_REFILL: ray stream + persistent warps that hand idle lanes new rays (kernels/trace_refill.cuh).  Same
    per-ray arithmetic, so every camera sample's radiance, the ray counts and the oracle parity are unchanged."""
    hs, cam, film, smp, prm = cornell_small
    sc = ctx.upload(hs.desc()); osc = O.OracleScene(hs.desc())
    f0, rad0, st0 = sc.render_pt_samples(cam, film, smp, prm)
    ctx.set_option(L.ARN_OPT_TRACE_REFILL, 1)
    try:
        f1, rad1, st1 = sc.render_pt_samples(cam, film, smp, prm)
        # a few wave / pipeline shapes (ragged last wave, one pipeline)
        ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, 5000); ctx.set_option(L.ARN_OPT_PIPELINES, 1)
        f2, rad2, st2 = sc.render_pt_samples(cam, film, smp, prm)
    finally:
        ctx.set_option(L.ARN_OPT_TRACE_REFILL, 0); ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, 0); ctx.set_option(L.ARN_OPT_PIPELINES, 0)
    assert np.array_equal(rad0, rad1) and np.array_equal(rad0, rad2)
    assert (st0.extend_rays, st0.shadow_rays, st0.mis_rays) == (st1.extend_rays, st1.shadow_rays, st1.mis_rays) == (st2.extend_rays, st2.shadow_rays, st2.mis_rays)
    assert np.allclose(f0, f1, rtol=1e-5, atol=1e-6)
    _, orad = osc.render_pt_samples(cam, film, smp, prm)
    assert np.array_equal(rad1[..., :3], orad[..., :3])
    sc.close(); osc.close()


def test_compressed_8_wide_walk_renders_bit_exact(ctx, cornell_small):
    """ARN_OPT_BVH_WIDTH = 8: quantised 8-wide nodes + leaf blob (kernels/cw8_build.cuh, traverse8).  Every camera sample of the
    Cornell render (spheres, any-hit shadow rays, light rays) equals the oracle's; closest hits on 100 K random rays equal the
    exact binary walk's."""
    import torch
    hs, cam, film, smp, prm = cornell_small
    d = hs.desc()
    osc = O.OracleScene(d)
    ctx.set_option(L.ARN_OPT_BVH_WIDTH, 8)
    try:
        sc = ctx.upload(d)
        f8, rad8, st8 = sc.render_pt_samples(cam, film, smp, prm)
        rng = np.random.default_rng(21)
        n = 100_000
        rays = np.zeros(n, api.RAY_DTYPE)
        rays["o"] = rng.uniform([-1.8, -1.3, 2.2], [1.8, 2.2, 5.8], (n, 3)).astype(np.float32)
        v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
        rays["d"] = v.astype(np.float32)
        rays["tmax"] = np.where(rng.random(n) < 0.3, rng.uniform(0.1, 4.0, n), np.inf).astype(np.float32)
        rd, hd = _dev(rays)
        sc.intersect_closest_dev(rd.data_ptr(), n, hd.data_ptr())
        ctx.synchronize()                                   # the *_dev queries are asynchronous on the context's stream
        h8 = hd.cpu().numpy().copy()
        hb = torch.empty_like(hd)
        sc.intersect_closest_counted_dev(rd.data_ptr(), n, hb.data_ptr())           # the exact binary walk
        a8 = sc.intersect_any(rays)
    finally:
        ctx.set_option(L.ARN_OPT_BVH_WIDTH, 0)
    _, orad = osc.render_pt_samples(cam, film, smp, prm)
    _, ost, _ = osc.render_pt(cam, film, smp, prm)
    assert np.array_equal(rad8[..., :3], orad[..., :3])
    assert (st8.extend_rays, st8.shadow_rays, st8.mis_rays) == (ost.extend_rays, ost.shadow_rays, ost.mis_rays)
    assert np.array_equal(h8, hb.cpu().numpy())
    assert np.array_equal(a8 != 0, hb.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)["prim_id"] >= 0)
    sc.close(); osc.close()


def _random_pyramid(rng, channels, scale):
    h, w = int(rng.choice([1, 2, 8, 16, 32])), int(rng.choice([2, 4, 16, 32]))
    lv = []
    while True:
        shape = (h, w, 3) if channels == 3 else (h, w)
        lv.append((scale * rng.random(shape)).astype(np.float32))
        if h == 1 and w == 1:
            break
        h, w = max(h // 2, 1), max(w // 2, 1)
    return lv


def _random_scene(seed, textured=False, wild=False):
    """A random scene for the shading fuzz: a closed room of random quads, a soup of triangles with and without shading normals,
    random materials of all four families with parameters out to the clamps, sphere emitters (clipped / transformed) and a random
    set of delta / distant lights, a random camera (perspective with or without lens, or orthographic).  `textured`: random image
    textures (RGB on kd / ks, Luma on sigma | roughness and as bump maps; trilinear and EWA, all wrap modes) on most materials."""
    rng = np.random.default_rng(seed)
    hs = api.HostScene()
    rgb_tex, luma_tex, bump_tex = [0], [0], [0]
    if textured:
        wraps = [L.ARN_WRAP_REPEAT, L.ARN_WRAP_CLAMP, L.ARN_WRAP_BLACK]
        def tex(channels, scale):
            return hs.add_texture(_random_pyramid(rng, channels, scale), trilinear=bool(rng.random() < 0.5), max_aniso=float(rng.choice([1.0, 4.0, 8.0, 16.0])),
                                  wrapping=wraps[int(rng.integers(0, 3))], scaling=tuple(float(x) for x in rng.uniform(0.3, 6, 2)), shifting=tuple(float(x) for x in rng.uniform(-1, 1, 2)))
        rgb_tex += [tex(3, 1.0) for _ in range(3)]
        luma_tex += [tex(1, float(rng.choice([0.6, 30.0]))) for _ in range(2)]
        bump_tex += [tex(1, float(rng.choice([0.01, 0.2]))) for _ in range(2)]
    mats = []
    for _ in range(10):
        kind = int(rng.integers(0, 4))
        col = lambda lo=0.0, hi=1.0: tuple(float(x) for x in rng.uniform(lo, hi, 3))
        rough = float(rng.choice([0.001, 0.01, 0.08, 0.3, 0.7, 1.0, 2.5]))
        if kind == 0:
            m = api.material(L.ARN_MAT_MATTE, kd=col(), sigma=float(rng.choice([0.0, 0.0, 0.3, 20.0, 90.0, 140.0])))
        elif kind == 1:
            m = api.material(L.ARN_MAT_PLASTIC, kd=col(), ks=col(), roughness=rough)
        elif kind == 2:
            m = api.material(L.ARN_MAT_GLASS, kd=col() if rng.random() < 0.7 else (0, 0, 0), ks=col() if rng.random() < 0.7 else (0, 0, 0), roughness=rough,
                             eta=float(rng.choice([1.0001, 1.2, 1.5, 2.4, 0.75])))
        else:
            m = api.material(L.ARN_MAT_TRANSLUCENT, kd=col(), ks=col(), roughness=rough, dissolve=float(rng.choice([0.0, 0.3, 0.8, 1.0])))
        if textured:
            m.kd_tex, m.ks_tex = int(rng.choice(rgb_tex)), int(rng.choice(rgb_tex))
            m.aux_tex, m.bump_tex = int(rng.choice(luma_tex)), int(rng.choice(bump_tex))
        mats.append(hs.add_material(m))
    pick = lambda: mats[int(rng.integers(0, len(mats)))]
    # room [-3, 3]^3, each wall one quad
    B = 3.0
    walls = [([-B, -B, -B], [B, -B, -B], [B, -B, B], [-B, -B, B]), ([-B, B, -B], [-B, B, B], [B, B, B], [B, B, -B]),
             ([-B, -B, -B], [-B, B, -B], [B, B, -B], [B, -B, -B]), ([-B, -B, B], [B, -B, B], [B, B, B], [-B, B, B]),
             ([-B, -B, -B], [-B, -B, B], [-B, B, B], [-B, B, -B]), ([B, -B, -B], [B, B, -B], [B, B, B], [B, -B, B])]
    for q in walls:
        hs.add_mesh(np.float32(q), np.uint32([0, 1, 2, 0, 2, 3]), pick())
    for _ in range(14):                                   # objects: small random meshes, half of them with shading normals / uvs
        c = rng.uniform(-2.2, 2.2, 3); nt = int(rng.integers(1, 12))
        pos = (c + rng.normal(scale=0.5, size=(nt * 3, 3))).astype(np.float32)
        idx = np.arange(nt * 3, dtype=np.uint32)
        nrm = uv = None
        if rng.random() < 0.5:
            nrm = rng.normal(size=(nt * 3, 3)).astype(np.float32); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        if rng.random() < 0.5:
            uv = rng.uniform(0, 1, (nt * 3, 2)).astype(np.float32)
        hs.add_mesh(pos, idx, pick(), normals=nrm, uvs=uv)
    for k in range(int(rng.integers(1, 4))):              # sphere emitters
        rad = float(rng.uniform(0.15, 0.6))
        t = np.eye(4, dtype=np.float32); t[3, :3] = rng.uniform(-2.3, 2.3, 3)
        if rng.random() < 0.5:
            a = rng.uniform(0, 6.28); t[1, 1], t[1, 2], t[2, 1], t[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
        full = rng.random() < 0.5
        hs.add_sphere(rad, -rad if full else -rad * float(rng.uniform(0.2, 0.9)), rad if full else rad * float(rng.uniform(0.2, 0.9)),
                      6.2831855 if full else float(rng.uniform(2.0, 6.0)), pick(), emission=tuple(float(x) for x in rng.uniform(2, 30, 3)), transform=t)
    for k in range(2):                                    # non-emissive spheres
        rad = float(rng.uniform(0.3, 0.8)); t = np.eye(4, dtype=np.float32); t[3, :3] = rng.uniform(-2, 2, 3)
        hs.add_sphere(rad, -rad, rad, 6.2831855, pick(), transform=t)
    if wild:                                              # general affine instances: rotation x non-uniform scale, clipped, some emissive
        rw = np.random.default_rng(seed + 77777)
        for k in range(4):
            q, _ = np.linalg.qr(rw.normal(size=(3, 3)))
            m = (q * rw.uniform(0.5, 1.8, 3)).astype(np.float32)
            t = np.eye(4, dtype=np.float32); t[:3, :3] = m; t[3, :3] = rw.uniform(-2.2, 2.2, 3)
            rad = float(rw.uniform(0.25, 0.6)); full = rw.random() < 0.5
            hs.add_sphere(rad, -rad if full else -rad * float(rw.uniform(0.2, 0.9)), rad if full else rad * float(rw.uniform(0.2, 0.9)),
                          6.2831855 if full else float(rw.uniform(2.0, 6.0)), pick(),
                          emission=tuple(float(x) for x in rw.uniform(2, 20, 3)) if k < 2 else None, transform=t)
    if rng.random() < 0.7:
        hs.add_light(api.point_light(rng.uniform(-2, 2, 3), rng.uniform(1, 20, 3)))
    if rng.random() < 0.7:
        hs.add_light(api.spot_light(rng.uniform(-2, 2, 3), rng.uniform(-1, 1, 3), rng.uniform(5, 40, 3), float(rng.uniform(0.4, 1.2)), float(rng.uniform(0.1, 0.35))))
    if rng.random() < 0.5:
        dv = rng.normal(size=3); hs.add_light(api.distant_light(rng.uniform(0.2, 2, 3), dv / np.linalg.norm(dv), 6.0))
    hs.build()
    eye = rng.uniform(-2.0, 2.0, 3); eye[2] = 2.6
    view_parent = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, -1, 0], [eye[0], eye[1], eye[2], 1]], np.float32)      # looking along -z
    w, h = 56, 40
    lens = (float(rng.uniform(0.02, 0.2)), float(rng.uniform(2, 5))) if rng.random() < 0.4 else None
    if rng.random() < 0.25:
        cam = api.make_ortho_camera(view_parent.reshape(-1), (-2.5, -1.8, 2.5, 1.8), 0.1, 100.0, w, h, lens=lens)
    else:
        parent_view = np.linalg.inv(view_parent.T).T.astype(np.float32)
        cam = api.make_camera(parent_view.reshape(-1), (-1.3, -0.95, 1.3, 0.95), 0.1, 100.0, float(rng.uniform(0.7, 1.4)), w, h, lens=lens)
    depth = int(rng.choice([1, 2, 5, 8, 12]))
    prm = api.make_pt_params(max_depth=depth, rr_threshold=float(rng.choice([0.05, 0.5, 0.0])))
    return hs, cam, api.make_film(w, h), api.make_sampler(2, 2, int(rng.choice([2, 8])), int(rng.integers(0, 1 << 30))), prm


@pytest.mark.parametrize("seed", [0, 2, 3, 4, 5, 6, 7, 9, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 24, 25, 26, 27])
def test_random_scenes_per_sample_radiance_is_bit_exact(ctx, seed):
    """Shading fuzz: every camera sample's radiance through the whole bounce loop equals the oracle's bit for bit on random scenes
    (materials of all four families out to their parameter clamps, partial / transformed sphere emitters, Point / Spot / Distant
    lights, thin lens / orthographic cameras, depths 1 .. 12, three Russian-roulette thresholds); so do the traced-ray counters."""
    hs, cam, film, smp, prm = _random_scene(seed)
    d = hs.desc()
    sc = ctx.upload(d); osc = O.OracleScene(d)
    _, grad, st = sc.render_pt_samples(cam, film, smp, prm)
    _, orad = osc.render_pt_samples(cam, film, smp, prm)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
    assert same.all(), f"seed {seed}: {(~same).sum()} of {same.size} samples differ, first {np.argwhere(~same)[:3].tolist()}: gpu {grad[~same][:2]} oracle {orad[~same][:2]}"
    _, ost, _ = osc.render_pt(cam, film, smp, prm)
    assert (st.extend_rays, st.shadow_rays, st.mis_rays, st.invalid_samples) == (ost.extend_rays, ost.shadow_rays, ost.mis_rays, ost.invalid_samples)
    assert (orad[..., :3].max(-1) > 0).mean() > 0.08
    sc.close(); osc.close()


def _probe_material(rng, kind):
    col = lambda: tuple(float(x) for x in rng.uniform(0, 1, 3))
    rough = float(rng.choice([0.0005, 0.001, 0.01, 0.08, 0.3, 0.7, 1.0, 2.5]))
    if kind == 0:
        m = api.material(L.ARN_MAT_MATTE, kd=col(), sigma=float(rng.choice([0.0, 0.3, 20.0, 90.0, 140.0])))
    elif kind == 1:
        m = api.material(L.ARN_MAT_PLASTIC, kd=col(), ks=col(), roughness=rough)
    elif kind == 2:
        m = api.material(L.ARN_MAT_GLASS, kd=col() if rng.random() < 0.8 else (0, 0, 0), ks=col() if rng.random() < 0.6 else (0, 0, 0), roughness=rough,
                         eta=float(rng.choice([1.0001, 1.2, 1.5, 2.4, 0.75])))
    else:
        m = api.material(L.ARN_MAT_TRANSLUCENT, kd=col(), ks=col(), roughness=rough, dissolve=float(rng.choice([0.0, 0.3, 0.8, 1.0])))
    m.alpha = O.load().arn_oracle_roughness_to_alpha(m.roughness)
    return m


@pytest.mark.parametrize("general_frame", [False, True])
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_bsdf_probe_is_bit_exact(ctx, kind, general_frame):
    """Kernel-level parity of Bsdf::{evaluate_sampled, evaluate, pdf} (material/bsdf.rs) for every material family: 12 random
    materials x 4096 (wo, u, wi) triples incl. grazing, axis-aligned and below-horizon directions, in the canonical frame and in
    random frames (dpdu not orthogonal to the shading normal, shading normal != geometric normal) — the device probe
    (arn_selftest_bsdf) returns the oracle's twelve floats bit for bit."""
    import ctypes as C
    rng = np.random.default_rng(4242 + kind + 10 * int(general_frame))
    n = 4096
    olib = O.load()
    for rep in range(12):
        m = _probe_material(rng, kind)
        def dirs():
            v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
            v[: n // 8, 2] *= 1e-3                                     # grazing (renormalised below)
            v[n // 8: n // 8 + 16] = np.float32([[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, 1, 0]] * 4)
            v /= np.linalg.norm(v, axis=1, keepdims=True)
            return np.ascontiguousarray(v, np.float32)
        wo, wi = dirs(), dirs()
        wi[n // 2: n // 2 + 512] = wo[n // 2: n // 2 + 512] * np.float32([-1, -1, 1])      # mirror directions: the microfacet peak
        u = np.ascontiguousarray(rng.uniform(0, 1, (n, 2)), np.float32); u[:8] = np.float32([[0, 0], [0.5, 0.5], [0.999999, 0.999999], [0.5, 0], [0, 0.5], [0.25, 0.75], [0.49999997, 0.1], [0.75, 1e-7]])
        fr = None
        if general_frame:
            ns = rng.normal(size=(n, 3)); ns /= np.linalg.norm(ns, axis=1, keepdims=True)
            ng = ns + rng.normal(scale=0.2, size=(n, 3)); ng /= np.linalg.norm(ng, axis=1, keepdims=True)
            dpdu = np.cross(ng, rng.normal(size=(n, 3))) * rng.uniform(0.1, 5, (n, 1))
            fr = np.ascontiguousarray(np.concatenate([dpdu, ns, ng], 1), np.float32)
        gout = np.zeros((n, 12), np.float32)
        rc = ctx.lib.arn_selftest_bsdf(ctx.c, C.byref(m), n, wo.ctypes.data, u.ctypes.data, wi.ctypes.data, fr.ctypes.data if fr is not None else None, gout.ctypes.data)
        assert rc == 0, ctx.error()
        oout = np.zeros((n, 12), np.float32)
        for i in range(n):
            olib.arn_oracle_bsdf_probe2(C.byref(m), wo[i].ctypes.data, u[i].ctypes.data, wi[i].ctypes.data, fr[i].ctypes.data if fr is not None else None, oout[i].ctypes.data)
        gb, ob = gout.view(np.uint32), oout.view(np.uint32)
        same = (gb == ob) | (np.isnan(gout) & np.isnan(oout))
        bad = np.argwhere(~same.all(axis=1))[:, 0]
        assert bad.size == 0, (f"material {rep} (type {m.type} rough {m.roughness} eta {m.eta} dissolve {m.dissolve} sigma {m.sigma} kd {list(m.kd)} ks {list(m.ks)}): {bad.size} of {n} probes differ; "
                               f"first {bad[0]}: wo {wo[bad[0]]} u {u[bad[0]]} wi {wi[bad[0]]}\n gpu    {gout[bad[0]]}\n oracle {oout[bad[0]]}")


@pytest.mark.parametrize("seed", range(100, 112))
def test_random_textured_scenes_per_sample_radiance_is_bit_exact(ctx, seed):
    """The shading fuzz with image textures: MipMap look-ups (trilinear / EWA, three wrap modes, pyramids from 1 x 2 to 32 x 32 texels),
    UV mappings, ray differentials through up to 12 bounces, bump mapping on meshes with and without shading normals / uvs and on
    rotated, clipped spheres — through the textured shade instance, every camera sample equals the oracle's."""
    hs, cam, film, smp, prm = _random_scene(seed, textured=True)
    d = hs.desc()
    assert d.n_textures == 7
    sc = ctx.upload(d); osc = O.OracleScene(d)
    _, grad, st = sc.render_pt_samples(cam, film, smp, prm)
    _, orad = osc.render_pt_samples(cam, film, smp, prm)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1) | np.all(np.isnan(grad) == np.isnan(orad), axis=-1) & np.all((grad == orad) | np.isnan(grad), axis=-1)
    assert same.all(), f"seed {seed}: {(~same).sum()} of {same.size} samples differ, first {np.argwhere(~same)[:3].tolist()}: gpu {grad[~same][:2]} oracle {orad[~same][:2]}"
    _, ost, _ = osc.render_pt(cam, film, smp, prm)
    assert (st.extend_rays, st.shadow_rays, st.mis_rays, st.invalid_samples) == (ost.extend_rays, ost.shadow_rays, ost.mis_rays, ost.invalid_samples)
    sc.close(); osc.close()


@pytest.mark.parametrize("seed", range(16))
def test_random_film_configurations_match_the_oracle(ctx, seed):
    """Film fuzz (k_accumulate_px / k_accumulate): random resolution, crop window, filter kind and radius (<= 4: warp-per-pixel
    gather, > 4: scatter), tile grid, samples per pixel (chunks of 32 per warp: counts around that), wave capacity (pixels whose
    samples straddle waves), rank / world / cell subdivision — the film equals the oracle's up to the order of the float additions."""
    rng = np.random.default_rng(900 + seed)
    w, h = int(rng.integers(20, 70)), int(rng.integers(16, 50))
    sx, sy = [(1, 1), (3, 1), (2, 2), (5, 7), (6, 6), (8, 4), (11, 3)][int(rng.integers(0, 7))]
    hs, cam, _, _, _ = scenes.cornell_scene(w, h, sx, sy)
    x0, y0 = int(rng.integers(0, w // 3)), int(rng.integers(0, h // 3))
    crop = (0, 0, w, h) if rng.random() < 0.4 else (x0, y0, int(rng.integers(x0 + 8, w + 1)), int(rng.integers(y0 + 8, h + 1)))
    kinds = [(0, 0.0, 0.0), (L.ARN_FILTER_BOX, 0.0, 0.0), (L.ARN_FILTER_TRIANGLE, 0.0, 0.0), (L.ARN_FILTER_GAUSSIAN, 0.7, 0.0), (L.ARN_FILTER_MITCHELL, 1.0 / 3, 1.0 / 3)]
    kind, fa, fb = kinds[int(rng.integers(0, 5))]
    radius = [(4.0, 4.0), (2.0, 2.0), (0.5, 0.5), (2.7, 1.3), (1.0, 3.9), (6.0, 5.0), (4.5, 2.0)][int(rng.integers(0, 7))]
    film = api.make_film(w, h, crop=crop, filter_radius=radius, filter_kind=kind, filter_a=fa, filter_b=fb)
    smp = api.make_sampler(sx, sy, 8, int(rng.integers(0, 1 << 30)))
    cw, ch = crop[2] - crop[0], crop[3] - crop[1]
    tiles = (int(rng.integers(1, min(cw, 9))), int(rng.integers(1, min(ch, 9))))
    world = int(rng.choice([1, 1, 2, 3])); rank = int(rng.integers(0, world)); subdiv = int(rng.choice([0, 0, 2, 4]))
    prm = api.make_pt_params(max_depth=3, tiles=tiles, rank=rank, world_size=world, subdiv=subdiv)
    d = hs.desc()
    sc = ctx.upload(d); osc = O.OracleScene(d)
    cfg = f"res {w}x{h} crop {crop} spp {sx}x{sy} filter {kind} r {radius} tiles {tiles} rank {rank}/{world} subdiv {subdiv}"
    wave = int(rng.choice([0, 1024, 4096]))
    try:
        of, ost, _ = osc.render_pt(cam, film, smp, prm)
    except Exception:
        # the reference panics here (film.rs:129: a tile of the grid laid out from (0, 0) misses the crop window): both sides refuse
        with pytest.raises(api.ArnError):
            sc.render_pt(cam, film, smp, prm)
        sc.close(); osc.close()
        return
    ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, wave)
    try:
        gf, st = sc.render_pt(cam, film, smp, prm)
    finally:
        ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, 0)
    assert st.camera_rays == ost.camera_rays, cfg                 # 0 when the rank owns no tile of a coarse grid: an empty film on both sides
    assert gf.shape == of.shape and np.isfinite(gf).all(), cfg
    scale = max(float(np.abs(of).max()), 1e-6)
    assert np.abs(gf - of).max() <= 1e-5 * scale, f"{cfg}: max diff {np.abs(gf - of).max()} of {scale}"
    sc.close(); osc.close()


@pytest.mark.parametrize("case", ["cornell", "zoo_lens", "random3", "textured104"])
def test_stratified_sampler_mode_is_bit_exact(ctx, case):
    """ARN_SAMPLER_STRATIFIED (the reference's sampler as intended; include/arn.h): the device's stratified draws — film jitter and lens
    in k_generate, light choice / light / scatter / BSDF / roulette draws in k_shade — are the oracle's, so every camera sample's radiance
    still is, bit for bit; and the mode does change the picture's samples (it is not silently the parity sampler)."""
    if case == "cornell":
        hs, cam, film, smp, prm = scenes.cornell_scene(64, 48, 4, 4)
    elif case == "zoo_lens":
        import test_gpu_parity as P
        hs, cam, film, smp, prm = P._material_zoo((0.15, 7.5))
    elif case == "random3":
        hs, cam, film, smp, prm = _random_scene(3)
    else:
        hs, cam, film, smp, prm = _random_scene(104, textured=True)
    strat = api.make_sampler(smp.sampledx, smp.sampledy, smp.ndim, smp.seed, mode=L.ARN_SAMPLER_STRATIFIED)
    d = hs.desc()
    sc = ctx.upload(d); osc = O.OracleScene(d)
    gf, grad, st = sc.render_pt_samples(cam, film, strat, prm)
    of, orad = osc.render_pt_samples(cam, film, strat, prm)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
    assert same.all(), f"{(~same).sum()} of {same.size} samples differ, first {np.argwhere(~same)[:3].tolist()}"
    _, ost, _ = osc.render_pt(cam, film, strat, prm)
    assert (st.extend_rays, st.shadow_rays, st.mis_rays, st.invalid_samples) == (ost.extend_rays, ost.shadow_rays, ost.mis_rays, ost.invalid_samples)
    assert np.abs(gf - of).max() <= 2e-5 * max(float(np.abs(of).max()), 1e-6)
    _, prad, _ = sc.render_pt_samples(cam, film, smp, prm)
    assert not np.array_equal(prad, grad)
    # the film jitter is stratified: per pixel the spp film positions fall into distinct cells of the sampledx x sampledy grid
    with pytest.raises(api.ArnError):
        sc.render_pt(cam, film, api.make_sampler(2, 2, 8, 0, mode=7), prm)
    sc.close(); osc.close()


@pytest.mark.parametrize("seed", range(200, 212))
def test_random_scenes_with_affine_sphere_instances_are_bit_exact(ctx, seed):
    """The shading fuzz with general affine sphere instances (rotation x non-uniform scale, clipped, two of them emissive): every
    accepted hit on one of them rewrites the traversal ray (bvh.rs:108-113), their light samples and pdfs live in the local frame
    (transformed.rs:120-146).  Per-sample radiance and ray counts equal the oracle's."""
    hs, cam, film, smp, prm = _random_scene(seed, textured=(seed % 3 == 0), wild=True)
    d = hs.desc()
    sc = ctx.upload(d); osc = O.OracleScene(d)
    _, grad, st = sc.render_pt_samples(cam, film, smp, prm)
    _, orad = osc.render_pt_samples(cam, film, smp, prm)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1) | (np.isnan(grad).any(-1) & np.isnan(orad).any(-1))
    assert same.all(), f"seed {seed}: {(~same).sum()} of {same.size} samples differ, first {np.argwhere(~same)[:3].tolist()}: gpu {grad[~same][:2]} oracle {orad[~same][:2]}"
    _, ost, _ = osc.render_pt(cam, film, smp, prm)
    assert (st.extend_rays, st.shadow_rays, st.mis_rays, st.invalid_samples) == (ost.extend_rays, ost.shadow_rays, ost.mis_rays, ost.invalid_samples)
    sc.close(); osc.close()


@pytest.mark.parametrize("scene", ["cornell", "box972", "box2352"])
def test_shared_memory_pair_walk_is_bit_exact(ctx, cornell_small, scene):
    """Small trees are walked by k_trace from pair records staged in shared memory (traverse2p: sign-selected LDS.128 per axis,
    packed FFMA2 conservative test, 4-byte stack entries).  Same leaves in the same order as the global-memory walk: every
    camera sample's radiance and the ray counts are those of ARN_OPT_SMEM_NODES = 1 (never) and of the oracle.  box972 fits
    the shared-memory budget (734 interior nodes), box2352 (1783) does not and takes the global-memory walk either way."""
    if scene == "cornell":
        hs, cam, film, smp, prm = cornell_small
    else:
        hs, cam, film, smp, prm = scenes.c4_box_scene(cells=9 if scene == "box972" else 14, res=64, sampledx=2, sampledy=2)
    d = hs.desc()
    assert (((d.n_nodes - 1) // 2) * 128 <= L.ARN_SMEM_NODE_BYTES) == (scene != "box2352")
    sc = ctx.upload(d); osc = O.OracleScene(d)
    f0, rad0, st0 = sc.render_pt_samples(cam, film, smp, prm)
    ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, 5000); ctx.set_option(L.ARN_OPT_PIPELINES, 1)      # ragged waves, thin launches
    try:
        f2, rad2, st2 = sc.render_pt_samples(cam, film, smp, prm)
        ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, 0); ctx.set_option(L.ARN_OPT_PIPELINES, 0)
        ctx.set_option(L.ARN_OPT_SMEM_NODES, 1)
        f1, rad1, st1 = sc.render_pt_samples(cam, film, smp, prm)
    finally:
        ctx.set_option(L.ARN_OPT_SMEM_NODES, 0); ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, 0); ctx.set_option(L.ARN_OPT_PIPELINES, 0)
    assert np.array_equal(rad0, rad1) and np.array_equal(rad0, rad2)
    assert (st0.extend_rays, st0.shadow_rays, st0.mis_rays) == (st1.extend_rays, st1.shadow_rays, st1.mis_rays) == (st2.extend_rays, st2.shadow_rays, st2.mis_rays)
    assert np.allclose(f0, f1, rtol=1e-5, atol=1e-6)
    _, orad = osc.render_pt_samples(cam, film, smp, prm)
    assert np.array_equal(rad0[..., :3], orad[..., :3])
    sc.close(); osc.close()


def test_shared_memory_walk_is_bit_exact_on_adversarial_rays(ctx):
    """The same rays against a tree small enough for the pair records: batches of >= 2^15 rays take the shared-memory walk
    (traverse2p: sign-selected LDS.128, FFMA2, truncated entry distances on the stack) in k_closest_batch / k_any_batch.  Ids and t
    equal the oracle's and those of the global-memory walk (ARN_OPT_SMEM_NODES = 1) bit for bit."""
    h, d, rays = _adversarial_case(24, 5)
    assert ((d.n_nodes - 1) // 2) * 128 <= L.ARN_SMEM_NODE_BYTES and rays.shape[0] >= 1 << 15
    osc = O.OracleScene(d); sc = ctx.upload(d)
    gh, ga = sc.intersect_closest(rays), sc.intersect_any(rays)
    ctx.set_option(L.ARN_OPT_SMEM_NODES, 1)
    try:
        gh1, ga1 = sc.intersect_closest(rays), sc.intersect_any(rays)
    finally:
        ctx.set_option(L.ARN_OPT_SMEM_NODES, 0)
    oh = osc.intersect_closest(rays)
    assert (oh["prim_id"] >= 0).sum() > rays.shape[0] // 10
    assert (oh["prim_id"] == d.n_prims - 1).sum() > 100, "the sphere should be reached"
    mism = np.nonzero((gh["prim_id"] != oh["prim_id"]) | (gh["t"] != oh["t"]))[0]
    assert mism.size == 0, f"{mism.size} mismatches, first rays {mism[:5]}: o {rays['o'][mism[:3]]} d {rays['d'][mism[:3]]} gpu {gh[mism[:3]]} oracle {oh[mism[:3]]}"
    assert np.array_equal(gh, gh1) and np.array_equal(ga, ga1)
    assert np.array_equal(ga != 0, oh["prim_id"] >= 0)
    sc.close(); osc.close()


@pytest.mark.parametrize("scale", [1e-3, 1.0, 3e4])
def test_shared_memory_walk_on_random_soups_with_degenerate_geometry(ctx, scale):
    """Traversal fuzz (flat, collinear, duplicated triangles, clipped and affine spheres, origins on box planes, tmax at the hit) on
    soups small enough for the shared-memory walk, at three coordinate scales."""
    for seed in range(3):
        h, d, rays = _soup_case(np.random.default_rng(100 + seed), scale, n_tri=1000, n=8000)
        assert ((d.n_nodes - 1) // 2) * 128 <= L.ARN_SMEM_NODE_BYTES and rays.shape[0] >= 1 << 15
        osc = O.OracleScene(d); sc = ctx.upload(d)
        gh, ga = sc.intersect_closest(rays), sc.intersect_any(rays)
        oh = osc.intersect_closest(rays)
        mism = np.nonzero((gh["prim_id"] != oh["prim_id"]) | (gh["t"] != oh["t"]))[0]
        assert mism.size == 0, f"seed {seed}: {mism.size} mismatches, first rays {mism[:5]}"
        assert np.array_equal(ga != 0, oh["prim_id"] >= 0)
        sc.close(); osc.close()




def test_programmatic_dependent_launch_is_bit_exact(ctx, cornell_small):
    """ARN_OPT_PDL: the kernels of a pipeline are launched with the programmatic-stream-serialization attribute and park at
    griddepcontrol.wait until their predecessor has completed (measured slower, off by default).  Same results."""
    hs, cam, film, smp, prm = cornell_small
    sc = ctx.upload(hs.desc())
    f0, rad0, st0 = sc.render_pt_samples(cam, film, smp, prm)
    ctx.set_option(L.ARN_OPT_PDL, 1); ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, 3000); ctx.set_option(L.ARN_OPT_PIPELINES, 3)
    try:
        f1, rad1, st1 = sc.render_pt_samples(cam, film, smp, prm)
    finally:
        ctx.set_option(L.ARN_OPT_PDL, 0); ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, 0); ctx.set_option(L.ARN_OPT_PIPELINES, 0)
    assert np.array_equal(rad0, rad1)
    assert (st0.extend_rays, st0.shadow_rays, st0.mis_rays) == (st1.extend_rays, st1.shadow_rays, st1.mis_rays)
    assert np.allclose(f0, f1, rtol=1e-5, atol=1e-6)
    sc.close()
