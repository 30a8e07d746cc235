"""GPU parity tests proper: CUDA path (through the C-ABI) vs the oracle on the same inputs."""
import math

import numpy as np
import pytest

import oracle_lib as O
from arendur_b200 import api, scenes, _lib as L

pytestmark = pytest.mark.gpu


def _camera_grid_rays(cam, w, h, step=1):
    xs, ys = np.meshgrid(np.arange(0, w, step) + 0.5, np.arange(0, h, step) + 0.5, indexing="xy")
    pf = np.zeros((xs.size, 4), np.float32)
    pf[:, 0], pf[:, 1] = xs.reshape(-1), ys.reshape(-1)
    return O.camera_rays(cam, pf)


def _assert_hits_equal(gh, oh):
    """Bit-exact prim ids and distances (north_star asks ids exact, t within 1e-5 relative)."""
    mism = np.nonzero(gh["prim_id"] != oh["prim_id"])[0]
    assert mism.size == 0, f"{mism.size} primitive-id mismatches, first at ray {mism[:5]}: gpu {gh['prim_id'][mism[:5]]} oracle {oh['prim_id'][mism[:5]]}"
    hit = oh["prim_id"] >= 0
    rel = np.abs(gh["t"][hit] - oh["t"][hit]) / oh["t"][hit]
    assert rel.size == 0 or rel.max() <= 1e-5, f"max relative t error {rel.max()}"
    assert np.array_equal(gh["t"][hit], oh["t"][hit]), "t is expected to be bit-exact (no FMA contraction)"
    assert np.all(np.isinf(gh["t"][~hit]))


def test_closest_hit_cornell_primary(ctx, cornell_small):
    hs, cam, film, smp, prm = cornell_small
    d = hs.desc()
    sc = ctx.upload(d)
    osc = O.OracleScene(d)
    rays = _camera_grid_rays(cam, film.res_x, film.res_y)
    _assert_hits_equal(sc.intersect_closest(rays), osc.intersect_closest(rays))
    sc.close(); osc.close()


def test_closest_hit_heightfield(ctx):
    """C2 at reduced size (128x128 cells = 32 768 triangles, 480x270 rays): bit-exact ids and t."""
    hs = api.HostScene()
    mat = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.7, 0.7, 0.7)))
    pos, idx = scenes.heightfield(128, -2.0, 2.0, 4.0, 0.15, 0x5EED)
    hs.add_mesh(pos, idx, mat)
    d = hs.build()
    w, h = 480, 270
    cam = api.make_camera(api.IDENTITY, (-16.0 / 9.0, -1.0, 16.0 / 9.0, 1.0), 0.1, 1000.0, math.pi / 2, w, h)
    rays = _camera_grid_rays(cam, w, h)
    sc = ctx.upload(d)
    osc = O.OracleScene(d)
    gh, oh = sc.intersect_closest(rays), osc.intersect_closest(rays)
    assert (oh["prim_id"] >= 0).sum() > 1000
    _assert_hits_equal(gh, oh)
    # any-hit agrees with closest-hit existence; finite tmax clips
    assert np.array_equal(sc.intersect_any(rays) != 0, oh["prim_id"] >= 0)
    clipped = rays.copy(); clipped["tmax"] = 3.9
    assert np.array_equal(sc.intersect_any(clipped), osc.intersect_any(clipped))
    sc.close(); osc.close()


def test_closest_hit_incoherent_rays(ctx, cornell_small):
    """Random origins inside the box, random directions: spheres, glass box and walls all get hit."""
    hs, cam, film, smp, prm = cornell_small
    d = hs.desc()
    rng = np.random.default_rng(7)
    n = 20000
    rays = np.zeros(n, api.RAY_DTYPE)
    rays["o"] = rng.uniform([-1.8, -1.3, 2.2], [1.8, 2.2, 5.8], (n, 3)).astype(np.float32)
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    rays["d"] = v.astype(np.float32)
    rays["tmax"] = np.inf
    # rays aimed at the two emissive spheres from inside the room (transformed-sphere path)
    tgt = np.where(rng.random(n // 4)[:, None] < 0.5, np.float32([-3, 0, -4.5]), np.float32([2, 2, -2.5]))
    dv = tgt + rng.normal(scale=0.8, size=(n // 4, 3)) - rays["o"][: n // 4]
    rays["d"][: n // 4] = (dv / np.linalg.norm(dv, axis=1, keepdims=True)).astype(np.float32)
    sc = ctx.upload(d)
    osc = O.OracleScene(d)
    gh, oh = sc.intersect_closest(rays), osc.intersect_closest(rays)
    assert (oh["prim_id"] >= 1112).sum() > 10, "test should reach the spheres"
    _assert_hits_equal(gh, oh)
    sc.close(); osc.close()


def test_empty_and_ragged_batches(ctx, cornell_small):
    hs, cam, film, smp, prm = cornell_small
    sc = ctx.upload(hs.desc())
    assert sc.intersect_closest(np.zeros(0, api.RAY_DTYPE)).shape == (0,)
    rays = _camera_grid_rays(cam, film.res_x, film.res_y)[:1001]      # not a multiple of the block size
    osc = O.OracleScene(hs.desc())
    _assert_hits_equal(sc.intersect_closest(rays), osc.intersect_closest(rays))
    sc.close(); osc.close()


def _rel_rmse(gpu_film, ref_film):
    g, _ = api.film_finalize(gpu_film)
    r, _ = O.film_finalize(ref_film)
    return float(np.sqrt(np.mean((g - r) ** 2)) / np.mean(r)), g, r


def test_render_cornell_matches_oracle(ctx, cornell_small):
    """Full PT loop, same sample sequence: per-pixel relative RMSE of the finalised image and
    identical ray counts up to rare libm-ulp decision flips (tolerances stated here)."""
    hs, cam, film, smp, prm = cornell_small
    d = hs.desc()
    sc = ctx.upload(d)
    osc = O.OracleScene(d)
    gf, st = sc.render_pt(cam, film, smp, prm)
    rf, ost, _ = osc.render_pt(cam, film, smp, prm)
    rmse, g, r = _rel_rmse(gf, rf)
    assert st.camera_rays == ost.camera_rays      # pixels covered by spawn_tiles (quirk A-15: 72 rows -> 64)
    for a, b, name in ((st.extend_rays, ost.extend_rays, "extend"), (st.shadow_rays, ost.shadow_rays, "shadow"), (st.mis_rays, ost.mis_rays, "mis")):
        assert abs(int(a) - int(b)) <= max(2, int(1e-5 * b)), f"{name} ray count gpu {a} vs oracle {b}"
    assert st.invalid_samples == ost.invalid_samples
    # every sample contributes bit-identical (L*w, w) terms; only the order of the float additions
    # differs (atomics vs the tile loop).  Stated tolerances: film sums within 2e-6 of the largest sum;
    # finalised image (sum / weight; the one-sided Lanczos weights sum to ~0 at some pixels, which
    # amplifies the last-bit differences there): per-pixel relative RMSE < 1e-3 overall and < 1e-5 over
    # the pixels whose weight sum is at least 1 % of the median.
    assert np.abs(gf - rf).max() <= 2e-6 * np.abs(rf).max(), np.abs(gf - rf).max() / np.abs(rf).max()
    assert rmse < 1e-3, f"relative RMSE {rmse}"
    ok = np.abs(rf[..., 3]) >= 0.01 * np.median(np.abs(rf[..., 3]))
    assert ok.mean() > 0.85        # rows 64..71 are never rendered (spawn_tiles quirk A-15) and carry ~0 weight
    rmse_ok = float(np.sqrt(np.mean((g[ok] - r[ok]) ** 2)) / np.mean(r[ok]))
    assert rmse_ok < 1e-5, f"relative RMSE over well-conditioned pixels {rmse_ok}"
    sc.close(); osc.close()


def test_per_sample_radiance_bit_exact(ctx, cornell_small):
    """calculate_lighting's result for every camera sample: bit-identical to the oracle (all
    traversals, BxDF sampling, NEE/MIS, Russian roulette).  Allowance: 1e-4 of the samples, for
    the ~1e-8-per-call cases where f64->f32 rounding of a transcendental differs between libm
    and CUDA."""
    hs, cam, film, smp, prm = cornell_small
    d = hs.desc()
    sc = ctx.upload(d)
    osc = O.OracleScene(d)
    gf, grad, st = sc.render_pt_samples(cam, film, smp, prm)
    rf, orad = osc.render_pt_samples(cam, film, smp, prm)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
    assert same.all(), f"{(~same).sum()} of {same.size} samples differ"
    assert (orad[..., :3].max(-1) > 0).mean() > 0.3, "most samples should carry radiance"
    sc.close(); osc.close()


def test_per_sample_radiance_depth_sweep(ctx):
    """max_depth 1, 2, 3 and 5 (min_depth = max_depth/2 moves the Russian-roulette onset)."""
    hs, cam, film, smp, _ = scenes.cornell_scene(48, 36, 2, 1)
    d = hs.desc()
    sc = ctx.upload(d)
    osc = O.OracleScene(d)
    for depth in (1, 2, 3, 5):
        prm = api.make_pt_params(max_depth=depth)
        _, grad, _ = sc.render_pt_samples(cam, film, smp, prm)
        _, orad = osc.render_pt_samples(cam, film, smp, prm)
        same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
        assert same.all(), f"depth {depth}: {(~same).sum()} of {same.size} samples differ"
    sc.close(); osc.close()


def test_render_tile_partition_sums_to_full_frame(ctx, cornell_small):
    """Multi-GPU partitioning on one device: the films of rank 0 and rank 1 of a 2-rank split add up to
    the 1-rank film (Film::merge_into semantics), and each rank matches the oracle's same partition."""
    hs, cam, film, smp, prm = cornell_small
    d = hs.desc()
    sc = ctx.upload(d)
    full, _ = sc.render_pt(cam, film, smp, prm)
    parts = []
    for rank in range(2):
        p = api.make_pt_params(max_depth=prm.max_depth, rank=rank, world_size=2)
        f, st = sc.render_pt(cam, film, smp, p)
        parts.append(f)
    assert np.allclose(parts[0] + parts[1], full, rtol=1e-4, atol=1e-5)
    osc = O.OracleScene(d)
    p1 = api.make_pt_params(max_depth=prm.max_depth, rank=1, world_size=2)
    rf, _, _ = osc.render_pt(cam, film, smp, p1)
    assert np.allclose(parts[1][..., 3], rf[..., 3], rtol=2e-5, atol=1e-6)
    sc.close(); osc.close()


def test_wide_equals_binary(ctx, cornell_small):
    """The product kernels walk the 4-wide collapse of the BVH; the counted kernels walk the reference's binary
    nodes (bvh.rs:97-128).  Same leaves in the same order => identical hit records, bit for bit, on rays
    that exercise ties, spheres and deep subtrees; and an identical film through the whole bounce loop."""
    import torch
    hs, cam, film, smp, prm = cornell_small
    sc = ctx.upload(hs.desc())
    rng = np.random.default_rng(11)
    n = 200_000
    rays = np.zeros(n, api.RAY_DTYPE)
    rays["o"] = rng.uniform([-1.8, -1.3, 2.2], [1.8, 2.2, 5.8], (n, 3)).astype(np.float32)
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    v[: n // 8, rng.integers(0, 3)] = 0.0                       # axis-parallel components: inv = +-inf
    rays["d"] = v.astype(np.float32)
    rays["tmax"] = np.where(rng.random(n) < 0.3, rng.uniform(0.1, 4.0, n), np.inf).astype(np.float32)
    dr = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
    h_wide = torch.empty(n * api.HIT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    h_bin = torch.empty_like(h_wide)
    ctx.set_option(L.ARN_OPT_BVH_WIDTH, 4)                      # auto would pick the binary walk for a tree this small
    sc.intersect_closest_dev(dr.data_ptr(), n, h_wide.data_ptr())
    ctr = sc.intersect_closest_counted_dev(dr.data_ptr(), n, h_bin.data_ptr())
    ctx.synchronize()
    assert ctr[0] > n
    assert torch.equal(h_wide, h_bin)
    assert (h_bin.cpu().numpy().view(api.HIT_DTYPE)["prim_id"] >= 0).sum() > n // 4
    # whole bounce loop: the counted render walks the binary nodes
    f_wide, rad_wide, st = sc.render_pt_samples(cam, film, smp, prm)
    ctx.set_option(L.ARN_OPT_COUNT_TRAVERSAL, 1)
    f_bin, rad_bin, stc = sc.render_pt_samples(cam, film, smp, prm)
    ctx.set_option(L.ARN_OPT_COUNT_TRAVERSAL, 0)
    assert stc.extend_nodes > 0 and st.extend_nodes == 0
    assert np.array_equal(rad_wide, rad_bin)                    # every camera sample, bit for bit
    assert np.allclose(f_wide, f_bin, rtol=1e-5, atol=1e-6)     # film: float atomics, order differs run to run
    # forced binary (uncounted) == forced wide, any-hit included
    a4 = sc.intersect_any(rays)
    ctx.set_option(L.ARN_OPT_BVH_WIDTH, 2)
    a2 = sc.intersect_any(rays)
    h2 = sc.intersect_closest(rays)
    ctx.set_option(L.ARN_OPT_BVH_WIDTH, 0)
    assert np.array_equal(a2, a4)
    assert h2.tobytes() == h_bin.cpu().numpy().tobytes()
    sc.close()


def _light_rig():
    # inside the Cornell room: a point light near the ceiling, a spot aimed at the glass box, a distant "sun"
    return [api.point_light((0.3, 1.6, 3.0), (2.0, 1.5, 1.0)),
            api.spot_light((-0.8, 1.9, 4.5), (0.4, -1.0, -0.6), (9.0, 9.0, 7.0), 0.7, 0.35),
            api.distant_light((0.8, 0.8, 1.0), (0.2, -1.0, -0.4), 6.0)]


def test_analytic_lights_bit_exact(ctx):
    """Point / Spot / Distant lights (lighting/pointlights.rs, distantlight.rs) next to the two emissive
    spheres: delta lights skip MIS, the distant light keeps the weight and traces its (fruitless) specular
    light rays.  Every camera sample bit-identical to the oracle; the same rays are traced."""
    hs, cam, film, smp, prm = scenes.cornell_scene(96, 72, 2, 2, lights=_light_rig())
    d = hs.desc()
    assert d.n_lights == 5
    sc = ctx.upload(d)
    osc = O.OracleScene(d)
    gf, grad, st = sc.render_pt_samples(cam, film, smp, prm)
    rf, orad = osc.render_pt_samples(cam, film, smp, prm)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
    assert same.all(), f"{(~same).sum()} of {same.size} samples differ"
    _, ost, _ = osc.render_pt(cam, film, smp, prm)
    assert (st.extend_rays, st.shadow_rays, st.mis_rays) == (ost.extend_rays, ost.shadow_rays, ost.mis_rays)
    # the rig changes the picture: compare with the sphere-lit render
    hs0, *_ = scenes.cornell_scene(96, 72, 2, 2)
    sc0 = ctx.upload(hs0.desc())
    _, grad0, _ = sc0.render_pt_samples(cam, film, smp, prm)
    assert np.abs(grad[..., :3] - grad0[..., :3]).mean() > 1e-3
    sc.close(); sc0.close(); osc.close()


def test_analytic_lights_only(ctx):
    """A scene without any emissive primitive (and without spheres at all): only a spot and a point light."""
    hs = api.HostScene()
    hs.add_light(api.spot_light((0, 3, 0), (0.1, -1, 0.05), (20, 18, 15), 0.6, 0.3))
    hs.add_light(api.point_light((2, 1, -1), (3, 3, 4)))
    matte = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.6, 0.7)))
    plastic = hs.add_material(api.material(L.ARN_MAT_PLASTIC, kd=(0.4, 0.3, 0.2), ks=(0.5, 0.5, 0.5), roughness=0.2))
    hs.add_mesh(np.float32([[-10, 0, -10], [10, 0, -10], [10, 0, 10], [-10, 0, 10]]), np.uint32([0, 2, 1, 0, 3, 2]), matte)
    hs.add_mesh(np.float32([[-1, 0.5, -1], [1, 0.5, -1], [1, 1.2, 1], [-1, 1.2, 1]]), np.uint32([0, 2, 1, 0, 3, 2]), plastic)
    hs.build()
    view_parent = np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, -1, 0, 0], [0, 5, 0, 1]], np.float32)
    parent_view = np.linalg.inv(view_parent.T).T.astype(np.float32)
    cam = api.make_camera(parent_view.reshape(-1), (-1, -1, 1, 1), 0.1, 100.0, 1.0, 64, 64)
    film, smp, prm = api.make_film(64, 64), api.make_sampler(2, 2, 8, 3), api.make_pt_params(max_depth=4)
    d = hs.desc()
    assert d.n_spheres == 0 and d.n_lights == 2
    sc = ctx.upload(d)
    osc = O.OracleScene(d)
    _, grad, st = sc.render_pt_samples(cam, film, smp, prm)
    _, orad = osc.render_pt_samples(cam, film, smp, prm)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
    assert same.all(), f"{(~same).sum()} of {same.size} samples differ"
    assert st.mis_rays == 0 and st.shadow_rays > 0 and (orad[..., :3].max(-1) > 0).mean() > 0.3
    sc.close(); osc.close()


def _material_zoo(lens=None):
    """A room of quads, one per material kind the reference has (material/{matte,plastic,glass,translucent}.rs),
    lit by a transformed partial sphere and a point light, seen through a pinhole or a thin lens."""
    hs = api.HostScene()
    hs.add_light(api.point_light((0.0, 2.5, 0.0), (3.0, 3.0, 3.0)))
    mats = [
        api.material(L.ARN_MAT_MATTE, kd=(0.6, 0.5, 0.4)),                                             # Lambert
        api.material(L.ARN_MAT_MATTE, kd=(0.4, 0.5, 0.6), sigma=25.0),                                 # Oren-Nayar
        api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.5, 0.5), sigma=200.0),                                # sigma clamped to 90
        api.material(L.ARN_MAT_PLASTIC, kd=(0.3, 0.4, 0.2), ks=(0.6, 0.6, 0.6), roughness=0.05),
        api.material(L.ARN_MAT_PLASTIC, kd=(0.5, 0.2, 0.2), ks=(0.3, 0.3, 0.3), roughness=0.6),
        api.material(L.ARN_MAT_GLASS, kd=(0.0, 0.0, 0.0), ks=(1.0, 1.0, 1.0), roughness=0.1, eta=1.5),   # Fresnel lobe only
        api.material(L.ARN_MAT_GLASS, kd=(0.7, 0.7, 0.7), ks=(0.0, 0.0, 0.0), roughness=0.3, eta=1.33),  # microfacet R + T only
        api.material(L.ARN_MAT_GLASS, kd=(0.6, 0.7, 0.8), ks=(0.9, 0.9, 0.9), roughness=0.02, eta=2.4),
        api.material(L.ARN_MAT_TRANSLUCENT, kd=(0.5, 0.6, 0.3), ks=(0.4, 0.4, 0.4), roughness=0.25, dissolve=0.6),
        api.material(L.ARN_MAT_TRANSLUCENT, kd=(0.5, 0.5, 0.5), ks=(0.2, 0.2, 0.2), roughness=0.5, dissolve=0.0),   # transmission only
        api.material(L.ARN_MAT_MATTE, kd=(0.0, 0.0, 0.0)),                                             # no lobes at all
    ]
    ids = [hs.add_material(m) for m in mats]
    # floor and back wall
    hs.add_mesh(np.float32([[-6, 0, -6], [6, 0, -6], [6, 0, 6], [-6, 0, 6]]), np.uint32([0, 2, 1, 0, 3, 2]), ids[0])
    hs.add_mesh(np.float32([[-6, 0, -4], [6, 0, -4], [6, 6, -4], [-6, 6, -4]]), np.uint32([0, 1, 2, 0, 2, 3]), ids[1])
    # tilted panels, one per remaining material, in two rows
    rng = np.random.default_rng(5)
    for k, mid in enumerate(ids[2:]):
        cx, cy, cz = -4.0 + 2.0 * (k % 5), 0.8 + 1.4 * (k // 5), -1.5 + 0.8 * (k // 5)
        t = rng.uniform(-0.4, 0.4, 3)
        quad = np.float32([[cx - 0.8, cy - 0.6 + t[0], cz], [cx + 0.8, cy - 0.6 + t[1], cz + t[2]], [cx + 0.8, cy + 0.6, cz + 0.3], [cx - 0.8, cy + 0.6, cz + 0.3 + t[0]]])
        hs.add_mesh(quad, np.uint32([0, 1, 2, 0, 2, 3]), mid)
    tr = np.float32([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [1.0, 4.0, 1.0, 1]])
    hs.add_sphere(0.7, -0.5, 0.7, 5.0, ids[1], emission=(14.0, 13.0, 11.0), transform=tr)
    hs.build()
    view_parent = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, -1, 0], [0, 2.0, 7.0, 1]], np.float32)    # looking along -z
    parent_view = np.linalg.inv(view_parent.T).T.astype(np.float32)
    cam = api.make_camera(parent_view.reshape(-1), (-1.2, -0.9, 1.2, 0.9), 0.1, 100.0, 0.9, 80, 60, lens=lens)
    return hs, cam, api.make_film(80, 60), api.make_sampler(2, 2, 8, 11), api.make_pt_params(max_depth=6)


@pytest.mark.parametrize("lens", [None, (0.15, 7.5)])
def test_material_zoo_bit_exact(ctx, lens):
    """Every material kind and lobe combination (Lambert, Oren-Nayar, Ashikhmin-Shirley/Beckmann plastic, Fresnel
    specular, Torrance-Sparrow reflection + transmission, translucent with and without its glossy lobe, lobe-less
    surfaces), thin-lens camera (perspective.rs:300-311): per-sample radiance bit-identical to the oracle."""
    hs, cam, film, smp, prm = _material_zoo(lens)
    d = hs.desc()
    sc = ctx.upload(d)
    osc = O.OracleScene(d)
    _, grad, st = sc.render_pt_samples(cam, film, smp, prm)
    _, orad = osc.render_pt_samples(cam, film, smp, prm)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
    assert same.all(), f"{(~same).sum()} of {same.size} samples differ"
    _, ost, _ = osc.render_pt(cam, film, smp, prm)
    assert (st.extend_rays, st.shadow_rays, st.mis_rays, st.invalid_samples) == (ost.extend_rays, ost.shadow_rays, ost.mis_rays, ost.invalid_samples)
    assert (orad[..., :3].max(-1) > 0).mean() > 0.2
    sc.close(); osc.close()


def test_c2_full_size_bit_exact(ctx):
    """BASELINE C2 at its full size: 1 002 528-triangle height field (1 537 561 BVH nodes), 1920x1080 pixel-centre
    rays.  Every (prim id, t) equals the oracle's bit for bit, through the 4-wide walk (what `auto` picks for a tree
    this size) and through the binary walk; any-hit == closest-hit existence; clipping tmax just below the hit
    distance turns every hit into a miss (size-independent properties)."""
    import torch
    hs, cam, film = scenes.c2_heightfield_scene()
    d = hs.desc()
    assert d.n_triangles == 1_002_528
    w, h = film.res_x, film.res_y
    xs, ys = np.meshgrid(np.arange(w) + 0.5, np.arange(h) + 0.5, indexing="xy")
    pf = np.zeros((w * h, 4), np.float32); pf[:, 0], pf[:, 1] = xs.reshape(-1), ys.reshape(-1)
    rays = O.camera_rays(cam, pf)
    sc = ctx.upload(d); osc = O.OracleScene(d)
    oh = osc.intersect_closest(rays)
    assert (oh["prim_id"] >= 0).sum() > 250_000          # the field covers 14 % of the 90-degree frame
    for width in (0, 4, 2):
        ctx.set_option(L.ARN_OPT_BVH_WIDTH, width)
        _assert_hits_equal(sc.intersect_closest(rays), oh)
    ctx.set_option(L.ARN_OPT_BVH_WIDTH, 0)
    assert np.array_equal(sc.intersect_any(rays) != 0, oh["prim_id"] >= 0)
    hit = oh["prim_id"] >= 0
    clipped = rays[hit].copy(); clipped["tmax"] = np.nextafter(oh["t"][hit], np.float32(0))
    # the accepted t must be < tmax (strict), so a ray that ends just before its hit sees nothing nearer
    gh = sc.intersect_closest(clipped)
    assert np.all((gh["prim_id"] < 0) | (gh["t"] < clipped["tmax"]))
    assert (gh["prim_id"] < 0).mean() > 0.99
    sc.close(); osc.close()


def test_c4_full_size(ctx):
    """BASELINE C4 at its full size: 20 000 172 triangles (closed box of six displaced height fields) + the two
    emissive spheres, 30 M-node reference SAH tree -> the 4-wide walk.  Incoherent rays from inside the box and the
    depth-8 bounce loop on a small frame agree with the oracle bit for bit; the device-built LBVH over the same
    components returns the same distances (ids up to exact-t ties)."""
    hs, cam, film, smp, prm = scenes.c4_box_scene(res=1024, sampledx=1, sampledy=1)
    d = hs.desc()
    assert d.n_triangles == 20_000_172
    sc = ctx.upload(d); osc = O.OracleScene(d)
    rng = np.random.default_rng(21)
    n = 200_000
    rays = np.zeros(n, api.RAY_DTYPE)
    rays["o"] = rng.uniform(-3.5, 3.5, (n, 3)).astype(np.float32)
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    rays["d"] = v.astype(np.float32); rays["tmax"] = np.inf
    gh, oh = sc.intersect_closest(rays), osc.intersect_closest(rays)
    assert (oh["prim_id"] >= 0).mean() > 0.99            # the box is closed
    _assert_hits_equal(gh, oh)
    assert np.array_equal(sc.intersect_any(rays) != 0, oh["prim_id"] >= 0)
    # bounce loop, depth 8, on a 96 x 96 frame of the same view
    cam = api.make_camera(np.float32([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 3.5, 1]]), (-1.0, -1.0, 1.0, 1.0), 0.1, 1000.0, 1.2707964, 96, 96)
    crop = api.make_film(96, 96)
    _, grad, st = sc.render_pt_samples(cam, crop, smp, prm)
    _, orad = osc.render_pt_samples(cam, crop, smp, prm)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
    assert same.all(), f"{(~same).sum()} of {same.size} samples differ"
    assert (orad[..., :3].max(-1) > 0).mean() > 0.5
    sc.close(); osc.close()
    # the same components under the device-built tree
    d2, ms = hs.build_gpu(ctx)
    sc2 = ctx.upload(d2)
    g2 = sc2.intersect_closest(rays)
    assert g2["t"].tobytes() == oh["t"].tobytes()
    assert (g2["prim_id"] != oh["prim_id"]).mean() < 1e-3
    sc2.close()


def test_sample_ranges_waves_tile_grids_and_roulette(ctx):
    """The knobs bench.py and the multi-GPU partition rely on: sample sub-ranges add up to the whole render, a wave
    capacity smaller than one pixel row splits pixels' samples across waves without changing any sample, other tile
    grids / rank counts partition the frame, and non-default Russian-roulette constants follow the oracle."""
    hs, cam, film, smp, _ = scenes.cornell_scene(48, 36, 3, 3)
    d = hs.desc()
    sc = ctx.upload(d); osc = O.OracleScene(d)
    prm = api.make_pt_params(max_depth=5)
    full, rad_full, st_full = sc.render_pt_samples(cam, film, smp, prm)
    # (1) sample ranges
    acc = np.zeros_like(full); rays = 0
    for b, e in ((0, 3), (3, 4), (4, 9)):
        p = api.make_pt_params(max_depth=5, spp_begin=b, spp_end=e)
        f, r, st = sc.render_pt_samples(cam, film, smp, p)
        assert np.array_equal(r, rad_full[:, :, b:e])
        acc += f; rays += st.extend_rays + st.shadow_rays + st.mis_rays
    assert rays == st_full.extend_rays + st_full.shadow_rays + st_full.mis_rays
    assert np.allclose(acc, full, rtol=1e-5, atol=1e-6)
    # (2) tiny waves: 1024 samples per wave = 113.8 pixels of 9 samples -> pixels straddle waves
    ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, 1024)
    f_small, rad_small, st_small = sc.render_pt_samples(cam, film, smp, prm)
    ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, 0)
    assert np.array_equal(rad_small, rad_full) and np.allclose(f_small, full, rtol=1e-5, atol=1e-6)
    assert st_small.kernel_launches > 10 * st_full.kernel_launches
    # (2b) 1, 2 and 8 concurrent wave pipelines over those tiny waves: same samples, same film up to addition order;
    # per-kernel timings only exist with serial launches
    for pipes in (1, 2, 8):
        ctx.set_option(L.ARN_OPT_PIPELINES, pipes)
        ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, 2048)
        f_p, rad_p, st_p = sc.render_pt_samples(cam, film, smp, prm)
        assert np.array_equal(rad_p, rad_full) and np.allclose(f_p, full, rtol=1e-5, atol=1e-6), pipes
        assert (st_p.extend_rays, st_p.shadow_rays, st_p.mis_rays) == (st_full.extend_rays, st_full.shadow_rays, st_full.mis_rays)
        assert (st_p.extend_ms > 0) == (pipes == 1)
    ctx.set_option(L.ARN_OPT_WAVE_CAPACITY, 0)
    ctx.set_option(L.ARN_OPT_PIPELINES, 0)
    with pytest.raises(api.ArnError):
        ctx.set_option(L.ARN_OPT_PIPELINES, 9)
    # (3) 5 x 4 tile grid over 3 ranks, against the oracle's same partition
    total = np.zeros_like(full)
    for rank in range(3):
        p = api.make_pt_params(max_depth=5, tiles=(5, 4), rank=rank, world_size=3)
        gf, st = sc.render_pt(cam, film, smp, p)
        rf, ost, _ = osc.render_pt(cam, film, smp, p)
        assert st.camera_rays == ost.camera_rays > 0
        assert np.abs(gf - rf).max() <= 2e-6 * np.abs(rf).max()
        total += gf
    p = api.make_pt_params(max_depth=5, tiles=(5, 4))
    whole, _ = sc.render_pt(cam, film, smp, p)
    assert np.allclose(total, whole, rtol=1e-5, atol=1e-6)
    # (4) roulette from the first bounce with a high threshold
    p = api.make_pt_params(max_depth=7, min_depth=1, rr_threshold=0.6)
    _, grad, st = sc.render_pt_samples(cam, film, smp, p)
    _, orad = osc.render_pt_samples(cam, film, smp, p)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
    assert same.all()
    _, ost, _ = osc.render_pt(cam, film, smp, p)
    assert (st.extend_rays, st.shadow_rays, st.mis_rays) == (ost.extend_rays, ost.shadow_rays, ost.mis_rays)
    sc.close(); osc.close()


def test_upload_and_render_argument_errors(ctx, cornell_small):
    """Error convention of the C-ABI: malformed scenes and arguments come back as negative codes
    with a message, never as a crash or a silently wrong render."""
    import ctypes as C
    hs, cam, film, smp, prm = cornell_small
    d = hs.desc()

    def clone():
        c = L.SceneDesc()
        C.memmove(C.byref(c), C.byref(d), C.sizeof(L.SceneDesc))
        return c

    def upload_fails(desc, code):
        with pytest.raises(api.ArnError) as e:
            ctx.upload(desc)
        assert e.value.code == code, (e.value.code, str(e.value))
        assert len(str(e.value)) > 10

    nodes = hs.nodes().copy()
    # child index past the end of the node array
    bad = nodes.copy(); bad[0, 6] = d.n_nodes + 5
    c1 = clone(); c1.nodes = bad.ctypes.data_as(C.POINTER(L.Node)); upload_fails(c1, L.ARN_E_INVALID)
    # a cycle: the root's second child is the root's first child's ancestor
    bad = nodes.copy()
    inner = [i for i in range(1, d.n_nodes) if (bad[i, 7] >> 2) == 0][0]
    bad[inner, 6] = 0x7fffffff
    c2 = clone(); c2.nodes = bad.ctypes.data_as(C.POINTER(L.Node)); upload_fails(c2, L.ARN_E_INVALID)
    # leaf range past the component list
    leaf = [i for i in range(d.n_nodes) if (bad[i, 7] >> 2) != 0][0]
    bad = nodes.copy(); bad[leaf, 6] = d.n_prims
    c3 = clone(); c3.nodes = bad.ctypes.data_as(C.POINTER(L.Node)); upload_fails(c3, L.ARN_E_INVALID)
    # triangle index out of range
    idx = np.ctypeslib.as_array(d.indices, (d.n_triangles * 3,)).copy(); idx[7] = d.n_vertices
    c4 = clone(); c4.indices = idx.ctypes.data_as(C.POINTER(C.c_uint32)); upload_fails(c4, L.ARN_E_INVALID)
    # a triangle as a light: unsupported (and useless in the reference: surface_area() == 0)
    lp = np.ctypeslib.as_array(d.light_prims, (d.n_lights,)).copy(); lp[0] = 0
    c5 = clone(); c5.light_prims = lp.ctypes.data_as(C.POINTER(C.c_uint32)); upload_fails(c5, L.ARN_E_UNSUPPORTED)
    # analytic light reference without a table
    lp = np.ctypeslib.as_array(d.light_prims, (d.n_lights,)).copy(); lp[0] = L.ARN_LIGHT_ANALYTIC | 3
    c6 = clone(); c6.light_prims = lp.ctypes.data_as(C.POINTER(C.c_uint32)); upload_fails(c6, L.ARN_E_INVALID)
    # a tree deeper than the traversal stack: a left-leaning chain of 70 interior nodes over 71 coincident triangles
    n = 71
    chain = np.zeros((2 * n - 1, 8), np.uint32); cf = chain.view(np.float32)
    cf[:, 0:3] = (0, 0, 3); cf[:, 3:6] = (1, 1, 3)
    # pre-order: interior i has first child i+1; subtree(i+1) has 2*(n-2-i)+1 nodes; second child = i + 1 + that
    for i in range(n - 1):
        sub_next = 2 * (n - 2 - i) + 1
        chain[i, 6] = sub_next + 1; chain[i, 7] = 0
    chain[n - 1, 6] = 0; chain[n - 1, 7] = (1 << 2) | 3                         # the deepest leaf: slot 0
    for j in range(n - 1):                                                      # the right-hand leaves, innermost first
        chain[n + j, 6] = 1 + j; chain[n + j, 7] = (1 << 2) | 3
    deep = api.HostScene()
    mat = deep.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.5, 0.5)))
    tri = np.float32([[0, 0, 3], [1, 0, 3], [0, 1, 3]])
    for _ in range(n):
        deep.add_mesh(tri, np.uint32([0, 1, 2]), mat)
    deep.add_light(api.point_light((0, 0, 0), (1, 1, 1)))
    dd = deep.build()
    c7 = L.SceneDesc(); C.memmove(C.byref(c7), C.byref(dd), C.sizeof(L.SceneDesc))
    order = np.arange(n, dtype=np.uint32)
    c7.nodes = chain.ctypes.data_as(C.POINTER(L.Node)); c7.n_nodes = 2 * n - 1; c7.order = order.ctypes.data_as(C.POINTER(C.c_uint32))
    upload_fails(c7, L.ARN_E_UNSUPPORTED)
    # render arguments
    sc = ctx.upload(d)
    for bad_prm, code in ((api.make_pt_params(max_depth=0), L.ARN_E_INVALID), (api.make_pt_params(max_depth=8, rank=2, world_size=2), L.ARN_E_INVALID),
                          (api.make_pt_params(max_depth=8, spp_begin=3, spp_end=2), L.ARN_E_INVALID)):
        with pytest.raises(api.ArnError) as e:
            sc.render_pt(cam, film, smp, bad_prm)
        assert e.value.code == code
    with pytest.raises(api.ArnError):
        sc.render_pt(cam, api.make_film(8, 8), smp, prm)                        # 8 x 8 window / 16 x 16 tiles: spawn_tiles divides by zero
    with pytest.raises(api.ArnError):
        sc.render_pt(cam, api.make_film(64, 48, filter_kind=9), smp, prm)
    # and the scene still renders afterwards
    f, st = sc.render_pt(cam, film, smp, prm)
    assert st.camera_rays > 0 and np.isfinite(f).all()
    sc.close()


def test_full_cornell_render_matches_reference_png(ctx):
    """The reference's only result artefact, cornellbox.png (README.md:10: cb.json at 1024 x 768, 32 x 32 = 1024 spp),
    against the product's render of the same configuration — same resolution, same sample count, same 8-bit
    finalisation — through the committed 64 x 48 box-filtered fixture.  The reference's RNG was unseeded, so the
    comparison is statistical, but at 1024 spp the noise of a 16 x 16 block mean is far below the tolerances."""
    import os
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cornellbox_ref_64x48.npy")).astype(np.float64)
    hs, cam, film, smp, prm = scenes.cornell_scene(1024, 768, 32, 32)
    sc = ctx.upload(hs.desc())
    f, st = sc.render_pt(cam, film, smp, prm)
    assert st.invalid_samples < 100 and st.camera_rays == 1024 * 768 * 1024     # invalid radiance becomes a black sample (pt.rs:152-156)
    _, rgb8 = api.film_finalize(f)
    mine = rgb8.astype(np.float64).reshape(48, 16, 64, 16, 3).mean(axis=(1, 3))
    # whole-image mean colour per channel (8-bit units)
    rel = np.abs(mine.mean((0, 1)) - ref.mean((0, 1))) / ref.mean((0, 1))
    assert np.all(rel < 0.002), rel                                       # measured: 0.008 %, 0.03 %, 0.03 %
    # every 16 x 16-pixel block against the reference block (floor of 50/255 for the dark ones)
    err = np.abs(mine - ref) / np.maximum(ref, 50.0)                       # measured: median 0.3 %, p99 1.8 %, max 3.5 %
    assert np.quantile(err, 0.99) < 0.025 and err.max() < 0.05, (np.quantile(err, 0.99), err.max())
    assert np.abs(mine - ref).mean() < 0.5                                 # 8-bit levels; measured 0.23 (max 2.4)
    sc.close()


def test_sphere_only_scene_bit_exact(ctx):
    """No triangle at all: 60 spheres — full, z-clipped, phi-clipped, bare and under rotation / translation / non-uniform
    scale — of every material kind, two of them emissive, plus a spot light (shape/sphere.rs:133-317,
    component/transformed.rs:70-158, component/shape.rs:74-168).  Hits and every camera sample equal the oracle's."""
    rng = np.random.default_rng(17)
    hs = api.HostScene()
    hs.add_light(api.spot_light((0, 6, 0), (0, -1, 0.1), (30, 30, 30), 1.0, 0.6))
    mats = [hs.add_material(m) for m in (
        api.material(L.ARN_MAT_MATTE, kd=(0.7, 0.6, 0.5)), api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.6, 0.7), sigma=20.0),
        api.material(L.ARN_MAT_PLASTIC, kd=(0.4, 0.3, 0.2), ks=(0.5, 0.5, 0.5), roughness=0.1),
        api.material(L.ARN_MAT_GLASS, kd=(0.6, 0.6, 0.6), ks=(0.9, 0.9, 0.9), roughness=0.05, eta=1.5),
        api.material(L.ARN_MAT_TRANSLUCENT, kd=(0.5, 0.5, 0.4), ks=(0.3, 0.3, 0.3), roughness=0.3, dissolve=0.5))]

    def rot(axis, a):
        c, s_ = math.cos(a), math.sin(a)
        m = np.eye(4, dtype=np.float64)
        i, j = [(1, 2), (0, 2), (0, 1)][axis]
        m[i, i], m[i, j], m[j, i], m[j, j] = c, -s_, s_, c
        return m

    for k in range(60):
        radius = float(rng.uniform(0.3, 0.9))
        zmin = -radius if k % 3 else float(rng.uniform(-0.8, -0.1)) * radius
        zmax = radius if k % 4 else float(rng.uniform(0.2, 0.9)) * radius
        phimax = 2 * math.pi if k % 5 else float(rng.uniform(1.0, 5.5))
        m = np.eye(4)
        m[:3, 3] = rng.uniform([-4, 0.5, -4], [4, 4, 4])
        if k % 2:
            m = m @ rot(int(rng.integers(0, 3)), float(rng.uniform(0, 6.28))) @ np.diag([*rng.uniform(0.6, 1.5, 3), 1.0])
        tr = None if k % 7 == 0 else m.T.astype(np.float32)              # rows = columns of the matrix; every 7th sphere is bare (at the origin)
        em = (9.0, 8.0, 7.0) if k in (3, 11) else None
        hs.add_sphere(radius if tr is not None else 0.25, zmin if tr is not None else -0.25, zmax if tr is not None else 0.25, phimax, mats[k % 5], emission=em, transform=tr)
    # a big ground sphere
    g = np.eye(4); g[1, 3] = -200.0
    hs.add_sphere(200.0, -200.0, 200.0, 2 * math.pi, mats[0], transform=g.T.astype(np.float32))
    hs.build()
    d = hs.desc()
    assert d.n_triangles == 0 and d.n_spheres == 61
    view_parent = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, -1, 0], [0, 2.0, 10.0, 1]], np.float32)
    cam = api.make_camera(np.linalg.inv(view_parent.T).T.astype(np.float32).reshape(-1), (-1.2, -0.9, 1.2, 0.9), 0.1, 1000.0, 1.0, 96, 72)
    film, smp, prm = api.make_film(96, 72), api.make_sampler(2, 2, 8, 5), api.make_pt_params(max_depth=6)
    sc = ctx.upload(d); osc = O.OracleScene(d)
    n = 100_000
    rays = np.zeros(n, api.RAY_DTYPE)
    rays["o"] = rng.uniform([-5, 0.2, -5], [5, 5, 8], (n, 3)).astype(np.float32)
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    rays["d"] = v.astype(np.float32); rays["tmax"] = np.inf
    gh, oh = sc.intersect_closest(rays), osc.intersect_closest(rays)
    assert (oh["prim_id"] >= 0).mean() > 0.4
    _assert_hits_equal(gh, oh)
    _, grad, st = sc.render_pt_samples(cam, film, smp, prm)
    _, orad = osc.render_pt_samples(cam, film, smp, prm)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
    assert same.all(), f"{(~same).sum()} of {same.size} samples differ"
    assert (orad[..., :3].max(-1) > 0).mean() > 0.05          # a dark scene: one spot cone and two small emitters
    sc.close(); osc.close()


def test_transformed_instance_of_an_emitter(ctx, tmp_path):
    """arencli's ComponentDesc::Transformed (arencli.rs:162-181): the instance of an emissive sphere glows when a path
    hits it directly but is not one of `Scene.lights` (never sampled, and a BSDF-sampled ray that reaches it is not
    `ptr::eq` to the sampled light).  Scene read from a JSON file; GPU == oracle per sample."""
    import json, os, shutil
    mini = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mini_scene")
    src = json.load(open(os.path.join(mini, "scene.json")))
    for f in ("room.obj", "room.mtl"):
        shutil.copy(os.path.join(mini, f), tmp_path / f)
    src["components"] += [
        {"name": "bulb", "value": {"Shaped": {"shape": {"Sphere": {"radius": 0.3, "zmin": -1.0, "zmax": 1.0, "phimax": 6.28}},
                                              "material": {"name": "lamp_matte", "value": None},
                                              "light": {"name": "bulb_light", "value": {"Constant": {"value": {"inner": [4.0, 4.0, 4.0]}}}}, "transform": None}}},
        {"name": "bulb2", "value": {"Transformed": {"transform": [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [-0.7, 1.0, 0.5, 1]], "original": "bulb"}}}]
    p = tmp_path / "scene.json"; p.write_text(json.dumps(src))
    hs = api.HostScene()
    cam, film, smp, prm, _ = hs.load_json(p, base_dir=tmp_path)
    d = hs.build()
    assert d.n_spheres == 3 and d.n_lights == 2
    sc = ctx.upload(d); osc = O.OracleScene(d)
    _, grad, st = sc.render_pt_samples(cam, film, smp, prm)
    _, orad = osc.render_pt_samples(cam, film, smp, prm)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
    assert same.all(), f"{(~same).sum()} of {same.size} samples differ"
    _, ost, _ = osc.render_pt(cam, film, smp, prm)
    assert (st.extend_rays, st.shadow_rays, st.mis_rays) == (ost.extend_rays, ost.shadow_rays, ost.mis_rays)
    sc.close(); osc.close()


def test_gpu_matches_committed_golden_vectors(ctx):
    """The CUDA path against the committed fixture tests/golden/oracle_cornell_64x48.npz (oracle output for the Cornell
    fixture at 64 x 48, 4 spp; generator committed next to it): hits, per-sample radiance, ray counts, film."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_cornell_64x48.npz"))
    hs, cam, film, smp, prm = scenes.cornell_scene(64, 48, 2, 2)
    sc = ctx.upload(hs.desc())
    f, rad, st = sc.render_pt_samples(cam, film, smp, prm)
    same = np.all(rad.view(np.uint32) == gold["radiance"].view(np.uint32), axis=-1)
    assert (~same).sum() <= 3, f"{(~same).sum()} samples differ from the golden vectors"
    assert np.abs(np.array([st.camera_rays, st.extend_rays, st.shadow_rays, st.mis_rays, st.invalid_samples]) - gold["rays"]).max() <= 8
    assert np.abs(f - gold["film"]).max() <= 2e-6 * np.abs(gold["film"]).max()
    xs, ys = np.meshgrid(np.arange(0, 64, 2) + 0.5, np.arange(0, 48, 2) + 0.5, indexing="xy")
    pf = np.zeros((xs.size, 4), np.float32); pf[:, 0], pf[:, 1] = xs.reshape(-1), ys.reshape(-1)
    hits = sc.intersect_closest(O.camera_rays(cam, pf))
    assert np.array_equal(hits["prim_id"], gold["prim_id"]) and hits["t"].tobytes() == gold["t"].tobytes()
    sc.close()
