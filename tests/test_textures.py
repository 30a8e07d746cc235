"""SURVEY.md §8(f) N4 — image textures, ray differentials, bump mapping.
CPU part: the oracle's restatement (oracle/texture.hpp) against answers derived by hand from the cited lines of
texturing/textures/image.rs, geometry/interaction.rs (the reference holds no test or fixture for textures: parity unpinned).
GPU part: the textured shade instance against the oracle, per camera sample, bit for bit."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_lib as O
from arendur_b200 import api, _lib as L


def _tex(levels, trilinear=True, wrapping=L.ARN_WRAP_REPEAT, max_aniso=8.0, scaling=(1.0, 1.0), shifting=(0.0, 0.0)):
    t = L.Texture()
    lv = [np.ascontiguousarray(a, np.float32) for a in levels]
    t.channels = 3 if lv[0].ndim == 3 else 1
    t.n_levels, t.trilinear, t.wrapping, t.max_aniso = len(lv), int(trilinear), wrapping, max_aniso
    t.scale_u, t.scale_v = scaling; t.shift_u, t.shift_v = shifting
    off = 0
    for i, a in enumerate(lv):
        t.level_h[i], t.level_w[i], t.level_offset[i] = a.shape[0], a.shape[1], off
        off += a.size
    return t, np.ascontiguousarray(np.concatenate([a.reshape(-1) for a in lv]), np.float32)


def _lookup(t, texels, uv, dxy=(0, 0, 0, 0)):
    lib = O.load()
    lib.arn_oracle_texture_lookup.argtypes = [C.POINTER(L.Texture), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    uvf, d, out = np.float32(uv), np.float32(dxy), np.zeros(3, np.float32)
    lib.arn_oracle_texture_lookup(C.byref(t), O._p(texels), O._p(uvf), O._p(d), O._p(out))
    return out


def test_triangle_filter_known_answers():
    """triangle_filter (image.rs:427-445): s = u * nx - 0.5; texel centres reproduce the texel, mid-points the mean; Repeat wraps
    on both sides, Black returns zeros outside, Clamp sends an index below zero to the FAR edge (`as usize` of a negative float)."""
    rng = np.random.default_rng(1)
    img = rng.random((4, 4, 3)).astype(np.float32)
    t, tx = _tex([img])
    for j in range(4):
        for i in range(4):
            assert np.array_equal(_lookup(t, tx, ((i + 0.5) / 4, (j + 0.5) / 4)), img[j, i])
    got = _lookup(t, tx, (1.0 / 4, 0.5 / 4))                 # s = 0.5: between texels 0 and 1 of row 0
    want = img[0, 0] * np.float32(0.5) + img[0, 1] * np.float32(0.5)      # (1-ds)(1-dt) = 0.5, ds(1-dt) = 0.5, dt = 0
    assert np.allclose(got, want, rtol=0, atol=1e-7)
    # u = 0.05: s = -0.3 -> floor -1: Repeat takes texel 3 with weight 0.3 and texel 0 with weight 0.7
    got = _lookup(t, tx, (0.05, 0.5 / 4))
    want = img[0, 3] * np.float32(1 - (np.float32(-0.3) - np.float32(-1.0))) + img[0, 0] * (np.float32(-0.3) - np.float32(-1.0))
    assert np.allclose(got, want, atol=2e-7)
    tb, _ = _tex([img], wrapping=L.ARN_WRAP_BLACK)
    assert np.allclose(_lookup(tb, tx, (0.05, 0.5 / 4)), img[0, 0] * np.float32(0.7), atol=2e-7)      # the texel at -1 is black
    tc, _ = _tex([img], wrapping=L.ARN_WRAP_CLAMP)
    assert np.allclose(_lookup(tc, tx, (0.05, 0.5 / 4)), img[0, 3] * np.float32(0.3) + img[0, 0] * np.float32(0.7), atol=2e-7)   # sic: -1 clamps to dx - 1


def test_find_level_is_levels_minus_one_times_log2_width():
    """find_level (image.rs:522-527) = (levels - 1) * log2(max(width, 1e-8)): below one texture width the level is negative and
    look_up_tri ends in triangle_filter(0); width 2 on a 3-level pyramid gives level 2 = the last; sqrt(2) gives exactly level 1."""
    levels = [np.full((4, 4, 3), 0.1, np.float32), np.full((2, 2, 3), 0.5, np.float32), np.full((1, 1, 3), 0.9, np.float32)]
    t, tx = _tex(levels)
    assert np.allclose(_lookup(t, tx, (0.3, 0.3), (0.5, 0.0, 0.0, 0.0)), 0.1)
    assert np.allclose(_lookup(t, tx, (0.3, 0.3), (2.0, 0.0, 0.0, 0.0)), 0.9)
    assert np.allclose(_lookup(t, tx, (0.3, 0.3), (math.sqrt(2.0), 0.0, 0.0, 0.0)), 0.5, atol=1e-6)
    w = 2.0 ** 0.75                                              # level 1.5: halfway between levels 1 and 2
    assert np.allclose(_lookup(t, tx, (0.3, 0.3), (w, 0.0, 0.0, 0.0)), 0.7, atol=1e-5)
    assert np.allclose(_lookup(t, tx, (0.3, 0.3), (0.0, 0.0, 0.0, 0.0)), 0.1)      # zero width -> 1e-8 -> far below zero


def test_ewa_filter_properties():
    """EWA (image.rs:447-519): a constant texture filters to the constant; a zero minor axis falls back to triangle_filter(0);
    the level is clamped to >= 0 so footprints below one texture width blend ewa(0) * 1 + ewa(1) * 0."""
    const = [np.full((8, 8, 3), 0.37, np.float32), np.full((4, 4, 3), 0.37, np.float32), np.full((2, 2, 3), 0.37, np.float32), np.full((1, 1, 3), 0.37, np.float32)]
    t, tx = _tex(const, trilinear=False)
    for d in ((0.05, 0.01, 0.0, 0.04), (0.2, 0.0, 0.0, 0.01), (0.01, 0.3, 0.25, 0.02)):
        assert np.allclose(_lookup(t, tx, (0.4, 0.6), d), 0.37, atol=1e-6)
    rng = np.random.default_rng(2)
    img = rng.random((8, 8, 3)).astype(np.float32)
    lv = [img, np.zeros((4, 4, 3), np.float32), np.zeros((2, 2, 3), np.float32), np.zeros((1, 1, 3), np.float32)]
    t, tx = _tex(lv, trilinear=False)
    tt, _ = _tex(lv, trilinear=True)
    assert np.array_equal(_lookup(t, tx, (0.3, 0.7), (0.0, 0.0, 0.0, 0.0)), _lookup(tt, tx, (0.3, 0.7)))      # minor == 0 -> triangle_filter(0)
    a = _lookup(t, tx, (0.3, 0.7), (0.03, 0.0, 0.0, 0.03))        # isotropic footprint of 0.24 texels: level clamped to 0, level 1 (zeros) weighs 0
    assert a.min() > 0.0 and np.all(a <= img.max())
    # the Gaussian weights are symmetric: a footprint centred on a texel of a left-right symmetric row keeps that symmetry
    sym = np.tile(np.float32([0.1, 0.4, 0.9, 0.9, 0.4, 0.1, 0.0, 0.0])[None, :, None], (8, 1, 3))
    ts, txs = _tex([sym, np.zeros((4, 4, 3), np.float32)], trilinear=False)
    l = _lookup(ts, txs, (2.5 / 8, 0.5), (0.2, 0.0, 0.0, 0.2)); r = _lookup(ts, txs, (3.5 / 8, 0.5), (0.2, 0.0, 0.0, 0.2))
    assert np.allclose(l, r, atol=1e-6)


def test_uv_mapping_scales_and_shifts():
    """UVMapping::map (mappings.rs:21-30): st = uv * scaling + shifting, the differentials scale with it."""
    rng = np.random.default_rng(3)
    img = rng.random((4, 4, 3)).astype(np.float32)
    t, tx = _tex([img], scaling=(2.0, 1.0), shifting=(0.25, 0.0))
    # uv = (0.0625, 0.125): st = (0.375, 0.125) = centre of texel (1, 0)
    assert np.array_equal(_lookup(t, tx, (0.0625, 0.125)), img[0, 1])


def test_compute_dxy_on_a_plane():
    """compute_dxy (interaction.rs:204-224): offset rays hit the plane z = 0; dpdx / dpdy are the hit-point offsets; (dudx, dudy)
    solve the 2x2 system the source builds from the two largest-normal-free coordinates (sic: dudy comes from the SAME solve)."""
    lib = O.load()
    lib.arn_oracle_compute_dxy.argtypes = [C.c_void_p] * 6
    pos, n = np.float32([0.5, 0.25, 0.0]), np.float32([0, 0, 1])
    dpdu, dpdv = np.float32([2, 0, 0]), np.float32([0, 4, 0])
    rays = np.float32([0.6, 0.25, 1.0, 0, 0, -1,   0.5, 0.45, 2.0, 0, 0, -1])     # rx lands at x + 0.1, ry at y + 0.2
    out = np.zeros(10, np.float32)
    lib.arn_oracle_compute_dxy(O._p(pos), O._p(n), O._p(dpdu), O._p(dpdv), O._p(rays), O._p(out))
    assert np.allclose(out[0:3], [0.1, 0, 0], atol=1e-6) and np.allclose(out[3:6], [0, 0.2, 0], atol=1e-6)
    # |n.z| largest: Matrix2::new(dpdu.x, dpdv.x, dpdu.y, dpdv.y) = columns (2, 0), (0, 4); inverse * (0.1, 0) = (0.05, 0)
    assert np.allclose(out[6:10], [0.05, 0.0, 0.0, 0.05], atol=1e-6)               # dudx, dvdx (= dvdxy.x = 0), dudy (= dudxy.y = 0), dvdy (= 0.2 / 4)


# ---------------------------------------------------------------- GPU parity
def _checker(n, a, b, cells=4):
    y, x = np.mgrid[0:n, 0:n]
    m = (((x * cells) // n + (y * cells) // n) % 2).astype(np.float32)[..., None]
    return (np.float32(a) * m + np.float32(b) * (1 - m)).astype(np.float32)


def _pyramid(img):
    """A box-filtered pyramid (the reference resizes with the `image` crate's Lanczos3; any pyramid is valid input)."""
    out = [np.ascontiguousarray(img, np.float32)]
    while out[-1].shape[0] > 1:
        a = out[-1]
        out.append(((a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2]) * np.float32(0.25)).astype(np.float32))
    return out


def textured_scene(res=(96, 72), spp=(2, 2), trilinear=True, wrapping=L.ARN_WRAP_REPEAT, bump=True):
    hs = api.HostScene()
    rng = np.random.default_rng(5)
    kd_img = _checker(32, (0.8, 0.2, 0.1), (0.1, 0.3, 0.8)) * (0.6 + 0.4 * rng.random((32, 32, 1)).astype(np.float32))
    t_kd = hs.add_texture(_pyramid(kd_img), trilinear=trilinear, wrapping=wrapping, scaling=(3.0, 2.0), shifting=(0.1, 0.0))
    t_ks = hs.add_texture(_pyramid(_checker(16, (0.9, 0.9, 0.9), (0.2, 0.2, 0.2), 8)), trilinear=trilinear, wrapping=wrapping)
    yy, xx = np.mgrid[0:64, 0:64].astype(np.float32)
    height = (0.02 * np.sin(xx * 0.7) * np.cos(yy * 0.5)).astype(np.float32)
    t_bump = hs.add_texture(_pyramid(height), trilinear=True, wrapping=wrapping, scaling=(4.0, 4.0))
    t_rough = hs.add_texture(_pyramid((0.05 + 0.4 * rng.random((8, 8))).astype(np.float32)), trilinear=trilinear, wrapping=wrapping)
    t_sig = hs.add_texture(_pyramid((20.0 * rng.random((8, 8))).astype(np.float32)), trilinear=trilinear, wrapping=wrapping)
    floor = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.5, 0.5), kd_tex=t_kd, bump_tex=t_bump if bump else 0))
    wall = hs.add_material(api.material(L.ARN_MAT_PLASTIC, kd=(0.4, 0.4, 0.4), ks=(0.5, 0.5, 0.5), roughness=0.2, ks_tex=t_ks, aux_tex=t_rough))
    ball = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.6, 0.6, 0.6), sigma=0.0, kd_tex=t_kd, aux_tex=t_sig, bump_tex=t_bump if bump else 0))
    glass = hs.add_material(api.material(L.ARN_MAT_GLASS, kd=(0.7, 0.7, 0.7), ks=(0.9, 0.9, 0.9), roughness=0.1, eta=1.5, kd_tex=t_kd))
    lightm = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.5, 0.5)))
    quad = np.uint32([0, 1, 2, 0, 2, 3])
    uv = np.float32([[0, 0], [1, 0], [1, 1], [0, 1]])
    up = np.float32([[0, 1, 0]] * 4)
    hs.add_mesh(np.float32([[-3, -1, 8], [3, -1, 8], [3, -1, 2], [-3, -1, 2]]), quad, floor, normals=up, uvs=uv)             # floor, with normals
    hs.add_mesh(np.float32([[-3, -1, 8], [3, -1, 8], [3, 3, 8], [-3, 3, 8]]), quad, wall, uvs=uv * np.float32([2, 1]))         # back wall, uvs only
    hs.add_mesh(np.float32([[-3, -1, 2], [-3, -1, 8], [-3, 3, 8], [-3, 3, 2]]), quad, glass, uvs=uv)                           # left wall: glass
    hs.add_mesh(np.float32([[3, -1, 2], [3, 3, 2], [3, 3, 8], [3, -1, 8]]), quad, floor)                                        # right wall: default uvs (0,0),(1,0),(1,1)
    tr = np.eye(4, dtype=np.float32); tr[3, 0:3] = (0.8, -0.2, 5.0)
    hs.add_sphere(0.8, -0.8, 0.8, 6.28, ball, transform=tr)
    tl = np.eye(4, dtype=np.float32); tl[3, 0:3] = (-0.5, 2.2, 4.0)
    hs.add_sphere(0.4, -0.4, 0.4, 6.28, lightm, emission=(18.0, 16.0, 12.0), transform=tl)
    hs.build()
    cam = api.make_camera(api.IDENTITY, (-1.0, -0.75, 1.0, 0.75), 0.1, 100.0, 1.2, res[0], res[1])
    return hs, cam, api.make_film(res[0], res[1]), api.make_sampler(spp[0], spp[1], 8, 7), api.make_pt_params(max_depth=5)


def test_oracle_textured_render_is_finite_and_texture_dependent():
    hs, cam, film, smp, prm = textured_scene(res=(48, 36), spp=(1, 1))
    osc = O.OracleScene(hs.desc())
    f, st, _ = osc.render_pt(cam, film, smp, prm)
    hs2, *_ = textured_scene(res=(48, 36), spp=(1, 1), bump=False)
    f2, st2, _ = O.OracleScene(hs2.desc()).render_pt(cam, film, smp, prm)
    assert np.isfinite(f).all() and f[..., :3].max() > 0
    assert not np.array_equal(f, f2), "bump mapping must change the picture"
    assert hs.desc().n_textures == 5


@pytest.mark.gpu
@pytest.mark.parametrize("trilinear,wrapping", [(True, L.ARN_WRAP_REPEAT), (False, L.ARN_WRAP_REPEAT), (True, L.ARN_WRAP_CLAMP), (False, L.ARN_WRAP_BLACK)])
def test_textured_scene_per_sample_radiance_bit_exact(ctx, trilinear, wrapping):
    """Image textures (trilinear and EWA, every wrap mode) on kd / ks / sigma / roughness, bump mapping on a mesh with normals and
    on a transformed sphere, ray differentials through 5 bounces incl. glass: every camera sample equals the oracle's."""
    hs, cam, film, smp, prm = textured_scene(trilinear=trilinear, wrapping=wrapping)
    d = hs.desc()
    sc = ctx.upload(d); osc = O.OracleScene(d)
    gf, grad, st = sc.render_pt_samples(cam, film, smp, prm)
    of, orad = osc.render_pt_samples(cam, film, smp, prm)
    _, ost, _ = osc.render_pt(cam, film, smp, prm)
    assert (st.extend_rays, st.shadow_rays, st.mis_rays) == (ost.extend_rays, ost.shadow_rays, ost.mis_rays)
    bad = np.argwhere(np.any(grad[..., :3] != orad[..., :3], axis=-1))
    assert bad.shape[0] == 0, f"{bad.shape[0]} of {grad.shape[0] * grad.shape[1] * grad.shape[2]} samples differ, first {bad[:5].tolist()}: gpu {grad[tuple(bad[0])]} oracle {orad[tuple(bad[0])]}"
    assert np.allclose(gf, of, rtol=2e-5, atol=2e-5)
    sc.close(); osc.close()


@pytest.mark.gpu
def test_texture_upload_validation(ctx):
    hs, cam, film, smp, prm = textured_scene(res=(32, 24), spp=(1, 1))
    d = hs.desc()
    import copy
    mats = (L.Material * d.n_materials)()
    C.memmove(mats, d.materials, C.sizeof(mats))
    mats[0].kd_tex = 99
    d2 = L.SceneDesc(); C.memmove(C.byref(d2), C.byref(d), C.sizeof(d)); d2.materials = mats
    with pytest.raises(api.ArnError) as e:
        ctx.upload(d2)
    assert e.value.code == L.ARN_E_INVALID
    mats[0].kd_tex = 3                      # a Luma texture where an RGB one is needed
    with pytest.raises(api.ArnError):
        ctx.upload(d2)
