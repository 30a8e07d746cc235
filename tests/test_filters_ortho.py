"""Film filters of sample/filters.rs other than the deserialised Lanczos default, and OrthoCam
(filming/ortho.rs) — SURVEY.md §8(f) N3.  The reference has no tests for them: the oracle is checked against
values computed by hand from the cited formulas; the GPU path is checked against the oracle."""
import math

import numpy as np
import pytest

import oracle_lib as O
from arendur_b200 import api, scenes, _lib as L

f32 = np.float32


def test_filter_known_answers():
    box = api.make_film(8, 8, filter_radius=(2.0, 3.0), filter_kind=L.ARN_FILTER_BOX)
    assert O.filter_eval(box, 1.7, -2.9) == 1.0                                        # filters.rs:55-57
    tri = api.make_film(8, 8, filter_radius=(2.0, 3.0), filter_kind=L.ARN_FILTER_TRIANGLE)
    assert O.filter_eval(tri, 0.5, -1.0) == f32(1.5) * f32(2.0)                        # (rx - |x|)(ry - |y|), :81-83
    assert O.filter_eval(tri, -2.0, 0.0) == 0.0
    # Gaussian (sic): exp(-a x^2) - (-a r^2) per axis — the stored "exp" is the exponent itself (:104-107,121-125)
    g = api.make_film(8, 8, filter_radius=(2.0, 2.0), filter_kind=L.ARN_FILTER_GAUSSIAN, filter_a=0.5)
    want = (math.exp(-0.5 * 0.25) + 0.5 * 4.0) * (math.exp(-0.5 * 1.0) + 0.5 * 4.0)
    assert abs(O.filter_eval(g, 0.5, -1.0) / want - 1) < 1e-6
    # Mitchell b = c = 1/3.  sic: only the constant term carries the 1/6 (operator precedence, filters.rs:158-168),
    # so this is not the Mitchell-Netravali kernel; the restatement follows the source, term by term
    m = api.make_film(8, 8, filter_radius=(2.0, 2.0), filter_kind=L.ARN_FILTER_MITCHELL, filter_a=1 / 3, filter_b=1 / 3)
    b = c = 1 / 3
    inner = lambda x: (12 - 9 * b - 6 * c) * x ** 3 + (-18 - 12 * b + 6 * c) * x ** 2 + (6 - 2 * b) / 6
    outer = lambda x: (-b - 6 * c) * x ** 3 + (6 * b + 30 * c) * x ** 2 - (12 * b + 48 * c) * x + (8 * b + 24 * c) / 6
    assert abs(O.filter_eval(m, 0.0, 0.0) - (8 / 9) ** 2) < 1e-6                       # inner(0)^2
    assert abs(O.filter_eval(m, 2.0, 0.0) / (outer(2.0) * inner(0.0)) - 1) < 1e-5      # 2 x / r = 2
    assert abs(O.filter_eval(m, -0.5, 1.5) / (inner(0.5) * outer(1.5)) - 1) < 1e-5     # |.| of the scaled offsets
    # Lanczos default tau = 3 and explicit tau; one-sidedness is covered in test_host_logic
    l0 = api.make_film(8, 8)
    l3 = api.make_film(8, 8, filter_a=3.0)
    l2 = api.make_film(8, 8, filter_a=2.0)
    assert O.filter_eval(l0, 0.7, 1.3) == O.filter_eval(l3, 0.7, 1.3) != O.filter_eval(l2, 0.7, 1.3)
    sinc = lambda x: math.sin(math.pi * x) / (math.pi * x)
    assert abs(O.filter_eval(l2, 0.7, 1.3) - sinc(0.35) * sinc(0.7) * sinc(0.65) * sinc(1.3)) < 1e-6


def test_ortho_camera_known_answers():
    """ortho_transform = scale(1,1,1/(f-n)) * translate(0,0,-n) (ortho.rs:57-66); raster (x, y) -> view (sx, sy, n):
    rays start ON the near plane at the screen position and run along +z (ortho.rs:180-183)."""
    ident = np.eye(4, dtype=f32)
    cam = O.ortho_camera_make(ident.reshape(-1), (-2.0, -1.0, 2.0, 1.0), 0.5, 100.0, 40, 20)
    pf = np.zeros((3, 4), f32)
    pf[0, :2] = (20.0, 10.0)          # centre of the raster
    pf[1, :2] = (0.0, 0.0)            # top-left: screen (pmin.x, pmax.y)
    pf[2, :2] = (40.0, 20.0)          # bottom-right: (pmax.x, pmin.y)
    rays = O.camera_rays(cam, pf)
    assert np.allclose(rays["o"], [[0, 0, 0.5], [-2, 1, 0.5], [2, -1, 0.5]], atol=1e-5)
    assert np.array_equal(rays["d"], np.tile(f32([0, 0, 1]), (3, 1)))
    # the product's host constructor builds the same matrices
    pcam = api.make_ortho_camera(ident.reshape(-1), (-2.0, -1.0, 2.0, 1.0), 0.5, 100.0, 40, 20)
    assert bytes(pcam) == bytes(cam)
    # lens (sic): the origin becomes (lens sample, 0) — pview is forgotten (ortho.rs:184-196)
    lcam = O.ortho_camera_make(ident.reshape(-1), (-2.0, -1.0, 2.0, 1.0), 0.5, 100.0, 40, 20, lens=(0.1, 5.0))
    pf[:, 2:] = 0.5                   # centre of the lens
    lr = O.camera_rays(lcam, pf)
    assert np.allclose(lr["o"], 0.0, atol=1e-7)
    assert np.allclose(lr["d"][1], f32([-2, 1, 5.5]) / np.linalg.norm([-2, 1, 5.5]), atol=1e-6)
    with pytest.raises(api.ArnError):
        api.make_ortho_camera(np.zeros(16, f32), (-1, -1, 1, 1), 0.1, 10.0, 8, 8)      # "matrix inversion failure"


def _scene(film, cam_kind):
    hs, cam, _, smp, prm = scenes.cornell_scene(film.res_x, film.res_y, 2, 2)
    if cam_kind == "ortho":
        # into the room (the perspective camera of cb.json looks along +z from the origin), tilted down a little
        c, s_ = math.cos(-0.25), math.sin(-0.25)
        view_parent = np.array([[1, 0, 0, 0], [0, c, s_, 0], [0, -s_, c, 0], [0, 0.2, 1.0, 1]], f32)
        cam = api.make_ortho_camera(view_parent.reshape(-1), (-1.6, -1.2, 1.6, 1.2), 0.1, 100.0, film.res_x, film.res_y)
    return hs, cam, smp, prm


@pytest.mark.gpu
@pytest.mark.parametrize("kind,a,b,radius", [
    (L.ARN_FILTER_BOX, 0.0, 0.0, (0.5, 0.5)), (L.ARN_FILTER_BOX, 0.0, 0.0, (2.0, 1.0)), (L.ARN_FILTER_TRIANGLE, 0.0, 0.0, (2.0, 2.0)),
    (L.ARN_FILTER_GAUSSIAN, 0.5, 0.0, (2.0, 2.0)), (L.ARN_FILTER_MITCHELL, 1 / 3, 1 / 3, (2.0, 2.0)), (L.ARN_FILTER_MITCHELL, 0.0, 0.5, (3.0, 4.0)),
    (L.ARN_FILTER_LANCZOS, 2.0, 0.0, (3.0, 3.0)), (L.ARN_FILTER_GAUSSIAN, 2.0, 0.0, (6.0, 5.0)),      # radius > 4: the scatter kernel
])
def test_film_filters_match_oracle(ctx, kind, a, b, radius):
    film = api.make_film(64, 48, filter_radius=radius, filter_kind=kind, filter_a=a, filter_b=b)
    hs, cam, smp, prm = _scene(film, "perspective")
    sc = ctx.upload(hs.desc()); osc = O.OracleScene(hs.desc())
    gf, st = sc.render_pt(cam, film, smp, prm)
    rf, ost, _ = osc.render_pt(cam, film, smp, prm)
    # same samples (bit-exact radiance, test_per_sample_radiance_bit_exact); the film differs by the order of the
    # float additions and by libdevice expf/sinf vs the oracle's correctly-rounded ones in the weights: 2e-6 of the peak
    assert np.abs(gf - rf).max() <= 2e-6 * np.abs(rf).max(), np.abs(gf - rf).max() / np.abs(rf).max()
    assert np.abs(rf[..., 3]).max() > 0
    sc.close(); osc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lens", [None, (0.05, 6.0)])
def test_ortho_camera_bit_exact(ctx, lens):
    film = api.make_film(64, 48)
    hs, cam, smp, prm = _scene(film, "ortho")
    if lens:
        cam.has_lens, cam.lens_radius, cam.focal_distance = 1, lens[0], lens[1]
    sc = ctx.upload(hs.desc()); osc = O.OracleScene(hs.desc())
    _, grad, st = sc.render_pt_samples(cam, film, smp, prm)
    _, orad = osc.render_pt_samples(cam, film, smp, prm)
    same = np.all(grad.view(np.uint32) == orad.view(np.uint32), axis=-1)
    assert same.all(), f"{(~same).sum()} of {same.size} samples differ"
    assert (orad[..., :3].max(-1) > 0).mean() > 0.1
    sc.close(); osc.close()


def test_crop_window_quirks_oracle():
    """Film::spawn_tiles lays its tiles out from (0, 0), not from crop.pmin, and unwraps the intersection of every grown
    tile with the crop window (film.rs:118-129): a crop window away from the origin panics; a slightly offset one
    renders the pixels [0, w) x [0, h) and keeps what falls inside the window."""
    hs, cam, _, smp, prm = scenes.cornell_scene(64, 48, 1, 1)
    osc = O.OracleScene(hs.desc())
    far = api.make_film(64, 48, crop=(32, 24, 64, 48))
    out = np.zeros((24, 32, 4), f32); st = L.Stats(); trav = np.zeros(3, np.uint64)
    import ctypes as C
    rc = osc.lib.arn_oracle_render_pt(osc.h, C.byref(cam), C.byref(far), C.byref(smp), C.byref(prm), out.ctypes.data_as(C.c_void_p),
                                      C.byref(st), trav.ctypes.data_as(C.c_void_p), 2)
    assert rc == L.ARN_E_INVALID
    near = api.make_film(64, 48, crop=(3, 2, 51, 34))                 # 48 x 32 window: 3 x 2 pixel tiles from (0, 0)
    f, st, _ = osc.render_pt(cam, near, smp, prm)
    assert st.camera_rays == 48 * 32 and f[..., 3].max() > 0
    # samples exist for pixels [0, 48) x [0, 32) only; the window [3, 51) x [2, 34) keeps their splats (radius 4 reaches its far edge)
    assert f[0, 0, 3] != 0 and f[-1, -1, 3] != 0
    small = api.make_film(64, 48, crop=(3, 2, 51, 34), filter_radius=(0.5, 0.5))
    f2, _, _ = osc.render_pt(cam, small, smp, prm)
    assert np.all(f2[-2:, :, 3] == 0) and np.all(f2[:, -3:, 3] == 0) and f2[0, 0, 3] != 0      # pixels 48..50 / 32..33 get no sample
    osc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("radius", [(4.0, 4.0), (2.7, 1.3)])
def test_crop_window_quirks_gpu(ctx, radius):
    hs, cam, _, smp, prm = scenes.cornell_scene(64, 48, 2, 2)
    sc = ctx.upload(hs.desc()); osc = O.OracleScene(hs.desc())
    with pytest.raises(api.ArnError) as e:
        sc.render_pt(cam, api.make_film(64, 48, crop=(32, 24, 64, 48), filter_radius=radius), smp, prm)
    assert e.value.code == L.ARN_E_INVALID
    for crop in ((3, 2, 51, 34), (0, 0, 60, 41), (1, 0, 64, 48)):       # offset window; sizes the 16 x 16 grid does not divide
        film = api.make_film(64, 48, crop=crop, filter_radius=radius)
        gf, st = sc.render_pt(cam, film, smp, prm)
        rf, ost, _ = osc.render_pt(cam, film, smp, prm)
        assert st.camera_rays == ost.camera_rays
        assert np.abs(gf - rf).max() <= 2e-6 * np.abs(rf).max(), (crop, np.abs(gf - rf).max() / np.abs(rf).max())
        assert np.array_equal(gf[..., 3] == 0, rf[..., 3] == 0), crop        # the same film pixels are touched (fractional radii: tile sinks)
    sc.close(); osc.close()
