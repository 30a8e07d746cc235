"""Host-side logic of the product (no GPU): film finalisation, light distribution, PNG output,
the ParitySampler contract, filter and warp helpers of the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
from arendur_b200 import api, _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_film_finalize_matches_oracle_and_reference_semantics():
    rng = np.random.default_rng(0)
    film = rng.uniform(0, 3, (17, 9, 4)).astype(np.float32)
    film[0, 0] = (1, 2, 3, 0)                  # wsum == 0 -> black (film.rs:339-341)
    film[0, 1] = (5, -1, 0.25, 1)              # clamp to [0, 1] then trunc(v * 255)
    film[0, 2] = (0.999, 0.5, 1.0, 1.0)
    g, g8 = api.film_finalize(film)
    o, o8 = O.film_finalize(film)
    assert np.array_equal(g.view(np.uint32), o.view(np.uint32)) and np.array_equal(g8, o8)
    assert g8[0, 0].tolist() == [0, 0, 0] and g8[0, 1].tolist() == [255, 0, 63] and g8[0, 2].tolist() == [254, 127, 255]


def test_light_distribution():
    lib = L.load()
    f = np.float32([994.57465, 685.0824, 0.0, 3.0])
    cdf, integ = np.zeros(5, np.float32), C.c_float()
    assert lib.arn_light_distribution(4, C.cast(api._ptr(f), L.c_float_p), C.cast(api._ptr(cdf), L.c_float_p), C.byref(integ)) == 0
    ocdf, ointeg = np.zeros(5, np.float32), C.c_float()
    O.load().arn_oracle_light_distribution(4, O._p(f), O._p(ocdf), C.byref(ointeg))
    assert np.array_equal(cdf, ocdf) and integ.value == ointeg.value and cdf[-1] == 1.0 and cdf[2] == cdf[3]
    # all-zero power: uniform cdf i/(n+1) (distribution.rs:44-49) and integral 0 (-> pdf 0)
    z = np.zeros(3, np.float32); cdf = np.zeros(4, np.float32)
    assert lib.arn_light_distribution(3, C.cast(api._ptr(z), L.c_float_p), C.cast(api._ptr(cdf), L.c_float_p), C.byref(integ)) == 0
    assert np.allclose(cdf, [0, 0.25, 0.5, 0.75]) and integ.value == 0.0
    neg = np.float32([1, -1])
    assert lib.arn_light_distribution(2, C.cast(api._ptr(neg), L.c_float_p), C.cast(api._ptr(cdf), L.c_float_p), C.byref(integ)) == L.ARN_E_INVALID   # assert!(curfunc >= 0)


def test_png_writer_roundtrip(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(1)
    film = rng.uniform(0, 1.2, (33, 47, 4)).astype(np.float32); film[..., 3] = 1.0
    path = tmp_path / "out.png"
    assert L.load().arn_save_png(str(path).encode(), api._ptr(film), 47, 33) == 0
    img = np.asarray(Image.open(path))
    _, rgb8 = api.film_finalize(film)
    assert img.shape == (33, 47, 3) and np.array_equal(img, rgb8)     # RGB8, rows top to bottom, no gamma (film.rs:380-391)


def test_parity_sampler_contract():
    """Draws are in [0,1), independent of the order in which pixels / samples are visited, and the
    1-D and 2-D streams have independent counters (sample/sink.rs:56-65)."""
    lib = O.load()
    out = np.zeros(8 + 2 * 26, np.float32)
    lib.arn_oracle_sampler_draws(0, 17, 5, 3, 8, 26, O._p(out))
    assert (out >= 0).all() and (out < 1).all()
    again = np.zeros_like(out); lib.arn_oracle_sampler_draws(0, 17, 5, 3, 8, 26, O._p(again))
    assert np.array_equal(out, again)
    other = np.zeros_like(out); lib.arn_oracle_sampler_draws(0, 17, 5, 4, 8, 26, O._p(other))
    assert not np.array_equal(out, other)
    # uniformity over many pixels (first 2-D draw = film jitter)
    vals = []
    for px in range(64):
        for s in range(16):
            o = np.zeros(2, np.float32); lib.arn_oracle_sampler_draws(0, px, 7, s, 0, 1, O._p(o)); vals.append(o.copy())
    v = np.array(vals)
    assert abs(v.mean() - 0.5) < 0.02 and abs(np.corrcoef(v[:, 0], v[:, 1])[0, 1]) < 0.08
    assert np.histogram(v[:, 0], bins=8, range=(0, 1))[0].min() > 80


def test_lanczos_filter_is_one_sided():
    """Quirk A-4: negative offsets get weight factor 1 per axis (sinc(x) = 1 for x < 1e-5)."""
    lz = O.load().arn_oracle_lanczos
    assert lz(-2.5, -0.5) == 1.0
    assert lz(0.0, 0.0) == 1.0
    x = 1.5
    expect = (np.sin(np.pi * x / 3) / (np.pi * x / 3)) * (np.sin(np.pi * x) / (np.pi * x))
    assert abs(lz(x, -1.0) - expect) < 1e-6 and lz(x, -1.0) < 0
    assert abs(lz(3.0, 0.0)) < 1e-6


def test_roughness_to_alpha_shared_by_host_and_oracle():
    hs = api.HostScene()
    for r in (0.968, 0.99, 0.92, 1.0, 1e-9):
        hs.add_material(api.material(L.ARN_MAT_PLASTIC, kd=(0.5, 0.5, 0.5), ks=(0.5, 0.5, 0.5), roughness=r))
    m = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(0.5, 0.5, 0.5)))
    hs.add_sphere(1.0, -1, 1, 6.28, m)
    d = hs.build()
    for i, r in enumerate((0.968, 0.99, 0.92, 1.0, 1e-9)):
        assert d.materials[i].alpha == O.load().arn_oracle_roughness_to_alpha(r)
    assert abs(d.materials[1].alpha - 1.61) < 0.01


def test_host_scene_argument_errors():
    hs = api.HostScene()
    with pytest.raises(api.ArnError):
        hs.build()                                             # no components
    m = hs.add_material(api.material(L.ARN_MAT_MATTE, kd=(1, 1, 1)))
    with pytest.raises(api.ArnError):
        hs.add_mesh(np.zeros((3, 3), np.float32), np.uint32([0, 1, 5]), m)       # index out of range
    with pytest.raises(api.ArnError):
        hs.add_mesh(np.zeros((3, 3), np.float32), np.uint32([0, 1, 2]), 7)       # unknown material
    # trailing indices that do not form a triangle are ignored (TriangleInstance iterator)
    hs.add_mesh(np.float32([[0, 0, 0], [1, 0, 0], [0, 1, 0]]), np.uint32([0, 1, 2, 0, 1]), m)
    assert hs.build().n_triangles == 1


def test_fast_transcendentals_host_sweep(tmp_path):
    """kernels/cr_math.cuh compiled for the host: on every 257th f32 bit pattern (16.7 M arguments per function) the short f64
    kernels either decline or return (float)libm_f64(x); the full 2^32 sweep (stride 1) takes ~2 minutes on 8 cores and is
    recorded in DESIGN.md."""
    import subprocess
    exe = str(tmp_path / "test_cr_math")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_cr_math.cpp")])
    r = subprocess.run([exe, "257"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:]


def test_stratified_sampler_mode_stratifies_every_dimension_it_covers():
    """ARN_SAMPLER_STRATIFIED (include/arn.h; the reference's sampler as intended): over the spp samples of a
    pixel every covered 1-D dimension visits each of the spp strata once, every covered 2-D dimension each cell of the
    sampledx x sampledy grid once; later dimensions and the whole parity mode are the plain hash draws."""
    import ctypes as C
    import numpy as np
    import oracle_lib as O
    from arendur_b200 import _lib as L
    lib = O.load()
    for sx, sy, ndim in ((4, 4, 8), (3, 5, 2), (1, 7, 4), (32, 32, 3)):
        n = sx * sy
        smp = O.make_sampler(sx, sy, ndim, seed=11, mode=L.ARN_SAMPLER_STRATIFIED)
        par = O.make_sampler(sx, sy, ndim, seed=11, mode=L.ARN_SAMPLER_PARITY)
        n1, n2 = ndim + 3, ndim + 3
        for (px, py) in ((0, 0), (17, 5)):
            draws = np.zeros((n, n1 + 2 * n2), np.float32); pdraws = np.zeros_like(draws); old = np.zeros_like(draws)
            for s in range(n):
                lib.arn_oracle_sampler_draws2(C.byref(smp), px, py, s, n1, n2, draws[s].ctypes.data)
                lib.arn_oracle_sampler_draws2(C.byref(par), px, py, s, n1, n2, pdraws[s].ctypes.data)
                lib.arn_oracle_sampler_draws(11, px, py, s, n1, n2, old[s].ctypes.data)
            assert np.array_equal(pdraws, old)                                       # mode 0 is the sampler every parity test pins
            assert ((draws >= 0) & (draws < 1)).all()
            for d in range(ndim):
                strata = np.floor(draws[:, d].astype(np.float64) * n).astype(int)
                assert sorted(strata) == list(range(n)), (sx, sy, d)
                x, y = draws[:, n1 + 2 * d].astype(np.float64), draws[:, n1 + 2 * d + 1].astype(np.float64)
                cells = np.floor(x * sx).astype(int) * sy + np.floor(y * sy).astype(int)
                assert sorted(cells) == list(range(n)), (sx, sy, d)
            assert np.array_equal(draws[:, ndim:n1], pdraws[:, ndim:n1])              # beyond ndim: the raw draws
            assert np.array_equal(draws[:, n1 + 2 * ndim:], pdraws[:, n1 + 2 * ndim:])
            if n > 1:
                assert not np.array_equal(draws[:, :ndim], pdraws[:, :ndim])
        # different pixels and dimensions get different permutations
        a = np.zeros(n1 + 2 * n2, np.float32); b = np.zeros_like(a)
        lib.arn_oracle_sampler_draws2(C.byref(smp), 3, 4, 0, n1, n2, a.ctypes.data); lib.arn_oracle_sampler_draws2(C.byref(smp), 4, 3, 0, n1, n2, b.ctypes.data)
        assert not np.array_equal(a, b)


def test_pair_records_of_small_trees_restate_the_nodes():
    """arn_scene_upload re-encodes a small tree for k_trace's shared-memory walk (kernels/traverse.cuh, traverse2p): one 128-byte
    record per interior node — per axis both children's planes in both orders, the children's reference words twice.  Checked
    against the 32-byte nodes it is built from; a tree beyond the shared-memory budget and a single-leaf tree give no records."""
    from arendur_b200 import scenes
    lib = L.load()
    hs, *_ = scenes.cornell_scene(64, 48, 2, 2)
    nodes = hs.nodes()                                                   # (n, 8) uint32 bit patterns of arn_node
    n = nodes.shape[0]
    cnt = C.c_uint32(0)
    assert lib.arn_selftest_pair_records(nodes.ctypes.data_as(C.c_void_p), n, None, C.byref(cnt)) == 0
    n_int = (n - 1) // 2
    assert cnt.value == n_int and n_int * 128 <= L.ARN_SMEM_NODE_BYTES
    rec = np.zeros((n_int, 32), np.uint32)
    assert lib.arn_selftest_pair_records(nodes.ctypes.data_as(C.c_void_p), n, rec.ctypes.data_as(C.c_void_p), C.byref(cnt)) == 0
    interior = np.nonzero((nodes[:, 7] >> 2) == 0)[0]
    pair_of = {int(i): k for k, i in enumerate(interior)}
    assert len(interior) == n_int and interior[0] == 0                   # the root's record is the first
    for i in interior:
        r = rec[pair_of[int(i)]]
        for slot, ch in enumerate((int(i) + 1, int(i) + int(nodes[i, 6]))):
            bmin, bmax = nodes[ch, 0:3], nodes[ch, 3:6]
            for a in range(3):
                assert r[a * 8 + slot * 2] == bmin[a] and r[a * 8 + slot * 2 + 1] == bmax[a]          # direction >= 0: (min, max)
                assert r[a * 8 + 4 + slot * 2] == bmax[a] and r[a * 8 + 4 + slot * 2 + 1] == bmin[a]  # direction < 0: (max, min)
            leaf = (nodes[ch, 7] >> 2) != 0
            w0 = int(nodes[ch, 6]) if leaf else pair_of[ch] * 128
            for rep in range(2):
                assert r[24 + rep * 4 + slot * 2] == w0 and r[25 + rep * 4 + slot * 2] == nodes[ch, 7]
    # too large for the budget / a single leaf: no records, and the count says so
    big, *_ = scenes.c4_box_scene(cells=14, res=16, sampledx=1, sampledy=1)
    bn = big.nodes()
    assert lib.arn_selftest_pair_records(bn.ctypes.data_as(C.c_void_p), bn.shape[0], None, C.byref(cnt)) == 0 and cnt.value == 0
    assert lib.arn_selftest_pair_records(nodes[interior[-1] + 1:].ctypes.data_as(C.c_void_p), 1, None, C.byref(cnt)) == 0 and cnt.value == 0
