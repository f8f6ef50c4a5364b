"""Pins the CPU oracle against everything the reference's own tests and fixtures hold for this path
(SURVEY.md §4, §8c, App. B).  CPU only."""
import json
import math
import os

import numpy as np
import pytest

import pyoracle as O

L = O.lib()
HERE = os.path.dirname(os.path.abspath(__file__))


def fromhex(s):
    return float.fromhex(s)


# ---- Low_discrepancy_sequence (low_discrepancy_sequence.ml:8-36) ------------------------------------
def test_lds_phi_alpha_known_answers():
    # SURVEY App. B, derived from the reference formulas with glibc pow
    assert L.orc_lds_phi(18) == fromhex("0x1.09c6b0a6941cdp+0")
    a = np.zeros(18)
    L.orc_lds_alpha(18, O.dptr(a))
    want = ["0x1.ed2abc0801724p-1", "0x1.db06cfac89271p-1", "0x1.c98db4fa98eb7p-1", "0x1.b8b9236c54c30p-1"]
    assert [x.hex() for x in a[:4]] == [fromhex(w).hex() for w in want]
    got = [L.orc_lds_get(O.dptr(a), 0, d) for d in range(4)]
    assert got == [0.46321666333890166, 0.42778634053372677, 0.3936592632203064, 0.36078749368096474]
    assert L.orc_lds_get(O.dptr(a), 179999, 0) == fromhex("0x1.ff62f9f200000p-2")
    assert L.orc_lds_get(O.dptr(a), 8294399 + 1023 * 1024, 17) == fromhex("0x1.e3427ac000000p-4")
    assert L.orc_lds_phi(34) == fromhex("0x1.05321cbbd543ep+0")
    b = np.zeros(34)
    L.orc_lds_alpha(34, O.dptr(b))
    assert b[0] == fromhex("0x1.f5d0b17313401p-1")
    assert L.orc_lds_get(O.dptr(b), 0, 0) == 0.4801078274700785
    assert L.orc_lds_phi(1) == fromhex("0x1.9e3779b97f4a8p+0")  # golden ratio
    c = np.zeros(1)
    L.orc_lds_alpha(1, O.dptr(c))
    assert L.orc_lds_get(O.dptr(c), 0, 0) == 0.1180339887498949
    assert L.orc_lds_phi(2) == 1.324717957244746  # plastic number
    d = np.zeros(2)
    L.orc_lds_alpha(2, O.dptr(d))
    assert [L.orc_lds_get(O.dptr(d), 0, k) for k in (0, 1)] == [0.2548776662466927, 0.06984029099805333]


def _integrate_1d(f, lower, upper, iterations):
    # low_discrepancy_sequence_test.ml:7-21 (Kahan-summed QMC estimate)
    a = np.zeros(1)
    L.orc_lds_alpha(1, O.dptr(a))
    s = c = 0.0
    for i in range(iterations):
        x = (upper - lower) * L.orc_lds_get(O.dptr(a), i, 0) + lower
        y = f(x) - c
        t = s + y
        c = t - s - y
        s = t
    return (upper - lower) / iterations * s


def test_lds_reference_1d_integrals():
    # low_discrepancy_sequence_test.ml:40-56, same iteration counts and tolerances
    assert abs(_integrate_1d(math.sin, 0.0, math.pi, 1000) - 2.0) < 1e-3
    assert abs(_integrate_1d(math.sin, -1.0, 1.0, 5000) - 0.0) < 1e-3
    assert abs(_integrate_1d(lambda x: math.sqrt(1.0 - x * x), 0.0, 1.0, 2000) - math.pi / 4) < 1e-3
    assert abs(_integrate_1d(math.exp, 0.0, 3.0, 2000) - math.expm1(3.0)) < 0.03


# ---- Filter_kernel.Binomial (filter_kernel.ml:49-85) --------------------------------------------------
def test_filter_binomial_weights():
    w = np.zeros(9)
    L.orc_filter_binomial(5, 1, O.dptr(w))
    w1 = [fromhex("0x1.d555555555556p-3"), fromhex("0x1.1555555555556p-1"), fromhex("0x1.d555555555556p-3")]
    assert np.array_equal(w.reshape(3, 3), np.outer(w1, w1))
    assert w[0] == 0.05251736111111112 and w[1] == 0.12413194444444448 and w[4] == 0.29340277777777785
    assert abs(w.sum() - 1.0) < 1e-15
    # energy kept by an edge pixel = 37/48
    assert abs((w.reshape(3, 3)[1:, :].sum()) - 37.0 / 48.0) < 1e-15


# ---- Tile (tile.ml:28-39; path_tracer_test.ml:34-70) ---------------------------------------------------
def _split(w, h, max_area):
    cap = 20000
    a = [np.zeros(cap, dtype=np.int32) for _ in range(4)]
    n = L.orc_tile_split(w, h, max_area, *[O.iptr(x) for x in a], cap)
    return [tuple(int(x[i]) for x in a) for i in range(n)]  # (row, col, w, h)


def test_tile_split_reference_test():
    tiles = _split(10, 5, 7)
    assert all(w * h <= 7 for _, _, w, h in tiles)
    pts = set()
    for r, c, w, h in tiles:
        for y in range(h):
            for x in range(w):
                assert (c + x, r + y) not in pts
                pts.add((c + x, r + y))
    assert pts == {(x, y) for x in range(10) for y in range(5)}


def test_tile_split_known_counts():
    from collections import Counter
    assert Counter((w, h) for _, _, w, h in _split(600, 300, 1024)) == {(37, 19): 96, (19, 37): 64, (38, 19): 64,
                                                                        (37, 18): 32}
    assert Counter((w, h) for _, _, w, h in _split(1024, 1024, 1024)) == {(32, 32): 1024}
    assert Counter((w, h) for _, _, w, h in _split(1920, 1080, 1024)) == {(30, 34): 1536, (30, 33): 512}
    assert Counter((w, h) for _, _, w, h in _split(3840, 2160, 1024)) == {(30, 34): 6144, (30, 33): 2048}


# ---- Film_tile (film_tile.ml:15-61; path_tracer_test.ml:72-119) ---------------------------------------
def test_film_tile_write_pixel_locus():
    w, h = 7, 8
    out = np.zeros((h + 2, w + 2, 3))
    L.orc_film_tile_write_pixel(w, h, 0, 0, O.dptr(np.ones(3)), O.dptr(out))
    nz = out > 0
    assert nz[0:3, 0:3].all()  # global (col-1..col+1, row-1..row+1)
    nz[0:3, 0:3] = False
    assert not nz.any()
    assert abs(out.sum() - 3.0) < 1e-14


# ---- Bbox.is_hit (bbox.ml:40-56; path_tracer_test.ml:121-130) -----------------------------------------
def test_bbox_reference_tests():
    mn, mx, o = np.zeros(3), np.ones(3), np.array([-5.0, 0.5, 0.5])
    hit, miss = np.array([1.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0])
    assert L.orc_bbox_is_hit(O.dptr(mn), O.dptr(mx), O.dptr(o), O.dptr(hit), 0.0, 5.01) == 1
    assert L.orc_bbox_is_hit(O.dptr(mn), O.dptr(mx), O.dptr(o), O.dptr(hit), 0.0, 4.99) == 0
    assert L.orc_bbox_is_hit(O.dptr(mn), O.dptr(mx), O.dptr(o), O.dptr(miss), 0.0, 1000.0) == 0


# ---- Shader_space (path_tracer_test.ml:132-142) ---------------------------------------------------------
def test_unit_square_to_hemisphere_is_normalized():
    rng = np.random.default_rng(0)
    out = np.zeros(3)
    for _ in range(101):
        u, v = rng.random(2)
        L.orc_unit_square_to_hemisphere(u, v, O.dptr(out))
        assert abs(out @ out - 1.0) < 1e-6 and out[2] >= 0


def test_shader_space_frame_maps_normal_to_z_and_inverts():
    rng = np.random.default_rng(1)
    out, back = np.zeros(3), np.zeros(3)
    for _ in range(200):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        L.orc_shader_space_rotate(O.dptr(n), O.dptr(n), 0, O.dptr(out))
        assert np.allclose(out, [0, 0, 1], atol=1e-12)
        v = rng.normal(size=3)
        L.orc_shader_space_rotate(O.dptr(n), O.dptr(v), 0, O.dptr(out))
        L.orc_shader_space_rotate(O.dptr(n), O.dptr(out), 1, O.dptr(back))
        assert np.allclose(back, v, atol=1e-12)
    for n in ([0.0, 0.0, 1.0], [0.0, 0.0, -1.0]):  # the two special cases of shader_space.ml:15-18
        n = np.array(n)
        L.orc_shader_space_rotate(O.dptr(n), O.dptr(n), 0, O.dptr(out))
        assert np.allclose(out, [0, 0, 1], atol=1e-15)


# ---- intersectors: analytic cases (unpinned by reference tests) -----------------------
def test_sphere_intersect_analytic_and_leaf_agreement():
    t = np.zeros(1)
    c, o, d = np.zeros(3), np.array([-5.0, 0.0, 0.0]), np.array([1.0, 0.0, 0.0])
    assert L.orc_sphere_intersect_scalar(O.dptr(c), 1.0, O.dptr(o), O.dptr(d), 0.0, 10.0, O.dptr(t)) == 1
    assert t[0] == 4.0  # bench/intersect_bench.ml's hit ray, analytic answer
    assert L.orc_sphere_intersect_scalar(O.dptr(c), 1.0, O.dptr(o), O.dptr(np.array([0.0, 1.0, 0.0])), 0.0, 10.0,
                                         O.dptr(t)) == 0
    o_in = np.array([0.5, 0.0, 0.0])  # inside, heading to the centre side: far root
    assert L.orc_sphere_intersect_scalar(O.dptr(c), 1.0, O.dptr(o_in), O.dptr(-d), 0.0, 10.0, O.dptr(t)) == 1
    assert t[0] == 1.5
    # scalar OCaml formula vs the Rust AVX kernel (emulated AND intrinsics, checked equal inside) agree to
    # rounding and pick the same nearest sphere
    rng = np.random.default_rng(2)
    for _ in range(500):
        n = int(rng.integers(1, 17))
        pad = (-n) % 4
        cs = rng.uniform(-5, 5, size=(n, 3))
        rs = rng.uniform(0.1, 1.5, size=n)
        xs, ys, zs, rr = (np.concatenate([a, np.full(pad, np.nan)]) for a in (cs[:, 0], cs[:, 1], cs[:, 2], rs))
        o = rng.uniform(-8, 8, size=3)
        dd = rng.normal(size=3)
        dd /= np.linalg.norm(dd)
        ts = np.zeros(1)
        idx = L.orc_spheres_intersect_simd(O.dptr(xs), O.dptr(ys), O.dptr(zs), O.dptr(rr), n + pad, O.dptr(o),
                                           O.dptr(dd), 0.0, 1e300, O.dptr(ts))
        assert idx != -2, "AVX intrinsics and lane-by-lane emulation disagree"
        best, bt = -1, 1e300
        for i in range(n):
            if L.orc_sphere_intersect_scalar(O.dptr(cs[i].copy()), rs[i], O.dptr(o), O.dptr(dd), 0.0, bt,
                                             O.dptr(t)):
                best, bt = i, t[0]
        assert idx == best
        if best >= 0:
            assert abs(ts[0] - bt) <= 1e-12 * max(1.0, abs(bt))


def test_triangle_intersect_analytic():
    a, b, c = np.array([0.0, 0, 0]), np.array([1.0, 0, 0]), np.array([0.0, 1, 0])
    o, d = np.array([0.25, 0.25, -2.0]), np.array([0.0, 0.0, 1.0])
    t, u, v = np.zeros(1), np.zeros(1), np.zeros(1)
    assert L.orc_triangle_intersect(*[O.dptr(x) for x in (a, b, c, o, d)], 0.0, 1e300, O.dptr(t), O.dptr(u),
                                    O.dptr(v)) == 1
    assert (t[0], u[0], v[0]) == (2.0, 0.25, 0.25)
    # two-sided
    assert L.orc_triangle_intersect(*[O.dptr(x) for x in (a, c, b, o, d)], 0.0, 1e300, O.dptr(t), O.dptr(u),
                                    O.dptr(v)) == 1
    # outside, parallel, behind, beyond t_max
    o2 = np.array([0.8, 0.8, -2.0])
    assert L.orc_triangle_intersect(*[O.dptr(x) for x in (a, b, c, o2, d)], 0.0, 1e300, O.dptr(t), O.dptr(u),
                                    O.dptr(v)) == 0
    dp = np.array([1.0, 0.0, 0.0])
    assert L.orc_triangle_intersect(*[O.dptr(x) for x in (a, b, c, o, dp)], 0.0, 1e300, O.dptr(t), O.dptr(u),
                                    O.dptr(v)) == 0
    assert L.orc_triangle_intersect(*[O.dptr(x) for x in (a, b, c, o, -d)], 0.0, 1e300, O.dptr(t), O.dptr(u),
                                    O.dptr(v)) == 0
    assert L.orc_triangle_intersect(*[O.dptr(x) for x in (a, b, c, o, d)], 0.0, 1.5, O.dptr(t), O.dptr(u),
                                    O.dptr(v)) == 0


# ---- the one golden artefact: shirley-spheres.png (layout-independent facts) ----------------------------
@pytest.fixture(scope="module")
def shirley_c1_oracle():
    import path_tracer_ocaml_b200 as P
    from helpers import make_params
    scene = P.shirley_spheres(600, 300)
    osc = O.OracleScene(scene.tables())
    img, cn = osc.render(make_params(scene, 600, 300, 32, 8), n_threads=os.cpu_count())
    return scene, osc, img, cn


def test_oracle_matches_golden_png_facts(shirley_c1_oracle):
    _, _, img, _ = shirley_c1_oracle
    facts = json.load(open(os.path.join(HERE, "golden", "shirley_png_facts.json")))
    # edge darkening of the 3x3 splat with out-of-image taps dropped (integrator.ml:115-117): sqrt(37/48)
    assert abs(img[0].mean() / img[1].mean() - facts["row0_over_row1"]) < 0.01
    assert abs(img[:, 0].mean() / img[:, 1].mean() - facts["col0_over_col1"]) < 0.01
    assert abs(img[0].mean() / img[1].mean() - math.sqrt(37 / 48)) < 0.01
    # sky rows: camera + background gradient + filter + gamma; the PNG holds trunc(255*v) (+-1 LSB)
    for y, rgb in facts["sky_row_mean_rgb"].items():
        got = np.floor(255.0 * img[int(y), 2:-2]).mean(0)
        assert np.abs(got - np.array(rgb)).max() <= 1.0, (y, got, rgb)
    # global brightness: small-sphere colours differ (unpinned PRNG) but only by a few LSB overall
    assert np.abs(255.0 * img.mean((0, 1)) - np.array(facts["mean_rgb"])).max() < 6.0


def test_oracle_tree_and_counters_are_reference_like(shirley_c1_oracle):
    scene, osc, _, cn = shirley_c1_oracle
    st = osc.tree_stats()
    assert scene.tables()["n_spheres"] == 530  # 4 + 526 kept (Base.Random.float rule, pinned by the golden PNG)
    assert sum(st["leaf_histogram"].values()) * 2 - 1 == st["nodes"]  # binary tree
    assert max(st["leaf_histogram"]) <= 16  # Simd_leaf.length_cutoff (lib.rs:13)
    assert cn.paths == 600 * 300 * 32
    assert cn.rays == sum(cn.rays_by_bounce) and cn.rays_by_bounce[0] == cn.paths
    assert cn.rays == cn.hits + cn.missed
    assert cn.paths == cn.missed + cn.absorbed + cn.exhausted


def test_scalar_and_simd_leaves_render_the_same_image():
    # `--no-simd` (Array_leaf, cutoff 4) vs the default Simd_leaf: same closest hits up to last-ulp rounding
    import path_tracer_ocaml_b200 as P
    from helpers import make_params
    scene = P.shirley_spheres(120, 60)
    p = make_params(scene, 120, 60, 4, 8)
    a, _ = O.OracleScene(scene.tables(), O.ORC_LEAF_SIMD, 16).render(p)
    b, _ = O.OracleScene(scene.tables(), O.ORC_LEAF_ARRAY, 4).render(p)
    assert np.sqrt(np.mean((a - b) ** 2)) < 5e-3
    assert np.mean(np.abs(a - b) < 1e-9) > 0.98
