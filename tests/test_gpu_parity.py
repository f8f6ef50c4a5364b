"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against
the CPU oracle on the same inputs.  Tolerances (BASELINE.json north_star: sample stream bit-exact; image
within a stated per-pixel relative tolerance and RMSE bound):
  * R2 stream, sample offsets, cx, cy: BIT-EXACT (uint64 compare of the doubles)
  * float64 device mode vs oracle: identical ray counts per bounce, image |diff| <= 1e-9
  * float32 device mode vs oracle at identical spp/bounces: RMSE <= 0.01 (<= 32 spp), <= 0.003 (>= 256 spp);
    >= 99% of pixel channels within |d| <= 0.02*ref + 1/255; |mean signed error| <= 1e-3 per channel
  * first-hit map: primitive id equal on >= 99.9% of camera rays; |dt|/t median <= 1e-6, 99th pct <= 2e-4
    (float32 evaluates c = |f|^2 - r^2 for the r = 1000 ground sphere with ~3e-5 relative error; the
    shading kernel re-projects the hit point onto the sphere, so this does not reach the image);
    arbitrary rays (intersect_batch): id equal >= 99.9%, median |dt|/t <= 2e-6, 99th pct |dt| <= 1e-3
"""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi, integrator
import pyoracle as O
from helpers import image_metrics, make_params, oracle_alpha, oracle_lds, resolve_numpy

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
NCPU = os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------
# stage 1: sample stream and camera rays — bit exact
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mb,offsets", [
    (8, np.arange(0, 600 * 300 + 32 * 32, dtype=np.int32)),                       # every offset of C1
    (16, np.arange(0, 1024 * 1024 + 256 * 256, 37, dtype=np.int32)),              # C2, strided
    (8, np.concatenate([np.arange(0, 3840 * 2160 + 1024 * 1024, 1009, dtype=np.int32),
                        np.array([0, 1, 8294399, 8294399 + 1023 * 1024], dtype=np.int32)])),  # C4, strided
    (0, np.arange(0, 1000, dtype=np.int32)), (64, np.arange(0, 5000, 13, dtype=np.int32))])
def test_r2_stream_bit_exact(mb, offsets):
    dev = integrator.r2_stream(mb, offsets)
    want = oracle_lds(oracle_alpha(2 + 2 * mb), offsets)
    assert dev.shape == want.shape
    assert np.array_equal(dev.view(np.uint64), want.view(np.uint64))
    # and the vectorised checker itself equals oracle.cpp's lds_get on a sample
    a = oracle_alpha(2 + 2 * mb)
    for i in range(0, len(offsets), max(1, len(offsets) // 50)):
        for d in (0, 2 * mb + 1):
            assert want[i, d] == O.lib().orc_lds_get(O.dptr(a), int(offsets[i]), d)


def test_r2_known_answers_on_device():
    # SURVEY App. B values, straight from the device
    out = integrator.r2_stream(8, np.array([0, 179999, 8294399 + 1023 * 1024], dtype=np.int32))
    assert list(out[0, :4]) == [0.46321666333890166, 0.42778634053372677, 0.3936592632203064, 0.36078749368096474]
    assert out[1, 0] == float.fromhex("0x1.ff62f9f200000p-2")
    assert out[2, 17] == float.fromhex("0x1.e3427ac000000p-4")


@pytest.mark.parametrize("W,H,spp", [(600, 300, 32), (37, 19, 3), (1, 1, 5), (3840, 2160, 1024)])
def test_raygen_bit_exact(W, H, spp):
    scene = P.shirley_spheres(W, H)
    integ = P.Integrator(scene, W, H, spp, 8)
    npix = W * H
    total = npix * spp
    n = min(total, 200000)
    for first in sorted({0, max(0, total - n), (total // 2 // npix) * npix}):
        first = min(first, total - n)
        pixel, offset, cx, cy, d = integ.raygen(first, n)
        k = first + np.arange(n, dtype=np.int64)
        pas = k // npix
        # enumeration: pass-major over the pixel list; offsets follow integrator.ml:98 (pass * spp, sic)
        assert np.array_equal(offset, (pixel + pas * spp).astype(np.int32))
        gx, gy = pixel % W, pixel // W
        s = oracle_lds(oracle_alpha(18), offset)
        want_cx = (gx.astype(np.float64) + s[:, 0]) * (1.0 / W)
        want_cy = 1.0 - ((gy.astype(np.float64) + s[:, 1]) * (1.0 / H))
        assert np.array_equal(cx.view(np.uint64), want_cx.view(np.uint64))
        assert np.array_equal(cy.view(np.uint64), want_cy.view(np.uint64))
        cam = np.array([scene.camera.lower_left_x, scene.camera.lower_left_y, scene.camera.view_x,
                        scene.camera.view_y])
        want_d = np.zeros(3)
        for i in range(0, n, max(1, n // 200)):
            O.lib().orc_camera_ray(O.dptr(cam), cx[i], cy[i], O.dptr(want_d))
            assert np.abs(d[i] - want_d).max() < 2e-7
    if total == n:  # small images: every pixel exactly spp times
        pixel, *_ = integ.raygen(0, total)
        assert np.array_equal(np.bincount(pixel, minlength=npix), np.full(npix, spp))


# ---------------------------------------------------------------------------------------------------
# stages 2+3: traversal and primitive tests
# ---------------------------------------------------------------------------------------------------
def _first_hit_check(scene, W, H, mb=8, prim_frac=0.999):
    integ = P.Integrator(scene, W, H, 1, mb)
    t_dev, p_dev = integ.first_hit()
    t_ref, p_ref, _, _ = O.OracleScene(scene.tables()).first_hit(integ.params)
    eq = p_dev == p_ref
    assert eq.mean() >= prim_frac, eq.mean()
    hit = eq & (p_ref >= 0)
    assert np.isnan(t_dev[eq & (p_ref < 0)]).all()
    rel = np.abs(t_dev[hit] - t_ref[hit]) / t_ref[hit]
    assert np.median(rel) <= 1e-6 and np.quantile(rel, 0.99) <= 2e-4, (np.median(rel), np.quantile(rel, 0.99))
    return eq.mean(), rel


def test_first_hit_map_shirley():
    _first_hit_check(P.shirley_spheres(600, 300), 600, 300)


def test_first_hit_map_cornell_mixed_shapes():
    _first_hit_check(P.cornell_box(256, 256), 256, 256, mb=16)


def test_first_hit_map_mesh():
    _first_hit_check(P.synthetic_mesh_scene(20000, 320, 180), 320, 180, prim_frac=0.995)


def test_first_hit_map_big_mesh_threaded_tree_build():
    # >= 65536 primitives: the host builder hands subtrees to a thread pool (bvh.cpp); same closest hits
    scene = P.synthetic_mesh_scene(150000, 240, 135)
    _first_hit_check(scene, 240, 135, prim_frac=0.99)
    st = scene.tree_stats()
    assert st["triangles"] >= 140000 and st["max_stack"] + 1 <= 97


@pytest.fixture(params=["gpu-sah", "gpu-lbvh"])
def gpu_builder(request, monkeypatch):
    monkeypatch.setenv("PTB_BUILDER", request.param)


def test_device_built_tree_first_hit_and_image_parity(gpu_builder):
    """gpu_bvh.cuh (binned SAH level by level, or Morton sort + radix tree, then the 4-wide collapse, all on the
    device): closest hits do not depend on the tree."""
    scene = P.synthetic_mesh_scene(60000, 240, 135)
    _first_hit_check(scene, 240, 135, prim_frac=0.99)
    st = scene.tree_stats()
    assert st["triangles"] >= 55000 and st["nodes"] > 1000 and st["max_stack"] + 1 <= 97, st
    integ = P.Integrator(scene, 240, 135, 8, 8)
    img = integ.render()
    ref, cn = O.OracleScene(scene.tables()).render(integ.params, n_threads=os.cpu_count())
    m = image_metrics(img, ref)
    assert m["rmse"] < 0.01 and m["within"] > 0.99, m
    assert abs(int(integ.stats.rays) - int(cn.rays)) / cn.rays < 2e-3
    with pytest.raises(P.PtbError):  # float64 validation mode needs the host-built tree
        integ.render(flags=capi.PTB_FLAG_F64)


def test_device_and_host_built_trees_render_the_same_image(monkeypatch):
    imgs = {}
    for b in ("host", "gpu", "gpu-lbvh"):
        monkeypatch.setenv("PTB_BUILDER", b)
        sc = P.synthetic_mesh_scene(30000, 160, 90)
        imgs[b] = P.Integrator(sc, 160, 90, 16, 8).render()
    for b in ("gpu", "gpu-lbvh"):
        m = image_metrics(imgs[b], imgs["host"])
        assert m["rmse"] < 2e-3, (b, m)  # same closest hits; float tie-breaks / atomics order differ


def _random_rays(rng, n, lo, hi):
    o = rng.uniform(lo, hi, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o.astype(np.float32), d.astype(np.float32)


@pytest.mark.parametrize("which", ["shirley", "cornell", "mesh", "one_sphere"])
def test_intersect_batch_matches_oracle(which):
    rng = np.random.default_rng(11)
    if which == "shirley":
        scene = P.shirley_spheres(600, 300)
        lo, hi = (-12, -3, -28), (12, 4, -4)
    elif which == "cornell":
        scene = P.cornell_box(64, 64)
        lo, hi = (-0.45, -0.45, -1.95), (0.45, 0.45, -1.05)
    elif which == "mesh":
        scene = P.synthetic_mesh_scene(5000, 64, 36)
        lo, hi = (-60, -60, -420), (60, 60, -280)
    else:
        scene = P.Scene()
        scene.set_textures([capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.5, 0.5, 0.5))])
        scene.set_materials([capi.Material(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=0.0)])
        scene.set_spheres([0.0], [0.0], [0.0], [1.0], [0])
        scene.set_background(capi.PTB_BG_CONSTANT, (1, 1, 1))
        lo, hi = (-3, -3, -3), (3, 3, 3)
    o, d = _random_rays(rng, 200000, lo, hi)
    t, prim = integrator.intersect_batch(scene, o, d)
    tr, pr, _ = O.OracleScene(scene.tables()).intersect_batch(o.astype(np.float64), d.astype(np.float64),
                                                              n_threads=NCPU)
    eq = prim == pr
    assert eq.mean() >= 0.999, eq.mean()
    assert (pr >= 0).mean() > 0.05  # the rays do hit something
    hit = eq & (pr >= 0)
    # float32 knows a ray origin's position relative to the r = 1000 ground sphere only to ulp(1000) = 6e-5,
    # so for arbitrary origins dt ~ 1e-4/cos(theta) is the float32 floor; it stays below the reference's own
    # 1e-3 scatter offset (shader_space.ml:53)
    adt = np.abs(t[hit] - tr[hit])
    assert np.median(adt / np.maximum(tr[hit], 1e-3)) <= 2e-6 and np.quantile(adt, 0.99) <= 1e-3, \
        (np.median(adt / np.maximum(tr[hit], 1e-3)), np.quantile(adt, 0.99))
    assert np.isnan(t[eq & (pr < 0)]).all()
    # t_min / t_max window, like spheres_intersect_native's arguments (lib.rs:55-56)
    t2, p2 = integrator.intersect_batch(scene, o, d, t_min=0.5, t_max=2.0)
    tr2, pr2, _ = O.OracleScene(scene.tables()).intersect_batch(o.astype(np.float64), d.astype(np.float64), 0.5,
                                                                2.0, n_threads=NCPU)
    assert (p2 == pr2).mean() >= 0.999
    ok = p2 >= 0
    assert ((t2[ok] >= 0.5) & (t2[ok] <= 2.0)).all()


@pytest.mark.parametrize("which", ["mesh", "soup", "sphere_cloud"])
def test_intersect_batch_axis_parallel_and_scaled_rays_on_global_memory_scenes(which):
    """Scenes too big for shared memory are traversed through quantised nodes (8-bit child boxes, decoded with the
    ray's reciprocal direction).  The rays that stress that decode: exactly axis-parallel directions (guarded
    reciprocals of 1e30 magnitude), directions with two zero components, and directions far from unit length."""
    rng = np.random.default_rng(5)
    if which == "mesh":
        scene = P.synthetic_mesh_scene(20000, 64, 36)
    elif which == "soup":
        m = 30000
        scene = P.Scene()
        scene.set_textures([capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.5, 0.5, 0.5))])
        scene.set_materials([capi.Material(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=1.0)])
        a = rng.uniform(-10, 10, size=(m, 3))
        v = np.concatenate([a, a + rng.normal(scale=0.3, size=(m, 3)), a + rng.normal(scale=0.3, size=(m, 3))])
        idx = np.stack([np.arange(m), np.arange(m) + m, np.arange(m) + 2 * m], axis=1).astype(np.int32)
        scene.set_triangles(v[:, 0], v[:, 1], v[:, 2], idx)
        scene.set_background(capi.PTB_BG_CONSTANT, (1.0, 1.0, 1.0))
    else:
        scene = sphere_cloud(20000, 64, 36, with_quad=True)
    t = scene.tables()
    pts = [np.stack([t[k][:t["n_vertices"]] for k in ("vx", "vy", "vz")], 1)] if t.get("n_vertices", 0) else []
    if t["n_spheres"]:
        pts.append(np.stack([t["xs"], t["ys"], t["zs"]], 1))
    pts = np.concatenate(pts)
    lo, hi = np.quantile(pts, 0.02, axis=0), np.quantile(pts, 0.98, axis=0)  # (the floor / ground primitives are huge)
    n = 60000
    o = rng.uniform(lo - 0.2 * (hi - lo), hi + 0.2 * (hi - lo), size=(n, 3))
    d = np.zeros((n, 3))
    axis = rng.integers(0, 3, n)
    d[np.arange(n), axis] = rng.choice([-1.0, 1.0], n)       # first third: exactly axis-parallel
    k = n // 3
    d[k:2 * k] = rng.normal(size=(k, 3))
    d[k:2 * k, 0] = 0.0                                       # second third: one zero component ...
    d[k + k // 2:2 * k, 1] = 0.0                              # ... or two
    d[2 * k:] = rng.normal(size=(n - 2 * k, 3)) * rng.choice([1e-3, 1.0, 1e3], size=(n - 2 * k, 1))  # scaled
    d[np.abs(d).sum(1) == 0] = (0.0, 0.0, 1.0)
    o32, d32 = o.astype(np.float32), d.astype(np.float32)
    tt, prim = integrator.intersect_batch(scene, o32, d32)
    tr, pr, _ = O.OracleScene(t).intersect_batch(o32.astype(np.float64), d32.astype(np.float64), n_threads=NCPU)
    for part, sl in (("axis-parallel", slice(0, k)), ("zero components", slice(k, 2 * k)), ("scaled", slice(2 * k, n))):
        eq = prim[sl] == pr[sl]
        assert eq.mean() >= 0.998, (which, part, eq.mean())
        assert (pr[sl] >= 0).mean() > 0.02, (which, part)
        assert np.isnan(tt[sl][eq & (pr[sl] < 0)]).all()


def test_intersect_batch_empty_and_single():
    scene = P.shirley_spheres(64, 32)
    t, prim = integrator.intersect_batch(scene, np.zeros((0, 3)), np.zeros((0, 3)))
    assert len(t) == 0 and len(prim) == 0
    t, prim = integrator.intersect_batch(scene, [[0, 0, 0]], [[0, 0, -1]])
    assert prim[0] >= 0 and t[0] > 0


# ---------------------------------------------------------------------------------------------------
# whole pipeline: image parity
# ---------------------------------------------------------------------------------------------------
def _render_pair(scene, W, H, spp, mb, flags=0, threads=NCPU):
    integ = P.Integrator(scene, W, H, spp, mb)
    img = integ.render(flags=flags)
    ref, cn = O.OracleScene(scene.tables()).render(integ.params, n_threads=threads,
                                                   flags=flags & ~capi.PTB_FLAG_F64)
    return img, ref, integ.stats, cn


@pytest.mark.parametrize("make,W,H,spp,mb", [
    (lambda: P.shirley_spheres(200, 100), 200, 100, 8, 8),
    (lambda: P.cornell_box(96, 96, ("constant", (1, 1, 1), None)), 96, 96, 8, 16),
    (lambda: P.cornell_box(96, 96, ("gradient", (1, 1, 1), (0.5, 0.7, 1.0))), 96, 96, 4, 16),
    (lambda: P.synthetic_mesh_scene(3000, 96, 54), 96, 54, 4, 8)])
def test_float64_device_mode_equals_oracle(make, W, H, spp, mb):
    img, ref, st, cn = _render_pair(make(), W, H, spp, mb, flags=capi.PTB_FLAG_F64, threads=1)
    # same decisions on every path: the per-bounce ray counts are identical
    assert list(st.rays_by_bounce[:mb]) == list(cn.rays_by_bounce[:mb])
    assert st.rays == cn.rays and st.paths == cn.paths
    # a handful of knife-edge paths may flip (libm last-ulp, different tree shape on exact ties)
    d = np.abs(img - ref)
    assert np.mean(d <= 1e-9) >= 0.9995 and np.sqrt(np.mean(d * d)) < 1e-4


def test_float32_image_parity_shirley_c1():
    """configs[0]: shirley_spheres --dimension=600,300 --samples-per-pixel=32 --max-ray-bounces=8"""
    img, ref, st, cn = _render_pair(P.shirley_spheres(600, 300), 600, 300, 32, 8)
    m = image_metrics(img, ref)
    assert m["rmse"] <= 0.01 and m["within"] >= 0.99 and m["bias"] <= 1e-3, m
    assert st.paths == 600 * 300 * 32
    assert abs(int(st.rays) - int(cn.rays)) / cn.rays < 1e-3
    for b in range(8):
        assert abs(int(st.rays_by_bounce[b]) - int(cn.rays_by_bounce[b])) <= 1e-3 * cn.rays_by_bounce[0]
    # the golden PNG's layout-independent facts hold for the device image too
    facts = json.load(open(os.path.join(HERE, "golden", "shirley_png_facts.json")))
    assert abs(img[0].mean() / img[1].mean() - facts["row0_over_row1"]) < 0.01
    assert abs(img[:, 0].mean() / img[:, 1].mean() - facts["col0_over_col1"]) < 0.01
    for y, rgb in facts["sky_row_mean_rgb"].items():
        got = np.floor(255.0 * img[int(y), 2:-2]).mean(0)
        assert np.abs(got - np.array(rgb)).max() <= 1.0


def test_float32_image_parity_cornell_geometry():
    """configs[1] geometry (18 triangles + 3 spheres, 16 bounces) at reduced size; white furnace
    background (the reference has no path-traced cornell image)."""
    img, ref, st, cn = _render_pair(P.cornell_box(256, 256, ("constant", (1, 1, 1), None)), 256, 256, 32, 16)
    m = image_metrics(img, ref)
    assert m["rmse"] <= 0.01 and m["within"] >= 0.99 and m["bias"] <= 1e-3, m
    assert abs(int(st.rays) - int(cn.rays)) / cn.rays < 2e-3
    # furnace: albedos <= 1 and a white environment, so no pixel can exceed 1 (+ rounding)
    assert img.max() <= 1.0 + 1e-4


def test_float32_image_parity_mesh():
    """configs[2] stand-in: synthetic mesh through the ganesha assembly, reduced size."""
    img, ref, st, cn = _render_pair(P.synthetic_mesh_scene(50000, 320, 180), 320, 180, 16, 8)
    m = image_metrics(img, ref)
    assert m["rmse"] <= 0.01 and m["within"] >= 0.985 and m["bias"] <= 1e-3, m


def sphere_cloud(n, W, H, with_quad=False, seed=7):
    """`n` small spheres of three materials in a slab in front of the camera (far more than shared memory holds:
    the traversal runs on the global-memory copy of the scene), optionally over a two-triangle floor."""
    rng = np.random.default_rng(seed)
    s = P.Scene()
    s.set_textures([capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.7, 0.4, 0.3)), capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.8, 0.8, 0.9)),
                    capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(1.0, 1.0, 1.0))])
    s.set_materials([capi.Material(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=0.0), capi.Material(kind=capi.PTB_MAT_METAL, texture=1, index=0.0),
                     capi.Material(kind=capi.PTB_MAT_DIELECTRIC, texture=2, index=1.5)])
    cam = P.scenes.Camera.create((0.0, 2.0, 9.0), (0.0, 0.5, 0.0), (0, 1, 0), W / H, 45.0)
    xs, ys, zs = rng.uniform(-6, 6, n), rng.uniform(0.05, 3.0, n), rng.uniform(-6, 3, n)
    cam.transform(xs, ys, zs)
    s.set_spheres(xs, ys, zs, rng.uniform(0.02, 0.09, n), rng.integers(0, 3, n).astype(np.int32))
    if with_quad:
        vx, vy, vz = np.array([-8.0, 8, 8, -8]), np.zeros(4), np.array([-8.0, -8, 5, 5])
        cam.transform(vx, vy, vz)
        s.set_triangles(vx, vy, vz, [0, 1, 2, 0, 2, 3], material=[0, 0])
    s.set_background(capi.PTB_BG_GRADIENT_Y, (1, 1, 1), (0.5, 0.7, 1.0))
    s.camera = cam
    return s


@pytest.mark.parametrize("with_quad", [False, True])
def test_global_memory_sphere_scene_render_parity(with_quad):
    """Spheres (and a mixed scene) too big for shared memory: the render pipeline on the global-memory traversal
    (256-bit load records, one fetch per step) against the oracle, float32 image and float64 decisions."""
    W, H = 192, 108
    _first_hit_check(sphere_cloud(20000, W, H, with_quad), W, H, prim_frac=0.998)
    img, ref, st, cn = _render_pair(sphere_cloud(20000, W, H, with_quad), W, H, 8, 8)
    m = image_metrics(img, ref)
    # (thousands of sub-pixel glass and metal spheres at 8 spp: a path that takes the other side of a silhouette in
    # float32 moves its pixel by several LSB, so fewer pixels agree closely than on the smooth scenes; the float64
    # run below is the exact check)
    assert m["rmse"] <= 0.012 and m["within"] >= 0.9 and m["bias"] <= 1e-3, m
    assert abs(int(st.rays) - int(cn.rays)) / cn.rays < 2e-3
    img, ref, st, cn = _render_pair(sphere_cloud(20000, W, H, with_quad), W, H, 2, 8, flags=capi.PTB_FLAG_F64, threads=NCPU)
    assert list(st.rays_by_bounce[:8]) == list(cn.rays_by_bounce[:8])
    # (identical decisions bounce by bounce; with ~7000 glass spheres a few last-ulp differences of libm still pick the
    # other branch of a reflect / refract draw, each of which moves the 9 pixels under its filter footprint)
    d = np.abs(img - ref)
    assert np.mean(d <= 1e-9) >= 0.995 and np.sqrt(np.mean(d * d)) < 2e-3, (np.mean(d <= 1e-9), np.sqrt(np.mean(d * d)))


def test_float32_high_spp_tolerance():
    img, ref, st, cn = _render_pair(P.shirley_spheres(160, 80), 160, 80, 256, 8)
    m = image_metrics(img, ref)
    assert m["rmse"] <= 0.003 and m["within"] >= 0.99, m


# ---------------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("W,H,spp,mb", [(1, 1, 1, 8), (37, 19, 1, 1), (5, 3, 7, 2), (64, 32, 2, 0),
                                        (33, 65, 3, 64)])
def test_edge_sizes_against_oracle(W, H, spp, mb):
    scene = P.shirley_spheres(W, H)
    img, ref, st, cn = _render_pair(scene, W, H, spp, mb, flags=capi.PTB_FLAG_F64, threads=1)
    assert st.rays == cn.rays
    assert np.abs(img - ref).max() <= 1e-9
    if mb == 0:  # integrator.ml:31-32: the path stops before intersecting anything
        assert st.rays == 0 and not img.any()
    img32 = P.Integrator(scene, W, H, spp, mb).render()
    assert np.sqrt(np.mean((img32 - ref) ** 2)) <= 0.02


@pytest.mark.parametrize("batch", [1024, 5000, 70001])
def test_small_batches_cross_pass_and_claim_boundaries(monkeypatch, batch):
    """Batches that are no multiple of the pixel count or of the segment size: the camera-sample enumeration wraps
    inside claims, chunks end inside segments, and the last batch is ragged.  Float64 mode must still take exactly
    the oracle's decisions."""
    monkeypatch.setenv("PTB_BATCH", str(batch))
    W, H, spp, mb = 61, 37, 9, 6
    scene = P.shirley_spheres(W, H)
    img, ref, st, cn = _render_pair(scene, W, H, spp, mb, flags=capi.PTB_FLAG_F64, threads=1)
    assert list(st.rays_by_bounce[:mb]) == list(cn.rays_by_bounce[:mb])
    assert st.paths == W * H * spp
    d = np.abs(img - ref)
    assert np.mean(d <= 1e-9) >= 0.9995
    img32 = P.Integrator(scene, W, H, spp, mb).render()
    assert np.sqrt(np.mean((img32 - ref) ** 2)) <= 0.02


def test_raw_and_unfiltered_outputs_compose():
    W, H, spp, mb = 120, 60, 4, 8
    scene = P.shirley_spheres(W, H)
    integ = P.Integrator(scene, W, H, spp, mb)
    sums = integ.render(flags=capi.PTB_FLAG_NO_FILTER | capi.PTB_FLAG_F64)
    raw = integ.render(flags=capi.PTB_FLAG_RAW_SUMS | capi.PTB_FLAG_F64)
    img = integ.render(flags=capi.PTB_FLAG_F64)
    w = np.zeros(9)
    O.lib().orc_filter_binomial(5, 1, O.dptr(w))
    assert np.abs(resolve_numpy(sums, spp, w.reshape(3, 3), raw=True) - raw).max() < 1e-12
    assert np.abs(np.sqrt(raw / spp) - img).max() < 1e-12
    ref, _ = O.OracleScene(scene.tables()).render(integ.params, flags=capi.PTB_FLAG_RAW_SUMS)
    assert np.abs(raw - ref).max() < 1e-9


def test_tile_shards_sum_to_the_whole_image():
    import torch
    W, H, spp, mb = 200, 120, 4, 8
    scene = P.shirley_spheres(W, H)
    whole = torch.zeros(H, W, 3, device="cuda")
    P.Integrator(scene, W, H, spp, mb).render_device(whole)
    for world in (2, 3, 8):
        acc = torch.zeros(H, W, 3, device="cuda")
        paths = 0
        for r in range(world):
            part = torch.zeros(H, W, 3, device="cuda")
            integ = P.Integrator(scene, W, H, spp, mb, tile_rank=r, tile_world=world)
            integ.render_device(part)
            paths += integ.stats.paths
            assert (part != 0).any()
            acc += part
        assert paths == W * H * spp
        # same samples, same arithmetic; only the float32 atomic accumulation order differs
        assert torch.allclose(acc, whole, rtol=2e-5, atol=1e-5)
    integ = P.Integrator(scene, W, H, spp, mb)
    img = integ.resolve_device(whole).cpu().numpy().astype(np.float64)
    assert np.sqrt(np.mean((img - integ.render()) ** 2)) < 1e-5


def test_errors_on_device_path():
    scene = P.shirley_spheres(64, 32)
    with pytest.raises(P.PtbError, match="max_bounces"):
        P.Integrator(scene, 64, 32, 1, 65).render()
    with pytest.raises(P.PtbError, match="positive"):
        P.Integrator(scene, 64, 32, 0, 8).render()
    fresh = P.shirley_spheres(64, 32)
    p = make_params(fresh, 64, 32, 1, 8)
    rc = P.lib().ptb_render(fresh.h, C.byref(p), capi.dptr(np.zeros((32, 64, 3))), None)
    assert rc == -4 and b"not committed" in P.lib().ptb_last_error()


# ---------------------------------------------------------------------------------------------------
# BASELINE.json's full size (3840x2160) through size-independent properties
# ---------------------------------------------------------------------------------------------------
def test_full_size_properties_c4():
    import torch
    W, H, spp, mb = 3840, 2160, 16, 8
    scene = P.shirley_spheres(W, H)
    integ = P.Integrator(scene, W, H, spp, mb)
    sums = torch.zeros(H, W, 3, device="cuda")
    integ.render_device(sums)
    st = integ.stats
    rb = list(st.rays_by_bounce[:mb])
    assert st.paths == W * H * spp and rb[0] == st.paths and st.rays == sum(rb)
    assert all(rb[i] >= rb[i + 1] for i in range(mb - 1)) and rb[-1] > 0
    # rays per path is a property of the scene, not of the resolution: compare with the small config
    small = P.Integrator(P.shirley_spheres(640, 360), 640, 360, 32, mb)  # same 16:9 aspect
    small.render()
    assert abs(st.rays / st.paths - small.stats.rays / small.stats.paths) < 0.02
    img = integ.resolve_device(sums)
    # edge darkening sqrt(37/48) of the 3x3 splat and (37/48) at the corners (pre-gamma ratio squared)
    r = (img[0, 8:-8].mean() / img[1, 8:-8].mean()).item()
    c = (img[8:-8, 0].mean() / img[8:-8, 1].mean()).item()
    assert abs(r - math.sqrt(37 / 48)) < 0.01 and abs(c - math.sqrt(37 / 48)) < 0.01
    # sky: the top rows are pure background, analytic after filter+gamma; compare with the oracle at a
    # strided set of pixels through orc_trace_sample (per-sample radiance, exact for sky paths)
    osc = O.OracleScene(scene.tables())
    s = sums.cpu().numpy().astype(np.float64)
    for gx, gy in [(10, 5), (1900, 3), (3800, 20), (2000, 40)]:
        want = sum(osc.trace_sample(integ.params, gx, gy, k) for k in range(spp))
        assert np.abs(s[gy, gx] - want).max() < 1e-4 * spp
    # linearity in the environment: doubling the background doubles every unfiltered sum
    t = scene.tables()
    scene2 = P.shirley_spheres(W, H)
    scene2.set_background(capi.PTB_BG_GRADIENT_Y, 2 * t["bg0"], 2 * t["bg1"])
    sums2 = torch.zeros(H, W, 3, device="cuda")
    P.Integrator(scene2, W, H, spp, mb).render_device(sums2)
    assert torch.allclose(sums2, 2 * sums, rtol=1e-4, atol=1e-4)
    # downsampled by 6.4x7.2 the image agrees with the small config's image (same scene, same camera)
    small_img = torch.from_numpy(small.render()).float()
    big = torch.nn.functional.adaptive_avg_pool2d(img.cpu().permute(2, 0, 1)[None] ** 2, (360, 640))[0]
    sm = small_img.permute(2, 0, 1) ** 2
    assert (big[:, 4:-4, 4:-4] - sm[:, 4:-4, 4:-4]).abs().mean().item() < 0.02


@pytest.mark.gpu
def test_single_process_multi_gpu_render_equals_one_gpu():
    """ptb_render_multi: tiles dealt to the GPUs of this process, peer-memory reduce on device 0.  Same image as
    one GPU up to float summation order (skipped on a one-GPU box; the 1-device call must work anywhere)."""
    W, H, spp, mb = 200, 120, 8, 8
    ref = P.Integrator(P.shirley_spheres(W, H), W, H, spp, mb).render()
    n = min(P.lib().ptb_device_count(), 4)
    for k in sorted({1, n}):
        sc = P.shirley_spheres(W, H)
        integ = P.Integrator(sc, W, H, spp, mb)
        img = integ.render_multi(k)
        m = image_metrics(img, ref)
        assert m["rmse"] < 2e-6 and m["max"] < 1e-4, (k, m)
        assert integ.stats.paths == W * H * spp
    with pytest.raises(P.PtbError):
        P.Integrator(P.shirley_spheres(W, H), W, H, spp, mb).render_multi(P.lib().ptb_device_count() + 1)
