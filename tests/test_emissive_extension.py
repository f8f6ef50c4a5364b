"""The extension behind BASELINE.json configs[1] ("diffuse+light sampling"): an emissive material
and an area-light mixture pdf, behind the hooks the reference already threads (Material.emit, Hit.emit, Pdf.t,
~diffuse_plus_light; material.ml:59, pdf.ml:3-17, integrator.ml:38-66).  It is beyond the reference, so its only
oracle is oracle/oracle.cpp's extension (marked EXT there); the CPU tests below pin that extension analytically, the
GPU tests compare the device with it."""
import os

import numpy as np
import pytest

import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi
import pyoracle as O
from helpers import image_metrics, make_params

NCPU = os.cpu_count() or 1


def lamp_over_floor(W, H, with_light_pdf, radiance=8.0):
    """A grey floor quad under a small emissive quad, black background, camera looking down at the floor."""
    s = P.Scene()
    s.set_textures([capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.6, 0.5, 0.4)),
                    capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(radiance, radiance, radiance))])
    s.set_materials([capi.Material(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=0.0),
                     capi.Material(kind=capi.PTB_MAT_EMISSIVE, texture=1, index=0.0)])
    cam = P.scenes.Camera.create((0.0, 1.5, 3.0), (0.0, 0.0, 0.0), (0, 1, 0), W / H, 50.0)
    # floor y = 0, [-2,2]^2; lamp y = 1, [-0.25,0.25]^2
    vx = np.array([-2.0, 2, 2, -2, -0.25, 0.25, 0.25, -0.25])
    vy = np.array([0.0, 0, 0, 0, 1, 1, 1, 1])
    vz = np.array([-2.0, -2, 2, 2, -0.25, -0.25, 0.25, 0.25])
    lo, lu, lv = np.array([[-0.25, 0.25, -0.25], [1.0, 1.0, 1.0], [-0.25, -0.25, 0.25]])
    px, py, pz = lo.copy(), lu.copy(), lv.copy()  # three points: origin, origin+u, origin+v (columns x, y, z)
    cam.transform(vx, vy, vz)
    cam.transform(px, py, pz)
    s.set_triangles(vx, vy, vz, [0, 1, 2, 0, 2, 3, 4, 5, 6, 4, 6, 7], material=[0, 0, 1, 1])
    s.set_background(capi.PTB_BG_CONSTANT, (0, 0, 0))
    if with_light_pdf:
        o = np.array([px[0], py[0], pz[0]])
        s.set_light_quad(o, np.array([px[1], py[1], pz[1]]) - o, np.array([px[2], py[2], pz[2]]) - o)
    s.camera = cam
    return s


def emissive_shell(W, H, radiance=(2.0, 1.0, 0.5), albedo=(0.5, 0.25, 0.75)):
    """A lambert sphere inside a big emissive cube (12 triangles): every path ends on the emitter after 0 or 1
    scatter.  (A cube, not a sphere: the reference's Sphere.intersect returns the near root q/a whenever the origin
    is inside, sphere.ml:52, so a ray that starts inside a sphere and heads away from its centre misses it.)"""
    s = P.Scene()
    s.set_textures([capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=albedo), capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=radiance)])
    s.set_materials([capi.Material(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=0.0),
                     capi.Material(kind=capi.PTB_MAT_EMISSIVE, texture=1, index=0.0)])
    cam = P.scenes.Camera.create((0.0, 0.0, 5.0), (0.0, 0.0, 0.0), (0, 1, 0), W / H, 40.0)
    xs, ys, zs = np.array([0.0]), np.array([0.0]), np.array([0.0])
    cam.transform(xs, ys, zs)
    s.set_spheres(xs, ys, zs, [1.0], [0])
    c = 50.0
    v = np.array([[x, y, z] for x in (-c, c) for y in (-c, c) for z in (-c, c)])  # index = 4*ix + 2*iy + iz
    vx, vy, vz = v[:, 0].copy(), v[:, 1].copy(), v[:, 2].copy()
    cam.transform(vx, vy, vz)
    quads = [(0, 1, 3, 2), (4, 5, 7, 6), (0, 1, 5, 4), (2, 3, 7, 6), (0, 2, 6, 4), (1, 3, 7, 5)]
    idx = [i for a, b, cc, d in quads for i in (a, b, cc, a, cc, d)]
    s.set_triangles(vx, vy, vz, idx, material=[1] * 12)
    s.set_background(capi.PTB_BG_CONSTANT, (0, 0, 0))
    s.camera = cam
    return s


# ---- CPU: the oracle's extension, analytically -----------------------------------------------------------------
def test_oracle_emissive_shell_is_analytic():
    W, H, spp = 48, 48, 4
    sc = emissive_shell(W, H)
    sums, cn = O.OracleScene(sc.tables()).render(make_params(sc, W, H, spp, 8, flags=capi.PTB_FLAG_NO_FILTER), n_threads=1)
    mean = sums / spp
    E, a = np.array([2.0, 1.0, 0.5]), np.array([0.5, 0.25, 0.75])
    # corner pixels see the emitter directly: emit0 + attn0 * emit = E (integrator.ml:40,43 with Absorb)
    assert np.allclose(mean[0, 0], E, rtol=0, atol=1e-12)
    # the centre sees the lambert sphere: one cosine-sampled scatter (pd = 1), then the emitter: albedo * E
    assert np.allclose(mean[H // 2, W // 2], a * E, rtol=0, atol=1e-12)
    assert cn.exhausted == 0 and cn.missed == 0 and cn.absorbed == cn.paths


def test_oracle_light_sampling_is_unbiased_and_reduces_noise():
    W, H, spp = 48, 32, 1024  # (low sample counts reuse offsets between neighbouring pixels: `pass * spp`, integrator.ml:98)
    imgs = {}
    for lit in (False, True):
        sc = lamp_over_floor(W, H, lit)
        sums, cn = O.OracleScene(sc.tables()).render(make_params(sc, W, H, spp, 4, flags=capi.PTB_FLAG_NO_FILTER),
                                                     n_threads=NCPU)
        imgs[lit] = sums / spp
    floor = imgs[False].sum(-1) > 0
    a, b = imgs[False][floor].mean(0), imgs[True][floor].mean(0)
    assert np.abs(a - b).max() / a.max() < 0.01, (a, b)  # same expectation (measured: 0.0009 here, 0.0013 at 4096 spp)
    # variance: against a 4096-spp render, the floor pixels of the light-sampled image are an order of magnitude
    # closer than those of the cosine-sampled one at the same sample count (measured: 0.00026 vs 0.0045)
    sc = lamp_over_floor(W, H, True)
    ref = O.OracleScene(sc.tables()).render(make_params(sc, W, H, 4096, 4, flags=capi.PTB_FLAG_NO_FILTER),
                                            n_threads=NCPU)[0] / 4096
    fl = floor & (ref[..., 0] < 4.0)  # not the lamp itself
    err = {lit: np.abs(imgs[lit] - ref)[fl].mean() for lit in (False, True)}
    assert err[True] < 0.2 * err[False], err


def test_lit_cornell_tables():
    sc = P.cornell_box_lit(64, 64)
    t = sc.tables()
    assert t["n_triangles"] == 20 and t["n_spheres"] == 3 and t["has_light"]
    kinds = [t["materials"][i].kind for i in range(t["n_materials"])]
    assert kinds.count(capi.PTB_MAT_EMISSIVE) == 1
    # the light quad is the emissive square: side 0.1, horizontal in world space => |u| = |v| = 0.1, u.v = 0
    u, v = t["light_u"], t["light_v"]
    assert abs(np.linalg.norm(u) - 0.1) < 1e-12 and abs(np.linalg.norm(v) - 0.1) < 1e-12 and abs(u @ v) < 1e-15
    assert not P.cornell_box(64, 64).tables()["has_light"]
    with pytest.raises(P.PtbError):
        sc.set_light_quad((0, 0, 0), (1, 0, 0), (2, 0, 0))  # degenerate


# ---- GPU: device vs the extended oracle ------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("make,W,H,spp,mb", [
    (lambda: emissive_shell(64, 64), 64, 64, 4, 8),
    (lambda: lamp_over_floor(96, 64, False), 96, 64, 8, 4),
    (lambda: lamp_over_floor(96, 64, True), 96, 64, 8, 4),
    (lambda: P.cornell_box_lit(96, 96), 96, 96, 8, 16),
    (lambda: lamp_over_floor(64, 48, True), 64, 48, 4, 1)])  # max_bounces = 1: the lamp's own emission must still be collected
def test_float64_device_mode_equals_extended_oracle(make, W, H, spp, mb):
    sc = make()
    integ = P.Integrator(sc, W, H, spp, mb)
    img = integ.render(flags=capi.PTB_FLAG_F64 | capi.PTB_FLAG_NO_FILTER)
    st = integ.stats
    ref, cn = O.OracleScene(sc.tables()).render(make_params(sc, W, H, spp, mb, flags=capi.PTB_FLAG_NO_FILTER), n_threads=1)
    assert list(st.rays_by_bounce[:mb]) == list(cn.rays_by_bounce[:mb])
    d = np.abs(img - ref)
    assert np.mean(d <= 1e-9 * (1 + np.abs(ref))) >= 0.999 and ref.max() > 0, (np.mean(d <= 1e-9), d.max())


@pytest.mark.gpu
def test_float32_parity_lit_scenes():
    for make, W, H, spp, mb in [(lambda: lamp_over_floor(160, 96, True), 160, 96, 64, 4),
                                (lambda: P.cornell_box_lit(128, 128), 128, 128, 64, 16)]:
        sc = make()
        integ = P.Integrator(sc, W, H, spp, mb)
        img = integ.render()
        ref, cn = O.OracleScene(sc.tables()).render(integ.params, n_threads=NCPU)
        m = image_metrics(img, ref)
        assert m["rmse"] <= 0.01 and m["within"] >= 0.98 and m["bias"] <= 2e-3, m
        assert abs(int(integ.stats.rays) - int(cn.rays)) / cn.rays < 2e-3


@pytest.mark.gpu
def test_emissive_shell_is_analytic_on_device():
    W, H, spp = 48, 48, 4
    sc = emissive_shell(W, H)
    mean = P.Integrator(sc, W, H, spp, 8).render(flags=capi.PTB_FLAG_NO_FILTER) / spp
    assert np.allclose(mean[0, 0], [2.0, 1.0, 0.5], atol=1e-5)
    assert np.allclose(mean[H // 2, W // 2], [1.0, 0.25, 0.375], atol=1e-5)
