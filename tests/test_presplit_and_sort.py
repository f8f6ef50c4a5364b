"""Triangle pre-splitting (csrc/presplit.hpp, both tree builders) and the ray sort of ptb_intersect_batch
(csrc/ray_sort.cuh).  Neither may change a closest hit: Shape_tree.intersect's result (shape_tree.ml:198-220) does not
depend on how the tree groups the shapes, and the FFI contract (sphere-intersect-rs/src/lib.rs:53-76) fixes only that
result i belongs to ray i.

CPU: the pieces of a triangle cover it, stay inside its box and respect the cell size.
GPU: dense soups (host-built and device-built trees) against the oracle and against the same scene without
pre-splitting, bit for bit; sorted against unsorted ray order, bit for bit, including rays that start outside the
scene, miss it, are axis-parallel, and a t_min / t_max window; an image of a pre-split soup against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi, integrator
from path_tracer_ocaml_b200.scenes import Camera
import pyoracle as O

NCPU = os.cpu_count() or 1
_dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731


def _pieces(v, cell, org, cap=4096):
    out = np.zeros(6 * cap)
    n = P.lib().ptb_presplit_boxes(_dp(np.ascontiguousarray(v, dtype=np.float64).ravel()), float(cell), _dp(np.asarray(org, dtype=np.float64)),
                                   _dp(out), cap)
    assert 1 <= n <= cap
    return out[:6 * n].reshape(n, 6)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_pieces_cover_the_triangle_and_respect_the_cell_size(seed):
    rng = np.random.default_rng(seed)
    for _ in range(300):
        a = rng.uniform(-10, 10, 3)
        v = np.stack([a, a + rng.normal(scale=rng.choice([0.05, 0.3, 2.0]), size=3), a + rng.normal(scale=0.3, size=3)])
        cell = float(rng.choice([0.05, 0.11, 0.25, 0.7, 5.0]))
        org = rng.uniform(-11, -10, 3)
        b = _pieces(v, cell, org)
        lo, hi = v.min(0), v.max(0)
        assert (b[:, :3] >= lo - 1e-12).all() and (b[:, 3:] <= hi + 1e-12).all()       # inside the triangle's own box
        # no piece longer than the cell — or, for a triangle more than 8 cells long, than a quarter of the triangle
        assert ((b[:, 3:] - b[:, :3]) <= max(cell, (hi - lo).max() / 4) * (1 + 1e-9)).all()
        assert len(b) <= 200
        if ((hi - lo) <= cell).all():
            assert len(b) == 1 and np.allclose(b[0, :3], lo) and np.allclose(b[0, 3:], hi)  # small triangles stay whole
        # every point of the triangle lies in the box of some piece (the property the traversal relies on)
        w = rng.dirichlet((1, 1, 1), size=400)
        w[:3] = np.eye(3)  # the corners themselves
        pts = w @ v
        tol = 1e-9 * (1 + np.abs(pts).max())
        inside = ((pts[:, None, :] >= b[None, :, :3] - tol) & (pts[:, None, :] <= b[None, :, 3:] + tol)).all(2).any(1)
        assert inside.all()


def test_pieces_of_axis_aligned_and_degenerate_triangles():
    org = np.zeros(3)
    # a big axis-aligned right triangle in the z = 1 plane, corners on grid planes
    b = _pieces([[0, 0, 1], [4, 0, 1], [0, 4, 1]], 1.0, org)
    assert (b[:, 2] == 1).all() and (b[:, 5] == 1).all()
    assert 10 <= len(b) <= 16  # the 10 cells the triangle covers, plus at most the ones its hypotenuse only touches
    area = ((b[:, 3] - b[:, 0]) * (b[:, 4] - b[:, 1])).sum()
    assert 8.0 - 1e-9 <= area <= 16.0  # the boxes cover the triangle (area 8) inside its 4 x 4 box
    # zero-area triangles: a segment and a point
    assert len(_pieces([[0, 0, 0], [3, 0, 0], [1.5, 0, 0]], 1.0, org)) >= 1
    assert len(_pieces([[2, 2, 2], [2, 2, 2], [2, 2, 2]], 1.0, org)) == 1
    assert P.lib().ptb_presplit_boxes(_dp(np.zeros(9)), 0.0, _dp(org), _dp(np.zeros(6)), 1) == -1  # PTB_E_INVALID


# ----------------------------------------------------------------------------------------------------------------------
def _soup(m, half, seed, edge=0.3, centre=(0.0, 0.0, 0.0)):
    rng = np.random.default_rng(seed)
    s = P.Scene()
    s.set_textures([capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.6, 0.5, 0.4)), capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.2, 0.5, 0.8))])
    s.set_materials([capi.Material(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=1.0),
                     capi.Material(kind=capi.PTB_MAT_METAL, texture=1, index=1.0)])
    a = rng.uniform(-half, half, size=(m, 3)) + np.asarray(centre)
    v = np.concatenate([a, a + rng.normal(scale=edge, size=(m, 3)), a + rng.normal(scale=edge, size=(m, 3))])
    idx = np.stack([np.arange(m), np.arange(m) + m, np.arange(m) + 2 * m], axis=1).astype(np.int32)
    s.set_triangles(v[:, 0], v[:, 1], v[:, 2], idx, material=(np.arange(m) % 2).astype(np.int32))
    s.set_background(capi.PTB_BG_CONSTANT, (1.0, 1.0, 1.0))
    return s


def _rays(n, half, seed):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-1.3 * half, 1.3 * half, size=(n, 3))  # a good part starts outside the scene's box
    d = rng.normal(size=(n, 3))
    k = n // 8
    d[:k] = 0.0
    d[np.arange(k), rng.integers(0, 3, k)] = rng.choice([-1.0, 1.0], k)  # axis-parallel
    d[k:2 * k, 0] = 0.0                                                  # one zero component
    o[2 * k:3 * k] = rng.uniform(3 * half, 4 * half, size=(k, 3))        # far outside, most of them miss the box
    return o.astype(np.float32), d.astype(np.float32)


@pytest.mark.gpu
@pytest.mark.parametrize("m,half,builder", [(30_000, 2.5, "host"), (250_000, 5.0, "device")])
def test_presplit_soup_matches_oracle_and_the_plain_tree(m, half, builder, monkeypatch):
    """A soup dense enough for the builders to pre-split (several boxes contain a random point): the tree holds more
    references than triangles, and every closest hit equals the one of the tree without pre-splitting bit for bit and
    the oracle's within the float32 tolerances of test_gpu_parity."""
    n = 200_000
    o, d = _rays(n, half, 11)
    monkeypatch.delenv("PTB_BVH_PRESPLIT", raising=False)
    s1 = _soup(m, half, 3)
    s1.commit(0)
    st = s1.tree_stats()
    assert st["triangles"] > 1.5 * m, st          # references, not triangles
    assert P.lib().ptb_scene_primitive_count(s1.h) == m
    t1, p1 = integrator.intersect_batch(s1, o, d)
    monkeypatch.setenv("PTB_BVH_PRESPLIT", "0")
    s0 = _soup(m, half, 3)
    s0.commit(0)
    assert s0.tree_stats()["triangles"] == m
    t0, p0 = integrator.intersect_batch(s0, o, d)
    assert np.array_equal(p0, p1)
    assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
    assert 0 <= p1.max() < m and (p1 >= 0).mean() > 0.3
    ns = 40_000
    tr, pr, _ = O.OracleScene(s1.tables()).intersect_batch(o[:ns].astype(np.float64), d[:ns].astype(np.float64), n_threads=NCPU)
    eq = p1[:ns] == pr
    assert eq.mean() >= 0.999, eq.mean()
    hit = eq & (pr >= 0)
    assert np.quantile(np.abs(t1[:ns][hit] - tr[hit]), 0.99) <= 1e-3
    assert np.isnan(t1[:ns][eq & (pr < 0)]).all()


@pytest.mark.gpu
def test_sorted_ray_order_returns_the_same_bits(monkeypatch):
    """PTB_BATCH_SORT=1 (spatial order in the traversal queue, results scattered back to the caller's index) against
    =0 on a global-memory scene: identical t and primitive for every ray, with a t_min / t_max window, a ray count
    that is no multiple of the segment size, through the device-resident and the chunked host-buffer entry points."""
    half = 5.0
    s = _soup(60_000, half, 5)
    s.commit(0)
    n = (1 << 18) + 77
    o, d = _rays(n, half, 12)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("PTB_BATCH_SORT", mode)
        out[mode] = integrator.intersect_batch(s, o, d, t_min=0.05, t_max=6.0)
        monkeypatch.setenv("PTB_BATCH_CHUNK", "50000")  # several chunks, two streams
        out[mode + "c"] = integrator.intersect_batch(s, o, d, t_min=0.05, t_max=6.0)
        monkeypatch.delenv("PTB_BATCH_CHUNK")
    t0, p0 = out["0"]
    assert (p0 >= 0).mean() > 0.2 and (p0 < 0).mean() > 0.05
    hit = p0 >= 0
    assert (t0[hit] >= 0.05).all() and (t0[hit] <= 6.0).all()
    for k in ("1", "0c", "1c"):
        assert np.array_equal(out[k][1], p0), k
        assert np.array_equal(out[k][0].view(np.uint32), t0.view(np.uint32)), k
    # the default picks the sort for this scene only if the builder pre-split it; either way the answers are these
    monkeypatch.delenv("PTB_BATCH_SORT")
    t2, p2 = integrator.intersect_batch(s, o, d, t_min=0.05, t_max=6.0)
    assert np.array_equal(p2, p0) and np.array_equal(t2.view(np.uint32), t0.view(np.uint32))


@pytest.mark.gpu
def test_render_of_a_presplit_soup_matches_oracle(monkeypatch):
    """The render pipeline on a pre-split tree: the per-slot material and uv tables follow the references, so the image
    of a two-material soup agrees with the oracle's like any other triangle scene."""
    monkeypatch.delenv("PTB_BVH_PRESPLIT", raising=False)
    s = _soup(20_000, 2.0, 9, centre=(0.0, 0.0, -7.0))  # camera space: the eye at the origin looks down -z
    s.camera = Camera.create(eye=(0.0, 0.0, 0.0), target=(0.0, 0.0, -1.0), up=(0.0, 1.0, 0.0), aspect=1.0, vertical_fov_deg=40.0)
    W = H = 128
    integ = P.Integrator(s, W, H, 16, 6)
    img = integ.render()
    assert s.tree_stats()["triangles"] > 20_000
    ref, cn = O.OracleScene(s.tables()).render(integ.params, n_threads=NCPU, flags=0)
    d = img - ref
    assert np.sqrt(np.mean(d * d)) <= 0.01
    assert abs(int(integ.stats.rays) - int(cn.rays)) / cn.rays < 2e-3


@pytest.mark.gpu
def test_a_surface_with_a_tilted_ground_plane_is_not_taken_for_a_soup(monkeypatch):
    """The pre-split trigger counts how many triangle boxes contain a random point, every box capped at 64 mean cells:
    a wavy surface mesh stays one box per triangle even with two huge tilted triangles under it, whose boxes alone
    fill the scene; the hits equal the oracle's either way."""
    monkeypatch.delenv("PTB_BVH_PRESPLIT", raising=False)
    g = 80
    u, v = np.meshgrid(np.linspace(-5, 5, g + 1), np.linspace(-5, 5, g + 1), indexing="ij")
    z = 0.3 * np.sin(1.7 * u) * np.cos(1.3 * v)
    verts = np.stack([u.ravel(), v.ravel(), z.ravel()], 1)
    q = (np.arange(g)[:, None] * (g + 1) + np.arange(g)[None, :]).ravel()
    faces = np.concatenate([np.stack([q, q + 1, q + g + 2], 1), np.stack([q, q + g + 2, q + g + 1], 1)])
    nv = len(verts)
    ground = np.array([[-6, -6, -3.0], [6, -6, -1.0], [6, 6, 1.0], [-6, 6, -1.0]])  # tilted: its triangles' boxes span the scene
    verts = np.concatenate([verts, ground])
    faces = np.concatenate([faces, [[nv, nv + 1, nv + 2], [nv, nv + 2, nv + 3]]]).astype(np.int32)
    m = len(faces)
    assert m >= 4096
    s = P.Scene()
    s.set_textures([capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.5, 0.5, 0.5))])
    s.set_materials([capi.Material(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=1.0)])
    s.set_triangles(verts[:, 0], verts[:, 1], verts[:, 2], faces)
    s.set_background(capi.PTB_BG_CONSTANT, (1.0, 1.0, 1.0))
    s.commit(0)
    assert s.tree_stats()["triangles"] == m
    o, d = _rays(50_000, 5.0, 21)
    t, p = integrator.intersect_batch(s, o, d)
    tr, pr, _ = O.OracleScene(s.tables()).intersect_batch(o.astype(np.float64), d.astype(np.float64), n_threads=NCPU)
    assert (p == pr).mean() >= 0.999 and (pr >= 0).mean() > 0.2
