"""CPU tests of the product's host side: the C-ABI library loads and exports every declared symbol,
the host helpers equal the oracle bit for bit, scene tables round-trip, errors mirror the
reference's failure sites, and the N>1 sharding logic works over gloo.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi
import pyoracle as O
from helpers import make_params, resolve_numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = P.lib()


def test_library_exports_every_declared_symbol():
    declared = set()
    for h in ("ptb200.h", "ptb200_scenes.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        declared |= set(re.findall(r"\b(ptb_[a-z0-9_]+)\s*\(", src))
    assert len(declared) >= 30
    raw = C.CDLL(capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in include/ but not exported"
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    assert L.ptb_leaf_size() >= 1 and b"sm_100a" in L.ptb_version()


@pytest.mark.parametrize("D", [1, 2, 18, 34, 130])
def test_lds_alpha_bit_equal_to_oracle(D):
    a, b = np.zeros(D), np.zeros(D)
    assert L.ptb_lds_alpha(D, capi.dptr(a)) == 0
    O.lib().orc_lds_alpha(D, O.dptr(b))
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_lds_alpha_rejects_dimension_zero():
    # Low_discrepancy_sequence.create: failwith "expected dimension >= 1" (…ml:27-31)
    assert L.ptb_lds_alpha(0, capi.dptr(np.zeros(1))) == -1
    assert b"dimension >= 1" in L.ptb_last_error()


@pytest.mark.parametrize("whm", [(10, 5, 7), (600, 300, 1024), (1024, 1024, 1024), (1920, 1080, 1024),
                                 (3840, 2160, 1024), (37, 19, 1024), (1, 1, 1024), (1000, 3, 50)])
def test_tile_split_equal_to_oracle(whm):
    w, h, m = whm
    cap = 20000
    a = [np.zeros(cap, dtype=np.int32) for _ in range(4)]
    b = [np.zeros(cap, dtype=np.int32) for _ in range(4)]
    n = L.ptb_tile_split(w, h, m, *[capi.iptr(x) for x in a], cap)
    k = O.lib().orc_tile_split(w, h, m, *[O.iptr(x) for x in b], cap)
    assert n == k and n > 0
    for x, y in zip(a, b):
        assert np.array_equal(x[:n], y[:n])
    assert L.ptb_tile_split(3840, 2160, 1024, *[capi.iptr(x) for x in a], 10) == -1000 - 8192


def test_filter_binomial_bit_equal_to_oracle():
    a, b = np.zeros(9), np.zeros(9)
    assert L.ptb_filter_binomial(5, 1, capi.dptr(a)) == 0
    O.lib().orc_filter_binomial(5, 1, O.dptr(b))
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


@pytest.mark.parametrize("cam", [((13.0, 2.0, 4.5), (0, 0, 0), (0, 1, 0), 2.0, 20.0),
                                 ((0.5, 0.5, -1.0), (0.5, 0.5, 0.0), (0, 1, 0), 1.0, 53.13010235415598),
                                 ((328.0, 70.282, 345.0), (328.0, 10.0, 0.0), (-0.00212272, 0.998201, -0.0599264),
                                  16 / 9, 30.0)])
def test_camera_bit_equal_to_oracle(cam):
    eye, tgt, up, aspect, fov = (np.array(v, dtype=np.float64) if isinstance(v, tuple) else v for v in cam)
    a, b = np.zeros(20), np.zeros(20)
    assert L.ptb_camera_create(capi.dptr(eye), capi.dptr(tgt), capi.dptr(up), aspect, fov, capi.dptr(a)) == 0
    O.lib().orc_camera_create(O.dptr(eye), O.dptr(tgt), O.dptr(up), aspect, fov, O.dptr(b))
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    rng = np.random.default_rng(3)
    pts = rng.uniform(-20, 20, size=(3, 64))
    p1 = [np.ascontiguousarray(r) for r in pts.copy()]
    p2 = [np.ascontiguousarray(r) for r in pts.copy()]
    L.ptb_camera_transform(capi.dptr(a[4:].copy()), *[capi.dptr(x) for x in p1], 64)
    O.lib().orc_camera_transform(O.dptr(b[4:].copy()), *[O.dptr(x) for x in p2], 64)
    for x, y in zip(p1, p2):
        assert np.array_equal(x.view(np.uint64), y.view(np.uint64))
    # the eye maps to the origin and the target onto the -z axis (camera.ml:14-27)
    e = [np.array([v]) for v in eye]
    L.ptb_camera_transform(capi.dptr(a[4:].copy()), *[capi.dptr(x) for x in e], 1)
    assert max(abs(v[0]) for v in e) < 1e-9
    t = [np.array([v]) for v in tgt]
    L.ptb_camera_transform(capi.dptr(a[4:].copy()), *[capi.dptr(x) for x in t], 1)
    assert abs(t[0][0]) < 1e-9 and abs(t[1][0]) < 1e-9 and t[2][0] < 0


def test_shirley_scene_tables():
    s = P.shirley_spheres(600, 300)
    t = s.tables()
    assert t["n_spheres"] == 530 and t["n_triangles"] == 0
    assert t["rs"][0] == 1000.0 and list(t["rs"][1:4]) == [1.0, 1.0, 1.0] and set(t["rs"][4:]) == {0.2}
    kinds = [t["materials"][m].kind for m in t["sphere_material"]]
    assert kinds[:4] == [capi.PTB_MAT_LAMBERTIAN, capi.PTB_MAT_DIELECTRIC, capi.PTB_MAT_METAL,
                         capi.PTB_MAT_LAMBERTIAN]
    frac = np.bincount(kinds[4:], minlength=3) / (len(kinds) - 4)
    assert abs(frac[0] - 0.8) < 0.06 and abs(frac[1] - 0.15) < 0.05 and abs(frac[2] - 0.05) < 0.04
    ground_tex = t["textures"][t["materials"][t["sphere_material"][0]].texture]
    assert (ground_tex.kind, ground_tex.width, ground_tex.height) == (capi.PTB_TEX_CHECKER, 1000, 2000)
    # camera space: big spheres sit in front of the camera (negative z), ~13.9 away
    d = np.sqrt(t["xs"][2] ** 2 + t["ys"][2] ** 2 + t["zs"][2] ** 2)
    assert abs(d - np.sqrt(13 ** 2 + 1 ** 2 + 4.5 ** 2)) < 1e-9 and t["zs"][2] < 0
    # same seed, same scene; another seed, another scene
    t2 = P.shirley_spheres(600, 300).tables()
    assert np.array_equal(t["xs"], t2["xs"])
    assert not np.array_equal(t["xs"], P.shirley_spheres(600, 300, seed=7).tables()["xs"][:530])


def test_cornell_and_mesh_scene_tables():
    s = P.cornell_box(256, 256)
    t = s.tables()
    assert (t["n_spheres"], t["n_triangles"], t["n_vertices"]) == (3, 18, 54)
    assert list(t["prim_order"][-3:]) == [0, 1, 2] and sorted(~t["prim_order"][:18]) == list(range(18))
    assert abs(s.camera.view_y - 1.0) < 1e-12  # fov = 2 atan 0.5 -> half height 0.5
    m = P.synthetic_mesh_scene(2000, 160, 90)
    tm = m.tables()
    assert tm["n_spheres"] == 0 and tm["n_triangles"] >= 1500 + 2
    floor_mat = tm["materials"][tm["tri_material"][-1]]
    assert tm["textures"][floor_mat.texture].kind == capi.PTB_TEX_CHECKER
    ys = tm["vy"][tm["indices"][-6:]]
    assert np.allclose(ys, ys[0])  # floor is flat at the mesh's camera-space bbox-min y
    assert abs(ys[0] - tm["vy"][tm["indices"][:-6]].min()) < 1e-9


def test_setter_validation_and_round_trip():
    s = P.Scene()
    bad = capi.Texture(kind=capi.PTB_TEX_CHECKER, width=4, height=4, even=5, odd=0)
    with pytest.raises(P.PtbError, match="checker rows"):
        s.set_textures([bad])
    with pytest.raises(P.PtbError, match="unknown material kind"):
        s.set_materials([capi.Material(kind=9, texture=0, index=0.0)])
    with pytest.raises(P.PtbError, match="vertex index out of bounds"):  # ganesha/bin/main.ml:82-84
        s.set_triangles([0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [0.0, 0.0, 0.0], [0, 1, 3])
    with pytest.raises(P.PtbError, match="non-empty list of shapes"):  # shape_tree.ml:254-255
        s.commit(0)
    s.set_textures([capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.1, 0.2, 0.3))])
    s.set_materials([capi.Material(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=0.0)])
    s.set_spheres([1.0, 2.0], [3.0, 4.0], [5.0, 6.0], [0.5, 0.25], [0, 0])
    s.set_background(capi.PTB_BG_CONSTANT, (0.25, 0.5, 0.75))
    t = s.tables()
    assert list(t["xs"]) == [1.0, 2.0] and list(t["rs"]) == [0.5, 0.25] and t["bg_kind"] == capi.PTB_BG_CONSTANT
    assert list(t["bg0"]) == [0.25, 0.5, 0.75] and tuple(t["textures"][0].rgb) == (0.1, 0.2, 0.3)
    s.set_spheres([1.0], [1.0], [1.0], [1.0], [3])
    with pytest.raises(P.PtbError, match="material row out of range"):
        s.commit(0)


def test_compute_entry_points_fail_loudly_without_a_device():
    if L.ptb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    s = P.shirley_spheres(64, 32)
    with pytest.raises(P.PtbError, match="no CUDA device"):
        s.commit(0)
    p = make_params(s, 64, 32, 1, 8)
    img = np.zeros((32, 64, 3))
    st = capi.Stats()
    assert L.ptb_render(s.h, C.byref(p), capi.dptr(img), C.byref(st)) == -4  # PTB_E_STATE: not committed
    off = np.zeros(4, dtype=np.int32)
    assert L.ptb_r2_stream(8, capi.iptr(off), 4, capi.dptr(np.zeros(72)), 0) == -2  # PTB_E_NO_DEVICE


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from path_tracer_ocaml_b200.distributed import render_sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    W, H, spp, mb = 96, 64, 2, 4
    scene = P.shirley_spheres(W, H)
    osc = O.OracleScene(scene.tables())
    w = np.zeros(9)
    O.lib().orc_filter_binomial(5, 1, O.dptr(w))

    def render_sums(r, n):  # the oracle stands in for the device renderer on CPU
        img, _ = osc.render(make_params(scene, W, H, spp, mb, rank=r, world=n, flags=capi.PTB_FLAG_NO_FILTER))
        return torch.from_numpy(img)

    img = render_sharded(render_sums, lambda s: resolve_numpy(s.numpy(), spp, w.reshape(3, 3)))
    if rank == 0:
        whole, _ = osc.render(make_params(scene, W, H, spp, mb))
        q.put(float(np.abs(img - whole).max()))
    dist.destroy_process_group()


def _gloo_band_worker(rank, world, port, q):
    """The production exchange (distributed.render_sharded_bands): reduce-scatter by row bands, halo-row swap,
    per-rank resolve of the own band, gather — with the oracle as the per-rank renderer and numpy as the resolve."""
    import torch
    import torch.distributed as dist
    from path_tracer_ocaml_b200.distributed import alloc_padded_sums, band_rows, render_sharded_bands
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    W, H, spp, mb = 64, 37, 2, 4  # 37 rows: the last band is padded
    scene = P.shirley_spheres(W, H)
    osc = O.OracleScene(scene.tables())
    w = np.zeros(9)
    O.lib().orc_filter_binomial(5, 1, O.dptr(w))
    sums = alloc_padded_sums(H, W, world, "cpu", dtype=torch.float64)
    img, _ = osc.render(make_params(scene, W, H, spp, mb, rank=rank, world=world, flags=capi.PTB_FLAG_NO_FILTER))
    sums[:H] = torch.from_numpy(img)

    def resolve_rows(local):  # an image of local's height; the caller drops the two halo rows
        return torch.from_numpy(resolve_numpy(local.numpy(), spp, w.reshape(3, 3)))

    band, full = render_sharded_bands(sums, H, resolve_rows)
    assert band.shape == (band_rows(H, world), W, 3)
    if rank == 0:
        whole, _ = osc.render(make_params(scene, W, H, spp, mb))
        q.put(float(np.abs(full.numpy() - whole).max()))
    else:
        assert full is None
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_reduce_scatter_resolve_gather_over_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_gloo_band_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    err = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-12


def test_tile_sharding_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards are disjoint, so the reduce adds each pixel's sum to zeros: exact up to the filter's sum order
    assert err < 1e-12


def test_tile_shards_partition_the_image():
    cap = 20000
    a = [np.zeros(cap, dtype=np.int32) for _ in range(4)]
    n = L.ptb_tile_split(600, 300, 1024, *[capi.iptr(x) for x in a], cap)
    for world in (1, 2, 3, 4, 8):
        cover = np.zeros((300, 600), dtype=np.int32)
        per_rank = []
        for r in range(world):
            cnt = 0
            for t in range(r, n, world):
                row, col, w, h = (int(x[t]) for x in a)
                cover[row:row + h, col:col + w] += 1
                cnt += w * h
            per_rank.append(cnt)
        assert (cover == 1).all()
        assert max(per_rank) - min(per_rank) <= 2 * 1024  # interleaving balances the shards
