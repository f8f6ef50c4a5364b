"""Host-interface behaviour that needs a device (`-m gpu`): the polled progress counter, the per-device lock that makes
the library safe to call from several host threads (the OCaml stubs release the runtime lock, ctypes drops the GIL),
re-commits, and the host-buffer paths of ptb_intersect_batch (page-locked, pageable + pinned for the call, staged)."""
import ctypes as C
import os
import threading
import time

import numpy as np
import pytest

import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi, integrator
from helpers import image_metrics

pytestmark = pytest.mark.gpu


def test_render_progress_counter_is_monotonic_and_ends_at_the_total(monkeypatch):
    """`update_progress` (integrator.ml:130,150; render_command.ml:86-104) as a poll from another thread."""
    monkeypatch.setenv("PTB_BATCH", str(1 << 16))  # many wavefront batches -> many counter updates
    W, H, spp = 320, 200, 64
    integ = P.Integrator(P.shirley_spheres(W, H), W, H, spp, 8)
    seen = []
    stop = threading.Event()

    def poll():
        done, total = C.c_uint64(), C.c_uint64()
        while not stop.is_set():
            capi.check(P.lib().ptb_render_progress(0, C.byref(done), C.byref(total)))
            seen.append((done.value, total.value))
            time.sleep(0.0005)

    th = threading.Thread(target=poll)
    th.start()
    integ.render()
    stop.set()
    th.join()
    done, total = C.c_uint64(), C.c_uint64()
    capi.check(P.lib().ptb_render_progress(0, C.byref(done), C.byref(total)))
    assert done.value == total.value == W * H * spp == integ.stats.paths
    during = [d for d, t in seen if t == W * H * spp]
    assert during == sorted(during) and len(set(during)) >= 3, sorted(set(during))[:10]  # advances batch by batch
    assert all(d % (1 << 16) == 0 or d == W * H * spp for d in during)


def test_two_host_threads_share_one_device():
    """Two threads render different scenes on the same GPU at the same time, repeatedly: the per-device lock
    serialises them, so each gets exactly the image it gets alone (up to the order of the float atomics)."""
    jobs = [(P.shirley_spheres(160, 90), 160, 90, 8, 8), (P.cornell_box(96, 96), 96, 96, 8, 16),
            (P.synthetic_mesh_scene(5000, 128, 72), 128, 72, 4, 8)]
    alone = [P.Integrator(sc, w, h, spp, mb).render() for sc, w, h, spp, mb in jobs]
    errs, results = [], {}

    def work(k):
        try:
            sc, w, h, spp, mb = jobs[k]
            integ = P.Integrator(sc, w, h, spp, mb)
            for rep in range(6):
                results[(k, rep)] = integ.render()
                if k == 0 and rep % 2 == 1:
                    sc.commit(0)  # re-commits interleave with the other threads' renders
        except Exception as e:  # pragma: no cover
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(jobs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for (k, rep), img in results.items():
        m = image_metrics(img, alone[k])
        assert m["rmse"] < 1e-5 and m["max"] < 1e-3, (k, rep, m)


def test_recommit_is_cheap_and_keeps_the_scene_usable():
    sc = P.shirley_spheres(200, 100)
    integ = P.Integrator(sc, 200, 100, 4, 8)
    a = integ.render()
    ms = [sc.commit(0) for _ in range(20)]
    b = integ.render()
    assert image_metrics(a, b)["max"] < 1e-3
    assert np.median(ms) < 5.0, ms  # BVH build of 530 spheres + one 54 KB copy (was 3.5 - 230 ms with 12 blocking copies)
    # a scene that is modified after the commit must be committed again
    sc.set_background(capi.PTB_BG_CONSTANT, (1, 1, 1))
    p = integ._p(0)
    rc = P.lib().ptb_render(sc.h, C.byref(p), capi.dptr(np.zeros((100, 200, 3))), None)
    assert rc == -4 and b"not committed" in P.lib().ptb_last_error()
    sc.commit(0)
    assert integ.render().mean() > a.mean()  # white sky is brighter


def _rays(n, seed=3):
    rng = np.random.default_rng(seed)
    o = rng.uniform((-12, -3, -28), (12, 4, -4), size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


def test_intersect_batch_host_buffer_paths_agree(monkeypatch):
    """Chunked + pipelined copies (2 streams) from page-locked buffers, from pageable buffers pinned for the call, and
    the staged copy with pinning switched off: bit-identical answers, equal to the one-chunk call."""
    scene = P.shirley_spheres(600, 300)
    n = 3_000_017  # not a multiple of anything
    o, d = _rays(n)
    monkeypatch.setenv("PTB_BATCH_CHUNK", str(1 << 24))
    t_one, p_one = integrator.intersect_batch(scene, o, d)  # a single chunk
    monkeypatch.setenv("PTB_BATCH_CHUNK", str(1 << 19))     # 6 chunks, the last ragged
    t_pg, p_pg = integrator.intersect_batch(scene, o, d)    # pageable numpy arrays: pinned for the call
    monkeypatch.setenv("PTB_BATCH_PIN", "0")
    t_st, p_st = integrator.intersect_batch(scene, o, d)    # staged by the driver
    monkeypatch.delenv("PTB_BATCH_PIN")
    po, pd = capi.pinned_empty((n, 3), np.float32), capi.pinned_empty((n, 3), np.float32)
    pt, pp = capi.pinned_empty(n, np.float32), capi.pinned_empty(n, np.int32)
    po[:], pd[:] = o, d
    capi.check(P.lib().ptb_intersect_batch(scene.h, capi.fptr(po), capi.fptr(pd), 0.0, 3.0e38, n, capi.fptr(pt),
                                           capi.iptr(pp), 0, None))
    for t, p in ((t_pg, p_pg), (t_st, p_st), (pt, pp)):
        assert np.array_equal(p, p_one)
        assert np.array_equal(np.asarray(t).view(np.uint32), t_one.view(np.uint32))
    assert (p_one >= 0).mean() > 0.3


def test_single_process_multi_gpu_odd_frame_sizes():
    """k_reduce_peers covers frames whose float count is not a multiple of 4 (19x18 -> 1026 floats, 41x25 -> 3075)."""
    n = min(P.lib().ptb_device_count(), 2)
    if n < 2:
        pytest.skip("needs two GPUs")
    for W, H in ((19, 18), (41, 25), (5, 1)):
        ref = P.Integrator(P.shirley_spheres(W, H), W, H, 8, 8).render()
        img = P.Integrator(P.shirley_spheres(W, H), W, H, 8, 8).render_multi(n)
        assert image_metrics(img, ref)["max"] < 1e-4, (W, H)


def test_pass_ranges_add_up_to_the_full_render_and_match_the_oracle():
    """ptb_params.pass_first / pass_count: consecutive ranges of sample passes are exactly the full render's samples
    (same R2 offsets pixel + pass * spp), so their per-pixel sums add up to the full sums and their ray counts to the
    full count; a leading range equals the oracle's render of its first passes."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import pyoracle as O
    W, H, spp, mb = 96, 48, 8, 6
    scene = P.shirley_spheres(W, H)
    integ = P.Integrator(scene, W, H, spp, mb)
    full = integ.render(flags=capi.PTB_FLAG_NO_FILTER).copy()
    full_rays, full_paths = int(integ.stats.rays), int(integ.stats.paths)
    parts, rays, paths = [], 0, 0
    for rng in ((0, 3), (3, 4), (7, 0)):  # (7, 0): count 0 = the rest
        parts.append(integ.render(flags=capi.PTB_FLAG_NO_FILTER, passes=rng).copy())
        rays, paths = rays + int(integ.stats.rays), paths + int(integ.stats.paths)
    assert rays == full_rays and paths == full_paths == W * H * spp
    assert np.abs(sum(parts) - full).max() <= 1e-4 * max(1.0, full.max())
    orc = O.OracleScene(scene.tables())
    ref3, cn3 = orc.render(integ.params, n_threads=4, pass_limit=3, flags=capi.PTB_FLAG_NO_FILTER)
    ref7, cn7 = orc.render(integ.params, n_threads=4, pass_limit=7, flags=capi.PTB_FLAG_NO_FILTER)
    assert np.sqrt(np.mean((parts[0] - ref3) ** 2)) < 0.02 and abs(parts[0].mean() - ref3.mean()) < 2e-3
    assert np.sqrt(np.mean((parts[1] - (ref7 - ref3)) ** 2)) < 0.03
    # a partial render to an image is normalised by its own passes: a preview, not a darker frame
    prev = integ.render(passes=(0, 2))
    assert abs(prev.mean() - integ.render().mean()) < 0.02
    # float64 mode: the first three passes take exactly the oracle's decisions
    integ.render(flags=capi.PTB_FLAG_F64 | capi.PTB_FLAG_NO_FILTER, passes=(0, 3))
    assert list(integ.stats.rays_by_bounce[:mb]) == list(cn3.rays_by_bounce[:mb])
    for bad in ((8, 0), (-1, 2), (5, 4)):
        with pytest.raises(capi.PtbError):
            integ.render(passes=bad)


def test_progressive_render_equals_the_one_call_render():
    W, H, spp, mb = 120, 60, 16, 8
    integ = P.Integrator(P.shirley_spheres(W, H), W, H, spp, mb)
    whole = integ.render().copy()
    whole_rays = int(integ.stats.rays)
    seen = []
    img = integ.render_progressive(5, on_preview=lambda im, done: seen.append((done, float(im.mean()))))
    assert [d for d, _ in seen] == [5, 10, 15, 16]
    assert int(integ.stats.rays) == whole_rays
    assert np.abs(img - whole).max() < 1e-4
    # every preview is a full-brightness image of the passes so far
    assert all(abs(m - seen[-1][1]) < 0.02 for _, m in seen)
