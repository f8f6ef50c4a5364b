#!/usr/bin/env python
"""(Lives under tests/ because it uses the oracle as the checker and as the CPU side of the sweep.)
Measures BASELINE.json configs[0..2] (C1 shirley 600x300x32, C2 cornell geometry 1024^2x256x16,
C3 synthetic ganesha mesh 1920x1080x256) on one GPU: device throughput at the full size plus image parity
against the oracle at a reduced size of the same scene; and configs[4] (C5), the intersect_batch sweep
(rays x spheres / triangles, coherent and incoherent rays) against the oracle's AVX2 leaf kernel / scalar
Moller-Trumbore on the host.  Writes one JSON document to stdout (progress on stderr).

  python tests/configs_bench.py [--quick] > gpurun_out/configs.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import path_tracer_ocaml_b200 as P  # noqa: E402
from path_tracer_ocaml_b200 import capi  # noqa: E402
from path_tracer_ocaml_b200.integrator import intersect_batch  # noqa: E402
import pyoracle as O  # noqa: E402


def log(*a):
    print(*a, file=sys.stderr, flush=True)


C_BOX, C_SPH, C_TRI, C_HIT, C_FILM = 25.0, 34.0, 46.0, 200.0, 27.0  # algorithmic lane-ops per unit
_PEAK = {}


def fp32_peak():
    """Measured FFMA issue rate of this GPU in lane-op/s (the roofline denominator, as in bench.py)."""
    if "v" not in _PEAK:
        import ctypes as C
        v = C.c_double(0)
        capi.check(P.lib().ptb_fp32_peak(0, C.byref(v), None))
        _PEAK["v"] = v.value * 1e12
    return _PEAK["v"]


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
    except Exception:
        return 6650e9


def bytes_per_ray(*keys):
    """ncu-measured DRAM bytes per ray of the traversal kernel (profiles/r2_triangle_bytes_per_ray.json), mean of keys."""
    try:
        e = json.load(open(os.path.join(ROOT, "profiles", "r2_triangle_bytes_per_ray.json")))["entries"]
        return float(np.mean([e[k]["dram_bytes_per_ray"] for k in keys]))
    except Exception:
        return None


# Lane-divergent loads (every lane of a warp at its own node / primitive record of a scene in global memory) go through
# the SM's L1 data pipe at ONE lane per cycle for 4-16 bytes and 0.75 lanes per cycle for 32 bytes, texture path
# included (scripts/microbench/gather.cu, profiles/r2_gather_microbench_48MB.log): at most 24 B per cycle per SM
# = 6.7 TB/s of useful record bytes on 148 SMs at 1.9 GHz, well below what L2 can deliver.
L1_GATHER_PEAK = 6.71e12
B_BOX, B_SPH, B_TRI = 16.0, 16.0, 36.0  # record bytes: a quarter of a 64-byte quantised node, (c, r), 9 floats


def roofline(rays_per_s, cn, b_dram, gather=False):
    """bound = min(P_issue / A_ray, BW_hbm / B_ray); A_ray from the oracle's reference-faithful counts of
    the same rays (box x 25 + sphere x 34 + triangle x 46 lane-ops), B_ray = DRAM bytes per ray measured with ncu.
    gather: the scene is traversed out of global memory by divergent lanes, which adds the L1 gather bound
    L1_GATHER_PEAK / (algorithmic record bytes per ray)."""
    rays = max(int(cn.rays), 1)
    a = (cn.box_tests * C_BOX + cn.sphere_tests * C_SPH + cn.tri_tests * C_TRI) / rays
    issue = fp32_peak() / a if a > 0 else None
    hbm = hbm_peak() / b_dram if b_dram else None
    b_rec = (cn.box_tests * B_BOX + cn.sphere_tests * B_SPH + cn.tri_tests * B_TRI) / rays
    l1 = L1_GATHER_PEAK / b_rec if gather and b_rec > 0 else None
    bounds = {"fp32_issue": issue, "hbm": hbm, "l1_gather": l1}
    name = min((k for k in bounds if bounds[k] is not None), key=lambda k: bounds[k])
    bound = bounds[name]
    return {"a_ray_lane_ops": a, "box_per_ray": cn.box_tests / rays, "sphere_per_ray": cn.sphere_tests / rays,
            "tri_per_ray": cn.tri_tests / rays, "b_ray_dram_bytes": b_dram, "record_bytes_per_ray": b_rec,
            "issue_bound_rays_per_s": issue, "hbm_bound_rays_per_s": hbm, "l1_gather_bound_rays_per_s": l1,
            "bound": name, "bound_rays_per_s": bound, "achieved_rays_per_s": rays_per_s, "frac": rays_per_s / bound}


def pinned_like(a):
    b = capi.pinned_empty(a.shape, a.dtype)
    b[...] = a
    return b


def render_config(name, make_scene, W, H, spp, mb, small, reps=3, bkeys=None):
    scene = make_scene(W, H)
    integ = P.Integrator(scene, W, H, spp, mb, device=0)
    tree = scene.tree_stats()
    best = None
    for _ in range(reps):
        integ.render(flags=capi.PTB_FLAG_PROFILE)
        st = integ.stats
        if best is None or st.ms_device < best["ms_device"]:
            best = {"ms_device": st.ms_device, "ms_trace": st.ms_trace, "ms_total": st.ms_total, "paths": int(st.paths),
                    "rays": int(st.rays)}
    out = {"config": name, "W": W, "H": H, "spp": spp, "max_bounces": mb, "tree": tree, "commit_ms": scene.commit_ms,
           "mpaths_per_s": best["paths"] / best["ms_device"] / 1e3, "mrays_per_s": best["rays"] / best["ms_device"] / 1e3,
           "rays_per_path": best["rays"] / best["paths"], **best}
    # parity at a reduced size of the SAME scene against the oracle (float64 CPU restatement)
    w, h, s = small
    sc2 = make_scene(w, h)
    i2 = P.Integrator(sc2, w, h, s, mb, device=0)
    img = i2.render()
    t0 = time.perf_counter()
    ref, cn = O.OracleScene(sc2.tables()).render(i2.params, n_threads=os.cpu_count())
    dt = time.perf_counter() - t0
    d = img - ref
    out["parity"] = {"W": w, "H": h, "spp": s, "rmse": float(np.sqrt(np.mean(d * d))),
                     "within_tol": float(np.mean(np.abs(d) <= 0.02 * np.abs(ref) + 1 / 255)),
                     "bias": float(np.abs(d.mean((0, 1))).max()), "rays_device": int(i2.stats.rays), "rays_oracle": int(cn.rays),
                     "oracle_mpaths_per_s": cn.paths / dt / 1e6, "oracle_threads": os.cpu_count()}
    # trace kernel alone against its bound (the oracle's per-ray test counts of the reduced-size render of the scene)
    out["roofline_trace"] = roofline(best["rays"] / (best["ms_trace"] * 1e-3), cn, bytes_per_ray(*bkeys) if bkeys else None)
    log(json.dumps(out))
    return out


def rays_for(rng, n, lo, hi, coherent):
    if coherent:  # origins on a sphere of radius 3x the scene extent, aimed at points in the scene box
        c, ext = 0.5 * (lo + hi), 0.5 * float(np.max(hi - lo))
        v = rng.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        o = c + 3.0 * ext * v[0] + 0.05 * ext * rng.normal(size=(n, 3))  # one viewpoint, jittered
        tgt = rng.uniform(lo, hi, size=(n, 3))
        d = tgt - o
    else:
        o = rng.uniform(lo, hi, size=(n, 3))
        d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.ascontiguousarray(o, dtype=np.float32), np.ascontiguousarray(d, dtype=np.float32)


def sweep(quick):
    """C5: rays N x primitives M.  Spheres: uniform centres in [-10,10]^3, r in U[0.05,0.5]; triangles: seeded
    soup of the same extent (edge ~0.3).  Seeds fixed (0xB200)."""
    rng = np.random.default_rng(0xB200)
    res = []
    n_list = [1 << 20, 1 << 24] if quick else [1 << 20, 1 << 22, 1 << 24, 1 << 26]
    scenes = [("spheres", m) for m in (4, 16, 64, 256, 1024, 4096)] + [("triangles", t) for t in
                                                                      ((10_000, 100_000) if quick else (10_000, 100_000, 1_000_000, 10_000_000))]
    lo, hi = np.full(3, -10.0), np.full(3, 10.0)
    for kind, m in scenes:
        s = P.Scene()
        s.set_textures([capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.5, 0.5, 0.5))])
        s.set_materials([capi.Material(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=1.0)])
        if kind == "spheres":
            c = rng.uniform(-10, 10, size=(m, 3))
            s.set_spheres(c[:, 0], c[:, 1], c[:, 2], rng.uniform(0.05, 0.5, size=m))
        else:
            a = rng.uniform(-10, 10, size=(m, 3))
            v = np.concatenate([a, a + rng.normal(scale=0.3, size=(m, 3)), a + rng.normal(scale=0.3, size=(m, 3))])
            idx = np.stack([np.arange(m), np.arange(m) + m, np.arange(m) + 2 * m], axis=1).astype(np.int32)
            s.set_triangles(v[:, 0], v[:, 1], v[:, 2], idx)
        s.set_background(capi.PTB_BG_CONSTANT, (1.0, 1.0, 1.0))
        t0 = time.perf_counter()
        s.commit(0)
        commit_ms = (time.perf_counter() - t0) * 1e3
        # the reference's tree over 10^7 triangles is out of reach for the CPU side of a sweep: GPU numbers only there
        osc = O.OracleScene(s.tables()) if m <= 1_000_000 else None
        import torch
        import ctypes as C
        for coherent in (True, False):
            ns_list = n_list if m <= 1_000_000 else [1 << 22]
            if kind == "spheres" and m in (64, 1024) and not quick:
                ns_list = n_list + [1 << 28]  # BASELINE's upper end on the ray axis
            for n in ns_list:
                base = min(n, 1 << 24)  # 2^28 rays = 2^24 distinct rays, 16 times over (generation time, not the device, bounds this)
                o0, d0 = rays_for(rng, base, lo, hi, coherent)
                o, d = capi.pinned_empty((n, 3), np.float32), capi.pinned_empty((n, 3), np.float32)
                for k in range(n // base):
                    o[k * base:(k + 1) * base], d[k * base:(k + 1) * base] = o0, d0
                t, p = capi.pinned_empty(n, np.float32), capi.pinned_empty(n, np.int32)
                # (a) through the C ABI with page-locked HOST buffers: chunk-pipelined copies + kernels, wall clock
                host_ms = None
                for _ in range(3):
                    st = capi.Stats()
                    capi.check(P.lib().ptb_intersect_batch(s.h, capi.fptr(o), capi.fptr(d), 0.0, 3.0e38, n, capi.fptr(t), capi.iptr(p), 0, C.byref(st)))
                    host_ms = st.ms_total if host_ms is None else min(host_ms, st.ms_total)
                # (b) device-resident: rays and results already in HBM
                m_dev = min(n, 1 << 26)
                do, dd = torch.from_numpy(o[:m_dev]).cuda(), torch.from_numpy(d[:m_dev]).cuda()
                dt_, dp_ = torch.empty(m_dev, device="cuda"), torch.empty(m_dev, dtype=torch.int32, device="cuda")
                dev_ms = None
                for _ in range(3):
                    st = capi.Stats()
                    capi.check(P.lib().ptb_intersect_batch_device(s.h, C.c_void_p(do.data_ptr()), C.c_void_p(dd.data_ptr()), 0.0, 3.0e38, m_dev,
                                                                  C.c_void_p(dt_.data_ptr()), C.c_void_p(dp_.data_ptr()), 0, None, C.byref(st)))
                    dev_ms = st.ms_device if dev_ms is None else min(dev_ms, st.ms_device)
                del do, dd, dt_, dp_
                row = {"prims": kind, "m": m, "rays": n, "coherent": coherent, "commit_ms": commit_ms,
                       "gpu_mrays_per_s_device": m_dev / dev_ms / 1e3, "device_rays": m_dev,
                       "gpu_mrays_per_s_host_buffers": n / host_ms / 1e3, "host_buffers": "page-locked (ptb_host_alloc), chunk-pipelined",
                       "hit_fraction": float(np.mean(p[:base] >= 0)), "tree": s.tree_stats()}
                if kind == "triangles" and row["tree"]["triangles"] > m:
                    row["tree_note"] = ("pre-split soup (csrc/presplit.hpp): the tree holds several box references per triangle and the rays "
                                        "are traced in spatial order (csrc/ray_sort.cuh, inside the timed region); the roofline's work per ray is "
                                        "counted on the reference's tree, one box per triangle, which this tree undercuts")
                if osc is not None:
                    # CPU side on a bounded sample: the oracle's reference-faithful tree + leaf kernels
                    ns = min(n, 1 << 18)
                    t0 = time.perf_counter()
                    t_ref, p_ref, cn = osc.intersect_batch(o[:ns], d[:ns], 0.0, 3.0e38, n_threads=os.cpu_count())
                    dt = time.perf_counter() - t0
                    row["cpu_mrays_per_s"] = ns / dt / 1e6
                    row["cpu_sample_rays"] = ns
                    row["cpu_threads"] = os.cpu_count()
                    row["prim_agreement"] = float(np.mean(p[:ns] == p_ref))
                    hit = (p[:ns] == p_ref) & (p_ref >= 0)
                    row["t_rel_err_p99"] = float(np.quantile(np.abs(t[:ns][hit] - t_ref[hit]) / t_ref[hit], 0.99)) if hit.any() else 0.0
                    if kind == "triangles":
                        class _C:  # per-ray counts of the sample, in the shape roofline() reads
                            rays, box_tests, sphere_tests, tri_tests = ns, cn.box_tests, cn.sphere_tests, cn.tri_tests
                        key = {(1_000_000, False): "soup_1m_incoherent", (1_000_000, True): "soup_1m_coherent"}.get((m, coherent))
                        # incoherent rays: every lane at its own record (coherent lanes share cache lines: the gather
                        # bound does not apply to them as stated)
                        row["roofline"] = roofline(m_dev / (dev_ms * 1e-3), _C, bytes_per_ray(key) if key else None, gather=not coherent)
                else:
                    row["cpu_note"] = "no CPU side: the reference tree over 10^7 triangles is not built in a sweep (tests/test_baseline_sizes.py checks this size against an Array_leaf scan and across device builders)"
                log(json.dumps(row))
                res.append(row)
                del o, d, t, p
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--skip-sweep", action="store_true")
    ap.add_argument("--mesh-faces", type=int, default=1_000_000)
    a = ap.parse_args()
    if P.lib().ptb_device_count() == 0:
        raise SystemExit("configs_bench.py: no CUDA device; libptb200 has no CPU path")
    doc = {"configs": [], "sweep": []}
    white = ("constant", (1.0, 1.0, 1.0), None)
    doc["configs"].append(render_config("C1 shirley_spheres 600x300 32spp 8b", lambda w, h: P.shirley_spheres(w, h), 600, 300, 32, 8,
                                        (600, 300, 32)))
    doc["configs"].append(render_config("C2 cornell geometry 1024x1024 256spp 16b, constant white background (SURVEY D1)",
                                        lambda w, h: P.cornell_box(w, h, white), 1024, 1024, 64 if a.quick else 256, 16,
                                        (256, 256, 16)))
    nf = 100_000 if a.quick else a.mesh_faces
    doc["configs"].append(render_config(f"C3 synthetic ganesha mesh ({nf} faces) 1920x1080 256spp 8b",
                                        lambda w, h: P.synthetic_mesh_scene(nf, w, h), 1920, 1080, 32 if a.quick else 256, 8,
                                        (320, 180, 16), bkeys=("c3_1m_bounce0", "c3_1m_bounce1")))
    if not a.quick:
        doc["configs"].append(render_config("C3 synthetic ganesha mesh (10000000 faces) 1920x1080 256spp 8b",
                                            lambda w, h: P.synthetic_mesh_scene(10_000_000 if w == 1920 else 1_000_000, w, h), 1920,
                                            1080, 256, 8, (320, 180, 16), bkeys=("c3_10m_bounce0", "c3_10m_bounce1")))
    doc["configs"].append(render_config("C2 as written: cornell geometry + emissive square + light sampling (extension), 1024x1024 256spp 16b",
                                        lambda w, h: P.cornell_box_lit(w, h), 1024, 1024, 64 if a.quick else 256, 16, (256, 256, 16)))
    if not a.skip_sweep:
        doc["sweep"] = sweep(a.quick)
    print(json.dumps(doc))


if __name__ == "__main__":
    main()
