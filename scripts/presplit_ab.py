"""Triangle pre-splitting (bvh.cpp, PTB_BVH_PRESPLIT) x ray sort (PTB_BATCH_SORT) on the C5 triangle soups, device-resident
ptb_intersect_batch_device, results compared with the plain tree's.  usage: python scripts/presplit_ab.py [m ...]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi
from configs_bench import rays_for

ms_list = [int(a) for a in sys.argv[1:]] or [1_000_000]
lo, hi = np.full(3, -10.0), np.full(3, 10.0)
n = 1 << 22
for m in ms_list:
    rng = np.random.default_rng(0xB200)
    a = rng.uniform(-10, 10, size=(m, 3))
    v = np.concatenate([a, a + rng.normal(scale=0.3, size=(m, 3)), a + rng.normal(scale=0.3, size=(m, 3))])
    idx = np.stack([np.arange(m), np.arange(m) + m, np.arange(m) + 2 * m], axis=1).astype(np.int32)
    rays = {c: rays_for(rng, n, lo, hi, c) for c in (False, True)}
    ref = {}
    for builder, pre in (("default", "0"), ("default", None)):
        os.environ.pop("PTB_BUILDER", None), os.environ.pop("PTB_BVH_PRESPLIT", None), os.environ.pop("PTB_BATCH_SORT", None)
        if builder == "host":
            os.environ["PTB_BUILDER"] = "host"
        if pre:
            os.environ["PTB_BVH_PRESPLIT"] = pre
        s = P.Scene()
        s.set_textures([capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.5, 0.5, 0.5))])
        s.set_materials([capi.Material(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=1.0)])
        s.set_triangles(v[:, 0], v[:, 1], v[:, 2], idx)
        s.set_background(capi.PTB_BG_CONSTANT, (1.0, 1.0, 1.0))
        t0 = time.perf_counter()
        s.commit(0)
        commit_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        s.commit(0)
        commit_ms = min(commit_ms, (time.perf_counter() - t0) * 1e3)
        print(f"m={m} builder={builder} presplit={pre}: commit {commit_ms:.0f} ms, tree {s.tree_stats()}", flush=True)
        for coherent in (False, True):
            o, d = rays[coherent]
            do, dd = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
            for sort in ("0", "1", "oct0", "auto"):
                os.environ.pop("PTB_BATCH_SORT", None), os.environ.pop("PTB_SORT_OCT", None)
                if sort == "oct0":
                    os.environ["PTB_BATCH_SORT"], os.environ["PTB_SORT_OCT"] = "1", "0"
                elif sort != "auto":
                    os.environ["PTB_BATCH_SORT"] = sort
                dt, dp = torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
                best = None
                for _ in range(3):
                    st = capi.Stats()
                    capi.check(P.lib().ptb_intersect_batch_device(s.h, C.c_void_p(do.data_ptr()), C.c_void_p(dd.data_ptr()), 0.0, 3.0e38, n,
                                                                  C.c_void_p(dt.data_ptr()), C.c_void_p(dp.data_ptr()), 0, None, C.byref(st)))
                    best = st.ms_device if best is None else min(best, st.ms_device)
                key = coherent
                if key not in ref:
                    ref[key] = (dt.clone(), dp.clone())
                agree = float((dp == ref[key][1]).float().mean())
                tsame = float(((dt == ref[key][0]) | (torch.isnan(dt) & torch.isnan(ref[key][0]))).float().mean())
                print(f"   {'coherent' if coherent else 'incoherent'} sort={sort}: {best:.3f} ms, {n / best / 1e6:.3f} Grays/s, "
                      f"prim agreement with the first tree {agree:.6f}, t equal {tsame:.6f}", flush=True)
        del s
