"""A/B of the ray sort of ptb_intersect_batch (ray_sort.cuh): every workload with PTB_BATCH_SORT=0 and =1 in one
process, results compared bit for bit.  usage: python scripts/sort_ab.py [quick]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi
from configs_bench import rays_for

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
rng = np.random.default_rng(0xB200)
lo, hi = np.full(3, -10.0), np.full(3, 10.0)


def scene(kind, m):
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
    s.commit(0)
    return s


def run(s, do, dd, n, mode):
    os.environ["PTB_BATCH_SORT"] = mode
    dt, dp = torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
    best = None
    for _ in range(3):
        st = capi.Stats()
        capi.check(P.lib().ptb_intersect_batch_device(s.h, C.c_void_p(do.data_ptr()), C.c_void_p(dd.data_ptr()), 0.0, 3.0e38, n,
                                                      C.c_void_p(dt.data_ptr()), C.c_void_p(dp.data_ptr()), 0, None, C.byref(st)))
        best = st.ms_device if best is None else min(best, st.ms_device)
    return best, dt, dp


workloads = [("triangles", 1_000_000), ("triangles", 100_000), ("triangles", 10_000), ("spheres", 4096)]
if not quick:
    workloads += [("triangles", 10_000_000), ("spheres", 1024)]
for kind, m in workloads:
    s = scene(kind, m)
    for coherent in (False, True):
        for n in ((1 << 22,) if quick or m >= 10_000_000 else (1 << 22, 1 << 24)):
            o, d = rays_for(rng, n, lo, hi, coherent)
            do, dd = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
            ms0, t0, p0 = run(s, do, dd, n, "0")
            ms1, t1, p1 = run(s, do, dd, n, "1")
            same = bool(torch.equal(p0, p1)) and bool(torch.equal(t0.view(torch.int32), t1.view(torch.int32)))
            print(f"{kind} m={m} {'coherent' if coherent else 'incoherent'} n=2^{n.bit_length() - 1}: unsorted {ms0:.3f} ms "
                  f"({n / ms0 / 1e6:.3f} Grays/s), sorted {ms1:.3f} ms ({n / ms1 / 1e6:.3f} Grays/s), x{ms0 / ms1:.2f}, identical {same}",
                  flush=True)
    del s
os.environ.pop("PTB_BATCH_SORT", None)
