"""One intersect_batch workload of the C5 sweep for profiling: `m` random triangles (seeded soup in [-10,10]^3, edge
~0.3) or the C3 mesh, 2^22 incoherent or coherent rays, device-resident (ptb_intersect_batch_device).
usage: python scripts/soup_probe.py [m] [incoherent|coherent] [soup|mesh]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ctypes as C
import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi
from configs_bench import rays_for

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
coherent = len(sys.argv) > 2 and sys.argv[2] == "coherent"
kind = sys.argv[3] if len(sys.argv) > 3 else "soup"
rng = np.random.default_rng(0xB200)
if kind == "soup":
    s = P.Scene()
    s.set_textures([capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.5, 0.5, 0.5))])
    s.set_materials([capi.Material(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=1.0)])
    a = rng.uniform(-10, 10, size=(m, 3))
    v = np.concatenate([a, a + rng.normal(scale=0.3, size=(m, 3)), a + rng.normal(scale=0.3, size=(m, 3))])
    idx = np.stack([np.arange(m), np.arange(m) + m, np.arange(m) + 2 * m], axis=1).astype(np.int32)
    s.set_triangles(v[:, 0], v[:, 1], v[:, 2], idx)
    s.set_background(capi.PTB_BG_CONSTANT, (1.0, 1.0, 1.0))
    lo, hi = np.full(3, -10.0), np.full(3, 10.0)
else:
    s = P.synthetic_mesh_scene(m, 64, 36)
    t = s.tables()
    lo = np.array([t["vx"][:-4].min(), t["vy"][:-4].min(), t["vz"][:-4].min()])
    hi = np.array([t["vx"][:-4].max(), t["vy"][:-4].max(), t["vz"][:-4].max()])
s.commit(0)
print("tree", s.tree_stats(), "commit ms", s.commit_ms)
n = 1 << 22
o, d = rays_for(rng, n, lo, hi, coherent)
do, dd = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
dt, dp = torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
for rep in range(3):
    st = capi.Stats()
    capi.check(P.lib().ptb_intersect_batch_device(s.h, C.c_void_p(do.data_ptr()), C.c_void_p(dd.data_ptr()), 0.0, 3.0e38, n,
                                                  C.c_void_p(dt.data_ptr()), C.c_void_p(dp.data_ptr()), 0, None, C.byref(st)))
    print(f"{kind} m={m} {'coherent' if coherent else 'incoherent'}: {st.ms_device:.3f} ms, {n / st.ms_device / 1e6:.3f} Grays/s, hit {float((dp >= 0).float().mean()):.3f}")
