"""intersect_batch on sphere sets (4096: global memory; 1024, 64: shared memory) for the current PTB_REFILL.
usage: PTB_REFILL=r python scripts/refill_probe.py"""
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

rng = np.random.default_rng(0xB200)
lo, hi = np.full(3, -10.0), np.full(3, 10.0)
n = 1 << 24
for m in (4096, 1024, 64):
    s = P.Scene()
    s.set_textures([capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.5, 0.5, 0.5))])
    s.set_materials([capi.Material(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=1.0)])
    c = rng.uniform(-10, 10, size=(m, 3))
    s.set_spheres(c[:, 0], c[:, 1], c[:, 2], rng.uniform(0.05, 0.5, size=m))
    s.set_background(capi.PTB_BG_CONSTANT, (1.0, 1.0, 1.0))
    s.commit(0)
    for coherent in (False, True):
        o, d = rays_for(rng, n, lo, hi, coherent)
        do, dd = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
        dt, dp = torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
        best = None
        for _ in range(3):
            st = capi.Stats()
            capi.check(P.lib().ptb_intersect_batch_device(s.h, C.c_void_p(do.data_ptr()), C.c_void_p(dd.data_ptr()), 0.0, 3.0e38, n,
                                                          C.c_void_p(dt.data_ptr()), C.c_void_p(dp.data_ptr()), 0, None, C.byref(st)))
            best = st.ms_device if best is None else min(best, st.ms_device)
        print(f"refill {os.environ.get('PTB_REFILL', 'default')}: {m} spheres {'coherent' if coherent else 'incoherent'}: {n / best / 1e6:.2f} Grays/s", flush=True)
