"""Render of a triangle soup (m random triangles of edge ~0.3 in a cube in front of the camera) with and without triangle
pre-splitting: device time, rays, image difference.  usage: python scripts/soup_render_probe.py [m] [spp]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi
from path_tracer_ocaml_b200.scenes import Camera

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
W, H = 1920, 1080
imgs = {}
for pre in ("0", None):
    os.environ.pop("PTB_BVH_PRESPLIT", None)
    if pre:
        os.environ["PTB_BVH_PRESPLIT"] = pre
    rng = np.random.default_rng(0xB200)
    s = P.Scene()
    s.set_textures([capi.Texture(kind=capi.PTB_TEX_SOLID, rgb=(0.7, 0.6, 0.5))])
    s.set_materials([capi.Material(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=1.0)])
    a = rng.uniform(-10, 10, size=(m, 3)) + np.array([0.0, 0.0, -32.0])
    v = np.concatenate([a, a + rng.normal(scale=0.3, size=(m, 3)), a + rng.normal(scale=0.3, size=(m, 3))])
    idx = np.stack([np.arange(m), np.arange(m) + m, np.arange(m) + 2 * m], axis=1).astype(np.int32)
    s.set_triangles(v[:, 0], v[:, 1], v[:, 2], idx)
    s.set_background(capi.PTB_BG_GRADIENT_Y, (1.0, 1.0, 1.0), (0.5, 0.7, 1.0))
    s.camera = Camera.create(eye=(0.0, 0.0, 0.0), target=(0.0, 0.0, -1.0), up=(0.0, 1.0, 0.0), aspect=W / H, vertical_fov_deg=40.0)
    integ = P.Integrator(s, W, H, spp, 8)
    best = None
    for _ in range(3):
        img = integ.render(flags=capi.PTB_FLAG_PROFILE)
        st = integ.stats
        if best is None or st.ms_device < best[0]:
            best = (st.ms_device, st.ms_trace, int(st.rays), int(st.paths))
    imgs[pre] = img
    print(f"m={m} presplit={'off' if pre else 'auto'}: tree {s.tree_stats()}, device {best[0]:.2f} ms (trace {best[1]:.2f}), "
          f"{best[3] / best[0] / 1e3:.0f} Mpaths/s, {best[2] / best[0] / 1e6:.2f} Grays/s, rays {best[2]}", flush=True)
d = imgs["0"] - imgs[None]
print(f"image difference between the two trees: rmse {float(np.sqrt(np.mean(d * d))):.2e}, max {float(np.abs(d).max()):.2e}")
