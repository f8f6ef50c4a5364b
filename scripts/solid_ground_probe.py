"""Diagnostic: the shirley scene with its 1000x2000 checker ground replaced by a solid colour, per-launch device
times through PTB_TIMING=1 (does the checker path explain the slow bounce-0 k_shade launch?).
usage: PTB_TIMING=1 python scripts/solid_ground_probe.py [solid|checker]"""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi

W, H, SPP, MB = 3840, 2160, 32, 8
src = P.shirley_spheres(W, H)
t = src.tables()
texs = [capi.Texture.from_buffer_copy(t["textures"][i]) for i in range(t["n_textures"])]
if (sys.argv[1] if len(sys.argv) > 1 else "solid") == "solid":
    for x in texs:
        if x.kind == capi.PTB_TEX_CHECKER:
            x.kind = capi.PTB_TEX_SOLID
            x.rgb[0], x.rgb[1], x.rgb[2] = 0.5, 0.5, 0.5
s = P.Scene()
s.set_textures(texs)
s.set_materials([capi.Material.from_buffer_copy(t["materials"][i]) for i in range(t["n_materials"])])
s.set_spheres(t["xs"], t["ys"], t["zs"], t["rs"], material=t["sphere_material"])
s.set_background(t["bg_kind"], t["bg0"], t["bg1"])
s.camera = src.camera
integ = P.Integrator(s, W, H, SPP, MB)
sums = torch.zeros(H, W, 3, dtype=torch.float32, device="cuda")
for i in range(2):
    sums.zero_()
    integ.render_device(sums, flags=capi.PTB_FLAG_PROFILE)
    torch.cuda.synchronize()
print("ms_device", integ.stats.ms_device, "ms_trace", integ.stats.ms_trace)
