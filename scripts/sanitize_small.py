"""Tiny end-to-end render of the three scene kinds for compute-sanitizer (memcheck / racecheck)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import path_tracer_ocaml_b200 as P
for name, sc, W, H, spp, mb in (("shirley", P.shirley_spheres(64, 32), 64, 32, 2, 8), ("cornell", P.cornell_box(32, 32), 32, 32, 2, 16),
                                ("mesh", P.synthetic_mesh_scene(20000, 48, 27), 48, 27, 2, 8)):
    img = P.Integrator(sc, W, H, spp, mb).render()
    print(name, img.shape, float(img.mean()))
