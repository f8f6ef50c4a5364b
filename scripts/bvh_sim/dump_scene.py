"""Writes /tmp/shirley_scene.bin for scripts/bvh_sim/bvh_sim.cpp: sphere tables (camera space), material kinds and
the camera of the shirley scene, read back through the C ABI (no GPU needed)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import path_tracer_ocaml_b200 as P

sc = P.shirley_spheres(3840, 2160)
t, c = sc.tables(), sc.camera
kinds = np.array([m.kind for m in t["materials"]], dtype=np.int32)[np.asarray(t["sphere_material"])]
with open("/tmp/shirley_scene.bin", "wb") as f:
    np.array([t["n_spheres"]], dtype=np.int64).tofile(f)
    np.array([c.lower_left_x, c.lower_left_y, c.view_x, c.view_y]).tofile(f)
    for k in ("xs", "ys", "zs", "rs"):
        np.asarray(t[k], dtype=np.float64).tofile(f)
    kinds.astype(np.int32).tofile(f)
print(t["n_spheres"], "spheres; kinds", np.bincount(kinds))
