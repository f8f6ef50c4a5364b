"""Writes oracle/scenes/shirley_spheres.npz: the shirley_spheres scene (shirley_spheres/bin/main.ml:26-110, Random.init
42) as plain tables in CAMERA space, so that the oracle can be run — by bench.py --impl reference and by anyone
checking it — without loading the product library.  Run in the build container after a change to the scene
generator:  python oracle/scenes/make_scene_files.py      (tests/test_bench_contract.py checks the file is current)

The tables are exactly what path_tracer_ocaml_b200.Scene.tables() returns for P.shirley_spheres (sphere centres do
not depend on the image aspect: Mat4.look_at has no aspect in it, camera.ml:14-27); the camera's four numbers are
recomputed per aspect by the oracle's own Camera.create (pyoracle.shirley_camera)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import path_tracer_ocaml_b200 as P  # noqa: E402


def dump(path, t):
    m, x = t["materials"], t["textures"]
    np.savez_compressed(
        path, xs=t["xs"], ys=t["ys"], zs=t["zs"], rs=t["rs"], sphere_material=t["sphere_material"],
        mat_kind=np.array([m[i].kind for i in range(t["n_materials"])], dtype=np.int32),
        mat_texture=np.array([m[i].texture for i in range(t["n_materials"])], dtype=np.int32),
        mat_index=np.array([m[i].index for i in range(t["n_materials"])]),
        tex_kind=np.array([x[i].kind for i in range(t["n_textures"])], dtype=np.int32),
        tex_whe=np.array([[x[i].width, x[i].height, x[i].even, x[i].odd] for i in range(t["n_textures"])], dtype=np.int32),
        tex_rgb=np.array([list(x[i].rgb) for i in range(t["n_textures"])]),
        bg_kind=np.int32(t["bg_kind"]), bg0=t["bg0"], bg1=t["bg1"], prim_order=t["prim_order"])


if __name__ == "__main__":
    dump(os.path.join(HERE, "shirley_spheres.npz"), P.shirley_spheres(16, 9).tables())
    print("wrote", os.path.join(HERE, "shirley_spheres.npz"))
