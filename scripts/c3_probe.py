"""BASELINE configs[2] stand-in for profiling: the seeded synthetic mesh through the ganesha assembly, 1920x1080, a few
samples per pixel, 8 bounces.  Prints rays per bounce so that ncu's per-launch bytes can be turned into bytes per ray.
usage: python scripts/c3_probe.py [faces] [spp]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi

faces = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
W, H = 1920, 1080
sc = P.synthetic_mesh_scene(faces, W, H)
integ = P.Integrator(sc, W, H, spp, 8)
for rep in range(2):
    integ.render(flags=capi.PTB_FLAG_PROFILE)
    st = integ.stats
print(json.dumps({"faces": faces, "spp": spp, "tree": sc.tree_stats(), "commit_ms": sc.commit_ms, "paths": int(st.paths),
                  "rays": int(st.rays), "rays_by_bounce": [int(v) for v in st.rays_by_bounce[:8]], "ms_device": st.ms_device,
                  "ms_trace": st.ms_trace, "mrays_per_s": st.rays / st.ms_device / 1e3}))
