"""Commit (tree build + upload) and render time of the synthetic ganesha mesh with the host and the device builder."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi
W, H, SPP = 1920, 1080, 256
for nf in (100000, 1000000, 4000000):
    for builder in ("host", "gpu-lbvh", "gpu-sah"):
        os.environ["PTB_BUILDER"] = builder
        sc = P.synthetic_mesh_scene(nf, W, H)
        ts = []
        for i in range(3):
            t0 = time.perf_counter(); sc.commit(0); ts.append((time.perf_counter() - t0) * 1e3)
        integ = P.Integrator(sc, W, H, SPP, 8)
        best = 1e30
        for i in range(2):
            integ.render(); best = min(best, integ.stats.ms_device)
        st = integ.stats
        print(f"{nf:8d} faces {builder:8s}: commit ms {min(ts):8.1f}  render {best:7.1f} ms  {st.paths/best/1e3:7.0f} Mpaths/s {st.rays/best/1e3:7.0f} Mrays/s  {sc.tree_stats()}", flush=True)
