import sys, time
sys.path.insert(0,'/root/repo')
import path_tracer_ocaml_b200 as P
for nf in (100000, 1000000, 4000000):
    sc = P.synthetic_mesh_scene(nf, 1920, 1080)
    ts=[]
    for i in range(3):
        t0=time.perf_counter(); sc.commit(0); ts.append((time.perf_counter()-t0)*1e3)
    print(nf, 'commit ms', [round(t,1) for t in ts], sc.tree_stats())
