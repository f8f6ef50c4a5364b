import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import path_tracer_ocaml_b200 as P
W,H,SPP,MB=3840,2160,int(sys.argv[1]) if len(sys.argv)>1 else 256,8
scene=P.shirley_spheres(W,H)
integ=P.Integrator(scene,W,H,SPP,MB,device=0)
host=torch.empty(H,W,3,dtype=torch.float64).pin_memory(); hn=host.numpy()
for i in range(3):
    t0=time.perf_counter(); scene.commit(0); t1=time.perf_counter(); integ.render(image=hn); t2=time.perf_counter()
    s=integ.stats
    print(f"commit {1e3*(t1-t0):.1f} ms render call {1e3*(t2-t1):.1f} ms: ms_total {s.ms_total:.1f} ms_device {s.ms_device:.1f} ms_d2h {s.ms_d2h:.1f}")
