"""Development check (GPU box): render C1 on the device in f32 and f64 and compare with the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi
import pyoracle as O

W, H, spp, mb = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (600, 300, 32, 8)))
scene = P.shirley_spheres(W, H)
integ = P.Integrator(scene, W, H, spp, mb, device=0)
print("commit ms", scene.commit_ms, scene.tree_stats())
osc = O.OracleScene(scene.tables())
t0 = time.time(); ref, cn = osc.render(integ.params, n_threads=os.cpu_count()); t_cpu = time.time() - t0
print(f"oracle: {t_cpu:.2f}s  {cn.paths/t_cpu/1e6:.2f} Mpaths/s  rays {cn.rays}")
t_ref, p_ref, _, _ = osc.first_hit(integ.params)
t_dev, p_dev = integ.first_hit()
eq = (p_ref == p_dev)
print("first hit: prim equal frac", eq.mean(), " mismatches", (~eq).sum())
hit = eq & (p_ref >= 0)
print("first hit: max rel dt", np.max(np.abs(t_dev[hit] - t_ref[hit]) / t_ref[hit]))
for flags, name in ((0, "f32"), (capi.PTB_FLAG_F64, "f64")):
    for it in range(3):
        img = integ.render(flags=flags)
    st = integ.stats
    d = img - ref
    rmse = np.sqrt(np.mean(d * d))
    within = np.mean(np.abs(d) <= 0.02 * ref + 1 / 255)
    print(f"{name}: ms_device {st.ms_device:.2f} ms_total {st.ms_total:.2f}  rays {st.rays} (oracle {cn.rays})  "
          f"Mpaths/s {st.paths/st.ms_device/1e3:.1f} Mrays/s {st.rays/st.ms_device/1e3:.1f}  launches {st.kernel_launches}")
    print(f"   rmse {rmse:.6f}  within-tol {within:.5f}  max|d| {np.abs(d).max():.4f}  bias {d.mean((0,1))}")
    print("   rays_by_bounce dev", list(st.rays_by_bounce[:mb]), "\n   rays_by_bounce orc", list(cn.rays_by_bounce[:mb]))
    np.save(os.path.join(ROOT, "gpurun_out", f"dev_{name}.npy"), img.astype(np.float32))
integ.render(flags=capi.PTB_FLAG_PROFILE)
print("profile: ms_device", integ.stats.ms_device, "ms_trace", integ.stats.ms_trace)
