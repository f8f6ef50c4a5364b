"""C2 (cornell geometry, 1024x1024, 16 bounces) device time for the current environment knobs (PTB_REFILL, ...).
usage: python scripts/cornell_probe.py [spp] [lit]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lit = len(sys.argv) > 2 and sys.argv[2] == "lit"
W = H = 1024
s = P.cornell_box_lit(W, H) if lit else P.cornell_box(W, H, ("constant", (1.0, 1.0, 1.0), None))
integ = P.Integrator(s, W, H, spp, 16)
best = None
for _ in range(3):
    integ.render(flags=capi.PTB_FLAG_PROFILE)
    st = integ.stats
    if best is None or st.ms_device < best[0]:
        best = (st.ms_device, st.ms_trace, int(st.rays))
print(f"cornell{' lit' if lit else ''} {spp} spp, PTB_REFILL={os.environ.get('PTB_REFILL', 'default')}: device {best[0]:.2f} ms, trace {best[1]:.2f} ms, "
      f"{W * H * spp / best[0] / 1e3:.0f} Mpaths/s, {best[2] / best[0] / 1e6:.2f} Grays/s")
