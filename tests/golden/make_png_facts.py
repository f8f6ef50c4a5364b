"""Extracts the LAYOUT-INDEPENDENT facts of the reference's only golden artefact,
/root/reference/shirley-spheres.png (README.md:3,7: --dimension=600,300 --samples-per-pixel=32
--max-ray-bounces=8), into a small fixture.  Run in the build container (the PNG does not travel to
the GPU box):  python tests/golden/make_png_facts.py

Two fixtures:
* shirley_png_facts.json — layout-independent facts: the sky rows (camera + background + filter + gamma + 8-bit
  quantisation), the edge darkening of the 3x3 splat (integrator.ml:115-117) and the horizon row;
* shirley_png_rgb8.npz — the decoded 300x600x3 uint8 pixels, so that the oracle (and the device) can be compared
  with the reference's output PIXEL BY PIXEL: with Base.Random.float's two-draw rule the restated scene is the
  reference's scene and the R2 sampler makes the image deterministic (tests/test_golden_png.py)."""
import json
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
g8 = np.asarray(Image.open("/root/reference/shirley-spheres.png").convert("RGB"))
np.savez_compressed(os.path.join(HERE, "shirley_png_rgb8.npz"), rgb8=g8)
g = g8.astype(np.float64)
H, W, _ = g.shape
# the sky is the smooth region at the top; a row belongs to it while it has no strong horizontal edges
rows = []
for y in range(H):
    if np.abs(np.diff(g[y, 2:-2], axis=0)).max() > 3:
        break
    rows.append(y)
n_sky = len(rows)
facts = {
    "source": "shirley-spheres.png (reference repo root)",
    "width": W, "height": H,
    "mean_rgb": g.mean((0, 1)).tolist(),
    "row0_over_row1": float(g[0].mean() / g[1].mean()),
    "col0_over_col1": float(g[:, 0].mean() / g[:, 1].mean()),
    "corner_over_inner": float(g[0, 0].mean() / g[1, 1].mean()),
    "n_smooth_sky_rows": n_sky,
    "sky_row_mean_rgb": {str(y): g[y, 2:-2].mean(0).tolist() for y in range(1, n_sky, 8)},
}
json.dump(facts, open(os.path.join(HERE, "shirley_png_facts.json"), "w"), indent=1)
print(json.dumps(facts, indent=1)[:1500])
