"""The reference's one end-to-end golden artefact, shirley-spheres.png (README.md:3,7: --dimension=600,300
--samples-per-pixel=32 --max-ray-bounces=8), compared PIXEL BY PIXEL.

Why this is possible: the R2 sampler makes the reference image deterministic (up to the float64 summation order
where three tiles' borders overlap, far below 8 bits), and the scene is reproducible once `Random.float` is
Base's (shirley_spheres/bin/main.ml opens Base): two 30-bit draws per float on top of OCaml 5's LXM generator.
The 8-bit rule of Bimage/stb is truncation: byte = floor(255·v).

What an exact match pins at once, against the reference itself rather than against our reading of it: the PRNG and
the scene generator (main.ml:56-101,250-253), Camera.create/ray, Mat4.look_at, the binned-SAH Shape_tree and its
ordered traversal, the Rust AVX sphere kernel, Sphere.hit/tex_coord, the checker texture, all three materials,
Shader_space/Quaternion, the R2 stream and its `pass*spp` offset rule, trace_path's dimension consumption, the
3x3 binomial splat, stitch (edge darkening), gamma and the PNG quantisation.  The fixture is the decoded PNG
(tests/golden/shirley_png_rgb8.npz, made by tests/golden/make_png_facts.py)."""
import os

import numpy as np
import pytest

import path_tracer_ocaml_b200 as P
import pyoracle as O
from helpers import make_params

HERE = os.path.dirname(os.path.abspath(__file__))
W, H, SPP, MB = 600, 300, 32, 8


def golden():
    return np.load(os.path.join(HERE, "golden", "shirley_png_rgb8.npz"))["rgb8"].astype(np.int32)


def quantise(img):
    return np.floor(255.0 * np.clip(img, 0.0, 1.0)).astype(np.int32)


@pytest.fixture(scope="module")
def c1_scene():
    return P.shirley_spheres(W, H)


def test_shirley_scene_is_the_references_scene(c1_scene):
    t = c1_scene.tables()
    assert t["n_spheres"] == 530 and t["n_triangles"] == 0  # ground + 3 big + 526 kept small spheres


def test_oracle_reproduces_the_golden_png_exactly(c1_scene):
    osc = O.OracleScene(c1_scene.tables())  # Simd_leaf, cutoff 16: the default binary (main.ml:223-226)
    img, _ = osc.render(make_params(c1_scene, W, H, SPP, MB), n_threads=os.cpu_count())
    d = np.abs(quantise(img) - golden()).max(-1)
    # measured: 100 % of the 180 000 pixels equal in all three bytes.  The bar leaves room for a last-ulp libm
    # difference between glibc versions flipping a byte at a truncation boundary, nothing else.
    assert (d == 0).mean() >= 0.999, (d == 0).mean()
    assert d.max() <= 1, d.max()


def test_oracle_no_simd_leaf_also_matches_the_png(c1_scene):
    # `--no-simd`: Array_leaf (cutoff 4) + the scalar Sphere.intersect (sphere.ml:35-54).  A different tree and a
    # differently rounded test, the same closest hits: the image may differ from the PNG only where a last-ulp t
    # difference flips a decision (checker cell, silhouette) on some of a pixel's 32 samples.
    osc = O.OracleScene(c1_scene.tables(), O.ORC_LEAF_ARRAY, 4)
    img, _ = osc.render(make_params(c1_scene, W, H, SPP, MB), n_threads=os.cpu_count())
    d = np.abs(quantise(img) - golden()).max(-1)
    assert (d <= 1).mean() >= 0.995, (d <= 1).mean()
    assert (d <= 8).mean() >= 0.9995


def test_wrong_float_rule_does_not_match_the_png(c1_scene):
    # negative control: another seed renders another sphere field — the match above is not vacuous
    other = P.shirley_spheres(W, H, seed=7)
    osc = O.OracleScene(other.tables())
    img, _ = osc.render(make_params(other, W, H, 4, MB), n_threads=os.cpu_count())
    d = np.abs(quantise(img) - golden()).max(-1)
    assert (d <= 8).mean() < 0.9


@pytest.mark.gpu
def test_device_image_matches_the_golden_png(c1_scene):
    """The float32 device pipeline against the REFERENCE's own output (not against the oracle)."""
    integ = P.Integrator(c1_scene, W, H, SPP, MB, device=0)
    img = integ.render()
    d = np.abs(quantise(img) - golden()).max(-1)
    # float32 paths diverge from the float64 reference at decision boundaries (checker cells, silhouettes,
    # Schlick test) on single samples of a pixel: +-1 LSB almost everywhere, a handful of pixels beyond
    assert (d <= 1).mean() >= 0.99, (d <= 1).mean()
    assert (d <= 4).mean() >= 0.998, (d <= 4).mean()
    assert np.abs(quantise(img).mean((0, 1)) - golden().mean((0, 1))).max() < 0.25


@pytest.mark.gpu
def test_device_float64_mode_matches_the_golden_png(c1_scene):
    """PTB_FLAG_F64 runs the same wavefront pipeline in float64 (own 4-wide tree, own summation order)."""
    integ = P.Integrator(c1_scene, W, H, SPP, MB, device=0)
    img = integ.render(flags=P.capi.PTB_FLAG_F64)
    d = np.abs(quantise(img) - golden())
    # Knife edge: the sky's blue channel is exactly 1.0 (lerp of 1.0 and 1.0, main.ml:104-110), so an interior sky pixel
    # is sqrt(sum of the nine filter weights) = 1 - a few 1e-16 or exactly 1, i.e. byte 254 or 255 depending on the
    # summation order (per-sample fma splat in the reference, one 3x3 gather of per-pixel sums here).  Values within
    # 1e-6 of a byte boundary are therefore allowed to land on either side; everything else must match exactly.
    x = 255.0 * np.clip(img, 0.0, 1.0)
    knife = np.abs(x - np.round(x)) < 1e-6
    assert (d[~knife] == 0).mean() >= 0.9995, (d[~knife] == 0).mean()
    assert d[knife].max() <= 1 and d.max() <= 2


@pytest.mark.gpu
def test_cli_twin_writes_the_golden_png(tmp_path, c1_scene):
    """`shirley_spheres --dimension=600,300 --samples-per-pixel=32 --max-ray-bounces=8` (README.md:7) through the
    C++ twin of render_command: the PNG it writes, decoded, against the reference's PNG."""
    import subprocess
    from PIL import Image
    exe = os.path.join(os.path.dirname(P.__file__), "bin", "shirley_spheres")
    out = str(tmp_path / "out.png")
    subprocess.check_call([exe, "--dimension=600,300", "--samples-per-pixel=32", "--max-ray-bounces=8",
                           "--no-progress", "-o", out])
    got = np.asarray(Image.open(out).convert("RGB")).astype(np.int32)
    d = np.abs(got - golden()).max(-1)
    assert (d <= 1).mean() >= 0.99 and (d <= 4).mean() >= 0.998
