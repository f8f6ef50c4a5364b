"""The reference-side binding (integration/ocaml/ptb_stubs.c, what `external` declarations in ptb.ml bind) is
compiled against include/ptb200.h and run on a mock of the OCaml C runtime (tests/mock_caml/): no OCaml toolchain
exists in this image, so this is how the stubs' marshalling is exercised.  CPU: compiles, links, marshals every
setter, and the device-less commit raises Failure.  GPU: the image rendered through the stubs equals the image the
ctypes wrapper renders from the same scene."""
import os
import subprocess

import numpy as np
import pytest

import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, "path_tracer_ocaml_b200", "lib")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("stubs") / "stub_driver")
    cmd = ["gcc", "-O1", "-std=gnu11", "-Wall", "-Werror", "-I", os.path.join(HERE, "mock_caml"), "-I",
           os.path.join(ROOT, "include"), os.path.join(ROOT, "integration", "ocaml", "ptb_stubs.c"),
           os.path.join(HERE, "mock_caml", "stub_driver.c"), "-L", LIBDIR, "-lptb200", "-Wl,-rpath," + LIBDIR, "-lm",
           "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_stubs_compile_marshal_and_fail_loudly_without_a_device(driver):
    if P.lib().ptb_device_count() > 0:
        pytest.skip("a CUDA device is present: the device-less failure cannot be provoked")
    r = subprocess.run([driver, "nogpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "devices 0" in r.stdout
    assert "EXCEPTION Failure:" in r.stdout and "no CUDA device" in r.stdout


def test_stubs_surface_abi_errors_as_failure(driver):
    r = subprocess.run([driver, "badrow"], capture_output=True, text=True)
    assert "EXCEPTION Failure:" in r.stdout and "material row out of range" in r.stdout, r.stdout + r.stderr


def _python_scene():
    s = P.Scene()
    T, M = capi.Texture, capi.Material
    s.set_textures([T(kind=capi.PTB_TEX_SOLID, rgb=(0.8, 0.3, 0.3)), T(kind=capi.PTB_TEX_SOLID, rgb=(0.9, 0.9, 0.9)),
                    T(kind=capi.PTB_TEX_SOLID, rgb=(0.2, 0.2, 0.2)),
                    T(kind=capi.PTB_TEX_CHECKER, width=10, height=20, even=1, odd=2)])
    s.set_materials([M(kind=capi.PTB_MAT_LAMBERTIAN, texture=3, index=1.0), M(kind=capi.PTB_MAT_LAMBERTIAN, texture=0, index=1.0),
                     M(kind=capi.PTB_MAT_METAL, texture=1, index=1.0), M(kind=capi.PTB_MAT_DIELECTRIC, texture=-1, index=1.5)])
    s.set_spheres([0.0, 0.0, -1.0, 1.0], [-100.5, 0.0, 0.0, 0.0], [-1.0, -1.2, -1.0, -1.0], [100.0, 0.5, 0.5, 0.5],
                  material=[0, 1, 3, 2])
    s.set_background(capi.PTB_BG_GRADIENT_Y, (1.0, 1.0, 1.0), (0.5, 0.7, 1.0))
    from path_tracer_ocaml_b200.scenes import Camera
    s.camera = Camera([-2.0, -1.0, 4.0, 2.0] + [0.0] * 16)
    return s


@pytest.mark.gpu
def test_stubs_render_the_same_image_as_the_ctypes_path(driver, tmp_path):
    out = str(tmp_path / "img.f64")
    r = subprocess.run([driver, "render", out], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
    # the ray straight ahead hits the middle sphere (slot order is the caller's: sphere 1) at t = 1.2 - 0.5
    line = [l for l in r.stdout.splitlines() if l.startswith("intersect")][0].split()
    assert abs(float(line[2]) - 0.7) < 1e-4 and int(line[4]) == 1 and int(line[6]) == 1 and int(line[8]) == -1
    img = np.fromfile(out, dtype=np.float64).reshape(32, 64, 3)
    ref = P.Integrator(_python_scene(), 64, 32, 16, 6).render()
    # same library, same scene, same samples: only the order of the float32 atomics differs
    assert np.abs(img - ref).max() < 1e-3 and img.std() > 0.05
