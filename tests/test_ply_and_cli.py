"""The ply_format subset (ply_format/src/ply.ml:340-352 as ganesha uses it, ganesha/bin/main.ml:50-60,182-185)
and the C++ twin of the scene binaries (render_command.ml:16-47,64-109).  The parser tests run on the CPU;
the CLI renders need the GPU."""
import os
import struct
import subprocess

import numpy as np
import pytest

import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi
from helpers import image_metrics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "path_tracer_ocaml_b200", "bin")


def _mesh():
    xyz, faces = P.synthetic_mesh(1500)
    return xyz.reshape(-1, 3), faces.reshape(-1, 3)


@pytest.mark.parametrize("vt,ct,it,extra", [("float", "uchar", "int", ()), ("double", "uint8", "uint", ()),
                                            ("float", "ushort", "ushort", ("nx", "ny", "nz")),
                                            ("float", "int", "int", ("confidence",))])
def test_ply_round_trip_of_every_supported_layout(tmp_path, vt, ct, it, extra):
    xyz, faces = _mesh()
    path = tmp_path / "m.ply"
    P.write_ply_mesh(path, xyz, faces, vertex_type=vt, count_type=ct, index_type=it, extra_vertex_props=extra)
    v, f = P.read_ply_mesh(path)
    assert np.array_equal(v, xyz) and np.array_equal(f, faces)
    v2, f2 = P.read_ply_mesh(open(path, "rb").read())  # the of_bigstring entry point (memory, not a path)
    assert np.array_equal(v2, xyz) and np.array_equal(f2, faces)


def test_ply_skips_elements_it_does_not_need():
    # a fixed-width element before `vertex` and an unrelated list element after the faces
    hdr = ("ply\nformat binary_little_endian 1.0\ncomment x\nobj_info y\nelement material 2\nproperty uchar red\n"
           "property short shininess\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
           "element face 1\nproperty list uchar int vertex_indices\nelement edge_list 2\nproperty list int uchar e\n"
           "end_header\n").encode()
    body = struct.pack("<BhBh", 1, 300, 2, -5)
    body += np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype="<f4").tobytes()
    body += struct.pack("<Biii", 3, 0, 1, 2)
    body += struct.pack("<iBB", 2, 7, 8) + struct.pack("<i", 0)
    v, f = P.read_ply_mesh(hdr + body)
    assert v.shape == (3, 3) and f.tolist() == [[0, 1, 2]] and v[1, 0] == 1.0


@pytest.mark.parametrize("data,code,needle", [
    (b"pl", capi.PTB_E_INVALID if hasattr(capi, "PTB_E_INVALID") else -1, "not enough bytes"),
    (b"plx\nformat binary_little_endian 1.0\nend_header\n", -1, 'start with "ply'),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex 0\n", -1, "end_header"),
    (b"ply\nelement vertex 0\nend_header\n", -1, "no format line"),
    (b"ply\nformat ascii 1.0\nelement vertex 0\nproperty float x\nend_header\n", -6, "to do: handle message format"),
    (b"ply\nformat binary_big_endian 1.0\nend_header\n", -6, "to do: handle message format"),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty quad x\nend_header\n", -1, "unrecognized type"),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty float x\nproperty list uchar int l\nend_header\n",
     -6, "mixed list/non-list"),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty int x\nproperty int y\nproperty int z\n"
     b"element face 0\nproperty list uchar int vertex_indices\nend_header\n" + b"\0" * 12, -1, "expected Floats"),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\n"
     b"element face 1\nproperty list uchar int vertex_indices\nend_header\n" + b"\0" * 48 + struct.pack("<Biiii", 4, 0, 1, 2, 3),
     -1, "exactly 3"),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
     b"element face 1\nproperty list uchar int vertex_indices\nend_header\n" + b"\0" * 36 + struct.pack("<Biii", 3, 0, 1, 9),
     -1, "out of range"),
    (b"ply\nformat binary_little_endian 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
     b"element face 1\nproperty list uchar int vertex_indices\nend_header\n" + b"\0" * 20, -1, "truncated"),
])
def test_ply_errors_mirror_the_reference_failure_sites(data, code, needle):
    with pytest.raises(P.PtbError) as e:
        P.read_ply_mesh(data)
    assert f"error {code}:" in str(e.value) and needle in str(e.value), str(e.value)


def test_cli_argument_errors_without_a_device():
    exe = os.path.join(BIN, "shirley_spheres")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "path_tracer_ocaml_b200", "csrc"), "all"])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 124 and "--dimension" in r.stderr  # Arg.required (render_command.ml:21-25)
    r = subprocess.run([exe, "-d", "64x32"], capture_output=True, text=True)
    assert r.returncode == 124 and "WIDTH,HEIGHT" in r.stderr
    r = subprocess.run([exe, "-d", "64,32", "--device=cpu"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU path" in r.stderr
    r = subprocess.run([os.path.join(BIN, "ptb_scenes"), "teapot", "-d", "8,8"], capture_output=True, text=True)
    assert r.returncode != 0


def _read_png(path):
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"))


@pytest.mark.gpu
def test_cli_shirley_writes_the_same_image_as_the_library(tmp_path):
    out = tmp_path / "s.png"
    r = subprocess.run([os.path.join(BIN, "shirley_spheres"), "--dimension=200,100", "--samples-per-pixel=4",
                        "--max-ray-bounces=8", "-o", str(out), "--no-progress"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for line in ("dim = 200 x 100;", "#spheres = ", "tree depth = ", "build time = ", "rendered in: "):
        assert line in r.stdout, r.stdout
    img = P.Integrator(P.shirley_spheres(200, 100), 200, 100, 4, 8).render()
    want = np.clip((img * 255.0).astype(np.int64), 0, 255).astype(np.uint8)  # truncating 8-bit (golden PNG facts)
    got = _read_png(out)
    assert got.shape == (100, 200, 3)
    # float atomics make the sums order-dependent in the last bits: at most an LSB on a handful of pixels
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1 and np.mean(got != want) < 1e-3


@pytest.mark.gpu
def test_cli_preview_every_rewrites_the_output_and_ends_on_the_same_image(tmp_path):
    """--preview-every=K: the output file is rewritten after every K sample passes; the last write is the image of
    the plain run (same samples; the batches' sums are added in float64 on the host)."""
    a, b = tmp_path / "plain.png", tmp_path / "prog.png"
    common = ["--dimension=160,80", "--samples-per-pixel=9", "--max-ray-bounces=6"]
    r1 = subprocess.run([os.path.join(BIN, "shirley_spheres")] + common + ["-o", str(a), "--no-progress"], capture_output=True, text=True)
    r2 = subprocess.run([os.path.join(BIN, "shirley_spheres")] + common + ["-o", str(b), "--preview-every=4"], capture_output=True, text=True)
    assert r1.returncode == 0 and r2.returncode == 0, r1.stderr + r2.stderr
    assert [l for l in r2.stderr.replace("\r", "\n").splitlines() if l.startswith("preview:")] == [
        f"preview: {k} of 9 passes written to {b}" for k in (4, 8, 9)]
    x, y = _read_png(a), _read_png(b)
    assert np.abs(x.astype(int) - y.astype(int)).max() <= 1 and np.mean(x != y) < 2e-3


@pytest.mark.gpu
def test_cli_ganesha_from_a_ply_file_equals_the_in_memory_mesh(tmp_path):
    xyz, faces = P.synthetic_mesh(20000)
    ply = tmp_path / "g.ply"
    P.write_ply_mesh(ply, xyz, faces, extra_vertex_props=("nx", "ny", "nz"))
    out = tmp_path / "g.ppm"
    r = subprocess.run([os.path.join(BIN, "ganesha"), "-d", "160,90", "--samples-per-pixel", "4", "--ganesha-ply", str(ply),
                        "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    got = np.frombuffer(raw[raw.index(b"255\n") + 4:], dtype=np.uint8).reshape(90, 160, 3)
    img = P.Integrator(P.ganesha(ply, 160, 90), 160, 90, 4, 8).render()
    ref = P.Integrator(P.mesh_scene(xyz, faces, 160, 90), 160, 90, 4, 8).render()
    assert image_metrics(img, ref)["rmse"] < 1e-6
    want = np.clip((img * 255.0).astype(np.int64), 0, 255).astype(np.uint8)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1


@pytest.mark.gpu
def test_cli_cornell_runs_with_both_backgrounds(tmp_path):
    for bg in ("white", "sky"):
        out = tmp_path / f"c_{bg}.png"
        r = subprocess.run([os.path.join(BIN, "cornell_box"), "-d", "64,64", "--samples-per-pixel=2", "--max-ray-bounces=16",
                            f"--background={bg}", "-o", str(out)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "#triangles = 18" in r.stdout and "#spheres = 3" in r.stdout
        assert _read_png(out).shape == (64, 64, 3)
