"""bench.py's driver contract, as far as it can be checked without a GPU: the reference arm prints exactly one JSON
line with the agreed keys, non-zero ranks of a torchrun launch of it stay silent, and our own arm fails loudly
(no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

import path_tracer_ocaml_b200 as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, timeout=timeout)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mpaths/s" and d["unit"] == "Mpaths/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert "configs[3]" in d["config"]["workload"] and set(d["config"]) == {"workload", "sharding", "l2", "paths_per_step"}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_loads_only_the_oracle():
    """The reference arm must be self-contained: liboracle.so + the committed scene file, never the product library."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0'];\n"
            "runpy.run_path(%r, run_name='__main__')\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'liboracle.so' in maps, 'oracle not loaded'\n"
            "assert 'libptb200' not in maps and 'path_tracer_ocaml_b200' not in sys.modules, 'product loaded'\n" % BENCH)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]


def test_oracle_scene_file_is_current():
    """oracle/scenes/shirley_spheres.npz (what the reference arm renders) is the product's shirley scene, bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import pyoracle as O
    a, b = O.load_scene_file("shirley_spheres"), P.shirley_spheres(3840, 2160).tables()
    for k in ("xs", "ys", "zs", "rs", "sphere_material", "bg0", "bg1", "prim_order"):
        assert np.array_equal(a[k], b[k]), k
    assert a["n_spheres"] == b["n_spheres"] == 530 and a["n_materials"] == b["n_materials"] and a["n_textures"] == b["n_textures"]
    for i in range(a["n_materials"]):
        assert (a["materials"][i].kind, a["materials"][i].texture, a["materials"][i].index) == \
               (b["materials"][i].kind, b["materials"][i].texture, b["materials"][i].index)
    for i in range(a["n_textures"]):
        assert list(a["textures"][i].rgb) == list(b["textures"][i].rgb) and a["textures"][i].kind == b["textures"][i].kind
    cam = P.shirley_spheres(3840, 2160).camera
    assert O.shirley_camera(3840 / 2160) == (cam.lower_left_x, cam.lower_left_y, cam.view_x, cam.view_y)


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_own_arm_has_no_cpu_fallback():
    if P.lib().ptb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = _run(["--gpus", "1", "--steps", "1", "--warmup", "0", "--spp", "1"])
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr or "no CPU path" in r.stderr
