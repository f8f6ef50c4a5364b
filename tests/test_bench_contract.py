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
    assert "configs[3]" in d["config"]["workload"] and "sample" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_own_arm_has_no_cpu_fallback():
    if P.lib().ptb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = _run(["--gpus", "1", "--steps", "1", "--warmup", "0", "--spp", "1"])
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr or "no CPU path" in r.stderr
