import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # both libraries are build products (git-ignored); make sure they exist before collection
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    if not os.path.exists(os.path.join(ROOT, "path_tracer_ocaml_b200", "lib", "libptb200.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "path_tracer_ocaml_b200", "csrc")])


def pytest_collection_modifyitems(config, items):
    import path_tracer_ocaml_b200 as P

    if P.lib().ptb_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (libptb200 has no CPU path)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
