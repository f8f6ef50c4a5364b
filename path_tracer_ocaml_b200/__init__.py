"""path_tracer_ocaml_b200 — host-side Python harness over libptb200 (C ABI + sm_100a CUDA kernels).

The product is the shared library built from ``csrc/`` (see ``include/ptb200.h``).  This package only
binds it with ctypes for the tests and ``bench.py``, and mirrors the reference's interface names
(``Integrator.create/render``, ``Render_command.Args``, the scene binaries) so tests read like the
reference's own.  There is no CPU fallback: importing works anywhere, but every compute call raises
if the CUDA extension is missing or no GPU is present.
"""
from . import capi  # noqa: F401
from .capi import lib, PtbError, Params, Stats, Texture, Material  # noqa: F401
from .scenes import (Scene, shirley_spheres, cornell_box, cornell_box_lit, synthetic_mesh_scene, synthetic_mesh, mesh_scene,  # noqa: F401
                     read_ply_mesh, write_ply_mesh, ganesha)
from .integrator import Integrator, Args  # noqa: F401
