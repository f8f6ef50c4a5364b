"""ctypes binding of include/ptb200.h and include/ptb200_scenes.h (no compute happens at import)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libptb200.so")  # the one product library; no override, no fallback

PTB_TEX_SOLID, PTB_TEX_CHECKER = 0, 1
PTB_MAT_LAMBERTIAN, PTB_MAT_METAL, PTB_MAT_DIELECTRIC, PTB_MAT_EMISSIVE = 0, 1, 2, 3
PTB_BG_CONSTANT, PTB_BG_GRADIENT_Y = 0, 1
PTB_FLAG_F64, PTB_FLAG_RAW_SUMS, PTB_FLAG_NO_FILTER, PTB_FLAG_PROFILE = 1, 2, 4, 8


class PtbError(RuntimeError):
    pass


class Texture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("even", C.c_int32),
                ("odd", C.c_int32), ("_pad", C.c_int32), ("rgb", C.c_double * 3)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("texture", C.c_int32), ("index", C.c_double)]


class Params(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("samples_per_pixel", C.c_int32),
                ("max_bounces", C.c_int32), ("lower_left_x", C.c_double), ("lower_left_y", C.c_double),
                ("view_x", C.c_double), ("view_y", C.c_double), ("tile_rank", C.c_int32),
                ("tile_world", C.c_int32), ("flags", C.c_int32), ("device", C.c_int32),
                ("pass_first", C.c_int32), ("pass_count", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("rays_by_bounce", C.c_uint64 * 64),
                ("kernel_launches", C.c_uint64), ("ms_total", C.c_double), ("ms_device", C.c_double),
                ("ms_trace", C.c_double), ("ms_h2d", C.c_double), ("ms_d2h", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


_P = C.POINTER
_dp, _fp, _ip, _vp = _P(C.c_double), _P(C.c_float), _P(C.c_int32), C.c_void_p

# name -> (restype, argtypes); every symbol the two headers declare
SYMBOLS = {
    "ptb_device_count": (C.c_int, []),
    "ptb_last_error": (C.c_char_p, []),
    "ptb_version": (C.c_char_p, []),
    "ptb_leaf_size": (C.c_int, []),
    "ptb_scene_create": (_vp, []),
    "ptb_scene_destroy": (None, [_vp]),
    "ptb_scene_set_textures": (C.c_int, [_vp, _P(Texture), C.c_int32]),
    "ptb_scene_set_materials": (C.c_int, [_vp, _P(Material), C.c_int32]),
    "ptb_scene_set_spheres": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _ip, C.c_int64]),
    "ptb_scene_set_triangles": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int64, _ip, _ip, _dp, C.c_int64]),
    "ptb_scene_set_background": (C.c_int, [_vp, C.c_int32, _dp, _dp]),
    "ptb_scene_set_light_quad": (C.c_int, [_vp, _dp, _dp, _dp]),
    "ptb_scene_commit": (C.c_int, [_vp, C.c_int32, _dp]),
    "ptb_scene_primitive_count": (C.c_int64, [_vp]),
    "ptb_scene_tree_stats": (C.c_int, [_vp, _ip]),
    "ptb_render": (C.c_int, [_vp, _P(Params), _dp, _P(Stats)]),
    "ptb_render_progress": (C.c_int, [C.c_int32, _P(C.c_uint64), _P(C.c_uint64)]),
    "ptb_host_alloc": (_vp, [C.c_uint64]),
    "ptb_host_free": (None, [_vp]),
    "ptb_scene_commit_multi": (C.c_int, [_vp, C.c_int32, _dp]),
    "ptb_render_multi": (C.c_int, [_vp, _P(Params), C.c_int32, _dp, _P(Stats)]),
    "ptb_render_device": (C.c_int, [_vp, _P(Params), _vp, _vp, _P(Stats)]),
    "ptb_resolve_device": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "ptb_intersect_batch": (C.c_int, [_vp, _fp, _fp, C.c_float, C.c_float, C.c_int64, _fp, _ip, C.c_int32,
                                      _P(Stats)]),
    "ptb_intersect_batch_device": (C.c_int, [_vp, _vp, _vp, C.c_float, C.c_float, C.c_int64, _vp, _vp,
                                             C.c_int32, _vp, _P(Stats)]),
    "ptb_r2_stream": (C.c_int, [C.c_int32, _ip, C.c_int64, _dp, C.c_int32]),
    "ptb_raygen": (C.c_int, [_P(Params), C.c_int64, C.c_int64, _ip, _ip, _dp, _dp, _fp]),
    "ptb_first_hit": (C.c_int, [_vp, _P(Params), _fp, _ip]),
    "ptb_fp32_peak": (C.c_int, [C.c_int32, _dp, _dp]),
    "ptb_lds_alpha": (C.c_int, [C.c_int32, _dp]),
    "ptb_tile_split": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _ip, _ip, _ip, _ip, C.c_int32]),
    "ptb_filter_binomial": (C.c_int, [C.c_int32, C.c_int32, _dp]),
    "ptb_presplit_boxes": (C.c_int, [_dp, C.c_double, _dp, _dp, C.c_int32]),
    "ptb_camera_create": (C.c_int, [_dp, _dp, _dp, C.c_double, C.c_double, _dp]),
    "ptb_camera_transform": (C.c_int, [_dp, _dp, _dp, _dp, C.c_int64]),
    # ptb200_scenes.h
    "ptb_scene_load_shirley": (C.c_int, [_vp, C.c_double, C.c_int32, _dp]),
    "ptb_scene_load_cornell": (C.c_int, [_vp, C.c_double, C.c_int32, _dp, _dp, _dp]),
    "ptb_scene_load_cornell_lit": (C.c_int, [_vp, C.c_double, _dp, _dp]),
    "ptb_scene_get_light_quad": (C.c_int, [_vp, _ip, _dp, _dp, _dp]),
    "ptb_scene_load_mesh": (C.c_int, [_vp, _fp, C.c_int64, _ip, C.c_int64, C.c_double, _dp]),
    "ptb_mesh_synthetic": (C.c_int, [C.c_int64, C.c_uint32, _fp, C.c_int64, _ip, C.c_int64,
                                     _P(C.c_int64), _P(C.c_int64)]),
    "ptb_ply_read_mesh": (C.c_int, [C.c_char_p, _P(_fp), _P(C.c_int64), _P(_ip), _P(C.c_int64)]),
    "ptb_ply_parse_mesh": (C.c_int, [C.c_char_p, C.c_int64, _P(_fp), _P(C.c_int64), _P(_ip), _P(C.c_int64)]),
    "ptb_free": (None, [_vp]),
    "ptb_scene_counts": (C.c_int, [_vp, _P(C.c_int64), _P(C.c_int64), _P(C.c_int64), _ip, _ip]),
    "ptb_scene_get_spheres": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _ip]),
    "ptb_scene_get_triangles": (C.c_int, [_vp, _dp, _dp, _dp, _ip, _ip, _dp]),
    "ptb_scene_get_materials": (C.c_int, [_vp, _P(Material), _P(Texture)]),
    "ptb_scene_get_background": (C.c_int, [_vp, _ip, _dp, _dp]),
    "ptb_scene_get_prim_order": (C.c_int, [_vp, _ip, C.c_int64]),
}

_lib = None


def lib():
    """Load libptb200.so; fail loudly if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PtbError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                           f"g.build()'` (make -C path_tracer_ocaml_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the header and the library disagree
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc < 0:
        raise PtbError(f"libptb200 error {rc}: {lib().ptb_last_error().decode()}")
    return rc


def dptr(a):
    return a.ctypes.data_as(_dp)


def fptr(a):
    return a.ctypes.data_as(_fp)


def iptr(a):
    return a.ctypes.data_as(_ip)


def pinned_empty(shape, dtype):
    """A numpy array over page-locked host memory from ptb_host_alloc (what a host that wants full PCIe speed
    hands to ptb_intersect_batch / ptb_render).  The memory is released when the array is garbage-collected."""
    import numpy as np
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) if not isinstance(shape, int) else int(shape)
    nbytes = max(n * dt.itemsize, 1)
    p = lib().ptb_host_alloc(nbytes)
    if not p:
        raise PtbError(f"ptb_host_alloc({nbytes}) failed: {lib().ptb_last_error().decode()}")
    buf = (C.c_char * nbytes).from_address(p)
    arr = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)
    import weakref
    weakref.finalize(buf, lib().ptb_host_free, p)
    return arr
