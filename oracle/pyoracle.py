"""ctypes binding of oracle/liboracle.so — the CPU ORACLE.  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")

ORC_LEAF_SIMD, ORC_LEAF_ARRAY = 0, 1


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
                ("pass_first", C.c_int32), ("pass_count", C.c_int32)]  # (the oracle renders whole frames: ignored)


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("paths", "rays", "box_tests", "sphere_tests", "tri_tests", "leaf_visits", "hits",
                 "scatter_lambert", "scatter_metal", "scatter_dielectric", "absorbed", "missed", "exhausted")]
    _fields_ += [("rays_by_bounce", C.c_uint64 * 64)]

    def as_dict(self):
        d = {n: int(getattr(self, n)) for n, _ in self._fields_[:13]}
        d["rays_by_bounce"] = [int(v) for v in self.rays_by_bounce]
        return d


_dp, _ip, _vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_lds_phi.restype = C.c_double
        L.orc_lds_phi.argtypes = [C.c_int]
        L.orc_lds_alpha.argtypes = [C.c_int, _dp]
        L.orc_lds_get.restype = C.c_double
        L.orc_lds_get.argtypes = [_dp, C.c_int64, C.c_int]
        L.orc_filter_binomial.argtypes = [C.c_int, C.c_int, _dp]
        L.orc_tile_split.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, _ip, C.c_int]
        L.orc_bbox_is_hit.argtypes = [_dp, _dp, _dp, _dp, C.c_double, C.c_double]
        L.orc_camera_create.argtypes = [_dp, _dp, _dp, C.c_double, C.c_double, _dp]
        L.orc_camera_transform.argtypes = [_dp, _dp, _dp, _dp, C.c_int64]
        L.orc_camera_ray.argtypes = [_dp, C.c_double, C.c_double, _dp]
        L.orc_unit_square_to_hemisphere.argtypes = [C.c_double, C.c_double, _dp]
        L.orc_film_tile_write_pixel.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp]
        L.orc_sphere_intersect_scalar.argtypes = [_dp, C.c_double, _dp, _dp, C.c_double, C.c_double, _dp]
        L.orc_spheres_intersect_simd.argtypes = [_dp, _dp, _dp, _dp, C.c_int, _dp, _dp, C.c_double, C.c_double, _dp]
        L.orc_triangle_intersect.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, _dp, _dp, _dp]
        L.orc_shader_space_rotate.argtypes = [_dp, _dp, C.c_int, _dp]
        L.orc_scene_create.restype = _vp
        L.orc_scene_destroy.argtypes = [_vp]
        L.orc_scene_set_textures.argtypes = [_vp, C.POINTER(Texture), C.c_int]
        L.orc_scene_set_materials.argtypes = [_vp, C.POINTER(Material), C.c_int]
        L.orc_scene_set_spheres.argtypes = [_vp, _dp, _dp, _dp, _dp, _ip, C.c_int64]
        L.orc_scene_set_triangles.argtypes = [_vp, _dp, _dp, _dp, C.c_int64, _ip, _ip, _dp, C.c_int64]
        L.orc_scene_set_background.argtypes = [_vp, C.c_int, _dp, _dp]
        L.orc_scene_set_light_quad.argtypes = [_vp, _dp, _dp, _dp]
        L.orc_scene_commit.argtypes = [_vp, C.c_int, C.c_int, _ip, C.c_int64]
        L.orc_scene_tree_depth.argtypes = [_vp]
        L.orc_scene_node_count.restype = C.c_int64
        L.orc_scene_node_count.argtypes = [_vp]
        L.orc_scene_leaf_histogram.argtypes = [_vp, _ip, _ip, C.c_int]
        L.orc_render.argtypes = [_vp, C.POINTER(Params), _dp, C.POINTER(Counters), C.c_int, C.c_int]
        L.orc_trace_sample.argtypes = [_vp, C.POINTER(Params), C.c_int, C.c_int, C.c_int, _dp, C.POINTER(Counters)]
        L.orc_intersect_batch.argtypes = [_vp, _dp, _dp, C.c_double, C.c_double, C.c_int64, _dp, _ip,
                                          C.POINTER(Counters), C.c_int]
        L.orc_intersect_batch_linear.argtypes = [_vp, _dp, _dp, C.c_double, C.c_double, C.c_int64, _dp, _ip, C.c_int]
        L.orc_first_hit.argtypes = [_vp, C.POINTER(Params), _dp, _ip, _dp, _dp]
        _lib = L
    return _lib


def dptr(a):
    return a.ctypes.data_as(_dp)


def iptr(a):
    return a.ctypes.data_as(_ip)


def d3(v):
    return np.ascontiguousarray(v, dtype=np.float64)


class OracleScene:
    """A scene handed to the oracle as plain tables (the same tables the product's C ABI takes)."""

    def __init__(self, tables, leaf_kind=None, length_cutoff=None, use_ref_order=True, commit=True):
        L = lib()
        self.h = L.orc_scene_create()
        t = tables
        texs = (Texture * max(t["n_textures"], 1))()
        mats = (Material * max(t["n_materials"], 1))()
        C.memmove(texs, t["textures"], C.sizeof(Texture) * t["n_textures"])
        C.memmove(mats, t["materials"], C.sizeof(Material) * t["n_materials"])
        L.orc_scene_set_textures(self.h, texs, t["n_textures"])
        L.orc_scene_set_materials(self.h, mats, t["n_materials"])
        self._keep = [d3(t[k]) for k in ("xs", "ys", "zs", "rs", "vx", "vy", "vz", "uv")]
        xs, ys, zs, rs, vx, vy, vz, uv = self._keep
        sm = np.ascontiguousarray(t["sphere_material"], dtype=np.int32)
        L.orc_scene_set_spheres(self.h, dptr(xs), dptr(ys), dptr(zs), dptr(rs), iptr(sm), t["n_spheres"])
        if t["n_triangles"]:
            idx = np.ascontiguousarray(t["indices"], dtype=np.int32)
            tm = np.ascontiguousarray(t["tri_material"], dtype=np.int32)
            L.orc_scene_set_triangles(self.h, dptr(vx), dptr(vy), dptr(vz), t["n_vertices"], iptr(idx), iptr(tm),
                                      dptr(uv), t["n_triangles"])
        L.orc_scene_set_background(self.h, t["bg_kind"], dptr(d3(t["bg0"])), dptr(d3(t["bg1"])))
        if t.get("has_light"):
            L.orc_scene_set_light_quad(self.h, dptr(d3(t["light_o"])), dptr(d3(t["light_u"])), dptr(d3(t["light_v"])))
        if leaf_kind is None:  # shirley default = Simd_leaf (main.ml:223-226); anything with triangles = Array_leaf
            leaf_kind = ORC_LEAF_SIMD if t["n_triangles"] == 0 else ORC_LEAF_ARRAY
        if length_cutoff is None:
            length_cutoff = 16 if leaf_kind == ORC_LEAF_SIMD else 4
        if not commit:  # only intersect_batch_linear may be used
            return
        order = np.ascontiguousarray(t["prim_order"], dtype=np.int32)
        if use_ref_order and len(order):
            rc = L.orc_scene_commit(self.h, leaf_kind, length_cutoff, iptr(order), len(order))
        else:
            rc = L.orc_scene_commit(self.h, leaf_kind, length_cutoff, None, 0)
        if rc != 0:
            raise RuntimeError(f"orc_scene_commit failed: {rc}")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @staticmethod
    def params(product_params):
        p = Params()
        C.memmove(C.byref(p), C.byref(product_params), C.sizeof(Params))  # same POD layout (ptb_params)
        return p

    def render(self, params, n_threads=1, pass_limit=0, flags=None):
        p = self.params(params)
        if flags is not None:
            p.flags = flags
        img = np.zeros((p.height, p.width, 3), dtype=np.float64)
        cn = Counters()
        rc = lib().orc_render(self.h, C.byref(p), dptr(img), C.byref(cn), n_threads, pass_limit)
        if rc != 0:
            raise RuntimeError(f"orc_render failed: {rc}")
        return img, cn

    def trace_sample(self, params, gx, gy, pas):
        p = self.params(params)
        rgb = np.zeros(3)
        lib().orc_trace_sample(self.h, C.byref(p), gx, gy, pas, dptr(rgb), None)
        return rgb

    def first_hit(self, params):
        p = self.params(params)
        t = np.zeros((p.height, p.width))
        prim = np.zeros((p.height, p.width), dtype=np.int32)
        cx, cy = np.zeros((p.height, p.width)), np.zeros((p.height, p.width))
        lib().orc_first_hit(self.h, C.byref(p), dptr(t), iptr(prim), dptr(cx), dptr(cy))
        return t, prim, cx, cy

    def intersect_batch(self, o, d, t_min=0.0, t_max=1.7976931348623157e308, n_threads=1):
        o, d = d3(o).reshape(-1), d3(d).reshape(-1)
        n = len(o) // 3
        t = np.zeros(n)
        prim = np.zeros(n, dtype=np.int32)
        cn = Counters()
        lib().orc_intersect_batch(self.h, dptr(o), dptr(d), t_min, t_max, n, dptr(t), iptr(prim), C.byref(cn),
                                  n_threads)
        return t, prim, cn

    def intersect_batch_linear(self, o, d, t_min=0.0, t_max=1.7976931348623157e308, n_threads=1):
        o, d = d3(o).reshape(-1), d3(d).reshape(-1)
        n = len(o) // 3
        t = np.zeros(n)
        prim = np.zeros(n, dtype=np.int32)
        lib().orc_intersect_batch_linear(self.h, dptr(o), dptr(d), t_min, t_max, n, dptr(t), iptr(prim), n_threads)
        return t, prim

    def tree_stats(self):
        L = lib()
        sizes, counts = np.zeros(64, dtype=np.int32), np.zeros(64, dtype=np.int32)
        n = L.orc_scene_leaf_histogram(self.h, iptr(sizes), iptr(counts), 64)
        return {"depth": L.orc_scene_tree_depth(self.h), "nodes": L.orc_scene_node_count(self.h),
                "leaf_histogram": {int(s): int(c) for s, c in zip(sizes[:n], counts[:n])}}


# ---- scenes the oracle can load on its own (no product library) ----------------------------------------------------
def load_scene_file(name):
    """Tables of oracle/scenes/<name>.npz in the dict layout OracleScene takes (written by make_scene_files.py)."""
    z = np.load(os.path.join(_HERE, "scenes", name + ".npz"))
    nm, nt = len(z["mat_kind"]), len(z["tex_kind"])
    mats, texs = (Material * max(nm, 1))(), (Texture * max(nt, 1))()
    for i in range(nm):
        mats[i].kind, mats[i].texture, mats[i].index = int(z["mat_kind"][i]), int(z["mat_texture"][i]), float(z["mat_index"][i])
    for i in range(nt):
        texs[i].kind = int(z["tex_kind"][i])
        texs[i].width, texs[i].height, texs[i].even, texs[i].odd = (int(v) for v in z["tex_whe"][i])
        for c in range(3):
            texs[i].rgb[c] = float(z["tex_rgb"][i][c])
    e = np.zeros(0)
    return {"n_spheres": len(z["rs"]), "n_vertices": 0, "n_triangles": 0, "xs": z["xs"], "ys": z["ys"], "zs": z["zs"],
            "rs": z["rs"], "sphere_material": z["sphere_material"], "vx": e, "vy": e, "vz": e, "uv": e,
            "indices": np.zeros(0, dtype=np.int32), "tri_material": np.zeros(0, dtype=np.int32),
            "materials": mats, "n_materials": nm, "textures": texs, "n_textures": nt,
            "bg_kind": int(z["bg_kind"]), "bg0": z["bg0"], "bg1": z["bg1"], "prim_order": z["prim_order"]}


def shirley_camera(aspect):
    """Camera.create for shirley_spheres (main.ml:26-31) through the oracle's own restatement: (llx, lly, vx, vy)."""
    out = np.zeros(20)
    lib().orc_camera_create(dptr(d3([13.0, 2.0, 4.5])), dptr(d3([0.0, 0.0, 0.0])), dptr(d3([0.0, 1.0, 0.0])), float(aspect),
                            20.0, dptr(out))
    return tuple(float(v) for v in out[:4])


def make_params(cam4, width, height, spp, max_bounces, rank=0, world=1, flags=0):
    return Params(width=width, height=height, samples_per_pixel=spp, max_bounces=max_bounces, lower_left_x=cam4[0],
                  lower_left_y=cam4[1], view_x=cam4[2], view_y=cam4[3], tile_rank=rank, tile_world=world, flags=flags,
                  device=0)
