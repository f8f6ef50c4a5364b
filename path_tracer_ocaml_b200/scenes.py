"""Scene handles: the data of the reference's scene binaries (shirley_spheres/bin/main.ml,
cornell-box/bin/main.ml, ganesha/bin/main.ml) loaded through the C ABI, plus table read-back so a
checker can be fed exactly the same input."""
import ctypes as C

import numpy as np

from . import capi
from .capi import check, dptr, fptr, iptr, lib

_fp_t, _ip_t = C.POINTER(C.c_float), C.POINTER(C.c_int32)


class Camera:
    """Camera.t (path_tracer/src/camera.ml:46-54): what Camera.ray and Camera.transform read."""

    def __init__(self, cam20):
        self.lower_left_x, self.lower_left_y, self.view_x, self.view_y = (float(v) for v in cam20[:4])
        self.look_at = np.array(cam20[4:20], dtype=np.float64)

    @staticmethod
    def create(eye, target, up, aspect, vertical_fov_deg):
        out = np.zeros(20)
        e, t, u = (np.asarray(v, dtype=np.float64) for v in (eye, target, up))
        check(lib().ptb_camera_create(dptr(e), dptr(t), dptr(u), float(aspect), float(vertical_fov_deg), dptr(out)))
        return Camera(out)

    def transform(self, xs, ys, zs):
        check(lib().ptb_camera_transform(dptr(self.look_at), dptr(xs), dptr(ys), dptr(zs), len(xs)))


class Scene:
    """Opaque ptb_scene handle + the camera that goes with it."""

    def __init__(self):
        self.h = lib().ptb_scene_create()
        if not self.h:
            raise capi.PtbError("ptb_scene_create failed")
        self.camera = None
        self.committed_on = None
        self.commit_ms = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().ptb_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- setters (mirror the C ABI one to one) ------------------------------------------------
    def set_textures(self, rows):
        arr = (capi.Texture * len(rows))(*rows)
        check(lib().ptb_scene_set_textures(self.h, arr, len(rows)))

    def set_materials(self, rows):
        arr = (capi.Material * len(rows))(*rows)
        check(lib().ptb_scene_set_materials(self.h, arr, len(rows)))

    def set_spheres(self, xs, ys, zs, rs, material=None):
        xs, ys, zs, rs = (np.ascontiguousarray(a, dtype=np.float64) for a in (xs, ys, zs, rs))
        m = None if material is None else np.ascontiguousarray(material, dtype=np.int32)
        check(lib().ptb_scene_set_spheres(self.h, dptr(xs), dptr(ys), dptr(zs), dptr(rs),
                                          None if m is None else iptr(m), len(rs)))

    def set_triangles(self, vx, vy, vz, indices, material=None, uv=None):
        vx, vy, vz = (np.ascontiguousarray(a, dtype=np.float64) for a in (vx, vy, vz))
        idx = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1)
        m = None if material is None else np.ascontiguousarray(material, dtype=np.int32)
        u = None if uv is None else np.ascontiguousarray(uv, dtype=np.float64).reshape(-1)
        check(lib().ptb_scene_set_triangles(self.h, dptr(vx), dptr(vy), dptr(vz), len(vx), iptr(idx),
                                            None if m is None else iptr(m), None if u is None else dptr(u),
                                            len(idx) // 3))

    def set_background(self, kind, c0, c1=None):
        a = np.asarray(c0, dtype=np.float64)
        b = None if c1 is None else np.asarray(c1, dtype=np.float64)
        check(lib().ptb_scene_set_background(self.h, kind, dptr(a), None if b is None else dptr(b)))

    def set_light_quad(self, origin, u, v):
        """Extension: diffuse_plus_light = Mix (Diffuse, Quad_light) (ptb200.h); origin=None removes the light."""
        if origin is None:
            check(lib().ptb_scene_set_light_quad(self.h, None, None, None))
            return
        o, a, b = (np.asarray(x, dtype=np.float64) for x in (origin, u, v))
        check(lib().ptb_scene_set_light_quad(self.h, dptr(o), dptr(a), dptr(b)))

    def commit(self, device=0):
        ms = C.c_double(0)
        check(lib().ptb_scene_commit(self.h, device, C.byref(ms)))
        self.committed_on, self.commit_ms = device, ms.value
        return ms.value

    def commit_multi(self, n_devices):
        """Replicates the committed scene on devices 0..n_devices-1 (single-process multi-GPU)."""
        ms = C.c_double(0)
        check(lib().ptb_scene_commit_multi(self.h, n_devices, C.byref(ms)))
        self.committed_on, self.commit_ms, self.n_devices = 0, ms.value, n_devices
        return ms.value

    def tree_stats(self):
        out = np.zeros(8, dtype=np.int32)
        check(lib().ptb_scene_tree_stats(self.h, iptr(out)))
        return dict(zip(("nodes", "depth", "max_stack", "spheres", "triangles", "leaves", "leaf_max"), map(int, out)))

    # ---- read-back --------------------------------------------------------------------------------
    def tables(self):
        L = lib()
        ns, nv, nt = C.c_int64(), C.c_int64(), C.c_int64()
        nm, ntex = C.c_int32(), C.c_int32()
        check(L.ptb_scene_counts(self.h, C.byref(ns), C.byref(nv), C.byref(nt), C.byref(nm), C.byref(ntex)))
        ns, nv, nt, nm, ntex = ns.value, nv.value, nt.value, nm.value, ntex.value
        t = {"n_spheres": ns, "n_vertices": nv, "n_triangles": nt}
        xs, ys, zs, rs = (np.zeros(ns) for _ in range(4))
        sm = np.zeros(ns, dtype=np.int32)
        if ns:
            check(L.ptb_scene_get_spheres(self.h, dptr(xs), dptr(ys), dptr(zs), dptr(rs), iptr(sm)))
        t.update(xs=xs, ys=ys, zs=zs, rs=rs, sphere_material=sm)
        vx, vy, vz = (np.zeros(nv) for _ in range(3))
        idx, tm, uv = np.zeros(3 * nt, dtype=np.int32), np.zeros(nt, dtype=np.int32), np.zeros(6 * nt)
        if nt:
            check(L.ptb_scene_get_triangles(self.h, dptr(vx), dptr(vy), dptr(vz), iptr(idx), iptr(tm), dptr(uv)))
        t.update(vx=vx, vy=vy, vz=vz, indices=idx, tri_material=tm, uv=uv)
        mats, texs = (capi.Material * max(nm, 1))(), (capi.Texture * max(ntex, 1))()
        check(L.ptb_scene_get_materials(self.h, mats, texs))
        t.update(materials=mats, n_materials=nm, textures=texs, n_textures=ntex)
        kind = C.c_int32()
        c0, c1 = np.zeros(3), np.zeros(3)
        check(L.ptb_scene_get_background(self.h, C.byref(kind), dptr(c0), dptr(c1)))
        t.update(bg_kind=kind.value, bg0=c0, bg1=c1)
        has = C.c_int32()
        lo, lu, lv = np.zeros(3), np.zeros(3), np.zeros(3)
        check(L.ptb_scene_get_light_quad(self.h, C.byref(has), dptr(lo), dptr(lu), dptr(lv)))
        t.update(has_light=bool(has.value), light_o=lo, light_u=lu, light_v=lv)
        n = L.ptb_scene_get_prim_order(self.h, None, 0)
        order = np.zeros(max(n, 1), dtype=np.int32)
        if n:
            L.ptb_scene_get_prim_order(self.h, iptr(order), n)
        t["prim_order"] = order[:n]
        return t


def shirley_spheres(width, height, seed=42):
    """shirley_spheres/bin/main.ml:26-110,250-260 (Random.init 42)."""
    s = Scene()
    cam = np.zeros(20)
    check(lib().ptb_scene_load_shirley(s.h, width / height, seed, dptr(cam)))
    s.camera = Camera(cam)
    return s


def cornell_box(width, height, background=("constant", (1.0, 1.0, 1.0), None)):
    """cornell-box/bin/main.ml geometry; the background is the caller's."""
    s = Scene()
    cam = np.zeros(20)
    kind = capi.PTB_BG_CONSTANT if background[0] == "constant" else capi.PTB_BG_GRADIENT_Y
    c0 = np.asarray(background[1], dtype=np.float64)
    c1 = None if background[2] is None else np.asarray(background[2], dtype=np.float64)
    check(lib().ptb_scene_load_cornell(s.h, width / height, kind, dptr(c0), None if c1 is None else dptr(c1),
                                       dptr(cam)))
    s.camera = Camera(cam)
    return s


def cornell_box_lit(width, height, radiance=(32.0, 32.0, 32.0)):
    """Extension (BASELINE.json configs[1] "diffuse+light sampling"): the cornell geometry in a closed black box,
    lit by an emissive square at the reference's light position, sampled through the mixture pdf."""
    s = Scene()
    cam = np.zeros(20)
    r = np.asarray(radiance, dtype=np.float64)
    check(lib().ptb_scene_load_cornell_lit(s.h, width / height, dptr(r), dptr(cam)))
    s.camera = Camera(cam)
    return s


def synthetic_mesh(target_faces, seed=0xB200):
    nv, nf = C.c_int64(), C.c_int64()
    check(lib().ptb_mesh_synthetic(target_faces, seed, None, 0, None, 0, C.byref(nv), C.byref(nf)))
    xyz = np.zeros(3 * nv.value, dtype=np.float32)
    faces = np.zeros(3 * nf.value, dtype=np.int32)
    check(lib().ptb_mesh_synthetic(target_faces, seed, fptr(xyz), nv.value, iptr(faces), nf.value, C.byref(nv),
                                   C.byref(nf)))
    return xyz, faces


def mesh_scene(xyz, faces, width, height):
    """ganesha/bin/main.ml assembly around an indexed mesh given in world space."""
    s = Scene()
    cam = np.zeros(20)
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1)
    faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1)
    check(lib().ptb_scene_load_mesh(s.h, fptr(xyz), len(xyz) // 3, iptr(faces), len(faces) // 3, width / height,
                                    dptr(cam)))
    s.camera = Camera(cam)
    return s


def synthetic_mesh_scene(target_faces, width, height, seed=0xB200):
    xyz, faces = synthetic_mesh(target_faces, seed)
    return mesh_scene(xyz, faces, width, height)


def read_ply_mesh(path_or_bytes):
    """Ply.of_bigstring subset + ganesha's Mesh.create access pattern (ply_format/src/ply.ml:340-352,
    ganesha/bin/main.ml:50-60): returns (xyz float32[nv,3], faces int32[nf,3]) of a binary-little-endian PLY."""
    xyz, faces = _fp_t(), _ip_t()
    nv, nf = C.c_int64(), C.c_int64()
    if isinstance(path_or_bytes, (bytes, bytearray)):
        b = bytes(path_or_bytes)
        check(lib().ptb_ply_parse_mesh(b, len(b), C.byref(xyz), C.byref(nv), C.byref(faces), C.byref(nf)))
    else:
        check(lib().ptb_ply_read_mesh(str(path_or_bytes).encode(), C.byref(xyz), C.byref(nv), C.byref(faces),
                                      C.byref(nf)))
    try:
        v = np.ctypeslib.as_array(xyz, shape=(max(nv.value, 1) * 3,))[:nv.value * 3].copy().reshape(-1, 3)
        f = np.ctypeslib.as_array(faces, shape=(max(nf.value, 1) * 3,))[:nf.value * 3].copy().reshape(-1, 3)
    finally:
        lib().ptb_free(xyz)
        lib().ptb_free(faces)
    return v, f


def ganesha(ply_path, width, height):
    """ganesha/bin/main.ml: the PLY mesh + floor + camera, through the restated ply_format subset."""
    xyz, faces = read_ply_mesh(ply_path)
    return mesh_scene(xyz, faces, width, height)


def write_ply_mesh(path, xyz, faces, vertex_type="float", count_type="uchar", index_type="int", extra_vertex_props=()):
    """Test/tooling helper: a binary-little-endian PLY with a `vertex` element (x,y,z [+ extra float props]) and
    a `face` element holding the list property `vertex_indices` — the layout of pbrt-v3-scenes' ganesha.ply."""
    xyz = np.asarray(xyz).reshape(-1, 3)
    faces = np.asarray(faces).reshape(-1, 3)
    vt = {"float": "<f4", "double": "<f8"}[vertex_type]
    ct = {"uchar": "u1", "uint8": "u1", "int": "<i4", "ushort": "<u2"}[count_type]
    it = {"int": "<i4", "uint": "<u4", "ushort": "<u2", "short": "<i2"}[index_type]
    hdr = ["ply", "format binary_little_endian 1.0", "comment written by path_tracer_ocaml_b200",
           f"element vertex {len(xyz)}"] + [f"property {vertex_type} {n}" for n in ("x", "y", "z")]
    hdr += [f"property float {n}" for n in extra_vertex_props]
    hdr += [f"element face {len(faces)}", f"property list {count_type} {index_type} vertex_indices", "end_header"]
    vrec = np.zeros(len(xyz), dtype=[("p", vt, 3)] + [(n, "<f4") for n in extra_vertex_props])
    vrec["p"] = xyz
    frec = np.zeros(len(faces), dtype=[("n", ct), ("i", it, 3)])
    frec["n"] = 3
    frec["i"] = faces
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode())
        f.write(vrec.tobytes())
        f.write(frec.tobytes())
