"""Integrator / Render_command mirror (path_tracer/src/integrator.mli:4-16;
render_command/src/render_command.ml:6-47) over the C ABI.  torch is used only for device memory,
streams and torch.distributed — the arithmetic is all inside libptb200."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import check, dptr, lib


@dataclass
class Args:
    """Render_command.Args.t (render_command.ml:6-14) + the device flag the port adds."""
    width: int
    height: int
    samples_per_pixel: int = 1       # --samples-per-pixel default (render_command.ml:34)
    output: str = "output.png"       # --output default (render_command.ml:30)
    no_progress: bool = False
    max_bounces: int = 8             # --max-ray-bounces default (render_command.ml:42)
    device: int = 0


class Integrator:
    """Integrator.create ~width ~height ~samples_per_pixel ~max_bounces ~camera + a committed scene
    instead of the `intersect`/`background` closures (integrator.ml:71-87)."""

    def __init__(self, scene, width, height, samples_per_pixel, max_bounces, camera=None, device=0,
                 tile_rank=0, tile_world=1):
        self.scene = scene
        cam = camera or scene.camera
        self.params = capi.Params(width=width, height=height, samples_per_pixel=samples_per_pixel,
                                  max_bounces=max_bounces, lower_left_x=cam.lower_left_x,
                                  lower_left_y=cam.lower_left_y, view_x=cam.view_x, view_y=cam.view_y,
                                  tile_rank=tile_rank, tile_world=tile_world, flags=0, device=device)
        self.stats = capi.Stats()
        if scene.committed_on != device:
            scene.commit(device)

    @classmethod
    def create(cls, **kw):
        return cls(**kw)

    def _p(self, flags, passes=None):
        p = capi.Params.from_buffer_copy(self.params)
        p.flags = flags
        if passes is not None:
            p.pass_first, p.pass_count = passes
        return p

    # Integrator.render with a HOST image (float64, 3*W*H, Bimage layout) — copies included.
    # passes = (first, count): only those sample passes, image normalised by `count` (ptb_params.pass_first/pass_count)
    def render(self, flags=0, image=None, passes=None):
        p = self._p(flags, passes)
        if image is None:
            image = np.empty((p.height, p.width, 3), dtype=np.float64)
        check(lib().ptb_render(self.scene.h, C.byref(p), dptr(image), C.byref(self.stats)))
        return image

    def render_progressive(self, every, on_preview=None, flags=0):
        """The render in batches of `every` sample passes.  After each batch `on_preview(image, passes_done)` gets the
        image of the passes finished so far (filtered, gamma applied) — the progressive output of the reference's
        photon-map binary (progressive_photon_map.ml:447-449, the PNG rewritten per iteration) offered for the path
        tracer.  Returns the final image: the same samples as render(), combined in float64 on the host."""
        spp = self.params.samples_per_pixel
        acc = np.zeros((self.params.height, self.params.width, 3), dtype=np.float64)
        rays = paths = 0
        for first in range(0, spp, every):
            count = min(every, spp - first)
            acc += self.render(flags=flags | capi.PTB_FLAG_RAW_SUMS, passes=(first, count))  # filtered sums of the batch
            rays, paths = rays + self.stats.rays, paths + self.stats.paths
            raw = flags & (capi.PTB_FLAG_RAW_SUMS | capi.PTB_FLAG_NO_FILTER)
            image = acc.copy() if raw else np.sqrt(acc / (first + count))  # integrator.ml:152-154
            if on_preview is not None:
                on_preview(image, first + count)
        self.stats.rays, self.stats.paths = rays, paths
        return image

    # single-process multi-GPU: tiles dealt to n_devices GPUs, peer-memory reduce on device 0
    def render_multi(self, n_devices, flags=0, image=None):
        p = self._p(flags)
        if getattr(self.scene, "n_devices", 1) < n_devices:
            self.scene.commit_multi(n_devices)
        if image is None:
            image = np.empty((p.height, p.width, 3), dtype=np.float64)
        check(lib().ptb_render_multi(self.scene.h, C.byref(p), n_devices, dptr(image), C.byref(self.stats)))
        return image

    # device-resident path: adds this rank's per-pixel sums into a torch float32 CUDA tensor
    def render_device(self, sums, flags=0, stream=None, passes=None):
        import torch
        assert sums.is_cuda and sums.dtype == torch.float32 and sums.is_contiguous()
        p = self._p(flags, passes)
        st = torch.cuda.current_stream(sums.device) if stream is None else stream
        check(lib().ptb_render_device(self.scene.h, C.byref(p), C.c_void_p(sums.data_ptr()),
                                      C.c_void_p(st.cuda_stream), C.byref(self.stats)))
        return sums

    def resolve_device(self, sums, out=None, flags=0, stream=None):
        import torch
        p = self.params
        if out is None:
            out = torch.empty_like(sums)
        st = torch.cuda.current_stream(sums.device) if stream is None else stream
        check(lib().ptb_resolve_device(C.c_void_p(sums.data_ptr()), C.c_void_p(out.data_ptr()), p.width, p.height,
                                       p.samples_per_pixel, flags, p.device, C.c_void_p(st.cuda_stream)))
        return out

    def resolve_rows_device(self, local, out=None, flags=0, stream=None):
        """Filter + gamma of a block of summed rows (rows, W, 3) — a band of the frame with its halo rows — as an
        image of that height: what each rank runs on its own band after the reduce-scatter (distributed.py)."""
        import torch
        p = self.params
        if out is None:
            out = torch.empty_like(local)
        st = torch.cuda.current_stream(local.device) if stream is None else stream
        check(lib().ptb_resolve_device(C.c_void_p(local.data_ptr()), C.c_void_p(out.data_ptr()), p.width, local.shape[0],
                                       p.samples_per_pixel, flags, p.device, C.c_void_p(st.cuda_stream)))
        return out

    def first_hit(self):
        p = self._p(0)
        t = np.empty((p.height, p.width), dtype=np.float32)
        prim = np.empty((p.height, p.width), dtype=np.int32)
        check(lib().ptb_first_hit(self.scene.h, C.byref(p), capi.fptr(t), capi.iptr(prim)))
        return t, prim

    def raygen(self, first, n):
        p = self._p(0)
        pixel, offset = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        cx, cy = np.empty(n), np.empty(n)
        d = np.empty((n, 3), dtype=np.float32)
        check(lib().ptb_raygen(C.byref(p), first, n, capi.iptr(pixel), capi.iptr(offset), dptr(cx), dptr(cy),
                               capi.fptr(d)))
        return pixel, offset, cx, cy, d


def intersect_batch(scene, origins, directions, t_min=0.0, t_max=3.4028234663852886e38, device=0, stats=None):
    """Batched spheres_intersect_native (shirley_spheres/bin/main.ml:162-170): nearest t (NaN = miss)
    and primitive index (-1 = miss) per ray."""
    o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1)
    d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1)
    n = len(o) // 3
    t = np.empty(n, dtype=np.float32)
    prim = np.empty(n, dtype=np.int32)
    st = stats if stats is not None else capi.Stats()
    if scene.committed_on != device:
        scene.commit(device)
    check(lib().ptb_intersect_batch(scene.h, capi.fptr(o), capi.fptr(d), t_min, t_max, n, capi.fptr(t),
                                    capi.iptr(prim), device, C.byref(st)))
    return t, prim


def r2_stream(max_bounces, offsets, device=0):
    off = np.ascontiguousarray(offsets, dtype=np.int32)
    D = 2 + 2 * max_bounces
    out = np.empty((len(off), D), dtype=np.float64)
    check(lib().ptb_r2_stream(max_bounces, capi.iptr(off), len(off), dptr(out), device))
    return out
