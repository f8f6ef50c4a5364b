"""Shared helpers for the parity tests."""
import ctypes as C

import numpy as np

import path_tracer_ocaml_b200 as P
from path_tracer_ocaml_b200 import capi
import pyoracle as O


def make_params(scene, W, H, spp, mb, rank=0, world=1, flags=0):
    c = scene.camera
    return capi.Params(width=W, height=H, samples_per_pixel=spp, max_bounces=mb, lower_left_x=c.lower_left_x,
                       lower_left_y=c.lower_left_y, view_x=c.view_x, view_y=c.view_y, tile_rank=rank,
                       tile_world=world, flags=flags, device=0)


def oracle_alpha(D):
    a = np.zeros(D)
    O.lib().orc_lds_alpha(D, O.dptr(a))
    return a


def oracle_lds(alpha, offsets):
    """Low_discrepancy_sequence.get for every (offset, dimension), vectorised with the SAME float64
    operations as oracle.cpp lds_get (mul, add, trunc, sub — numpy does not fuse)."""
    n1 = (1 + np.asarray(offsets, dtype=np.int64)).astype(np.float64)[:, None]
    x = 0.5 + alpha[None, :] * n1
    return x - np.trunc(x)


def image_metrics(img, ref):
    d = img - ref
    return {
        "rmse": float(np.sqrt(np.mean(d * d))),
        "within": float(np.mean(np.abs(d) <= 0.02 * np.abs(ref) + 1.0 / 255.0)),
        "bias": np.abs(d.mean((0, 1))).max(),
        "max": float(np.abs(d).max()),
    }


def resolve_numpy(sums, spp, weights3x3, raw=False):
    """Film_tile.write_pixel splat + stitch (film_tile.ml:23-38, integrator.ml:114-128) + gamma, in numpy:
    out[y+dy, x+dx] += w[dy+1, dx+1] * S[y, x], destinations outside the image dropped."""
    H, W, _ = sums.shape
    out = np.zeros_like(sums)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            w = weights3x3[dy + 1, dx + 1]
            ys0, ys1 = max(0, -dy), min(H, H - dy)
            xs0, xs1 = max(0, -dx), min(W, W - dx)
            out[ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx] += w * sums[ys0:ys1, xs0:xs1]
    return out if raw else np.sqrt(out * (1.0 / spp))
