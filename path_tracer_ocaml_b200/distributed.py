"""Image-tile sharding across ranks: rank r renders the tiles t of the reference's Tile.split list
with t mod world == r into a zeroed full-frame buffer of per-pixel sample sums.  The only exchange of the whole
render follows, and it is cut so that no rank does serial work for the others:

  1. reduce-scatter of the sums by ROW BANDS (rank k ends up with the summed rows of band k; this also merges the
     1-pixel filter halos between tiles of different ranks),
  2. neighbouring ranks swap one boundary row (the 3x3 reconstruction filter of film_tile.ml:23-38 reaches one row
     into the next band),
  3. every rank resolves (filter + gamma, integrator.ml:114-128,152-154) ITS band,
  4. the resolved bands are gathered on rank 0 — or, when the image is wanted in host memory, every rank copies its
     band straight into one shared page-locked host image over its own PCIe link (SharedHostImage).

`render_sharded` (one reduce to rank 0, resolve there) is the simple form, kept for the CPU tests and as the
reference point; both run over NCCL on GPUs and over gloo on CPU (where the oracle stands in for the device).
"""
import os

import torch
import torch.distributed as dist


def render_sharded(render_sums, resolve, group=None, dst=0):
    """render_sums(rank, world) -> tensor of per-pixel sums (this rank's tiles only, zero elsewhere);
    resolve(sums) -> image, called on rank `dst` only.  Returns the image on dst, None elsewhere."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    sums = render_sums(rank, world)
    if world > 1:
        dist.reduce(sums, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return resolve(sums) if rank == dst else None


def band_rows(height, world):
    """Rows per band: the frame is padded to world * rows so that every band has the same size."""
    return -(-height // world)


def alloc_padded_sums(height, width, world, device, dtype=torch.float32):
    """Zeroed sums buffer of world * band_rows rows; its first `height` rows are the frame a renderer writes."""
    return torch.zeros(band_rows(height, world) * world, width, 3, dtype=dtype, device=device)


def render_sharded_bands(sums_padded, height, resolve_rows, group=None, gather_to=0, timers=None):
    """sums_padded: this rank's per-pixel sums in a buffer from alloc_padded_sums (already rendered).
    resolve_rows(local) -> image of `local`, a (rows + 2, W, 3) block of summed rows with one halo row above and
    below (zeros at the frame's top and bottom: the filter drops taps outside the image, integrator.ml:115-117).
    Returns (band_image, full_image): this rank's resolved band (rows, W, 3) and, on rank `gather_to`, the whole
    (height, W, 3) image (None elsewhere, and None everywhere if gather_to is None).
    timers: optional dict that receives (start, end) event pairs per phase when the tensors are on a GPU."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rows = band_rows(height, world)
    Wd = sums_padded.shape[1]
    cuda = sums_padded.is_cuda

    def mark(name, which):
        if timers is not None and cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            timers.setdefault(name, [None, None])[which] = ev

    local = torch.zeros(rows + 2, Wd, 3, dtype=sums_padded.dtype, device=sums_padded.device)
    band = local[1:rows + 1]
    mark("reduce", 0)
    if world > 1:
        dist.reduce_scatter_tensor(band, sums_padded, op=dist.ReduceOp.SUM, group=group)
        ops = []
        if rank > 0:  # my first row goes up, the row above my band comes down
            ops += [dist.P2POp(dist.isend, local[1], rank - 1, group), dist.P2POp(dist.irecv, local[0], rank - 1, group)]
        if rank < world - 1:
            ops += [dist.P2POp(dist.isend, local[rows], rank + 1, group), dist.P2POp(dist.irecv, local[rows + 1], rank + 1, group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    else:
        band.copy_(sums_padded[:rows])
    mark("reduce", 1)
    mark("resolve", 0)
    band_img = resolve_rows(local)[1:rows + 1]
    mark("resolve", 1)
    full = None
    mark("gather", 0)
    if gather_to is not None:
        if world > 1:
            parts = None
            if rank == gather_to:
                whole = torch.empty(rows * world, Wd, 3, dtype=band_img.dtype, device=band_img.device)
                parts = list(whole.view(world, rows, Wd, 3).unbind(0))
            dist.gather(band_img.contiguous(), parts, dst=gather_to, group=group)
            if rank == gather_to:
                full = whole[:height]
        else:
            full = band_img[:height]
    mark("gather", 1)
    return band_img, full


class SharedHostImage:
    """One page-locked host image shared by the ranks of a node: rank 0 creates a file in /dev/shm, every rank maps it
    and registers the mapping with CUDA, and each rank copies its resolved band into its rows — eight 12 MB copies over
    eight PCIe links instead of one 100 MB copy over rank 0's."""

    def __init__(self, height, width, world, rank, dtype=torch.float32, tag=None, group=None):
        self.rows = band_rows(height, world)
        self.height, self.width, self.world, self.rank = height, width, world, rank
        tag = tag or os.environ.get("MASTER_PORT", "0")
        self.path = f"/dev/shm/ptb200_image_{tag}_{os.getuid()}"
        n = self.rows * world * width * 3
        if rank == 0:
            torch.from_file(self.path, shared=True, size=n, dtype=dtype).zero_()
        if world > 1:
            dist.barrier(group)
        self.tensor = torch.from_file(self.path, shared=True, size=n, dtype=dtype).view(self.rows * world, width, 3)
        self.registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), self.tensor.numel() * self.tensor.element_size(), 0)
            self.registered = int(rc) == 0
        if world > 1:
            dist.barrier(group)
        if rank == 0:
            os.unlink(self.path)  # the mappings keep it alive; nothing is left behind in /dev/shm

    def put_band(self, band_img, non_blocking=True):
        self.tensor[self.rank * self.rows:(self.rank + 1) * self.rows].copy_(band_img, non_blocking=non_blocking and self.registered)

    def image(self):
        return self.tensor[:self.height]

    def close(self):
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self.registered = False


def render_sharded_gpu(scene, width, height, samples_per_pixel, max_bounces, device, flags=0):
    """The production multi-GPU render: one process per GPU over NCCL; returns the image on rank 0."""
    from .integrator import Integrator

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    integ = Integrator(scene, width, height, samples_per_pixel, max_bounces, device=device, tile_rank=rank,
                       tile_world=world)
    sums = alloc_padded_sums(height, width, world, f"cuda:{device}")
    integ.render_device(sums, flags=flags)
    _, img = render_sharded_bands(sums, height, lambda loc: integ.resolve_rows_device(loc, flags=flags))
    return img, integ.stats
