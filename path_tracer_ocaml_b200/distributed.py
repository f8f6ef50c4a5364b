"""Image-tile sharding across ranks: rank r renders the tiles t of the reference's
Tile.split list with t mod world == r into a zeroed full-frame buffer of per-pixel sample sums; ONE
reduce(sum) to rank 0 merges the shards, then rank 0 runs the filter+gamma resolve once.  No
collective touches the data path before that.  The same function drives NCCL on GPUs and gloo in the
CPU tests (where the per-rank renderer is the oracle)."""
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


def render_sharded_gpu(scene, width, height, samples_per_pixel, max_bounces, device, flags=0):
    """The production multi-GPU render: one process per GPU, NCCL reduce of float32[3*W*H]."""
    from .integrator import Integrator

    holder = {}

    def render_sums(rank, world):
        integ = Integrator(scene, width, height, samples_per_pixel, max_bounces, device=device, tile_rank=rank,
                           tile_world=world)
        holder["integ"] = integ
        sums = torch.zeros(height, width, 3, dtype=torch.float32, device=f"cuda:{device}")
        integ.render_device(sums, flags=flags)
        return sums

    def resolve(sums):
        return holder["integ"].resolve_device(sums, flags=flags)

    img = render_sharded(render_sums, resolve)
    return img, holder["integ"].stats
