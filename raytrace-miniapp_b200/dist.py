"""Image-tile sharding of one create_image call over the ranks of a torch.distributed group.

One process per GPU.  The unit of sharding is the source pixel (rows of the image are the
sharded axis): rank r traces the contiguous pixel range tile_bounds(n, world, r) and owns those
image rows exclusively (ASE).  The exchange step that follows the kernels is the one the full
application performs over MPI (intensity_step_struct::sum_reduce,
src/RayTraceStructures.cpp:1603-1646), restated for device-resident partials:
  row-cyclic ASE (default): every rank writes its rows compactly, all_gather of the compact rows
                           (1/world of the image per rank), un-permute into the image,
                           all_reduce(sum) of I_ang
  contiguous tiles, ASE   : all_gather of the equal-sized image tiles + all_reduce(sum) of I_ang
  seeded                  : all_reduce(sum) of the full image and of I_ang (scatter binning)
over NCCL / NVLink on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def tile_bounds(n_pixels, world, rank):
    """Equal contiguous tiles of ceil(n/world) pixels; the last ones may be short or empty."""
    per = (n_pixels + world - 1) // world
    lo = min(n_pixels, rank * per)
    return lo, min(n_pixels, lo + per), per


def exchange(image, I_ang, n_pixels, nv, method, group=None):
    """In-place exchange of the partial results.  image: flat [n_pixels*nv] tensor holding this
    rank's owned rows (ASE) or its partial sums (seeded); I_ang: flat partial sums."""
    world = dist.get_world_size(group)
    if world == 1:
        return image, I_ang
    if method == 1:
        rank = dist.get_rank(group)
        lo, hi, per = tile_bounds(n_pixels, world, rank)
        if per * world == n_pixels:
            dist.all_gather_into_tensor(image, image[lo * nv:hi * nv].clone(), group=group)
        else:  # ragged: gather padded tiles, then trim
            pad = torch.zeros(per * nv, dtype=image.dtype, device=image.device)
            pad[:(hi - lo) * nv] = image[lo * nv:hi * nv]
            full = torch.empty(world * per * nv, dtype=image.dtype, device=image.device)
            dist.all_gather_into_tensor(full, pad, group=group)
            image.copy_(full[:n_pixels * nv])
    else:
        dist.all_reduce(image, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(I_ang, op=dist.ReduceOp.SUM, group=group)
    return image, I_ang


def exchange_rows(image, I_ang, group=None):
    """Exchange for the row-cyclic decomposition: every rank holds a full-size image with only
    its own rows filled (the others zero), so the gather is a sum; x + 0 is exact, hence the
    result is bit-identical to the single-device image in ASE mode."""
    if dist.get_world_size(group) > 1:
        dist.all_reduce(image, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(I_ang, op=dist.ReduceOp.SUM, group=group)
    return image, I_ang


def rows_per_rank(n_rows, world):
    return (n_rows + world - 1) // world


def unpermute_rows(gathered, image, n_rows, row_elems, world):
    """Host-side statement of rtb200_unpermute_rows for identity pixel maps (CPU tests): block r
    of `gathered` holds the rows r, r + world, ... of the image, compactly."""
    per = rows_per_rank(n_rows, world)
    g = gathered.view(world, per, row_elems)
    img = image.view(n_rows, row_elems)
    for r in range(world):
        mine = len(range(r, n_rows, world))
        img[r::world] = g[r, :mine]
    return image


class RowGather:
    """Buffers of the row-cyclic ASE exchange (allocated once, reused every step)."""

    def __init__(self, info, world, device):
        self.per = rows_per_rank(info["sny"], world)
        n = self.per * info["snx"] * info["nv"]
        self.part = torch.empty(n, dtype=torch.float64, device=device)
        self.gathered = torch.empty(n * world, dtype=torch.float64, device=device)


def sharded_create_image(ctx, problem, image, I_ang, group=None, stream=None, cyclic=True, rows=None):
    """Stage-free step: `ctx` already holds the staged problem.  Zeroes the buffers, traces this
    rank's share on `stream` (default: torch's current stream) and exchanges.  Asynchronous.
    cyclic=True: image rows rank, rank + world, ... (balanced).  With `rows` (a RowGather; ASE
    traced by pixel owners) each rank writes its rows compactly and the exchange is an
    all_gather of 1/world of the image per rank + un-permute; without it every rank fills a
    full-size buffer and the exchange is an all_reduce (seeded: the partials overlap).
    cyclic=False: one contiguous tile per rank (exchange = all_gather of the tiles)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    I_ang.zero_()
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    if cyclic and rows is not None:
        ctx.launch_rows_compact(rank, world, rows.part, I_ang, stream=st)
        dist.all_gather_into_tensor(rows.gathered, rows.part, group=group)
        image.zero_()
        ctx.unpermute_rows(rows.gathered, world, rows.per, image, stream=st)
        dist.all_reduce(I_ang, op=dist.ReduceOp.SUM, group=group)
        return image, I_ang
    image.zero_()
    if cyclic:
        ctx.launch_rows(rank, world, image, I_ang, stream=st)
        return exchange_rows(image, I_ang, group)
    n = ctx.staged_pixels
    lo, hi, _ = tile_bounds(n, world, rank)
    ctx.launch(lo, hi, image, I_ang, stream=st)
    return exchange(image, I_ang, n, problem.euv_beam.nv, problem.method, group)
