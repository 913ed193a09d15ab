"""Multi-GPU decomposition of a render.

The path shards by SAMPLE INDEX: samples are independent given (pixel, sample index, seed)
(reference init_sampler_state, pt.cu:378-399), so every rank holds the whole scene + BVH,
renders a contiguous slice of sample indices of every pixel into SUM accumulators
(Renderer.set_film_mode("sum") + set_sample_offset(first)), and ONE reduce of the
accumulation buffers (NCCL over NVLink on the GPUs; gloo in the CPU tests) followed by a
division by the total sample count yields the same image as a single-GPU render of all
samples, up to fp32 summation order.  There is no per-bounce communication.

Slices are multiples of 16 samples whenever possible so that each 4x4 CMJ pattern
(cmj.cu:71-80: index = n_spp % 16, scramble hashed from n_spp / 16) stays on one rank.

Multi-frame batches (rtcamp8-style, BASELINE config 5) shard by FRAME instead:
frame f -> rank f mod world, no collective.
"""

CMJ_PATTERN = 16


def sample_slice(total_spp, rank, world):
    """(first_sample, n_samples) of `rank`: contiguous, disjoint, covering [0, total_spp)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    blocks = (total_spp + CMJ_PATTERN - 1) // CMJ_PATTERN
    b0 = rank * blocks // world
    b1 = (rank + 1) * blocks // world
    first = min(b0 * CMJ_PATTERN, total_spp)
    last = min(b1 * CMJ_PATTERN, total_spp)
    return first, last - first


def frames_for_rank(n_frames, rank, world):
    return list(range(rank, n_frames, world))


def reduce_film(dist, sums, total_spp, dst=0):
    """Single exchange step: sum the per-rank accumulators onto `dst` and turn the sums into
    means there.  `sums` is a torch tensor (cuda for NCCL, cpu for gloo); in-place."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(sums, dst=dst, op=dist.ReduceOp.SUM)
        if dist.get_rank() != dst:
            return sums
    sums.mul_(1.0 / float(total_spp))
    if sums.dim() >= 1 and sums.shape[-1] == 4:
        # every rank's film wrote alpha = 1, so the reduced alpha is world / total_spp: reset it (k_scale_layers does the same)
        sums[..., 3] = 1.0
    return sums


def render_sharded(render_slice_sums, dist, total_spp, dst=0):
    """render_slice_sums(first, n) -> tensor of per-pixel SUMS over samples [first, first+n).
    Returns the mean image on rank `dst` (other ranks get their reduced-away buffer)."""
    rank = dist.get_rank() if (dist is not None and dist.is_initialized()) else 0
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    first, n = sample_slice(total_spp, rank, world)
    sums = render_slice_sums(first, n)
    return reduce_film(dist, sums, total_spp, dst)


def init_core_communicator(renderer, dist):
    """Makes `renderer` one rank of the torch.distributed world INSIDE the C++ core (fr_comm_init ->
    ncclCommInitRank): rank 0's 128-byte communicator id travels through the process group, after that the
    sample-sharded render + reduce is one call, Renderer.render_sharded (fr_render_sharded).  Collective."""
    from . import api
    rank, world = dist.get_rank(), dist.get_world_size()
    ident = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    renderer.comm_init(ident[0], rank, world)
    return ident[0]
