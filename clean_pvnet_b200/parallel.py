"""Multi-GPU use of the voting layer: one process per GPU, images sharded across ranks.

The path is embarrassingly parallel over images: there is NO exchange inside the
algorithm, so ranks run the single-GPU op on their contiguous slice of the batch and the only
collective is an all_gather of the [B/G, K, 2] keypoints (+ [B/G, K, 2, 2] covariances) -- a few KB
over NCCL/NVLink.  The philox sampling stream is keyed by the GLOBAL image index (`img_base`), so
the gathered result is identical for any number of ranks.
"""
import torch
import torch.distributed as dist


def shard_bounds(total, world_size, rank):
    """Contiguous split of `total` images: the first (total % world) ranks get one extra."""
    base, rem = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_ragged(local, total, group=None, async_op=False):
    """all_gather of per-rank slices [n_r, ...] (n_r from shard_bounds) into [total, ...].

    async_op=True returns (finish, work): the collective is enqueued on NCCL's own stream (after the producer of
    `local` on the current stream) so it overlaps whatever the caller launches next; call `work.wait()` and then
    `finish()` to get the gathered tensor."""
    world = dist.get_world_size(group)
    if world == 1:
        return (lambda: local, None) if async_op else local
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    even = all(hi - lo == nmax for lo, hi in sizes)
    if even:
        pad = local.contiguous()
    else:
        pad = local.new_zeros((nmax,) + tuple(local.shape[1:]))
        pad[: local.shape[0]] = local
    out = local.new_empty((world * nmax,) + tuple(local.shape[1:]))
    if hasattr(dist, "all_gather_into_tensor") and local.is_cuda:
        work = dist.all_gather_into_tensor(out, pad, group=group, async_op=async_op)
    else:
        work = dist.all_gather(list(out.view(world, nmax, *local.shape[1:]).unbind(0)), pad, group=group,
                               async_op=async_op)

    def finish():
        if even:
            return out
        o = out.view(world, nmax, *local.shape[1:])
        return torch.cat([o[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)

    if async_op:
        return finish, work
    return finish()


def sharded_ransac_voting_layer_v3(mask_local, vertex_local, round_hyp_num, total_images, inlier_thresh=0.999,
                                   min_num=5, max_num=30000, seed=0, group=None, op=None):
    """Runs ransac_voting_layer_v3 on this rank's images (a shard_bounds slice of a `total_images`
    batch) and returns the keypoints of the WHOLE batch on every rank.  `op` defaults to the CUDA
    operator; tests inject a stand-in to exercise the sharding logic on CPU/gloo."""
    if op is None:
        from .ransac_voting_gpu import ransac_voting_layer_v3 as op
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(total_images, world, rank)
    assert mask_local.shape[0] == hi - lo, "local batch does not match shard_bounds"
    local = op(mask_local, vertex_local, round_hyp_num, inlier_thresh=inlier_thresh, min_num=min_num,
               max_num=max_num, seed=seed, img_base=lo)
    return all_gather_ragged(local, total_images, group)
