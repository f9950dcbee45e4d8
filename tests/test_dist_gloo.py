"""world_size-2 gloo test (CPU) of the N>1 path's host logic: shard bounds, the ragged all_gather and
global-image-index seeding.  The compute op is injected: on CPU the oracle stands in for the CUDA
operator (tests may use the oracle as the checker; the product never does)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_op(mask, vertex, hn, inlier_thresh=0.999, min_num=5, max_num=30000, seed=0, img_base=0):
    import pvnet_oracle
    out = pvnet_oracle.ransac_voting_layer_v3(mask.numpy(), vertex.numpy(), hn, inlier_thresh=inlier_thresh,
                                              min_num=min_num, max_num=max_num, seed=seed, img_base=img_base)
    return torch.from_numpy(out)


def _oracle_dist_op(mask, vertex, mean, round_hyp_num=256, min_hyp_num=4096, inlier_thresh=0.99, min_num=5, max_num=30000,
                    seed=0, img_base=0):
    import pvnet_oracle
    _, cov = pvnet_oracle.estimate_voting_distribution_with_mean(mask.numpy(), vertex.numpy(), mean.numpy(), round_hyp_num=round_hyp_num,
                                                                  min_hyp_num=min_hyp_num, inlier_thresh=inlier_thresh, min_num=min_num,
                                                                  max_num=max_num, seed=seed, img_base=img_base)
    return mean, torch.from_numpy(cov)


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clean_pvnet_b200 import parallel, synth
        mask, vertex, _ = synth.make_inputs("tiny", device="cpu", seed=77, B=total)
        lo, hi = parallel.shard_bounds(total, world, rank)
        out = parallel.sharded_ransac_voting_layer_v3(mask[lo:hi], vertex[lo:hi], 16, total, inlier_thresh=0.99,
                                                      max_num=300, seed=5, op=_oracle_op)
        if rank == 0:
            q.put(out.numpy())
    finally:
        dist.destroy_process_group()


def _layer_worker(rank, world, port, total, steps, depth, q):
    """ShardedVotingLayer over gloo: `steps` pipelined calls with at most `depth` in flight, results asked for out of
    order -- the same bookkeeping (sequence numbers, deferred completion, ragged un-padding) the peer path uses on GPUs."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clean_pvnet_b200 import parallel, synth
        mask, vertex, _ = synth.make_inputs("tiny", device="cpu", seed=78, B=total)
        lo, hi = parallel.shard_bounds(total, world, rank)
        layer = parallel.ShardedVotingLayer(total, vertex.shape[3], depth=depth, gather="collective", op=_oracle_op)
        assert layer.mode == "collective"
        pend = [layer(mask[lo:hi], vertex[lo:hi], 16, inlier_thresh=0.99, max_num=300, seed=100 + s) for s in range(steps)]
        assert len(layer.inflight) <= depth                    # older calls were completed by the layer itself
        assert [p.seq for p in pend] == list(range(1, steps + 1))
        outs = [None] * steps
        for s in reversed(range(steps)):                       # out of order on purpose
            outs[s] = pend[s].result().numpy()
        # the other half of the un_pnp pair on its own channel: covariances of the whole batch on every rank
        cov = layer.distribution(mask[lo:hi], vertex[lo:hi], pend[0].local, round_hyp_num=16, min_hyp_num=64, max_num=300, seed=7,
                                 op=_oracle_dist_op)
        covs = cov.result().numpy()
        layer.check()
        assert not layer.inflight
        assert np.array_equal(pend[0].local.numpy(), outs[0][lo:hi])
        assert np.array_equal(cov.local.numpy(), covs[lo:hi]) and covs.shape == (total, vertex.shape[3], 2, 2)
        if rank == 0:
            q.put((np.stack(outs), covs))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("total,depth", [(5, 2), (4, 4)])
def test_sharded_layer_pipelined_calls(total, depth):
    steps = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_layer_worker, args=(r, 2, port, total, steps, depth, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, covs = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from clean_pvnet_b200 import synth
    mask, vertex, _ = synth.make_inputs("tiny", device="cpu", seed=78, B=total)
    for s in range(steps):
        want = _oracle_op(mask, vertex, 16, inlier_thresh=0.99, max_num=300, seed=100 + s).numpy()
        assert np.array_equal(got[s], want), s
    _, want_cov = _oracle_dist_op(mask, vertex, torch.from_numpy(got[0]), round_hyp_num=16, min_hyp_num=64, max_num=300, seed=7)
    assert np.array_equal(covs, want_cov.numpy())


def test_shard_bounds():
    from clean_pvnet_b200.parallel import shard_bounds
    for total in (0, 1, 5, 16, 128, 131):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gather_equals_single_process():
    total = 5       # ragged: ranks get 3 and 2 images
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from clean_pvnet_b200 import synth
    mask, vertex, _ = synth.make_inputs("tiny", device="cpu", seed=77, B=total)
    want = _oracle_op(mask, vertex, 16, inlier_thresh=0.99, max_num=300, seed=5).numpy()
    assert got.shape == (total, vertex.shape[3], 2)
    assert np.array_equal(got, want)
