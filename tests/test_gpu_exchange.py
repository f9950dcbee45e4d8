"""GPU: the multi-GPU result exchange (pvb_exchange, csrc/exchange.cu + the refit kernel's exchange tail) on ONE device.

Two (or three) "ranks" live in this process, each with its own receive ring in HBM; the rings are connected by raw base
pointers (pvb_exchange_connect_ptrs) instead of CUDA IPC handles, everything else -- the peer stores from the refit
kernel, the flags, the wait kernel, the ring schedule -- is the code that runs across GPUs.  bench.py --gpus N asserts the
same equality across real NVLink peers ("gather_check")."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(pvb, world, nmax, K, depth):
    from clean_pvnet_b200 import parallel
    exs = [parallel.PeerExchange(nmax * K * 8, 2 * depth, rank=r, world=world, device="cuda:0") for r in range(world)]
    for e in exs:
        e.connect_local(exs)
    return exs


def _gathered(ex, seq, world, nmax, K, sizes):
    from clean_pvnet_b200 import parallel
    buf = torch.empty(world * ex.bytes_per_rank, dtype=torch.uint8, device="cuda:0")
    ex.wait(seq, buf, timeout_s=5.0, floats_per_rank=[(hi - lo) * K * 2 for lo, hi in sizes])
    rows = buf.view(world, ex.bytes_per_rank)[:, : nmax * K * 8]
    out = rows.reshape(-1).view(torch.float32).view(world * nmax, K, 2)
    return parallel._unpad(out, sizes, nmax, (K, 2))


@pytest.mark.parametrize("world,total", [(2, 6), (3, 7)])
def test_pushed_results_equal_the_single_gpu_result(pvb, world, total):
    from clean_pvnet_b200 import parallel, synth
    mask, vertex, _ = synth.make_inputs("small", device="cuda:0", seed=31, B=total)
    K = vertex.shape[3]
    sizes = [parallel.shard_bounds(total, world, r) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    depth = 2
    exs = _make(pvb, world, nmax, K, depth)
    try:
        steps = 7                                   # > 2*depth: every slot of the ring is reused
        pending = []
        for s in range(1, steps + 1):
            if s > depth:                           # ring discipline: wait of call s-depth before call s, on every rank
                seq0 = s - depth
                want = pvb.ransac_voting_layer_v3(mask, vertex, 64, inlier_thresh=0.99, seed=500 + seq0, max_num=700)
                for r in range(world):
                    got = _gathered(exs[r], seq0, world, nmax, K, sizes)
                    assert torch.equal(got, want), (seq0, r)
                pending.remove(seq0)
            for r, (lo, hi) in enumerate(sizes):
                local = pvb.ransac_voting_layer_v3(mask[lo:hi], vertex[lo:hi], 64, inlier_thresh=0.99, seed=500 + s, max_num=700,
                                                   img_base=lo, _exchange=(exs[r].handle, s))
                assert local.shape == (hi - lo, K, 2)
            pending.append(s)
        for seq0 in pending:
            want = pvb.ransac_voting_layer_v3(mask, vertex, 64, inlier_thresh=0.99, seed=500 + seq0, max_num=700)
            for r in range(world):
                assert torch.equal(_gathered(exs[r], seq0, world, nmax, K, sizes), want), (seq0, r)
        for e in exs:
            e.check()
    finally:
        for e in exs:
            e.close()


def test_pushed_covariances_equal_the_single_gpu_result(pvb):
    """pvb_estimate_voting_distribution_push: 3 "ranks", ragged shards, 4 floats per (image, keypoint)."""
    from clean_pvnet_b200 import parallel, synth
    world, total = 3, 5
    mask, vertex, _ = synth.make_inputs("small", device="cuda:0", seed=35, B=total)
    K = vertex.shape[3]
    sizes = [parallel.shard_bounds(total, world, r) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    mean = pvb.ransac_voting_layer_v3(mask, vertex, 64, inlier_thresh=0.99, seed=1, max_num=700)
    _, want = pvb.estimate_voting_distribution_with_mean(mask, vertex, mean, round_hyp_num=64, min_hyp_num=256, max_num=700, seed=9)
    exs = [parallel.PeerExchange(nmax * K * 16, 4, rank=r, world=world, device="cuda:0") for r in range(world)]
    try:
        for e in exs:
            e.connect_local(exs)
        for r, (lo, hi) in enumerate(sizes):
            pvb.estimate_voting_distribution_with_mean(mask[lo:hi], vertex[lo:hi], mean[lo:hi], round_hyp_num=64, min_hyp_num=256,
                                                       max_num=700, seed=9, img_base=lo, _exchange=(exs[r].handle, 1))
        for r in range(world):
            buf = torch.empty(world * exs[r].bytes_per_rank, dtype=torch.uint8, device="cuda:0")
            exs[r].wait(1, buf, timeout_s=5.0, floats_per_rank=[(hi - lo) * K * 4 for lo, hi in sizes])
            rows = buf.view(world, exs[r].bytes_per_rank)[:, : nmax * K * 16]
            got = parallel._unpad(rows.reshape(-1).view(torch.float32).view(world * nmax, K, 2, 2), sizes, nmax, (K, 2, 2))
            assert torch.equal(got, want), r
            exs[r].check()
    finally:
        for e in exs:
            e.close()


def test_wait_times_out_instead_of_hanging(pvb):
    """A rank that never publishes: the wait kernel gives up after timeout_s, poisons its output and the status call raises."""
    from clean_pvnet_b200 import synth
    mask, vertex, _ = synth.make_inputs("small", device="cuda:0", seed=32, B=2)
    K = vertex.shape[3]
    exs = _make(pvb, 2, 1, K, 2)
    try:
        pvb.ransac_voting_layer_v3(mask[:1], vertex[:1], 64, inlier_thresh=0.99, seed=1, max_num=700, _exchange=(exs[0].handle, 1))
        buf = torch.zeros(2 * exs[0].bytes_per_rank, dtype=torch.uint8, device="cuda:0")
        exs[0].wait(1, buf, timeout_s=0.2)          # rank 1 never ran call 1
        torch.cuda.synchronize()
        got = buf.view(torch.float32).view(2, -1)[:, : K * 2]
        assert torch.isnan(got).all()
        with pytest.raises(RuntimeError, match="timed out"):
            exs[0].check()
    finally:
        for e in exs:
            e.close()


def test_argument_checks(pvb):
    from clean_pvnet_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.pvb_exchange_create(2, 2, 4, 64, ctypes.byref(h)) == _lib.PVB_ERR_INVALID      # rank out of range
    assert lib.pvb_exchange_create(0, 2, 1, 64, ctypes.byref(h)) == _lib.PVB_ERR_INVALID      # slots < 2
    _lib.check(lib.pvb_exchange_create(0, 2, 4, 64, ctypes.byref(h)))
    try:
        from clean_pvnet_b200 import synth
        mask, vertex, _ = synth.make_inputs("small", device="cuda:0", seed=33, B=1)
        with pytest.raises(RuntimeError, match="not connected"):
            pvb.ransac_voting_layer_v3(mask, vertex, 64, _exchange=(h, 1))
    finally:
        lib.pvb_exchange_destroy(h)


def test_sharded_layer_end_to_end_single_rank(pvb):
    """ShardedVotingLayer's peer path on hardware with a world of one: the push inside the refit kernel, the wait kernels on
    the side stream, the ring discipline's event waits and .result() -- 11 pipelined calls through a 4-slot ring, every
    gathered tensor equal to the plain call's result.  (Real peers: bench.py --gpus N, `gather_check`.)"""
    import socket
    import torch.distributed as dist
    from clean_pvnet_b200 import parallel, synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        mask, vertex, _ = synth.make_inputs("small", device="cuda:0", seed=41, B=4)
        layer = parallel.ShardedVotingLayer(4, vertex.shape[3], depth=2, gather="peer", device="cuda:0")
        assert layer.mode == "peer"
        pend = [layer(mask, vertex, 64, inlier_thresh=0.99, max_num=700, seed=900 + i) for i in range(11)]
        assert len(layer.inflight) <= 2
        for i in (10, 0, 5, 3):                                   # any order, any number of times
            want = pvb.ransac_voting_layer_v3(mask, vertex, 64, inlier_thresh=0.99, max_num=700, seed=900 + i)
            assert torch.equal(pend[i].result(), want), i
            assert torch.equal(pend[i].local, want), i
            assert pend[i].result() is pend[i].result()
        # the covariance channel: pushed from the covariance kernel, gathered the same way
        mean = pend[3].local
        c = layer.distribution(mask, vertex, mean, round_hyp_num=64, min_hyp_num=256, max_num=700, seed=77)
        _, want_cov = pvb.estimate_voting_distribution_with_mean(mask, vertex, mean, round_hyp_num=64, min_hyp_num=256, max_num=700,
                                                                 seed=77)
        assert torch.equal(c.result(), want_cov) and torch.equal(c.local, want_cov)
        layer.check()
        layer.close()
    finally:
        dist.destroy_process_group()
