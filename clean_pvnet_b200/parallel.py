"""Multi-GPU use of the voting layer: one process per GPU, images sharded across ranks.

The path is embarrassingly parallel over images: there is NO exchange inside the
algorithm, so ranks run the single-GPU op on their contiguous slice of the batch and the only cross-GPU
traffic is every rank's [B/G, K, 2] keypoints becoming visible on every rank -- a few KB.  The reference has
nothing here (torch.nn.DataParallel in the trainer only, lib/train/trainers/trainer.py:5,11).
The philox sampling stream is keyed by the GLOBAL image index (`img_base`), so the gathered result is
identical for any number of ranks.

Two ways to make the results visible:

  gather="peer" (default on GPUs)   the exchange is FUSED INTO THE REFIT KERNEL: the thread that produces an (image,
      keypoint) result stores it straight into every peer's receive ring over NVLink as self-validating 8-byte words
      {float, seq} (csrc/exchange.cu, vote.cu "exchange tail"): no fence, no flag, nothing waits in the producing call and
      no collective kernel is launched; a tiny wait kernel, enqueued `depth` calls later (or when the result is asked
      for), polls the rank's OWN memory until every word carries the call's sequence number and writes the floats out.
      Round 1's per-call NCCL all_gather cost 0.30 ms of a 0.97 ms step at 8 GPUs (NCCL's kernel must become co-resident
      on all GPUs and spins holding SM slots until the slowest rank arrives); this path has no such rendezvous.
  gather="collective"               one torch.distributed all_gather per call (NCCL on GPUs, gloo in the CPU tests of the
      host logic).  Kept as the portable fallback and as the baseline the fused path is measured against.

Ring discipline of the peer path (what makes reuse of the receive slots safe without acknowledgements): the ring has
2*depth slots; call s writes slot (s-1) % (2*depth) of every peer; a rank starts call s only after its own wait of call
s-depth has FINISHED (ShardedVotingLayer: an event wait on the compute stream; the wait kernels themselves poll on a side
stream).  A rank that executes call s has therefore seen every peer's call s-depth complete, and each peer started that
call only after its own wait of call s-2*depth had finished: the slot call s overwrites has been copied out everywhere.
"""
import ctypes

import torch
import torch.distributed as dist


def shard_bounds(total, world_size, rank):
    """Contiguous split of `total` images: the first (total % world) ranks get one extra."""
    base, rem = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_ragged(local, total, group=None, async_op=False):
    """all_gather of per-rank slices [n_r, ...] (n_r from shard_bounds) into [total, ...].

    async_op=True returns (finish, work): the collective is enqueued on the backend's own stream (after the producer of
    `local` on the current stream); call `work.wait()` and then `finish()` to get the gathered tensor."""
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
        return _unpad(out, sizes, nmax, local.shape[1:])

    if async_op:
        return finish, work
    return finish()


def _unpad(out, sizes, nmax, tail):
    """[world*nmax, ...] with every rank's block padded to nmax rows -> [total, ...]."""
    if all(hi - lo == nmax for lo, hi in sizes):
        return out
    o = out.view(len(sizes), nmax, *tail)
    return torch.cat([o[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


class PeerExchange:
    """A pvb_exchange (include/pvnet_vote_b200.h): this rank's receive ring, mapped into every peer through CUDA IPC.
    Construction is collective over `group` (the 64-byte IPC handles travel through all_gather_object, once)."""

    def __init__(self, bytes_per_rank, slots, group=None, device=None, rank=None, world=None):
        from . import _lib
        self._lib_mod = _lib
        self.lib = _lib.load()
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.slots = int(slots)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pvb_exchange_create(self.rank, self.world, self.slots, int(bytes_per_rank), ctypes.byref(h)))
        self.handle = h
        self.bytes_per_rank = int(self.lib.pvb_exchange_bytes_per_rank(h))

    def connect_ipc(self):
        """Collective: exchanges the IPC handles and maps every peer's ring."""
        if self.world == 1:
            return self
        mine = ctypes.create_string_buffer(self._lib_mod.PVB_IPC_HANDLE_BYTES)
        with torch.cuda.device(self.device):
            self._lib_mod.check(self.lib.pvb_exchange_get_handle(self.handle, mine))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine.raw), group=self.group)
        blob = ctypes.create_string_buffer(b"".join(handles), self.world * self._lib_mod.PVB_IPC_HANDLE_BYTES)
        with torch.cuda.device(self.device):
            self._lib_mod.check(self.lib.pvb_exchange_connect(self.handle, blob))
        return self

    def connect_local(self, peers):
        """Ranks that live in ONE process (tests, single-process multi-GPU): raw base pointers instead of IPC handles."""
        arr = (ctypes.c_void_p * self.world)(*[self.lib.pvb_exchange_base(p.handle) for p in peers])
        self._lib_mod.check(self.lib.pvb_exchange_connect_ptrs(self.handle, arr))
        return self

    def wait(self, seq, out, timeout_s=10.0, floats_per_rank=None):
        """Enqueues the wait kernel of call `seq` on the current stream; `out` = uint8 [world*bytes_per_rank] on this device.
        floats_per_rank: how many floats each rank publishes per call (ragged shards); None = bytes_per_rank/4 each."""
        counts = None
        if floats_per_rank is not None:
            counts = (ctypes.c_int32 * self.world)(*[int(n) for n in floats_per_rank])
        self._lib_mod.check(self.lib.pvb_exchange_wait(self.handle, int(seq), out.data_ptr(), counts, float(timeout_s),
                                                       torch.cuda.current_stream(self.device).cuda_stream))

    def check(self):
        """Synchronises; raises if any wait timed out."""
        self._lib_mod.check(self.lib.pvb_exchange_status(self.handle, torch.cuda.current_stream(self.device).cuda_stream))

    def close(self):
        if self.handle:
            with torch.cuda.device(self.device):
                self.lib.pvb_exchange_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pending:
    """Result of one sharded call: `.result()` returns the result of the WHOLE batch ([total, K, 2] keypoints or
    [total, K, 2, 2] covariances) on this rank's device (no host sync); `.local` is this rank's own slice, available at once."""

    def __init__(self, channel, seq, local):
        self._channel, self.seq, self.local = channel, seq, local
        self._gathered = None       # the gathered tensor (peer path: written by the wait kernel on the side stream)
        self._event = None          # peer path: recorded on the side stream after the wait kernel
        self._consumed = False      # peer path: the caller's stream has been made to wait for _event
        self._work = None           # collective path
        self._finish = None

    def result(self):
        self._channel.complete(self)
        return self._gathered


class _Channel:
    """One stream of sharded calls whose per-rank results have the shape [n_r, K, *unit]: its own receive ring (peer path),
    sequence numbers, in-flight list and side stream.  The keypoints and the covariances of a ShardedVotingLayer are two
    channels."""

    def __init__(self, owner, unit, use_peer, gather):
        self.o = owner
        self.unit = tuple(unit)                     # (2,) keypoints, (2, 2) covariances
        self.nf = 1
        for u in self.unit:
            self.nf *= u                            # floats per (image, keypoint)
        self.seq = 0
        self.inflight = []                          # oldest first
        self.exchange = None
        self.wait_stream = None
        self.error = None
        o = owner
        if use_peer:
            ok = 1
            try:
                self.exchange = PeerExchange(o.nmax * o.K * self.nf * 4, 2 * o.depth, group=o.group, device=o.device)
                self.exchange.connect_ipc()
            except Exception as e:          # no P2P between the GPUs, IPC unavailable, ...
                self.error = str(e)
                ok = 0
            # all ranks must agree on the path
            flag = torch.tensor([ok], dtype=torch.int32, device=o.device if dist.get_backend(o.group) == "nccl" else "cpu")
            if o.world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=o.group)
            if int(flag.item()) == 0:
                if gather == "peer":
                    raise RuntimeError(f"peer exchange unavailable: {self.error or 'failed on another rank'}")
                if self.exchange is not None:
                    self.exchange.close()
                self.exchange = None
        if self.exchange is not None:
            self.wait_stream = torch.cuda.Stream(device=o.device)
            self.counts = [(hi - lo) * o.K * self.nf for lo, hi in o.sizes]
            self.row_bytes = o.nmax * o.K * self.nf * 4

    def launch(self, run_local):
        """run_local(exchange_arg) -> this rank's result tensor; exchange_arg is (handle, seq) on the peer path, else None."""
        o = self.o
        self.seq += 1
        if self.exchange is not None:
            cur = torch.cuda.current_stream(o.device)
            # ring discipline: this rank's wait of call seq-depth must have finished before call seq may overwrite the
            # peers' slots (module docstring); that wait has been polling on the side stream for `depth` calls already
            while len(self.inflight) >= o.depth:
                cur.wait_event(self.inflight.pop(0)._event)
            local = run_local((self.exchange.handle, self.seq))
            p = Pending(self, self.seq, local)
            pushed = torch.cuda.Event()
            pushed.record(cur)
            with torch.cuda.stream(self.wait_stream):
                self.wait_stream.wait_event(pushed)
                buf = torch.empty(o.world * self.exchange.bytes_per_rank, dtype=torch.uint8, device=o.device)
                self.exchange.wait(p.seq, buf, o.timeout_s, self.counts)
                p._event = torch.cuda.Event()
                p._event.record(self.wait_stream)
            p._buf = buf
        else:
            while len(self.inflight) >= o.depth:
                self.complete(self.inflight[0])
            local = run_local(None)
            p = Pending(self, self.seq, local)
            p._finish, p._work = all_gather_ragged(local, o.total, o.group, async_op=True)
        self.inflight.append(p)
        return p

    def complete(self, p):
        """Make p's gathered tensor valid on the current stream."""
        o = self.o
        if self.exchange is not None:
            if not p._consumed:
                cur = torch.cuda.current_stream(o.device)
                cur.wait_event(p._event)
                p._buf.record_stream(cur)                       # allocated on the side stream, consumed here
                rows = p._buf.view(o.world, self.exchange.bytes_per_rank)[:, : self.row_bytes]
                flat = rows.reshape(o.world * self.row_bytes)   # a view unless bytes_per_rank was padded to 16
                out = flat.view(torch.float32).view(o.world * o.nmax, o.K, *self.unit)
                p._gathered = _unpad(out, o.sizes, o.nmax, (o.K,) + self.unit)
                p._consumed = True
            return
        while self.inflight and self.inflight[0].seq <= p.seq:   # collectives complete in call order
            q = self.inflight.pop(0)
            if q._work is not None:
                q._work.wait()
            q._gathered = q._finish()

    def drain(self):
        if self.exchange is not None:
            cur = torch.cuda.current_stream(self.o.device)
            for q in self.inflight:
                cur.wait_event(q._event)
            self.inflight.clear()
        elif self.inflight:
            self.complete(self.inflight[-1])


class ShardedVotingLayer:
    """ransac_voting_layer_v3 (and estimate_voting_distribution_with_mean) over a batch sharded across the ranks of `group`.

        layer = ShardedVotingLayer(total_images=128, K=17)           # collective (peer rings are mapped here)
        p = layer(mask_local, vertex_local, 512, inlier_thresh=0.99, seed=s)   # launches; never blocks on a peer
        kpt = p.result()                                             # [128, 17, 2] on every rank
        c = layer.distribution(mask_local, vertex_local, p.local, seed=s)      # the other half of resnet18.py:71-72
        var = c.result()                                             # [128, 17, 2, 2] on every rank

    Peer path: every call pushes its results from inside the producing kernel (refit / covariance); its wait kernel goes
    onto a SIDE stream right away (after an event recorded behind the call), so polling for the slowest peer never sits
    between two steps on the compute stream.  The ring discipline -- call s may start only when this rank's wait of call
    s-depth has finished -- is an event wait on the compute stream that is already satisfied in the steady state.
    `.result()` makes the caller's stream wait for the call's wait kernel.  `op` / `gather="collective"` let the CPU tests
    drive the same bookkeeping over gloo with a stand-in operator (one all_gather per call, at most `depth` outstanding)."""

    def __init__(self, total_images, K, group=None, depth=4, gather="auto", device=None, op=None, timeout_s=10.0,
                 with_distribution=True):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.total, self.K, self.depth = int(total_images), int(K), int(depth)
        if self.depth < 1:
            raise ValueError("depth must be >= 1")
        self.sizes = [shard_bounds(self.total, self.world, r) for r in range(self.world)]
        self.nmax = max(hi - lo for lo, hi in self.sizes)
        self.lo, self.hi = self.sizes[self.rank]
        self.op = op
        self.timeout_s = timeout_s
        if gather not in ("auto", "peer", "collective"):
            raise ValueError("gather must be 'auto', 'peer' or 'collective'")
        use_peer = gather == "peer" or (gather == "auto" and op is None and torch.cuda.is_available())
        if use_peer and min(hi - lo for lo, hi in self.sizes) == 0:
            if gather == "peer":
                raise ValueError("gather='peer' needs at least one image on every rank")
            use_peer = False
        self.device = None
        if use_peer or (op is None and torch.cuda.is_available()):
            self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._kpt = _Channel(self, (2,), use_peer, gather)
        self._cov = _Channel(self, (2, 2), use_peer and self._kpt.exchange is not None, gather) if with_distribution else None
        self.mode = "peer" if self._kpt.exchange is not None else "collective"
        self.gather_error = self._kpt.error

    # the keypoint channel's bookkeeping, exposed for tests and tools
    @property
    def inflight(self):
        return self._kpt.inflight

    @property
    def exchange(self):
        return self._kpt.exchange

    def __call__(self, mask_local, vertex_local, round_hyp_num, inlier_thresh=0.999, min_num=5, max_num=30000, seed=0, **kw):
        if mask_local.shape[0] != self.hi - self.lo:
            raise ValueError("local batch does not match shard_bounds")
        op = self.op
        if op is None:
            from .ransac_voting_gpu import ransac_voting_layer_v3 as op

        def run(ex):
            extra = dict(kw, _exchange=ex) if ex is not None else kw
            return op(mask_local, vertex_local, round_hyp_num, inlier_thresh=inlier_thresh, min_num=min_num,
                      max_num=max_num, seed=seed, img_base=self.lo, **extra)
        return self._kpt.launch(run)

    def distribution(self, mask_local, vertex_local, mean_local, round_hyp_num=256, min_hyp_num=4096, inlier_thresh=0.99,
                     min_num=5, max_num=30000, seed=0, op=None, **kw):
        """estimate_voting_distribution_with_mean on this rank's images; `.result()` is the covariance of the whole batch."""
        if self._cov is None:
            raise RuntimeError("constructed with with_distribution=False")
        if mask_local.shape[0] != self.hi - self.lo:
            raise ValueError("local batch does not match shard_bounds")
        if op is None:
            from .ransac_voting_gpu import estimate_voting_distribution_with_mean as op

        def run(ex):
            extra = dict(kw, _exchange=ex) if ex is not None else kw
            return op(mask_local, vertex_local, mean_local, round_hyp_num=round_hyp_num, min_hyp_num=min_hyp_num,
                      inlier_thresh=inlier_thresh, min_num=min_num, max_num=max_num, seed=seed, img_base=self.lo, **extra)[1]
        return self._cov.launch(run)

    def drain(self):
        """Make everything launched so far complete on the current stream (no host sync)."""
        self._kpt.drain()
        if self._cov is not None:
            self._cov.drain()

    def check(self):
        """Host sync + error check of the peer path (a timed-out wait fills its result with NaN and raises here)."""
        self.drain()
        for ch in (self._kpt, self._cov):
            if ch is not None and ch.exchange is not None:
                ch.exchange.check()

    def close(self):
        for ch in (self._kpt, self._cov):
            if ch is not None and ch.exchange is not None:
                ch.exchange.close()
                ch.exchange = None


def sharded_ransac_voting_layer_v3(mask_local, vertex_local, round_hyp_num, total_images, inlier_thresh=0.999,
                                   min_num=5, max_num=30000, seed=0, group=None, op=None):
    """One-shot form: runs ransac_voting_layer_v3 on this rank's images (a shard_bounds slice of a `total_images`
    batch) and returns the keypoints of the WHOLE batch on every rank through one collective.  For repeated calls use
    ShardedVotingLayer (peer-memory exchange, no collective in the steady state).  `op` defaults to the CUDA operator;
    tests inject a stand-in to exercise the sharding logic on CPU/gloo."""
    if op is None:
        from .ransac_voting_gpu import ransac_voting_layer_v3 as op
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(total_images, world, rank)
    assert mask_local.shape[0] == hi - lo, "local batch does not match shard_bounds"
    local = op(mask_local, vertex_local, round_hyp_num, inlier_thresh=inlier_thresh, min_num=min_num,
               max_num=max_num, seed=seed, img_base=lo)
    return all_gather_ragged(local, total_images, group)
