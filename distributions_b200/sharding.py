"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed for the plumbing).

Two partitionings:

* row shards  -- rows are independent given frozen statistics (examples/mixture/main.py:143-160):
  caches and prior are replicated, each rank scores + samples its own contiguous row range.
  NO data-path collective.
* row shards, update row -- batched add_value / remove_value: every rank accumulates its rows, ONE
  all-reduce(sum) of the [F][4][G] accumulators, every rank merges the global sums (row_sharded_update).
* feature shards of one cross-cat kind -- scores[n][g] = prior[g] + sum_f s_f[n][g] is a sum over
  features: rank k scores its features into a partial [rows][G] (the prior is added by exactly one
  rank), ONE reduce-scatter(sum) over row blocks leaves rank k with the fully reduced rows of block k,
  which it samples with the stand-alone sampler.  Rows are processed in tiles so that the
  reduce-scatter of tile i (NCCL, its own stream) overlaps the scoring of tile i+1.

The compute is injected as callables so the same orchestration runs on CUDA (C-ABI kernels) and in the
world_size=2 gloo tests on CPU.
"""
import torch
import torch.distributed as dist


def row_shard(n_rows, rank, world):
    """contiguous [lo, hi) row range of `rank`; ranges differ by at most one row"""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def feature_shard(n_features, rank, world):
    """indices of the features owned by `rank` (round-robin keeps model kinds balanced)"""
    return list(range(rank, n_features, world))


def block_rows(n_rows, world):
    """rows per reduce-scatter block: every rank's block has the same (padded) size"""
    return (n_rows + world - 1) // world


def reduce_scatter_rows(partial, out_block, group=None):
    """partial: [world * block][G] (this rank's partial scores, zero-padded), out_block: [block][G]
    receives sum over ranks of block `rank`.  Uses reduce_scatter_tensor where the backend has it
    (NCCL), all_reduce + slice otherwise (gloo)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    block = out_block.shape[0]
    assert partial.shape[0] == world * block
    try:
        if dist.get_backend(group) == "gloo":
            raise RuntimeError("gloo: no reduce_scatter_tensor")
        dist.reduce_scatter_tensor(out_block, partial, op=dist.ReduceOp.SUM, group=group)
    except RuntimeError:
        tmp = partial.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        out_block.copy_(tmp[rank * block:(rank + 1) * block])
    return out_block


def feature_sharded_score_sample(score_partial, sample_block, n_rows, n_groups, u, device, tile_rows=65536, group=None,
                                 comm_stream=None):
    """Score + sample `n_rows` rows whose FEATURES are sharded over the ranks of `group`.

    score_partial(lo, hi, out)   writes this rank's partial scores of rows [lo, hi) into out[(hi-lo)][G]
                                 (prior included on exactly one rank)
    sample_block(scores, u, out) samples rows of a fully reduced [b][G] block with uniforms u[b]
    u                            [n_rows] uniforms (replicated)

    Returns (assign_local, (lo, hi)): the indices of the rows this rank owns.  Row tiles of
    world * block rows are reduce-scattered so that rank k owns rows [t0 + k*block, t0 + (k+1)*block) of
    every tile.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    tile_rows = max(world, min(tile_rows, n_rows))
    block = block_rows(tile_rows, world)
    tile_rows = block * world
    owned_rows, owned_assign = [], []
    use_cuda = device.type == "cuda"
    cur = torch.cuda.current_stream(device) if use_cuda else None
    bufs = [torch.zeros((tile_rows, n_groups), dtype=torch.float32, device=device) for _ in range(2)]
    outs = [torch.zeros((block, n_groups), dtype=torch.float32, device=device) for _ in range(2)]
    pending = None  # (buffer index, t0, event)
    t0 = 0
    it = 0

    def finish(p):
        bi, pt0, ev = p
        if use_cuda and ev is not None:
            cur.wait_event(ev)
        lo = min(pt0 + rank * block, n_rows)
        hi = min(lo + block, min(pt0 + tile_rows, n_rows))
        if hi > lo:
            a = torch.empty(hi - lo, dtype=torch.int32, device=device)
            sample_block(outs[bi][:hi - lo], u[lo:hi], a)
            owned_rows.append((lo, hi))
            owned_assign.append(a)

    while t0 < n_rows:
        bi = it & 1
        hi = min(t0 + tile_rows, n_rows)
        buf = bufs[bi]
        if hi - t0 < tile_rows:
            buf.zero_()
        score_partial(t0, hi, buf[:hi - t0])
        if use_cuda and comm_stream is not None:
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(ready)
                reduce_scatter_rows(buf, outs[bi], group)
                done = torch.cuda.Event()
                done.record(comm_stream)
        else:
            reduce_scatter_rows(buf, outs[bi], group)
            done = None
        if pending is not None:
            finish(pending)  # sample the previous tile while this tile's reduce-scatter is in flight
        pending = (bi, t0, done)
        t0 = hi
        it += 1
    if pending is not None:
        finish(pending)
    return owned_assign, owned_rows


def row_sharded_update(accumulate, merge, xchg, sign=+1, group=None):
    """The update row (batched add_value / remove_value) under ROW shards: statistics are replicated, every
    rank holds a slice of the rows, so the per-group accumulators must be summed over the ranks before the merge.

    accumulate(xchg)    writes this rank's accumulators of its own rows into xchg ([F][4][G] float64)
    merge(xchg, sign)   merges summed accumulators into this rank's replica of the statistics
    One all-reduce(sum) in between; afterwards all replicas hold the same statistics (integers exactly; the
    float64 sums in the reduction order of the backend, identical on every rank)."""
    accumulate(xchg)
    dist.all_reduce(xchg, op=dist.ReduceOp.SUM, group=group)
    merge(xchg, sign)
    return xchg


def recommended_row_shards(world):
    """row_shards for PeerFeatureShards on `world` GPUs of one NVLink domain, from the measured c3 strong scaling
    (profiles/r02_bench/bench_n{2,4,8}.json): two feature shards per group hide the push completely (0.96 at 2 GPUs) and
    send (world / 2) x fewer bytes than pure feature sharding, whose per-lane remote stores top out at 210-340 GB/s
    (0.69 at 4 GPUs, 0.34 at 8); 2 x 2 reaches 0.88, 2 x 4 0.83."""
    return max(1, world // 2)


class PeerFeatureShards:
    """Feature shards with the reduction fused into the score kernel over NVLink peer memory.

    The `world` ranks form `row_shards` groups of `feature_shards = world / row_shards` ranks (pure feature
    sharding: row_shards = 1).  A group scores the contiguous row range row_shard(n_rows, rs, row_shards); inside it
    rank j holds the features feature_shard(F, j, feature_shards) and OWNS the row block j of the group's range.
    Every owner has `feature_shards` slots [block][G] (x 2, alternating by step); rank j's score kernel stores its
    partial rows straight into slot j of the owning rank (CUDA IPC mapping, dist_b200_score_push_batch, 16-byte
    stores), so the transfer overlaps the math tile by tile and no partial [N][G] tile is written to local HBM.
    Pushers then raise an epoch flag in the owners' memory and the owner's sampler waits on its flags ON THE DEVICE
    (dist_b200_peer_signal / _wait): a step is four stream-ordered launches and never blocks the host.  Slot
    buffers alternate between steps, which makes the flag chain sufficient: a pusher writing buffer b at step e + 2
    has waited at e + 1 for every rank's e + 1 push, which each rank enqueued behind its own step-e sampler.
    The owner samples the fixed-order sum of its slots (dist_b200_sample_from_slots; deterministic, unlike atomics).
    Needs one process per GPU on one NVLink domain.
    """
    launches_per_step = 4  # score + push kernel, signal, wait, slot-sum sampler

    def __init__(self, ctx, n_rows, n_groups, group=None, row_shards=1):
        self.ctx, self.group = ctx, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        assert self.world % row_shards == 0
        self.fs = self.world // row_shards           # feature shards = ranks per group
        self.rs_index, self.index = divmod(self.rank, self.fs)
        self.row_lo, self.row_hi = row_shard(n_rows, self.rs_index, row_shards)
        self.n_rows, self.G = self.row_hi - self.row_lo, n_groups
        self.block = block_rows(self.n_rows, self.fs)
        self.slot_floats = self.block * n_groups
        self.buf_floats = self.fs * self.slot_floats
        flag_bytes = 256
        self.mine, handle = ctx.peer_alloc(4 * 2 * self.buf_floats + flag_bytes)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        peers = range(self.rs_index * self.fs, (self.rs_index + 1) * self.fs)
        self.bases = [self.mine if r == self.rank else ctx.peer_open(handles[r]) for r in peers]
        # this rank's slot inside every owner's buffer, per alternating buffer; the owners' flag arrays
        self.slot_ptrs = [[b + 4 * (k * self.buf_floats + self.index * self.slot_floats) for b in self.bases] for k in range(2)]
        self.flag_ptrs = [b + 4 * 2 * self.buf_floats for b in self.bases]
        self.epoch = 0
        dist.barrier(group=group)  # every mapping exists before the first push

    def features(self, n_features):
        """indices of the features this rank scores"""
        return feature_shard(n_features, self.index, self.fs)

    def rows(self):
        """global [lo, hi) of the rows this rank's GROUP scores (the rows its columns must cover)"""
        return self.row_lo, self.row_hi

    def owned(self):
        """global [lo, hi) of the rows this rank samples"""
        lo = min(self.index * self.block, self.n_rows)
        return self.row_lo + lo, self.row_lo + min(lo + self.block, self.n_rows)

    def step(self, features, columns, prior, u, assign_out, stream=None):
        """features / columns: this rank's shard (>= 1 feature), columns covering rows(); prior: device [G], applied by
        the first rank of each group; u: device uniforms of the owned rows; assign_out: device int32 [rows owned].
        Returns the owned row range.  Everything is enqueued on `stream`; nothing blocks the host."""
        assert len(features) >= 1, "every rank needs at least one feature of the kind"
        self.epoch += 1
        k = self.epoch & 1
        self.ctx.score_push_batch(features, columns, self.n_rows, 0, prior if self.index == 0 else None,
                                  self.slot_ptrs[k], self.block, stream=stream)
        self.ctx.peer_signal(self.flag_ptrs, self.index, self.epoch, stream=stream)
        self.ctx.peer_wait(self.flag_ptrs[self.index], self.fs, self.epoch, stream=stream)
        lo, hi = self.owned()
        if hi > lo:
            self.ctx.sample_from_slots(self.mine + 4 * k * self.buf_floats, self.fs, self.slot_floats, hi - lo, self.G, u,
                                       assign_out, stream=stream)
        return lo, hi

    def close(self):
        import torch
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # nobody is still pushing into a buffer that is about to go away
        for r, b in enumerate(self.bases):
            if r != self.index:
                self.ctx.peer_close(b)
        self.ctx.peer_free(self.mine)
