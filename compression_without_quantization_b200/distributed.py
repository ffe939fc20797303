"""Multi-GPU: groups are independent (CIS:233-245, CGS:273-284), so after the (cheap, replicated)
prologue + grouping every rank encodes one CONTIGUOUS range of groups, balanced on the work each group
costs, and the only data-path collective is one all-gather of the per-group indices (NCCL over
NVLink when the tensors are on CUDA; gloo on CPU in the tests).  One process per GPU, torch.distributed
for the plumbing.
"""
import atexit
import ctypes

import numpy as np
import torch
import torch.distributed as dist


def _cuts_on_device(work, world):
    """world - 1 cut positions (int64 tensor on work's device) of the contiguous equal-work split: cut r closes range r at
    the first group whose work prefix reaches r/world of the total.  Nothing here waits for the host."""
    c = torch.cumsum(work.to(torch.float64), 0)
    frac = torch.arange(1, world, dtype=torch.float64, device=c.device) / world
    return torch.searchsorted(c, c[-1] * frac, right=False) + 1


def _bounds_from_cuts(cuts, n, world):
    b = [0] + [int(min(max(int(x), 0), n)) for x in cuts] + [n]
    for i in range(1, len(b)):
        b[i] = max(b[i], b[i - 1])
    return b


def balanced_ranges(work, world):
    """Split groups 0..len(work)-1 into `world` contiguous ranges with near-equal total work.

    work: 1-D tensor / array of non-negative per-group costs.  Returns world+1 ascending bounds
    (python ints), bounds[0] = 0, bounds[-1] = n_groups.  Deterministic: every rank computes the same."""
    w = work if isinstance(work, torch.Tensor) else torch.as_tensor(np.asarray(work))
    n = int(w.numel())
    if world <= 1 or n == 0:
        return [0] + [n] * max(world, 1)
    return _bounds_from_cuts(_cuts_on_device(w, world).cpu().tolist(), n, world)


def all_gather_ranges(local, bounds, group=None):
    """Every rank owns values[bounds[r]:bounds[r+1]] (passed as `local`, 1-D or [len, width]); returns the
    full concatenation on every rank.  Ranges are padded to the longest one so a single
    all_gather_into_tensor moves everything."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    lens = [bounds[r + 1] - bounds[r] for r in range(world)]
    assert local.shape[0] == lens[rank], "local slice does not match this rank's range"
    width = local.shape[1:]
    longest = max(max(lens), 1)
    send = torch.zeros((longest,) + tuple(width), dtype=local.dtype, device=local.device)
    send[:lens[rank]] = local
    recv = torch.empty((world * longest,) + tuple(width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return torch.cat([recv[r * longest:r * longest + lens[r]] for r in range(world)], 0)


def row_blocks(sizes):
    """Philox blocks per candidate of a group of `sizes` dims (3 blocks per 16 dims + tail)."""
    nq = (sizes + 3) // 4
    return 3 * (nq // 4) + torch.clamp(nq % 4, max=3)


def is_work(job):
    """Per-group cost of the importance sampler: candidates x Philox blocks per candidate."""
    sizes = job.groups.starts[1:] - job.groups.starts[:-1]
    return job.K.to(torch.int64) * torch.clamp(row_blocks(sizes), min=1)


def greedy_work(job):
    sizes = job.groups.starts[1:] - job.groups.starts[:-1]
    return torch.clamp(row_blocks(sizes), min=1)


class P2PIndexExchange:
    """The index exchange of a sharded run over NVLink peer memory (csrc/cwq_p2p.cu) instead of an all-gather: the rank
    that found an index stores it into every rank's index array, raises a flag on every peer, and a one-block kernel
    waits for the flags of all ranks.  Nothing of it involves the host.  One instance per process; buffers are
    allocated (and their IPC handles exchanged, a host-side collective) on first use and when the capacity grows."""

    def __init__(self, group=None, timeout_seconds=2.0):
        self.group, self.timeout = group, float(timeout_seconds)
        self.capacity, self.epoch = 0, 0
        self.local, self.peers, self.opened = None, [], []
        atexit.register(self.close)

    def _world_rank(self):
        if dist.is_initialized():
            return dist.get_world_size(self.group), dist.get_rank(self.group)
        return 1, 0

    def ensure(self, capacity_groups):
        """Collective: every rank calls it with the same capacity."""
        from . import _lib
        world, rank = self._world_rank()
        if world > _lib.P2P_MAX_WORLD:
            raise _lib.CwqError("peer-memory exchange supports up to {} ranks".format(_lib.P2P_MAX_WORLD))
        if capacity_groups <= self.capacity:
            return self
        torch.cuda.synchronize()              # nobody may still be reading the buffers that are about to go
        if world > 1:
            dist.barrier(self.group)
        self.close()
        cap = 1 << max(int(capacity_groups - 1).bit_length(), 20)
        nbytes = _lib.load().cwq_p2p_buffer_bytes(cap)
        ptr, handle = ctypes.c_void_p(0), (ctypes.c_uint8 * 64)()
        _lib.call("cwq_p2p_alloc", nbytes, ctypes.cast(ctypes.pointer(ptr), ctypes.c_void_p), ctypes.cast(handle, ctypes.c_void_p))
        self.local, self.nbytes, self.capacity = ptr.value, nbytes, cap
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, bytes(handle), group=self.group)
        self.peers = []
        for r in range(world):
            if r == rank:
                self.peers.append(self.local)
                continue
            p = ctypes.c_void_p(0)
            h = (ctypes.c_uint8 * 64).from_buffer_copy(handles[r])
            _lib.call("cwq_p2p_open", ctypes.cast(h, ctypes.c_void_p), ctypes.cast(ctypes.pointer(p), ctypes.c_void_p))
            self.peers.append(p.value)
            self.opened.append(p.value)
        self.peer_array = (ctypes.c_void_p * world)(*self.peers)
        self.epoch = 0
        if world > 1:
            dist.barrier(self.group)
        return self

    def exchange(self, job, bounds_dev):
        """Queue publish + signal + wait + copy-out behind the job's scoring kernel; afterwards (in stream order) the
        job's index array holds the indices of all groups."""
        from . import _lib, engine as E
        world, rank = self._world_rank()
        self.epoch += 1
        cap = job._score_cap
        _lib.call("cwq_p2p_exchange_indices", job._score_ws, cap, E._ptr(bounds_dev[rank:rank + 2]),
                  ctypes.cast(self.peer_array, ctypes.c_void_p), world, rank, self.epoch, self.capacity, self.timeout,
                  E._ptr(job.index), E._stream())

    def status(self):
        """0, or 1 + the rank whose flag did not arrive within the timeout (synchronises the current stream)."""
        from . import _lib
        from . import engine as E
        if not self.local:
            return 0
        view = E.raw_device_tensor(self.local, _lib.P2P_HEADER_BYTES, E.cuda_device())
        return int(view[_lib.P2P_STATUS_OFFSET:_lib.P2P_STATUS_OFFSET + 4].view(torch.int32).cpu()[0])

    def close(self):
        from . import _lib
        try:
            for p in self.opened:
                _lib.load().cwq_p2p_close(ctypes.c_void_p(p))
            if self.local:
                _lib.load().cwq_p2p_free(ctypes.c_void_p(self.local))
        except Exception:
            pass
        self.opened, self.peers, self.local, self.capacity = [], [], None, 0


class ISShard:
    """Sharded encode stage of the importance sampler: this rank scores one contiguous, work-balanced range of groups and
    one all-gather gives every rank all indices.

    On CUDA the GPU never waits for the host: `plan_async(job, side)` computes the range cuts on the device behind the
    grouping and starts the copy of (cuts, group count, per-call group offsets) to the host on `side`;
    `score_async(job)` scores this rank's range straight from the device-side cuts; `gather(job)` waits for the copy --
    which arrives while the scoring kernel runs -- because the all-gather needs its sizes on the host.
    `shard(job)` runs the three in order.  Jobs whose group tables are already on the host (and the CPU stand-ins of the
    gloo tests) take the plain path: ranges from the host-side tables, then the same encode and all-gather."""

    def __init__(self, group=None, p2p=None):
        self.group = group
        self.p2p = p2p           # a P2PIndexExchange: indices travel through peer memory instead of an all-gather

    def _world_rank(self):
        if dist.is_initialized():
            return dist.get_world_size(self.group), dist.get_rank(self.group)
        return 1, 0

    def __call__(self, job):
        world, rank = self._world_rank()
        g = job.groups
        if hasattr(g, "synced") and not g.synced and job.K.is_cuda:
            self.plan_async(job, torch.cuda.current_stream())
            self.score_async(job)
            self.gather(job)
        else:
            bounds = balanced_ranges(is_work(job), world)
            job.encode(bounds[rank], bounds[rank + 1])
            self._all_gather(job, bounds)

    def _all_gather(self, job, bounds):
        _, rank = self._world_rank()
        # in place: the finishing stages of a fused job read job.index through the raw pointer taken at allocation
        job.index.copy_(all_gather_ranges(job.index[bounds[rank]:bounds[rank + 1]], bounds, self.group))
        job.shard_bounds = bounds

    def plan_async(self, job, side):
        world, _ = self._world_rank()
        g = job.groups
        job._bounds_dev = job.shard_ranges_device(world)                     # world + 1 entries, computed on the device
        if self.p2p is not None:
            self.p2p.ensure(int(job.K.numel()))
        info = torch.cat([job._bounds_dev, g.seg_group_off.to(torch.int64)])
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        host = torch.empty(info.shape, dtype=info.dtype, device="cpu", pin_memory=True)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            host.copy_(info, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(side)
        info.record_stream(side)
        job._plan = (host, copied, world)

    def score_async(self, job):
        _, rank = self._world_rank()
        job.encode_device_range(job._bounds_dev[rank:rank + 2])

    def finish_plan(self, job):
        """Wait for the copy started by plan_async; installs the host-side group tables and returns the bounds."""
        host, copied, world = job._plan
        copied.synchronize()
        v = host.numpy()
        job.groups.set_host(v[world + 1:].copy())
        job.sync_groups()
        return [int(x) for x in v[:world + 1]]

    def gather(self, job):
        if self.p2p is None:
            self._all_gather(job, self.finish_plan(job))
            return
        # peer-memory exchange: queued at once (the GPU side needs nothing from the host); the host-side tables are
        # installed afterwards for the caller's benefit (result sizes), then the index array is re-pointed and trimmed
        self.p2p.exchange(job, job._bounds_dev)
        job.shard_bounds = self.finish_plan(job)


_p2p_exchanges = {}


def make_is_shard(group=None, p2p=False):
    """`shard` object for code_grouped_importance_sample_batch: encode this rank's range, exchange the indices
    (p2p=False: one NCCL all-gather; p2p=True: stores into every rank's index array over NVLink peer memory)."""
    if not p2p:
        return ISShard(group)
    if group not in _p2p_exchanges:
        _p2p_exchanges[group] = P2PIndexExchange(group)
    return ISShard(group, _p2p_exchanges[group])


def make_greedy_shard(group=None):
    def shard(job):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        bounds = balanced_ranges(greedy_work(job), world)
        job.encode(bounds[rank], bounds[rank + 1])
        ns = job.n_steps
        idx = job.indices.view(-1, ns)
        job.indices = all_gather_ranges(idx[bounds[rank]:bounds[rank + 1]], bounds, group).contiguous().view(-1)
        # the running samples of the other ranks' groups are rebuilt from the gathered indices
        job.decode()
        job.best_from_decode = True
        job.shard_bounds = bounds
    return shard
