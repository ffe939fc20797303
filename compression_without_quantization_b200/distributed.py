"""Multi-GPU: groups are independent (CIS:233-245, CGS:273-284), so after the (cheap, replicated)
prologue + grouping every rank encodes one CONTIGUOUS range of groups, balanced on the work each group
costs, and the only data-path collective is one all-gather of the per-group indices (NCCL over
NVLink when the tensors are on CUDA; gloo on CPU in the tests).  One process per GPU, torch.distributed
for the plumbing.
"""
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


def all_gather_ranges_(values, bounds, group=None):
    """In place: `values` (1-D or [n, width]) holds this rank's rows [bounds[rank], bounds[rank+1]); on return it holds
    every rank's rows.  The ranges are padded to the longest one so that ONE all_gather_into_tensor moves everything; the
    rank's rows are copied once into their slot of the receive buffer, the collective gathers in place (send buffer ==
    that slot, no separate send tensor), and one batched copy (torch.cat with out=) lays the ragged ranges back out."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return values
    rank = dist.get_rank(group)
    lens = [bounds[r + 1] - bounds[r] for r in range(world)]
    longest = max(max(lens), 1)
    recv = torch.empty((world, longest) + tuple(values.shape[1:]), dtype=values.dtype, device=values.device)
    recv[rank, :lens[rank]].copy_(values[bounds[rank]:bounds[rank + 1]])
    # NCCL gathers in place when the send buffer is the rank's slot of the receive buffer; gloo wants a separate input
    send = recv[rank] if values.is_cuda else recv[rank].clone()
    dist.all_gather_into_tensor(recv.view(-1), send.reshape(-1), group=group)
    if bounds[world] == values.shape[0] and bounds[0] == 0 and values.is_contiguous():
        torch.cat([recv[r, :lens[r]] for r in range(world)], 0, out=values)
    else:
        for r in range(world):
            if r != rank and lens[r]:
                values[bounds[r]:bounds[r + 1]].copy_(recv[r, :lens[r]])
    return values


def all_gather_ranges(local, bounds, group=None):
    """Every rank owns values[bounds[r]:bounds[r+1]] (passed as `local`, 1-D or [len, width]); returns the
    full concatenation on every rank (all_gather_ranges_ on a fresh array)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    assert local.shape[0] == bounds[rank + 1] - bounds[rank], "local slice does not match this rank's range"
    full = torch.empty((bounds[world] - bounds[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    rel = [b - bounds[0] for b in bounds]
    full[rel[rank]:rel[rank + 1]].copy_(local)
    return all_gather_ranges_(full, rel, group)


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


class ISShard:
    """Sharded encode stage of the importance sampler: this rank scores one contiguous, work-balanced range of groups and
    one all-gather gives every rank all indices.

    On CUDA the GPU never waits for the host: `plan_async(job, side)` computes the range cuts on the device behind the
    grouping and starts the copy of (cuts, group count, per-call group offsets) to the host on `side`;
    `score_async(job)` scores this rank's range straight from the device-side cuts; `gather(job)` waits for the copy --
    which arrives while the scoring kernel runs -- because the all-gather needs its sizes on the host.
    `shard(job)` runs the three in order.  Jobs whose group tables are already on the host (and the CPU stand-ins of the
    gloo tests) take the plain path: ranges from the host-side tables, then the same encode and all-gather."""

    def __init__(self, group=None, results="own", slice_uploads=True):
        """results: "own" -- the batch API returns, on every rank, the results of this rank's calls (a contiguous run of
        calls balanced on dims) and None for the others, so that download and formatting are sharded too; "all" -- every
        rank downloads and formats everything.  slice_uploads: with host inputs every rank uploads 1/world of each array
        and the ranks all-gather the rest over NVLink instead of every rank pulling the whole batch over PCIe."""
        self.group = group
        self.results = results
        self.slice_uploads = slice_uploads

    def call_range(self, sizes):
        """[first, last + 1) of the calls whose results this rank returns."""
        world, rank = self._world_rank()
        if self.results != "own" or world == 1:
            return 0, len(sizes)
        key = (tuple(sizes), world)
        if getattr(self, "_call_range_key", None) != key:       # a codec / benchmark repeats its batches
            self._call_range_key = key
            self._call_bounds = balanced_ranges(torch.as_tensor(np.asarray(sizes, dtype=np.int64)), world)
        return self._call_bounds[rank], self._call_bounds[rank + 1]

    def can_slice_uploads(self, targets, proposals):
        world, _ = self._world_rank()
        if not self.slice_uploads or world == 1:
            return False
        def host_f32(a):
            return (isinstance(a, torch.Tensor) and a.device.type == "cpu" and a.dtype == torch.float32) or \
                   (isinstance(a, np.ndarray) and a.dtype == np.float32)
        return all(host_f32(t.loc) and host_f32(t.scale) for t in targets) and \
            all(host_f32(p.loc) and host_f32(p.scale) for p in proposals)

    def upload_sliced(self, targets, proposals, dev):
        """The four flat input arrays of a run of calls on the device: this rank copies its 1/world slice of each from the
        host into one send block, ONE all-gather brings every rank's block over NVLink (a collective per array costs more
        in launch latency than the bytes do), and one strided device copy lays the blocks out as four contiguous arrays."""
        from . import engine as E
        world, rank = self._world_rank()
        lists = ([t.loc for t in targets], [t.scale for t in targets], [p.loc for p in proposals], [p.scale for p in proposals])
        spans = [E.host_span(arrays) for arrays in lists]
        n = int(spans[0].numel())
        m = (n + world - 1) // world
        recv = torch.empty((world, 4, m), dtype=torch.float32, device=dev)
        lo, hi = min(n, rank * m), min(n, rank * m + m)
        if hi > lo:
            for k in range(4):
                recv[rank, k, :hi - lo].copy_(spans[k][lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(recv.view(-1), recv[rank].reshape(-1), group=self.group)
        full = recv.permute(1, 0, 2).contiguous().view(4, world * m)
        return tuple(full[k, :n] for k in range(4))

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
        all_gather_ranges_(job.index, bounds, self.group)
        job.shard_bounds = bounds

    def plan_async(self, job, side):
        world, _ = self._world_rank()
        g = job.groups
        job._bounds_dev = job.shard_ranges_device(world)                     # world + 1 entries, computed on the device
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        nb, ns = world + 1, int(g.seg_group_off.numel())
        host = torch.empty(nb + ns, dtype=torch.int64, device="cpu", pin_memory=True)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            host[:nb].copy_(job._bounds_dev, non_blocking=True)
            host[nb:].copy_(g.seg_group_off, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(side)
        job._bounds_dev.record_stream(side)
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
        self._all_gather(job, self.finish_plan(job))


def make_is_shard(group=None, results="own", slice_uploads=True):
    """`shard` object for code_grouped_importance_sample_batch: encode this rank's range of groups, then one NCCL
    all-gather of the per-group indices (see ISShard for `results` and `slice_uploads`)."""
    return ISShard(group, results, slice_uploads)


def make_greedy_shard(group=None):
    def shard(job):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        bounds = balanced_ranges(greedy_work(job), world)
        job.encode(bounds[rank], bounds[rank + 1])
        ns = job.n_steps
        idx = job.indices.view(-1, ns)
        all_gather_ranges_(idx, bounds, group)
        # the running samples of the other ranks' groups are rebuilt from the gathered indices
        job.decode()
        job.best_from_decode = True
        job.shard_bounds = bounds
    return shard
