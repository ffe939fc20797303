"""Multi-GPU: groups are independent (CIS:233-245, CGS:273-284), so after the (cheap, replicated)
prologue + grouping every rank encodes one CONTIGUOUS range of groups, balanced on the work each group
costs, and the only data-path collective is one all-gather of the per-group indices (NCCL over
NVLink when the tensors are on CUDA; gloo on CPU in the tests).  One process per GPU, torch.distributed
for the plumbing.
"""
import numpy as np
import torch
import torch.distributed as dist


def balanced_ranges(work, world):
    """Split groups 0..len(work)-1 into `world` contiguous ranges with near-equal total work.

    work: 1-D tensor / array of non-negative per-group costs.  Returns world+1 ascending bounds
    (python ints), bounds[0] = 0, bounds[-1] = n_groups.  Deterministic: every rank computes the same."""
    w = work if isinstance(work, torch.Tensor) else torch.as_tensor(np.asarray(work))
    n = int(w.numel())
    if world <= 1 or n == 0:
        return [0] + [n] * max(world, 1)
    c = torch.cumsum(w.to(torch.float64), 0)
    total = float(c[-1].item())
    targets = torch.tensor([total * r / world for r in range(1, world)], dtype=torch.float64, device=c.device)
    cuts = torch.searchsorted(c, targets, right=False) + 1  # first prefix reaching the target closes the range
    b = [0] + [int(min(max(int(x), 0), n)) for x in cuts.cpu().tolist()] + [n]
    for i in range(1, len(b)):
        b[i] = max(b[i], b[i - 1])
    return b


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


def is_work(job):
    """Per-group cost of the importance sampler: candidates x Philox blocks per candidate."""
    sizes = job.groups.starts[1:] - job.groups.starts[:-1]
    return job.K.to(torch.int64) * torch.clamp((sizes + 3) // 4, min=1)


def greedy_work(job):
    sizes = job.groups.starts[1:] - job.groups.starts[:-1]
    return torch.clamp((sizes + 3) // 4, min=1)


def make_is_shard(group=None):
    """`shard` callable for code_grouped_importance_sample_batch: encode this rank's range, all-gather indices."""
    def shard(job):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        bounds = balanced_ranges(is_work(job), world)
        job.encode(bounds[rank], bounds[rank + 1])
        job.index = all_gather_ranges(job.index[bounds[rank]:bounds[rank + 1]], bounds, group).contiguous()
        job.shard_bounds = bounds
    return shard


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
