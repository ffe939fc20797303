"""Greedy coder: the reference's entry points (code/coded_greedy_sampler.py) on the B200 engine.

    code_greedy_sample            CGS:29-89
    decode_greedy_sample          CGS:93-167
    code_grouped_greedy_sample    CGS:170-296
    decode_grouped_greedy_sample  CGS:299-364

Same duck-typing and array conventions as coded_importance_sampler.py.  `adaptive`,
`backfitting_steps` and `use_log_prob` are accepted and unused, as in the reference (CGS:177-179).
"""
import numpy as np
import torch

from . import engine as E
from ._lib import CwqError
from .coded_importance_sampler import _check_dtype

__all__ = ["code_greedy_sample", "decode_greedy_sample", "code_grouped_greedy_sample",
           "decode_grouped_greedy_sample", "code_grouped_greedy_sample_batch", "decode_grouped_greedy_sample_batch"]


def code_greedy_sample(t_loc, t_scale, p_loc, p_scale, n_bits_per_step, n_steps, seed, rho=1.):
    """One group (CGS:29-89) -> (best_sample f32[G], bitcode bytes: n_steps indices, each
    n_bits_per_step bits LSB first)."""
    dev = E.cuda_device()
    tl, ts, pl, ps = (E.as_device_f32(x, dev) for x in (t_loc, t_scale, p_loc, p_scale))
    G = tl.numel()
    if G == 0:
        return np.zeros(0, np.float32), ("0" * (int(n_bits_per_step) * int(n_steps))).encode("utf-8")
    spec = E.SegmentSpec(G, seed, 1.0, 0, float("inf"), False)
    # The shard prior is N(p_loc / n_steps, rho p_scale / sqrt(n_steps)) (CGS:42-45).  Everything is divided by p_scale only;
    # the proposal loc enters as the per-step shard loc and the result is p_scale * best.
    job = E.GreedyJob(tl, ts, torch.zeros_like(pl), ps, [spec], n_steps, n_bits_per_step, rho, forced_starts=[np.array([0, G])],
                      shard_loc=E.greedy_shard_loc(pl, ps, n_steps))
    job.prepare().encode().unrescale().pack()
    n_bits = job.n_steps * job.n_bits_per_step
    bits = job.bits[:(n_bits + 7) // 8].cpu().numpy()
    return job.sample.cpu().numpy(), E.bits_to_ascii(bits, n_bits)


def _unpack_fixed(bitcode, count, bits_per_index, dev):
    packed, n_bits = E.ascii_to_bits(bitcode)
    if n_bits < count * bits_per_index:
        raise CwqError("bitcode too short: {} bits for {} indices of {} bits".format(n_bits, count, bits_per_index))
    out = torch.zeros(max(count, 1), dtype=torch.int32, device=dev)
    if count and bits_per_index:
        b = torch.from_numpy(packed).to(dev)
        E._lib.call("cwq_fixed_unpack", E._ptr(b), count, bits_per_index, E._ptr(out), E._stream())
    return out[:count]


def decode_greedy_sample(sample_index, p_loc, p_scale, n_bits_per_step, n_steps, seed, rho=1.):
    """One group (CGS:93-167) -> f32[G]."""
    dev = E.cuda_device()
    pl, ps = E.as_device_f32(p_loc, dev), E.as_device_f32(p_scale, dev)
    G = pl.numel()
    if G == 0:
        return np.zeros(0, np.float32)
    return decode_grouped_greedy_sample_batch([sample_index], [[0, G]], [_P(torch.zeros_like(pl), ps)], n_bits_per_step, n_steps,
                                              seed, rho, shard_loc=E.greedy_shard_loc(pl, ps, n_steps))[0]


class _P:
    def __init__(self, loc, scale):
        self.loc, self.scale, self.dtype = loc, scale, np.float32


def code_grouped_greedy_sample_batch(targets, proposals, n_steps, n_bits_per_step, seed, max_group_size_bits=12,
                                     rho=1., shard=None):
    """Many reference calls in one launch sequence; per call the tuple code_grouped_greedy_sample returns."""
    dev = E.cuda_device()
    ns = len(targets)
    per = (lambda v: list(v) if isinstance(v, (list, tuple)) else [v] * ns)
    seeds, mg = per(seed), per(max_group_size_bits)
    for t, p in zip(targets, proposals):
        _check_dtype(t, "Target")
        _check_dtype(p, "Proposal")
    q_loc = [E.as_device_f32(t.loc, dev, "Target") for t in targets]
    q_scale = [E.as_device_f32(t.scale, dev, "Target") for t in targets]
    p_loc = [E.as_device_f32(p.loc, dev, "Proposal") for p in proposals]
    p_scale = [E.as_device_f32(p.scale, dev, "Proposal") for p in proposals]
    sizes = [int(x.numel()) for x in p_loc]
    n_bits_per_group = n_bits_per_step * n_steps  # CGS:189
    specs = [E.SegmentSpec(sizes[i], seeds[i], n_bits_per_group, mg[i], float("inf"), False) for i in range(ns)]
    cat = (lambda xs: xs[0] if len(xs) == 1 else torch.cat(xs))
    job = E.GreedyJob(cat(q_loc), cat(q_scale), cat(p_loc), cat(p_scale), specs, n_steps, n_bits_per_step, rho).prepare()
    if shard is None:
        job.encode()
    else:
        shard(job)
    job.unrescale().pack()
    sample = job.sample.cpu().numpy()
    starts = job.groups.starts.cpu().numpy()
    bits_per_group = job.n_steps * job.n_bits_per_step
    total_bits = job.groups.n_groups * bits_per_group
    bits = job.bits[:(total_bits + 7) // 8].cpu().numpy()
    out = []
    for s in range(ns):
        g0, g1 = job.segment_group_range(s)
        a = job.table.starts[s]
        code = E.bits_to_ascii(bits, (g1 - g0) * bits_per_group, g0 * bits_per_group).decode("utf-8")
        out.append((sample[a:a + sizes[s]].copy(), code, [int(x) - a for x in starts[g0:g1 + 1]]))
    return out


def code_grouped_greedy_sample(sess, target, proposal, n_steps, n_bits_per_step, seed, max_group_size_bits=12,
                               adaptive=True, backfitting_steps=0, use_log_prob=False, rho=1.):
    """CGS:170-296 -> (sample f32[N], bitcode str of '0'/'1', group_start_indices list ending in N)."""
    _check_dtype(target, "Target")
    _check_dtype(proposal, "Proposal")
    num_dimensions = int(np.prod(np.shape(proposal.loc))) if not hasattr(proposal.loc, "numel") else int(proposal.loc.numel())
    if num_dimensions == 0:
        # CGS:228-252 with an empty range: starts [0, 0]; the empty group still emits n_steps zero indices
        return np.zeros(0, np.float32), "0" * (int(n_bits_per_step) * int(n_steps)), [0, 0]
    return code_grouped_greedy_sample_batch([target], [proposal], n_steps, n_bits_per_step, seed, max_group_size_bits, rho)[0]


def decode_grouped_greedy_sample_batch(bitcodes, group_start_indices, proposals, n_bits_per_step, n_steps, seed, rho=1.,
                                       shard_loc=None):
    """Batch of decode calls; group_start_indices[s] must already end with the number of dimensions.  shard_loc: see
    engine.GreedyJob (device tensor over all dims of the batch, or None)."""
    dev = E.cuda_device()
    ns = len(proposals)
    seeds = list(seed) if isinstance(seed, (list, tuple)) else [seed] * ns
    for p in proposals:
        _check_dtype(p, "Proposal")
    p_loc = [E.as_device_f32(p.loc, dev, "Proposal") for p in proposals]
    p_scale = [E.as_device_f32(p.scale, dev, "Proposal") for p in proposals]
    sizes = [int(x.numel()) for x in p_loc]
    specs = [E.SegmentSpec(sizes[i], seeds[i], 1.0, 0, float("inf"), False) for i in range(ns)]
    table = E.SegmentTable(specs, dev)
    groups = E._forced_groups(group_start_indices, table, dev)
    n_steps, b = int(n_steps), int(n_bits_per_step)
    parts = []
    for s in range(ns):
        ng = len(group_start_indices[s]) - 1
        parts.append(_unpack_fixed(bitcodes[s], ng * n_steps, b, dev))
    indices = parts[0] if ns == 1 else torch.cat(parts)
    cat = (lambda xs: xs[0] if len(xs) == 1 else torch.cat(xs))
    pl, ps = cat(p_loc), cat(p_scale)
    sample = torch.empty(table.n, dtype=torch.float32, device=dev)
    E._lib.call("cwq_greedy_decode", E._ptr(groups.starts), E._ptr(groups.group_seed), E._ptr(groups.gid), E._ptr(indices),
                table.n, n_steps, E.greedy_shard_scale(rho, n_steps), E._ptr(shard_loc), E._ptr(pl), E._ptr(ps), E._ptr(sample),
                E._stream())
    host = E.to_host([sample])[0]
    return [host[table.starts[s]:table.starts[s] + sizes[s]] for s in range(ns)]


def decode_grouped_greedy_sample(sess, bitcode, group_start_indices, proposal, n_bits_per_step, n_steps, seed,
                                 adaptive=True, rho=1.):
    """CGS:299-364 -> f32[N].  N is appended to the caller's group_start_indices list (CGS:323).  Unlike
    the reference, which silently stops at the first group without bits (CGS:346-347) and then fails on
    a shape mismatch, a bitcode shorter than n_groups * n_steps * n_bits_per_step raises CwqError."""
    _check_dtype(proposal, "Proposal")
    num_dimensions = int(np.prod(np.shape(proposal.loc))) if not hasattr(proposal.loc, "numel") else int(proposal.loc.numel())
    group_start_indices += [num_dimensions]
    if num_dimensions == 0:
        return np.zeros(0, np.float32)
    return decode_grouped_greedy_sample_batch([bitcode], [group_start_indices], [proposal], n_bits_per_step, n_steps,
                                              seed, rho)[0]
