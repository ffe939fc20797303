"""Importance-sampling coder: the reference's entry points (code/coded_importance_sampler.py) on the
B200 engine.  Names, argument order, defaults, return types and error messages follow the reference so
`pln.py` can import this module in place of the original:

    code_importance_sample            CIS:29-79
    decode_importance_sample          CIS:82-109
    code_grouped_importance_sample    CIS:112-274
    decode_grouped_importance_sample  CIS:277-363

`target` / `proposal` are duck-typed exactly as in the reference (only .loc, .scale, .dtype are read);
their arrays may be numpy, torch (CPU or CUDA) or any DLPack producer.  `sess` is accepted and ignored.
`*_batch` variants run many reference calls in one launch sequence (same results per call).
"""
import time

import numpy as np
import torch

from . import _lib as L
from . import engine as E
from ._lib import CwqError
from .binary_io import elias_delta_code, elias_delta_decode

__all__ = ["code_importance_sample", "decode_importance_sample", "code_grouped_importance_sample",
           "decode_grouped_importance_sample", "code_grouped_importance_sample_batch",
           "decode_grouped_importance_sample_batch"]


class _Dist:
    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale
        self.dtype = np.float32


def _check_dtype(dist, what):
    if not E.is_float32(getattr(dist, "dtype", None)):
        raise Exception("{} datatype must be float32!".format(what))


def _u16_view(t):
    return t.cpu().numpy().view(np.uint16)


# ----------------------------------------------------------------------------------------------
# per-group entry points
# ----------------------------------------------------------------------------------------------
def code_importance_sample(t_loc, t_scale, p_loc, p_scale, n_coding_bits, seed, return_index_only=False):
    """One group (CIS:29-79): K = ceil(exp(KL(t||p))) candidates from p, index of the largest
    importance weight.  Returns (best_sample f32[1, G], Elias-delta bitcode bytes) or, with
    return_index_only, (best_sample, index + 1).  n_coding_bits is unused, as in the reference."""
    dev = E.cuda_device()
    tl, ts, pl, ps = (E.as_device_f32(x, dev) for x in (t_loc, t_scale, p_loc, p_scale))
    G = tl.numel()
    if G == 0:
        best, index = np.zeros((1, 0), np.float32), 0
    else:
        spec = E.SegmentSpec(G, seed, 1.0, 0, float("inf"), True)
        job = E.ISJob(tl, ts, pl, ps, [spec], forced_starts=[np.array([0, G])]).prepare().encode().decode()
        index = int(job.host_indices(0)[0])
        best = job.sample.cpu().numpy().reshape(1, G)
    if return_index_only:
        return best, np.int64(index + 1)
    return best, elias_delta_code(index + 1).encode("utf-8")


def decode_importance_sample(sample_index, p_loc, p_scale, seed, use_index=False):
    """One group (CIS:82-109).  With use_index the 1-based index is given directly and only the sample
    f32[1, G] is returned; otherwise sample_index is the remaining bitstream and the result is
    (sample, code_length, index, samples[index + 1, G]) as in the reference."""
    dev = E.cuda_device()
    pl, ps = E.as_device_f32(p_loc, dev), E.as_device_f32(p_scale, dev)
    G = pl.numel()
    if use_index:
        index = int(sample_index) - 1
        code_length = None
    else:
        num, code_length = elias_delta_decode(sample_index)
        index = num - 1
    s32 = int(seed) & 0xFFFFFFFF
    if use_index:
        z = E.draw_candidates(s32, index, 1, G, dev)
        return (ps[None] * z + pl[None]).cpu().numpy()
    z = E.draw_candidates(s32, 0, index + 1, G, dev)   # the reference materialises rows 0..index (CIS:104-109)
    samples = (ps[None] * z + pl[None]).cpu().numpy()
    return samples[-1:], np.int64(code_length), np.int64(index), samples


# ----------------------------------------------------------------------------------------------
# grouped entry points
# ----------------------------------------------------------------------------------------------
def _batch_specs(targets, proposals, seeds, n_bits_per_group, max_group_size_bits, dim_kl_bit_limit):
    ns = len(targets)

    def per(v):
        return list(v) if isinstance(v, (list, tuple)) else [v] * ns

    seeds, nb, mg, dl = per(seeds), per(n_bits_per_group), per(max_group_size_bits), per(dim_kl_bit_limit)
    f32 = (np.float32, torch.float32)
    for t, p in zip(targets, proposals):
        if getattr(t, "dtype", None) not in f32:      # the common dtypes without the general test
            _check_dtype(t, "Target")
        if getattr(p, "dtype", None) not in f32:
            _check_dtype(p, "Proposal")
    sizes = [int(p.loc.numel()) if hasattr(p.loc, "numel") else int(np.size(p.loc)) for p in proposals]
    return [E.SegmentSpec(sizes[i], seeds[i], nb[i], mg[i], dl[i], True) for i in range(ns)]


def _upload(targets, proposals, dev):
    return (E.concat_to_device([t.loc for t in targets], dev, "Target"),
            E.concat_to_device([t.scale for t in targets], dev, "Target"),
            E.concat_to_device([p.loc for p in proposals], dev, "Proposal"),
            E.concat_to_device([p.scale for p in proposals], dev, "Proposal"))


def _encode_batch(targets, proposals, seeds, n_bits_per_group, max_group_size_bits, dim_kl_bit_limit, stage,
                  shard=None):
    """Shared driver.  stage: 'groups' | 'indices' | 'full'.  Returns the job with its group tables synchronised."""
    dev = E.cuda_device()
    specs = _batch_specs(targets, proposals, seeds, n_bits_per_group, max_group_size_bits, dim_kl_bit_limit)
    q_loc, q_scale, p_loc, p_scale = _upload(targets, proposals, dev)
    job = E.ISJob(q_loc, q_scale, p_loc, p_scale, specs)
    fused = job.n <= E.ASYNC_MAX_DIMS and (shard is None or hasattr(shard, "plan_async"))
    if fused:
        # one library call per stage set, group count on the device: nothing below waits for the host until the
        # results are copied back (a sharded run waits once, for the ranges and the all-gather sizes)
        if shard is None:
            stages = {"groups": L.STAGE_GROUP, "indices": L.STAGE_GROUP | L.STAGE_SCORE,
                      "full": L.STAGE_GROUP | L.STAGE_SCORE | L.STAGE_FINISH}[stage]
            job.launch(stages)
            if stage == "full":
                return job               # _host_views fetches the counts and the results with two synchronisations in all
            return job.sync_groups()
        job.launch(L.STAGE_GROUP)
        if stage != "groups":
            shard(job)
        if stage == "full":
            job.launch(L.STAGE_FINISH)
        return job.sync_groups()
    job.prepare(sync=True)
    if stage == "groups":
        return job.sync_groups()
    if shard is None:
        job.encode()
    else:
        shard(job)
    if stage == "indices":
        return job.sync_groups()
    job.decode().pack()
    return job.sync_groups()


def _segment_result(job, s, host, return_indices):
    """What the reference call of segment s returns, cut out of the host copies (which cover the calls host["calls"])."""
    g0, g1 = job.segment_group_range(s)
    a, b = job.table.starts[s], job.table.starts[s] + job.table.specs[s].n
    d0, gb, o_base, bit_base = host["dim0"], host["group0"], host["outlier0"], host["bit0"]
    sample = host["sample"][a - d0:b - d0]
    starts = host["starts"][g0 - gb:g1 + 1 - gb] - a
    o0, o1 = int(host["seg_outlier_off"][s]) - o_base, int(host["seg_outlier_off"][s + 1]) - o_base
    outlier_extras = (host["outlier_idx"][o0:o1], host["outlier_q"][o0:o1])
    if return_indices:
        indices = [np.int64(i) + 1 for i in host["index"][g0 - gb:g1 - gb]]
        return sample, indices, starts, outlier_extras
    bit0, bit1 = int(host["bit_off"][g0 - gb]) - bit_base, int(host["bit_off"][g1 - gb]) - bit_base
    bitcode = host["ascii"][bit0:bit1].tobytes()
    return sample, bitcode, starts, outlier_extras


def _host_views(job, with_bits, calls=None):
    """Device->host copies of everything the caller gets back for the calls [calls[0], calls[1]) of the job (default: all):
    one small copy for the counts, then one batch of copies for the payload.  No kernel runs on this stream unless the
    '0'/'1' expansion has not been queued yet."""
    s0, s1 = (0, job.table.count) if calls is None else calls
    if not job.groups.synced:
        # nothing is known on the host yet (a fused encode just queued): queue the '0'/'1' expansion on capacity bounds and
        # bring the group count, the per-call offsets and the total bit count over in ONE small copy
        cap = job._cap
        if with_bits and not hasattr(job, "ascii"):
            job.expand_code_capacity()
        first = E.to_host([job.groups.seg_group_off, job.seg_outlier_off] + ([job.bit_off[cap:cap + 1]] if with_bits else []))
        job.groups.set_host(first[0].copy())
        job.sync_groups()
        pre = (first[1].copy(), int(first[2][0]) if with_bits else 0)
    else:
        pre = None
    g0, g1 = job.segment_group_range(s0)[0], job.segment_group_range(s1 - 1)[1]
    d0, d1 = job.table.starts[s0], job.table.starts[s1 - 1] + job.table.specs[s1 - 1].n
    if pre is not None and s0 == 0 and s1 == job.table.count:
        seg_out = pre[0]
        bit0, bit1 = 0, pre[1]
    else:
        small = [job.seg_outlier_off] + ([job.bit_off[g0:g0 + 1], job.bit_off[g1:g1 + 1]] if with_bits else [])
        small = E.to_host(small)                                 # copies only
        seg_out = small[0].copy()
        bit0, bit1 = (int(small[1][0]), int(small[2][0])) if with_bits else (0, 0)
    o0, o1 = int(seg_out[s0]), int(seg_out[s1])
    names = ["sample", "starts", "index", "outlier_idx", "outlier_q"]
    tensors = [job.sample[d0:d1], job.groups.starts[g0:g1 + 1], job.index[g0:g1], job.outlier_idx[o0:o1], job.outlier_q[o0:o1]]
    if with_bits:
        # packed bits -> the reference's '0'/'1' characters on the device, once for the whole batch
        names += ["bit_off", "ascii"]
        if hasattr(job, "ascii"):
            ascii_dev = job.ascii[bit0:bit1]
        else:
            total = int(E.to_host([job.bit_off[job.groups.n_groups:job.groups.n_groups + 1]])[0][0])
            ascii_dev = E.device_bits_to_ascii(job.bits, total)[bit0:bit1]
        tensors += [job.bit_off[g0:g1 + 1], ascii_dev]
    host = dict(zip(names, E.to_host(tensors)))
    host["index"] = host["index"].view(np.uint32)
    host["outlier_q"] = host["outlier_q"].view(np.uint16)
    host.update(seg_outlier_off=seg_out, dim0=d0, group0=g0, outlier0=o0, bit0=bit0, calls=(s0, s1))
    return host


# A batch above this many dims is cut into chunks of whole calls that flow through a three-stage pipeline
# (upload on a copy stream | kernels on the current stream | download + host formatting), so that host<->device
# copies and the host-side bit formatting of one chunk overlap the scoring kernel of the next.
PIPELINE_MIN_DIMS = 1 << 21
PIPELINE_CHUNKS = 4
# a sharded run slices the upload of a chunk across the ranks (+ one all-gather) only above this many dims: below it every
# rank copies the whole chunk itself, which is faster than the collective's launch latency (the small FIRST chunk)
SLICE_MIN_DIMS = 1 << 20


def _chunk_weights(world):
    """Relative sizes of the chunks of a pipelined batch.  One GPU: 1:11:11:1 -- the upload of the first chunk and the
    download + formatting of the last are the only stages nothing overlaps, so both are as small as whole calls allow
    (one image of the 24-image batch; measured with tools/profile_pipeline.py: 108.37 ms end to end against 108.51 for
    1:7:7:1, 108.91 for 1:3:3:1 and 109.41 for two chunks, device time 107.5).  Sharded runs move 1/world of the bytes per
    rank while every extra chunk still costs its fixed stages, so fewer chunks win (measured on 2 and 8 GPUs): 1:6:1 on
    2-3 ranks, 1:3 from 4 ranks on."""
    if world <= 1:
        if PIPELINE_CHUNKS == 4:
            return [1.0, 11.0, 11.0, 1.0]
        return [1.0] + [3.0] * (PIPELINE_CHUNKS - 2) + [1.0] if PIPELINE_CHUNKS > 2 else [1.0] * max(PIPELINE_CHUNKS, 1)
    if PIPELINE_CHUNKS != 4:                       # an explicit setting (tests, tools) applies to every world size
        return [1.0] + [3.0] * (PIPELINE_CHUNKS - 2) + [1.0] if PIPELINE_CHUNKS > 2 else [1.0] * max(PIPELINE_CHUNKS, 1)
    return [1.0, 6.0, 1.0] if world < 4 else [1.0, 3.0]


def _plan_chunks(sizes, n_chunks, weights=None):
    """Consecutive runs of calls: list of (first, last + 1), cut where the running dim count passes the chunk weights'
    cumulative shares (default: a third of the inner chunks' dims for the first and the last chunk)."""
    if weights is not None:
        n_chunks = len(weights)
    if n_chunks <= 1 or len(sizes) <= 1:
        return [(0, len(sizes))]
    if weights is None:
        weights = [1.0] + [3.0] * (n_chunks - 2) + [1.0] if n_chunks > 2 else [1.0, 1.0]
    total, wsum = float(sum(sizes)), float(sum(weights))
    targets = np.cumsum(weights)[:-1] * total / wsum
    bounds, acc, start = [], 0, 0
    for i, n in enumerate(sizes):
        acc += n
        if len(bounds) < len(targets) and acc >= targets[len(bounds)] and i + 1 < len(sizes):
            bounds.append((start, i + 1))
            start = i + 1
    bounds.append((start, len(sizes)))
    return bounds


_side_streams = {}


def _copy_streams(dev):
    """Upload, download and 'peek' streams of a device, created once: torch's caching allocator keeps a pool per
    stream, so fresh streams on every call would turn every buffer into a new cudaMalloc.  The peek stream carries the
    small copies that tell the host a chunk's group count; they wait for a grouping queued far ahead and must not sit
    in front of the downloads of an earlier chunk."""
    key = (dev.type, dev.index)
    if key not in _side_streams:
        _side_streams[key] = tuple(torch.cuda.Stream(device=dev) for _ in range(3))
    return _side_streams[key]


def _encode_batch_pipelined(targets, proposals, specs, chunks, return_indices, shard=None, trace=None):
    """trace: optional list that receives (label, host seconds since entry) marks (tools/profile_pipeline.py)."""
    t_entry = time.perf_counter()

    def mark(label):
        if trace is not None:
            trace.append((label, time.perf_counter() - t_entry))

    dev = E.cuda_device()
    compute = torch.cuda.current_stream()
    up, down, peek = _copy_streams(dev)
    up.wait_stream(compute)
    staged = []
    # a sharded run: every rank uploads 1/world of each chunk and the ranks all-gather the rest over NVLink; every rank
    # collects (downloads + formats) the results of its own calls only, unless the shard asks for all of them
    own = (0, len(specs))
    sliced = False
    if shard is not None and hasattr(shard, "call_range"):
        own = shard.call_range([sp.n for sp in specs])
        sliced = shard.can_slice_uploads(targets, proposals)

    def upload(i):                                   # stage 1: a chunk's inputs on the copy stream
        a, b = chunks[i]
        with torch.cuda.stream(up):
            if sliced and sum(sp.n for sp in specs[a:b]) >= SLICE_MIN_DIMS:
                tensors = shard.upload_sliced(targets[a:b], proposals[a:b], dev)
            else:
                tensors = _upload(targets[a:b], proposals[a:b], dev)
            ev = torch.cuda.Event()
            ev.record(up)
        staged.append((tensors, ev))

    upload(0)                                        # the other uploads are queued once chunk 0 is on its way (below)
    n = len(chunks)
    jobs, done, results = [None] * n, [None] * n, [None] * n

    def group(i):                                    # stage 2a: prologue + grouping (+ this rank's range, sharded runs)
        (a, b), (tensors, ev) = chunks[i], staged[i]
        compute.wait_event(ev)
        for t in tensors:
            t.record_stream(compute)
        jobs[i] = E.ISJob(*tensors, specs[a:b]).launch(L.STAGE_GROUP)
        # the group count (sharded: also the ranges) starts travelling to the host on the side stream
        if shard is not None:
            shard.plan_async(jobs[i], peek)
        else:
            jobs[i].groups.fetch_async(peek)

    def score(i):                                    # stage 2b: scoring; the group count / the rank's range stay on the device
        if shard is None:
            jobs[i].launch(L.STAGE_SCORE)
        else:
            shard.score_async(jobs[i])

    def finish(i):                                   # stage 2c: (all-gather,) decode, outliers, packing, '0'/'1' expansion
        job = jobs[i]
        if shard is not None:
            shard.gather(job)                        # host wait: the copy made right after THIS chunk's grouping
        else:
            job.groups.finish_fetch()
            job.sync_groups()
        job.launch(L.STAGE_FINISH)
        if not return_indices:
            job.expand_code_async()                  # sized by the group count the host has just learnt
        done[i] = torch.cuda.Event()
        done[i].record(compute)

    def collect(i):                                  # stage 3: download (copies only) + formatting while later chunks are scored
        job = jobs[i]
        a, b = chunks[i]
        lo, hi = max(own[0], a) - a, min(own[1], b) - a          # this rank's calls inside the chunk
        results[i] = [None] * (b - a)
        if hi > lo:
            down.wait_event(done[i])
            with torch.cuda.stream(down):
                host = _host_views(job, with_bits=not return_indices, calls=(lo, hi))
            for s in range(lo, hi):
                results[i][s] = _segment_result(job, s, host, return_indices)
            job.release()                # downloads synchronised: the buffers may serve the next batch of this shape
        jobs[i] = None

    # Order on the compute stream: G0 S0 | G1 F0 S1 | G2 F1 S2 | ... : the grouping of the next chunk is queued behind the
    # scoring of the current one and ahead of its finishing stages.  The host learns the group count of chunk i (needed to
    # size the all-gather of a sharded run and the expanded code) from a side-stream copy made right after G(i), i.e.
    # while chunk i-1 is still being scored, so the GPU never waits for it; download + formatting of chunk i overlap the
    # scoring of chunk i+1 and need no kernel (a kernel would queue behind the resident scoring CTAs).
    mark("upload 0 queued")
    group(0)
    mark("group 0 queued")
    score(0)
    mark("chunk 0 queued")
    for i in range(1, n):
        upload(i)
    mark("uploads queued")
    for i in range(n):
        if i + 1 < n:
            group(i + 1)
        finish(i)
        mark("finish %d queued" % i)
        if i + 1 < n:
            score(i + 1)
        if i >= 1:
            collect(i - 1)
            mark("chunk %d collected" % (i - 1))
    collect(n - 1)
    mark("chunk %d collected" % (n - 1))
    compute.wait_stream(down)
    return [r for chunk in results for r in chunk]


def code_grouped_importance_sample_batch(targets, proposals, seed, n_bits_per_group, max_group_size_bits=4,
                                         dim_kl_bit_limit=12, return_indices=False, shard=None):
    """Many reference calls in one launch sequence.  seed / n_bits_per_group / max_group_size_bits /
    dim_kl_bit_limit may be scalars or per-call lists.  Returns a list with, per call, exactly what
    code_grouped_importance_sample returns.  `shard` (see distributed.py) replaces the single-GPU
    encode stage by a sharded one."""
    if len(targets) > 1 and (shard is None or hasattr(shard, "plan_async")):
        specs = _batch_specs(targets, proposals, seed, n_bits_per_group, max_group_size_bits, dim_kl_bit_limit)
        sizes = [sp.n for sp in specs]
        if PIPELINE_MIN_DIMS <= sum(sizes) <= E.ASYNC_MAX_DIMS and min(sizes) > 0:
            world = shard._world_rank()[0] if shard is not None else 1
            chunks = _plan_chunks(sizes, PIPELINE_CHUNKS, _chunk_weights(world))
            if len(chunks) > 1:
                return _encode_batch_pipelined(list(targets), list(proposals), specs, chunks, return_indices, shard)
    job = _encode_batch(targets, proposals, seed, n_bits_per_group, max_group_size_bits, dim_kl_bit_limit, "full",
                        shard=shard)
    host = _host_views(job, with_bits=not return_indices)
    out = [_segment_result(job, s, host, return_indices) for s in range(job.table.count)]
    job.release()
    return out


def code_grouped_importance_sample(sess, target, proposal, seed, n_bits_per_group, max_group_size_bits=4,
                                   dim_kl_bit_limit=12, return_group_indices_only=False, return_indices=False,
                                   return_indices_only=False):
    """CIS:112-274.  Returns (sample f32[N], bitcode bytes of '0'/'1', group_start_indices int64
    ndarray ending in N, (outlier_indices int64[M], outlier_samples uint16[M]))."""
    _check_dtype(target, "Target")
    _check_dtype(proposal, "Proposal")
    num_dimensions = int(np.prod(np.shape(proposal.loc))) if not hasattr(proposal.loc, "numel") else int(proposal.loc.numel())
    if num_dimensions == 0:
        # CIS:183-204 with an empty range: starts [0, 0], one empty group, K = 1, index 0
        starts = np.array([0, 0], dtype=np.int64)
        empty = (np.zeros(0, np.int64), np.zeros(0, np.uint16))
        if return_group_indices_only:
            return starts, []
        if return_indices_only:
            return (np.int64(1),)
        if return_indices:
            return np.zeros(0, np.float32), [np.int64(1)], starts, empty
        return np.zeros(0, np.float32), b"1", starts, empty
    stage = "groups" if return_group_indices_only else ("indices" if return_indices_only else "full")
    job = _encode_batch([target], [proposal], seed, n_bits_per_group, max_group_size_bits, dim_kl_bit_limit, stage)
    if return_group_indices_only:
        # CIS:192: the KL (bits) of every group closed inside the loop, i.e. all but the last
        kls = job.total_kl.cpu().numpy()
        return job.host_starts(0), [k / np.log(2) for k in kls[:-1]]
    if return_indices_only:
        return tuple(np.int64(i) + 1 for i in job.host_indices(0))
    host = _host_views(job, with_bits=not return_indices)
    out = _segment_result(job, 0, host, return_indices)
    job.release()
    return out


def decode_grouped_importance_sample_batch(bitcodes, group_start_indices, proposals, seed, outlier_indices,
                                           outlier_samples, use_indices=False):
    """Batch of decode calls; group_start_indices[s] must already end with the number of dimensions."""
    dev = E.cuda_device()
    ns = len(proposals)
    seeds = list(seed) if isinstance(seed, (list, tuple)) else [seed] * ns
    for p in proposals:
        _check_dtype(p, "Proposal")
    p_loc = [E.as_device_f32(p.loc, dev, "Proposal") for p in proposals]
    p_scale = [E.as_device_f32(p.scale, dev, "Proposal") for p in proposals]
    sizes = [int(x.numel()) for x in p_loc]
    specs = [E.SegmentSpec(sizes[i], seeds[i], 1.0, 0, float("inf"), True) for i in range(ns)]
    table = E.segment_table(specs, dev)          # cached per (specs, device, stream): a codec decodes image after image
    groups = E._forced_groups(group_start_indices, table, dev)
    idx_parts = []
    for s in range(ns):
        ng = len(group_start_indices[s]) - 1
        if use_indices:
            part = np.asarray([int(v) - 1 for v in bitcodes[s]], dtype=np.int64)
            if part.size != ng:
                raise CwqError("need one index per group")
            idx_parts.append(part.astype(np.uint32))
        else:
            part, _ = E.elias_delta_parse(bitcodes[s], ng)
            idx_parts.append(part)
    index = torch.from_numpy(np.concatenate(idx_parts).view(np.int32)).to(dev)
    cat = (lambda xs: xs[0] if len(xs) == 1 else torch.cat(xs))
    pl, ps = cat(p_loc), cat(p_scale)
    sample = torch.empty(table.n, dtype=torch.float32, device=dev)
    E._lib.call("cwq_is_decode", E._ptr(groups.starts), E._ptr(groups.group_seed), E._ptr(groups.gid), E._ptr(index), table.n,
                E._ptr(pl), E._ptr(ps), E._ptr(sample), E._stream())
    pos, q = [], []
    for s in range(ns):
        oi = np.asarray(outlier_indices[s], dtype=np.int64).reshape(-1)
        oq = np.asarray(outlier_samples[s]).reshape(-1).astype(np.uint16)
        if oi.size != oq.size:
            raise CwqError("outlier_indices and outlier_samples differ in length")
        if oi.size and (oi.min() < 0 or oi.max() >= sizes[s]):
            raise CwqError("outlier index out of range")
        pos.append(oi + table.starts[s])
        q.append(oq)
    pos, q = np.concatenate(pos), np.concatenate(q)
    if pos.size:
        pos_d = torch.from_numpy(pos).to(dev)
        q_d = torch.from_numpy(q.view(np.int16)).to(dev)
        E._lib.call("cwq_outlier_decode", E._ptr(sample), E._ptr(pos_d), E._ptr(q_d), int(pos.size), E._stream())
    host = E.to_host([sample])[0]
    return [host[table.starts[s]:table.starts[s] + sizes[s]] for s in range(ns)]


def decode_grouped_importance_sample(sess, bitcode, group_start_indices, proposal, n_bits_per_group, seed,
                                     outlier_indices, outlier_samples, use_indices=False):
    """CIS:277-363 -> f32[N].  As in the reference, N is appended to the caller's group_start_indices
    list (CIS:294) and n_bits_per_group is unused.  use_indices follows the evident intent of
    CIS:88-96 (bitcode is then a sequence of 1-based indices); the shipped branch cannot run."""
    _check_dtype(proposal, "Proposal")
    num_dimensions = int(np.prod(np.shape(proposal.loc))) if not hasattr(proposal.loc, "numel") else int(proposal.loc.numel())
    group_start_indices.append(num_dimensions)
    if num_dimensions == 0:
        return np.zeros(0, np.float32)
    return decode_grouped_importance_sample_batch([bitcode], [group_start_indices], [proposal], seed, [outlier_indices],
                                                  [outlier_samples], use_indices)[0]
