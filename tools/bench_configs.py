#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs that are not the headline bench line:
configs[0] (one 768x512 image, greedy), configs[1] (same image, importance sampler), configs[2]
(1M-dim sweep at 8/12/16-bit budgets) and a bounded slice of configs[4] (3840x2160 greedy).
Prints one JSON line per config; device-resident timing with CUDA events, encode pipeline only."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from compression_without_quantization_b200 import engine as E  # noqa: E402
from compression_without_quantization_b200.synthetic import pln_level_sizes, synthetic_posterior  # noqa: E402


def timed(fn, reps, warmup=3):
    for _ in range(warmup):     # kernels loaded, allocator pools at their steady-state size
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def is_config(name, arrays, specs, reps):
    dev = E.cuda_device()
    t = [torch.from_numpy(np.concatenate(a)).to(dev) for a in zip(*arrays)]

    def run():
        return E.ISJob(t[0], t[1], t[2], t[3], specs).launch(7)   # group + score + finish, one library call
    ms, job = timed(run, reps)
    job.sync_groups()
    sizes = (job.groups.starts[1:] - job.groups.starts[:-1]).double()
    K = job.K.double()
    cd = float((K * sizes).sum())
    print(json.dumps({"config": name, "sampler": "importance", "dims": int(t[0].numel()), "groups": job.groups.n_groups,
                      "mean_K": float(K.mean()), "max_K": float(K.max()), "candidate_dims": cd, "ms": ms,
                      "groups_per_s": job.groups.n_groups / ms * 1e3, "candidates_per_s": float(K.sum()) / ms * 1e3,
                      "candidate_dims_per_s": cd / ms * 1e3}), flush=True)


def greedy_config(name, arrays, specs, n_steps, n_bits, reps, g_limit=None, mode=0):
    dev = E.cuda_device()
    t = [torch.from_numpy(np.concatenate(a)).to(dev) for a in zip(*arrays)]
    job = E.GreedyJob(t[0], t[1], t[2], t[3], specs, n_steps, n_bits).prepare()
    ng = job.groups.n_groups
    g1 = ng if g_limit is None else min(ng, g_limit)

    def run():
        job.encode(0, g1, mode=mode)
        return job
    ms, _ = timed(run, reps, warmup=1)
    sizes = (job.groups.starts[1:g1 + 1] - job.groups.starts[:g1]).double()
    cd = float(sizes.sum()) * n_steps * (1 << n_bits)
    print(json.dumps({"config": name, "sampler": "greedy", "dims": int(t[0].numel()), "groups": ng, "groups_timed": g1,
                      "mean_group_size": float(sizes.mean()), "candidate_dims": cd, "ms": ms, "groups_per_s": g1 / ms * 1e3,
                      "candidates_per_s": g1 * n_steps * (1 << n_bits) / ms * 1e3, "candidate_dims_per_s": cd / ms * 1e3}),
          flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    n1, n2 = pln_level_sizes(512, 768)
    want = (lambda k: not args.only or k in args.only.split(","))
    if want("c1"):
        arrays = [synthetic_posterior(n2, 1000), synthetic_posterior(n1, 0)]
        specs = [E.SegmentSpec(n2, 42, 20, 2, 16, True), E.SegmentSpec(n1, 42, 20, 4, 16, True)]
        is_config("configs[1]: one 768x512 image, IS 20 bits", arrays, specs, args.reps)
    if want("c2"):
        for bits in (8, 12, 16):
            arrays = [synthetic_posterior(1 << 20, 77)]
            specs = [E.SegmentSpec(1 << 20, 42, bits, 4, bits - 2, True)]
            is_config("configs[2]: 1M dims, IS {} bits".format(bits), arrays, specs, args.reps)
    if want("big"):   # non-default group sizes: 32- and 64-dim groups (shared-memory staged path) and 128 (generic path)
        for m in (5, 6, 7):
            arrays = [synthetic_posterior(1 << 20, 78)]
            arrays = [(a[0] * 0.3, np.exp(-np.abs(np.log(a[1])) * 0.2).astype(np.float32), a[2], a[3]) for a in arrays]
            specs = [E.SegmentSpec(1 << 20, 42, 20, m, 16, True)]
            is_config("1M low-KL dims, IS 20 bits, max_group_size_bits {}".format(m), arrays, specs, args.reps)
    if want("c0"):
        arrays = [synthetic_posterior(n1, 0)]
        specs = [E.SegmentSpec(n1, 42, 30 * 14, 12, float("inf"), False)]
        greedy_config("configs[0]: one 768x512 image level 1, greedy 30 x 2^14", arrays, specs, 30, 14, max(1, args.reps // 3))
    if want("c4"):
        m1, _ = pln_level_sizes(2160, 3840)
        arrays = [synthetic_posterior(m1, 5)]
        specs = [E.SegmentSpec(m1, 42, 30 * 14, 12, float("inf"), False)]
        greedy_config("configs[4] slice: one 3840x2160 image level 1, greedy 30 x 2^14, first 1184 groups", arrays, specs, 30, 14, 1,
                      g_limit=1184)
    if want("c4modes"):
        m1, _ = pln_level_sizes(2160, 3840)
        arrays = [synthetic_posterior(m1, 5)]
        specs = [E.SegmentSpec(m1, 42, 30 * 14, 12, float("inf"), False)]
        for mode in (1, 2):
            greedy_config("configs[4] one 3840x2160 image, all 7535 groups, mode {}".format(mode), arrays, specs, 30, 14, 1, mode=mode)
            greedy_config("configs[4] slice 1184 groups, mode {}".format(mode), arrays, specs, 30, 14, 1, g_limit=1184, mode=mode)


if __name__ == "__main__":
    main()
