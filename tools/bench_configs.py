#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs that are not the headline bench line: configs[0] (one 768x512
image, greedy), configs[1] (same image, importance sampler), configs[2] (1M-dim sweep at 8/12/16-bit budgets) and the
drop-in latency of one image -- the same numbers `bench.py` carries in its `secondary` block, one JSON line per entry.
`--big` adds non-default group sizes (32 / 64 / 128 dims per group) and `--c4` a bounded slice of configs[4]."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from compression_without_quantization_b200 import engine as E  # noqa: E402
from compression_without_quantization_b200.synthetic import pln_level_sizes, synthetic_posterior  # noqa: E402


def timed(fn, reps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--c4", action="store_true")
    args = ap.parse_args()
    dev = E.cuda_device()
    n1, n2 = pln_level_sizes(512, 768)
    for name, v in bench.secondary_measurements(dev, n1, n2).items():
        print(json.dumps(dict(config=name, **v)), flush=True)
    if args.big:   # 32- and 64-dim groups (shared-memory staged scoring path) and 128 (generic path); fused front end up to 128
        for m in (5, 6, 7):
            a = synthetic_posterior(1 << 20, 78)
            a = (a[0] * 0.3, np.exp(-np.abs(np.log(a[1])) * 0.2).astype(np.float32), a[2], a[3])
            t = [torch.from_numpy(x).to(dev) for x in a]
            specs = [E.SegmentSpec(1 << 20, 42, 20, m, 16, True)]
            ms, job = timed(lambda: E.ISJob(t[0], t[1], t[2], t[3], specs).launch(7), 5)
            job.sync_groups()
            cd = float((job.K.double() * (job.groups.starts[1:] - job.groups.starts[:-1]).double()).sum())
            print(json.dumps({"config": "1M low-KL dims, IS 20 bits, max_group_size_bits {}".format(m), "groups": job.groups.n_groups,
                              "step_ms": ms, "candidate_dims_per_sec": cd / ms * 1e3}), flush=True)
    if args.c4:
        m1, _ = pln_level_sizes(2160, 3840)
        a = synthetic_posterior(m1, 5)
        t = [torch.from_numpy(x).to(dev) for x in a]
        job = E.GreedyJob(t[0], t[1], t[2], t[3], [E.SegmentSpec(m1, 42, 30 * 14, 12, float("inf"), False)], 30, 14).prepare()
        g1 = min(job.groups.n_groups, 1184)
        ms, _ = timed(lambda: job.encode(0, g1), 1, warmup=1)
        cd = float((job.groups.starts[g1] - job.groups.starts[0]).item()) * 30 * (1 << 14)
        print(json.dumps({"config": "configs[4] slice: one 3840x2160 image, greedy 30 x 2^14, first 1184 groups", "step_ms": ms,
                          "candidate_dims_per_sec": cd / ms * 1e3}), flush=True)


if __name__ == "__main__":
    main()
