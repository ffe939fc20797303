#!/usr/bin/env python
"""Per-step device times of the fused importance-sampler pipeline on the small configurations (BASELINE configs[1] and
configs[2]): every step timed on its own with CUDA events (min / median / max), the scoring kernel's share from the
library's own event pair, and the host time spent enqueuing a step."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from compression_without_quantization_b200 import _lib, engine as E  # noqa: E402
from compression_without_quantization_b200.synthetic import pln_level_sizes, synthetic_posterior  # noqa: E402


def run(name, arrays, specs, steps=40):
    dev = E.cuda_device()
    lib = _lib.load()
    t = [torch.from_numpy(np.concatenate(a)).to(dev) for a in zip(*arrays)]
    for _ in range(5):
        job = E.ISJob(t[0], t[1], t[2], t[3], specs).launch(7)
    torch.cuda.synchronize()
    # throughput: steps queued back to back (the GPU is never idle between them)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        job = E.ISJob(t[0], t[1], t[2], t[3], specs).launch(7)
    e1.record()
    torch.cuda.synchronize()
    b2b_ms = e0.elapsed_time(e1) / steps
    # per-kernel event times inside one step
    lib.cwq_profile_enable(2)
    for _ in range(4):
        job = E.ISJob(t[0], t[1], t[2], t[3], specs).launch(7)
    torch.cuda.synchronize()
    names = ["score", "tiles", "chain", "groups", "prefix", "index", "finish_dims", "pack"]
    per = {}
    for i in range(lib.cwq_profile_count()):
        v = ctypes.c_float(0)
        lib.cwq_profile_get_ms(i, ctypes.cast(ctypes.pointer(v), ctypes.c_void_p))
        per.setdefault(names[lib.cwq_profile_kind(i)], []).append(v.value * 1e3)
    kernels_us = {k: round(float(np.median(v)), 1) for k, v in per.items()}
    lib.cwq_profile_enable(1)
    ms, host = [], []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        h0 = time.perf_counter()
        e0.record()
        job = E.ISJob(t[0], t[1], t[2], t[3], specs).launch(7)
        e1.record()
        host.append(1e3 * (time.perf_counter() - h0))
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    score = []
    for i in range(lib.cwq_profile_count()):
        v = ctypes.c_float(0)
        lib.cwq_profile_get_ms(i, ctypes.cast(ctypes.pointer(v), ctypes.c_void_p))
        score.append(v.value)
    lib.cwq_profile_enable(0)
    job.sync_groups()
    sizes = (job.groups.starts[1:] - job.groups.starts[:-1]).double()
    cd = float((job.K.double() * sizes).sum())
    ms, host, score = np.array(ms), np.array(host), np.array(score)
    print(json.dumps({"config": name, "dims": int(t[0].numel()), "groups": job.groups.n_groups, "candidate_dims": cd,
                      "step_ms_back_to_back": b2b_ms, "kernels_us": kernels_us,
                      "step_ms": {"min": float(ms.min()), "median": float(np.median(ms)), "max": float(ms.max())},
                      "score_kernel_ms_median": float(np.median(score)), "host_enqueue_ms_median": float(np.median(host)),
                      "cd_per_s_step": cd / np.median(ms) * 1e3, "cd_per_s_kernel": cd / np.median(score) * 1e3}), flush=True)


def main():
    n1, n2 = pln_level_sizes(512, 768)
    if "--batch" in sys.argv:
        arrays, specs = [], []
        for img in range(24):
            arrays += [synthetic_posterior(n2, 1000 + img), synthetic_posterior(n1, img)]
            specs += [E.SegmentSpec(n2, 42, 20, 2, 16, True), E.SegmentSpec(n1, 42, 20, 4, 16, True)]
        run("configs[3]: 24 x (768x512), IS 20 bits", arrays, specs, steps=5)
        return
    run("configs[1]: one 768x512 image, IS 20 bits", [synthetic_posterior(n2, 1000), synthetic_posterior(n1, 0)],
        [E.SegmentSpec(n2, 42, 20, 2, 16, True), E.SegmentSpec(n1, 42, 20, 4, 16, True)])
    for bits in (8, 12, 16):
        run("configs[2]: 1M dims, IS {} bits".format(bits), [synthetic_posterior(1 << 20, 77)],
            [E.SegmentSpec(1 << 20, 42, bits, 4, bits - 2, True)])


if __name__ == "__main__":
    main()
