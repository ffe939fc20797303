#!/usr/bin/env python
"""BASELINE.json configs[4]: a batch of synthetic 3840x2160 images (level-1 latents, 4 147 200 dims each), greedy sampler
30 steps x 2^14 candidates, groups sharded over the ranks of one box (contiguous ranges + all-gather of indices).

    python tools/bench_greedy_4k.py --images 8                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_greedy_4k.py --images 64

Inputs are generated on the device with the SURVEY.md 8d recipe (same seed on every rank -> identical batches).
One warm-up step on one image, then --steps timed steps (device timing, barrier + synchronize on both sides, max over
ranks).  Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from compression_without_quantization_b200 import distributed as D, engine as E  # noqa: E402
from compression_without_quantization_b200.synthetic import pln_level_sizes  # noqa: E402


def device_posterior(n, seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    q_scale = torch.exp(-torch.abs(torch.randn(n, generator=g, device=dev) * 0.7))
    q_loc = 1.2 * torch.randn(n, generator=g, device=dev) * (torch.rand(n, generator=g, device=dev) < 0.35)
    return q_loc.float(), q_scale.float(), torch.zeros(n, device=dev), torch.ones(n, device=dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--n-steps", type=int, default=30)
    ap.add_argument("--n-bits", type=int, default=14)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    n1, _ = pln_level_sizes(2160, 3840)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(n_images):
        parts = [device_posterior(n1, 20261018 + i, dev) for i in range(n_images)]
        t = [torch.cat([p[j] for p in parts]) for j in range(4)]
        del parts
        specs = [E.SegmentSpec(n1, 42, args.n_steps * args.n_bits, 12, float("inf"), False) for _ in range(n_images)]
        job = E.GreedyJob(t[0], t[1], t[2], t[3], specs, args.n_steps, args.n_bits).prepare()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world > 1:
            D.make_greedy_shard()(job)
        else:
            job.encode()
        job.unrescale().pack()
        e1.record()
        barrier()
        return job, e0.elapsed_time(e1)

    step(1)                                  # warm-up: kernels loaded, NCCL communicator built
    times, job = [], None
    for _ in range(args.steps):
        job, ms = step(args.images)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    ms = sum(times) / len(times)
    ng = job.groups.n_groups
    cd = float(job.n) * args.n_steps * (1 << args.n_bits)
    # position-weighted checksums of the index stream, the packed bitstream and the decoded latents: equal across rank counts
    def checksum(x):
        x = x.reshape(-1)
        w = (torch.arange(x.numel(), device=x.device, dtype=torch.int64) % 65521) + 1
        return int((x.to(torch.int64) * w).sum().item()) & 0xFFFFFFFFFFFF
    digest = "{:012x}-{:012x}-{:012x}".format(checksum(job.indices), checksum(job.bits), checksum(job.sample.view(torch.int32)))
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    issue_peak = sms * 128 * 1.965e9
    if rank == 0:
        print(json.dumps({"config": "configs[4]: {} x 3840x2160 level-1 latents, greedy {} x 2^{}".format(args.images, args.n_steps, args.n_bits),
                          "n_gpus": world, "images": args.images, "dims": job.n, "groups": ng, "ms_per_step": ms,
                          "groups_per_s": ng / ms * 1e3, "candidates_per_s": ng * args.n_steps * (1 << args.n_bits) / ms * 1e3,
                          "candidate_dims_per_s": cd / ms * 1e3, "steps": args.steps, "digest": digest,
                          # 18.5 algorithmic thread-instructions per candidate-dim (greedy); 15.7 executed (ncu,
                          # profiles/r01_greedy_ncu_summary.md); peak = SMs x 128 lanes x 1965 MHz per GPU
                          "roofline_frac": cd / ms * 1e3 * 18.5 / (issue_peak * world),
                          "roofline_frac_executed": cd / ms * 1e3 * 15.7 / (issue_peak * world),
                          "gpu_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                          "scaling": "weak when --images grows with the ranks (8 per GPU), strong for a fixed --images"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
