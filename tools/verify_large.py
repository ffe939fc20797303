#!/usr/bin/env python
"""Maximum-size check (BASELINE.json configs[4]: 64 x 3840x2160 images, 265 420 800 level-1 dims in ONE job).

The O(N) stages run over the whole batch (offsets beyond 2^31 bytes in every per-dim array); the candidate loops run on
the LAST groups of the batch only (a full encode is ~80 s of GPU time) and must reproduce, bit for bit, what the last
image gives when it is encoded alone: segments are independent, so any offset / index-width bug shows up as a difference.
Greedy: prologue, grouping, encode of the last groups, replay decode.  Importance sampler: prologue, grouping, K,
scoring of the last groups, decode + packing over everything."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from compression_without_quantization_b200 import engine as E  # noqa: E402
from compression_without_quantization_b200.synthetic import pln_level_sizes  # noqa: E402


def device_posterior(n, seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    q_scale = torch.exp(-torch.abs(torch.randn(n, generator=g, device=dev) * 0.7))
    q_loc = 1.2 * torch.randn(n, generator=g, device=dev) * (torch.rand(n, generator=g, device=dev) < 0.35)
    return q_loc.float(), q_scale.float(), torch.zeros(n, device=dev), torch.ones(n, device=dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--tail-groups", type=int, default=24)
    args = ap.parse_args()
    dev = E.cuda_device()
    n1, _ = pln_level_sizes(2160, 3840)
    parts = [device_posterior(n1, 20261018 + i, dev) for i in range(args.images)]
    big = [torch.cat([p[j] for p in parts]) for j in range(4)]
    last = [p.clone() for p in parts[-1]]
    del parts
    out = {"images": args.images, "dims": int(big[0].numel()), "bytes_per_f32_array": int(big[0].numel()) * 4}
    tg = args.tail_groups

    # ---- greedy ----
    t0 = time.perf_counter()
    gspec = lambda: E.SegmentSpec(n1, 42, 30 * 14, 12, float("inf"), False)
    jb = E.GreedyJob(*big, [gspec() for _ in range(args.images)], 30, 14).prepare()
    js = E.GreedyJob(*last, [gspec()], 30, 14).prepare()
    ngb, ngs = jb.groups.n_groups, js.groups.n_groups
    off = jb.table.starts[-1]
    seg0 = int(jb.groups.seg_group_off_host[-2])
    ok = {"greedy_group_count": ngb - seg0 == ngs,
          "greedy_starts": bool(torch.equal(jb.groups.starts[seg0:] - off, js.groups.starts)),
          "greedy_kl": bool(torch.equal(jb.kl[off:], js.kl)),
          "greedy_seeds": bool(torch.equal(jb.groups.group_seed[seg0:], js.groups.group_seed))}
    jb.encode(ngb - tg, ngb)
    js.encode(ngs - tg, ngs)
    d0b, d0s = int(jb.groups.starts[ngb - tg]), int(js.groups.starts[ngs - tg])
    ok["greedy_indices"] = bool(torch.equal(jb.indices[(ngb - tg) * 30:], js.indices[(ngs - tg) * 30:]))
    ok["greedy_best"] = bool(torch.equal(jb.best[d0b:], js.best[d0s:]))
    jb.decode(); js.decode()
    ok["greedy_replay"] = bool(torch.equal(jb.sample[d0b:], js.sample[d0s:]))
    jb.pack(); js.pack()
    nb = tg * 30 * 14
    bits_b = E.device_bits_to_ascii(jb.bits, ngb * 420)[-nb:]
    bits_s = E.device_bits_to_ascii(js.bits, ngs * 420)[-nb:]
    ok["greedy_bits"] = bool(torch.equal(bits_b, bits_s))
    torch.cuda.synchronize()
    out.update(greedy_groups=ngb, greedy_seconds=time.perf_counter() - t0, greedy_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
    del jb, js
    torch.cuda.empty_cache()

    # ---- importance sampler ----
    t0 = time.perf_counter()
    ispec = lambda: E.SegmentSpec(n1, 42, 20, 4, 16, True)
    ib = E.ISJob(*big, [ispec() for _ in range(args.images)]).prepare(sync=True)
    isml = E.ISJob(*last, [ispec()]).prepare(sync=True)
    ngb, ngs = ib.groups.n_groups, isml.groups.n_groups
    seg0 = int(ib.groups.seg_group_off_host[-2])
    ok.update({"is_group_count": ngb - seg0 == ngs,
               "is_starts": bool(torch.equal(ib.groups.starts[seg0:] - off, isml.groups.starts)),
               "is_K": bool(torch.equal(ib.K[seg0:], isml.K)), "is_coef": bool(torch.equal(ib.coef[2 * off:], isml.coef)),
               "is_gid": bool(torch.equal(ib.groups.gid[off:] - seg0, isml.groups.gid))})
    tgi = 200 * tg
    ib.encode(ngb - tgi, ngb)
    isml.encode(ngs - tgi, ngs)
    ok["is_indices"] = bool(torch.equal(ib.index[ngb - tgi:], isml.index[ngs - tgi:]))
    ib.decode().pack()
    isml.decode().pack()
    d0b, d0s = int(ib.groups.starts[ngb - tgi]), int(isml.groups.starts[ngs - tgi])
    ok["is_sample"] = bool(torch.equal(ib.sample[d0b:], isml.sample[d0s:]))
    lb = ib.bit_off[ngb - tgi:] - ib.bit_off[ngb - tgi]
    ls = isml.bit_off[ngs - tgi:] - isml.bit_off[ngs - tgi]
    ok["is_code_lengths"] = bool(torch.equal(lb, ls))
    torch.cuda.synchronize()
    out.update(is_groups=ngb, is_seconds=time.perf_counter() - t0, is_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30, checks=ok,
               all_ok=all(ok.values()))
    print(json.dumps(out), flush=True)
    sys.exit(0 if out["all_ok"] else 1)


if __name__ == "__main__":
    main()
