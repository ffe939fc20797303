#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MIRACLE relative-entropy coding engine.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): a batch of 24 synthetic
768x512 images (Kodak-shaped), both PLN latent levels of each (N1 = 196 608 dims at 20 bits / <=16 dims
per group, N2 = 2 304 dims at 20 bits / <=4 dims per group), coded with the importance sampler.  A
"step" is one full encode of that batch: per-dim prologue -> KL-budget grouping -> candidate scoring /
argmax -> regeneration of the chosen latents -> outliers -> Elias-delta bitstream.

  value  : encoded latent groups / s, whole job, inputs resident in HBM (CUDA events, max over ranks)
  e2e    : same metric through the public API (code_grouped_importance_sample_batch) with HOST inputs:
           pinned host -> device copies of the four parameter arrays and device -> host copies of latents,
           group starts, outliers and the bitstream are inside the timed region
  N > 1  : strong scaling -- the batch's groups are cut into N contiguous work-balanced ranges, one per
           rank, and the per-group indices are all-gathered (NCCL) before the replicated decode/pack

`--impl reference` times the CPU restatement of the reference path (oracle/, all host cores) on a bounded
sample of the same workload: TensorFlow is not installed, so the reference itself cannot run here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "encoded_latent_groups_per_sec"
UNIT = "groups/s"
SEED = 42                      # miracle_arguments.py:147
N_BITS_PER_GROUP = 20          # {first,second}_level_n_bits_per_group
DIM_KL_BIT_LIMIT = 16
MAX_GROUP_SIZE_BITS = (4, 2)   # first / second level
# algorithmic thread-instructions per candidate-dim: 10 INT + 5.5 FP32 + 2 SFU
W_INSTR_IS = 17.5
SFU_PER_CD = 2.0


def workload(n_images, height=512, width=768):
    """Host arrays of the batch: per image level 2 then level 1 (the order pln.py codes them)."""
    from compression_without_quantization_b200.synthetic import pln_level_sizes, synthetic_posterior
    n1, n2 = pln_level_sizes(height, width)
    segs = []
    for img in range(n_images):
        segs.append((synthetic_posterior(n2, 1000 + img), MAX_GROUP_SIZE_BITS[1]))
        segs.append((synthetic_posterior(n1, img), MAX_GROUP_SIZE_BITS[0]))
    return segs, n1, n2


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window_begin(self):
        self.t0 = time.time()

    def window_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        rows = [r for t, r in self.rows if self.t0 is None or (self.t0 <= t <= (self.t1 or t) + 0.25)]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


def cpu_reference_sample(segs, threads, target_seconds=15.0):
    """Time the oracle port (oracle/cwq_oracle.c, pthreads over groups) on a bounded sample: the first
    groups of the first image's level-1 latents.  Returns groups/s etc. of the CPU path."""
    from oracle import oracle as O
    O.build()
    (q_loc, q_scale, p_loc, p_scale), mg = segs[1]
    t0 = time.perf_counter()
    t_loc, t_scale, kl, mask = O.is_prologue(q_loc, q_scale, p_loc, p_scale, DIM_KL_BIT_LIMIT)
    starts = O.group(kl, N_BITS_PER_GROUP, mg, True)
    total, K = O.group_totals(kl, starts)
    a, b = O.is_coef(t_loc, t_scale)
    t_pro = time.perf_counter() - t0
    sizes = np.diff(starts)
    pilot = min(4 * threads, starts.size - 1)
    t0 = time.perf_counter()
    O.is_encode_groups(a, b, starts, K, SEED, 0, pilot, threads)
    t_pilot = max(time.perf_counter() - t0, 1e-6)
    cd_pilot = float(np.sum(K[:pilot].astype(np.float64) * sizes[:pilot]))
    rate = cd_pilot / t_pilot
    cum = np.cumsum(K.astype(np.float64) * sizes)
    n_sample = int(np.searchsorted(cum, rate * target_seconds)) + 1
    n_sample = max(pilot, min(n_sample, starts.size - 1))
    t0 = time.perf_counter()
    O.is_encode_groups(a, b, starts, K, SEED, 0, n_sample, threads)
    dt = time.perf_counter() - t0
    # charge the O(N) prologue + grouping pro rata to the sampled groups
    dt_total = dt + t_pro * n_sample / (starts.size - 1)
    cand = float(np.sum(K[:n_sample].astype(np.float64)))
    cd = float(cum[n_sample - 1])
    return {"groups": n_sample, "seconds": dt_total, "groups_per_s": n_sample / dt_total, "candidates_per_s": cand / dt_total,
            "candidate_dims_per_s": cd / dt_total,
            "sample": "first {} of {} level-1 groups of image 0 ({:.3g} candidate-dims), oracle C port, {} threads".format(
                n_sample, starts.size - 1, cd, threads)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    segs, n1, n2 = workload(1)
    threads = os.cpu_count() or 1
    per_step = []
    info = None
    budget = float(os.environ.get("CWQ_BENCH_REFERENCE_SECONDS", "0")) or max(2.0, min(15.0, 120.0 / max(1, args.warmup + args.steps)))
    for i in range(args.warmup + args.steps):
        info = cpu_reference_sample(segs, threads, target_seconds=budget)
        if i >= args.warmup:
            per_step.append(info)
    gps = float(np.mean([x["groups_per_s"] for x in per_step]))
    line = {
        "impl": "reference", "metric": METRIC, "value": gps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([x["seconds"] for x in per_step])),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "24x(768x512) PLN latents, importance sampler, 20 bits/group (BASELINE configs[3]); "
                               "each step = a bounded sample of it", "l2": "n/a (CPU)"},
        "candidates_per_sec": float(np.mean([x["candidates_per_s"] for x in per_step])),
        "candidate_dims_per_sec": float(np.mean([x["candidate_dims_per_s"] for x in per_step])),
        "cpu_baseline": {"value": gps, "unit": UNIT, "cores": threads, "kind": "port", "sample": info["sample"],
                         "note": "CPU restatement of the reference (TensorFlow/TFP not installed, reference cannot run)"},
        "e2e": {"value": gps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    import compression_without_quantization_b200 as cwq
    from compression_without_quantization_b200 import _lib, distributed as D, engine as E
    from compression_without_quantization_b200.synthetic import Normal
    lib = _lib.load()

    segs, n1, n2 = workload(args.images)
    specs = [E.SegmentSpec(a[0].size, SEED, N_BITS_PER_GROUP, mg, DIM_KL_BIT_LIMIT, True) for a, mg in segs]
    host = [torch.from_numpy(np.concatenate([a[j] for a, _ in segs])).pin_memory() for j in range(4)]
    devt = [h.to(dev) for h in host]
    n_total = int(host[0].numel())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    shard_device = D.make_is_shard() if world > 1 else None

    def step_device():
        job = E.ISJob(devt[0], devt[1], devt[2], devt[3], specs)
        if world > 1:
            job.launch(_lib.STAGE_GROUP)
            shard_device(job)          # this rank's range is scored from device-side cuts; indices exchanged with the peers
            job.launch(_lib.STAGE_FINISH)
        else:
            job.launch(_lib.STAGE_GROUP | _lib.STAGE_SCORE | _lib.STAGE_FINISH)   # no host synchronisation inside the step
        return job

    # ---- device-resident throughput ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # started before the warm-up so nvidia-smi is already streaming in the timed window
    job = None
    for _ in range(max(args.warmup, 3)):
        job = step_device()
    barrier()
    job.sync_groups()
    n_groups = job.groups.n_groups
    sizes = (job.groups.starts[1:] - job.groups.starts[:-1]).to(torch.float64)
    Kd = job.K.to(torch.float64)
    candidates = float(Kd.sum().item())
    cand_dims = float((Kd * sizes).sum().item())
    if world > 1:
        b = job.shard_bounds
        my_cd = float((Kd[b[rank]:b[rank + 1]] * sizes[b[rank]:b[rank + 1]]).sum().item())
    else:
        my_cd = cand_dims

    lib.cwq_profile_enable(1)
    launches0 = lib.cwq_launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.window_begin()
    wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        job = step_device()
    ev1.record()
    barrier()
    wall = time.perf_counter() - wall0
    sampler.window_end()
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.cwq_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    score_ms = []
    import ctypes
    for i in range(lib.cwq_profile_count()):
        v = ctypes.c_float(0)
        _lib.check(lib.cwq_profile_get_ms(i, ctypes.cast(ctypes.pointer(v), ctypes.c_void_p)), "cwq_profile_get_ms")
        score_ms.append(v.value)
    lib.cwq_profile_enable(0)
    t = torch.tensor([ms_total, float(np.mean(score_ms)) if score_ms else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t[0].item()) / args.steps
    score_ms_max = float(t[1].item())

    # ---- end to end through the public API, host buffers in, host results out ----
    e2e = None
    if not args.no_e2e:
        off = np.cumsum([0] + [a[0].size for a, _ in segs])
        targets = [Normal(host[0][off[i]:off[i + 1]], host[1][off[i]:off[i + 1]]) for i in range(len(segs))]
        proposals = [Normal(host[2][off[i]:off[i + 1]], host[3][off[i]:off[i + 1]]) for i in range(len(segs))]
        mgs = [mg for _, mg in segs]
        shard = D.make_is_shard() if world > 1 else None

        def step_e2e():
            return cwq.code_grouped_importance_sample_batch(targets, proposals, SEED, N_BITS_PER_GROUP, mgs,
                                                            DIM_KL_BIT_LIMIT, shard=shard)
        for _ in range(2):
            res = step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = step_e2e()
        barrier()
        dt = (time.perf_counter() - t0) / args.steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        d2h = sum(r[0].nbytes + len(r[1]) // 8 + r[2].nbytes + r[3][0].nbytes + r[3][1].nbytes for r in res)
        e2e = {"value": n_groups / dt, "unit": UNIT, "h2d_bytes_per_step": 16 * n_total, "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * dt,
               "api": "code_grouped_importance_sample_batch (48 reference calls: 24 images x 2 levels), pinned host tensors in, "
                      "numpy latents / '0'-'1' bitstreams / group starts / outliers out"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    f_max = 1.965e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        try:
            f_max = float(json.load(open(peaks_path)).get("sm_max_mhz", 1965.0)) * 1e6
        except Exception:
            pass
    # the dominant kernel: k_is_score.  Algorithmic work = candidate-dims x 17.5 thread-instructions
    # (10 INT + 5.5 FP32 + 2 SFU per draw+score); peak = SMs x 128 lanes x max SM clock.
    score_s = score_ms_max * 1e-3 if score_ms_max > 0 else ms_per_step * 1e-3
    kernel_cd = my_cd if world > 1 else cand_dims
    cd_per_s_kernel = kernel_cd / score_s
    issue_peak = sms * 128 * f_max
    sfu_peak = sms * 16 * f_max
    achieved = cd_per_s_kernel * W_INSTR_IS
    f_obs = (clocks["sm_mhz"] * 1e6) if clocks and clocks.get("sm_mhz") else None
    roofline = {
        "bound": "fp32-issue", "kernel": "k_is_score", "achieved": achieved / 1e12, "peak": issue_peak / 1e12,
        "unit": "Tinstr/s", "frac": achieved / issue_peak,
        "sfu_frac": cd_per_s_kernel * SFU_PER_CD / sfu_peak,
        "frac_at_observed_clock": (achieved / (sms * 128 * f_obs)) if f_obs else None,
        # ideal-independent-chains microbenchmark of this kernel's instruction mix: 23.6 cycles per warp-candidate-dim
        # per sub-partition (tools/microbench/overlap.cu "v2 mix", profiles/r01_pipe_overlap_v2_mix.log)
        "frac_of_measured_mix_ceiling": cd_per_s_kernel / (sms * 4 * 32 / 23.6 * f_max),
        "candidate_dims_per_sec_kernel": cd_per_s_kernel, "kernel_ms": score_ms_max,
        "kernel_share_of_step": score_ms_max / ms_per_step if ms_per_step > 0 else None,
        # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on a 2-image
        # launch (4 043 264 B read, 0 B written; profiles/r01_is_score_ncu_summary.md), scaled to this launch's images
        "traffic": 4043264.0 * args.images / 2.0 / (world if world > 1 else 1),
        # thread-instructions the kernel executes per candidate-dim (ncu smsp__inst_executed x 32 / candidate-dims); the
        # 17.5 of `achieved` is SURVEY.md 8d's algorithmic count, kept as the reporting unit
        "executed_instr_per_candidate_dim": 16.1,
        "algorithmic_bytes": 8.0 * n_total / world + 16.0 * n_groups / world,
        "peak_source": "{} SMs x 128 FP32 lanes x {:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json); the path is on-chip "
                       "issue/SFU bound, not HBM or tensor bound (SURVEY.md 8d); pipe widths: profiles/r01_pipes_microbench*.log"
                       .format(sms, f_max / 1e6),
        "hbm_note": "HBM traffic of the kernel is ~8 B/dim + 16 B/group against >=1e3 candidate evaluations per dim",
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:   # the CPU leg is reported at N = 1 only
        threads = os.cpu_count() or 1
        c = cpu_reference_sample(segs, threads, target_seconds=15.0)
        cpu = {"value": c["groups_per_s"], "unit": UNIT, "cores": threads, "kind": "port", "sample": c["sample"],
               "candidates_per_sec": c["candidates_per_s"], "candidate_dims_per_sec": c["candidate_dims_per_s"],
               "note": "CPU restatement of the reference (TensorFlow/TFP not installed, the reference itself cannot run)"}

    s = ms_per_step * 1e-3
    line = {
        "metric": METRIC, "value": n_groups / s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "{}x(768x512) PLN latents (N1={}, N2={} per image), importance sampler, n_bits_per_group=20, "
                               "max_group_size_bits=4/2, dim_kl_bit_limit=16, seed=42 (BASELINE configs[3])".format(args.images, n1, n2),
                   "dims": n_total, "groups": n_groups, "candidates": candidates, "candidate_dims": cand_dims,
                   "parallelism": "groups sharded in {} contiguous work-balanced ranges + {} of indices".format(
                       world, "all-gather"),
                   "l2": "compute-bound path: inputs are 76 MB (< L2) but are read once per step by the O(N) stages only; "
                         "every step rewrites >126 MB of scratch (next/exit/gid/starts) between reads"},
        "candidates_per_sec": candidates / s, "candidate_dims_per_sec": cand_dims / s,
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
