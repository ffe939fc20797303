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
           rank, and the per-group indices are all-gathered (NCCL) before the replicated decode/pack; end to
           end every rank uploads 1/N of the inputs (all-gathered over NVLink) and returns its own calls
  result_digest / digest_matches_single : sha256 over every call's latents, bitstream, group starts and
           outliers of the e2e step; at N > 1 rank 0 also runs the unsharded encode and compares per call

`--impl reference` times the CPU restatement of the reference path (oracle/, all host cores) on a bounded
sample of the same workload: TensorFlow is not installed, so the reference itself cannot run here.
"""
import argparse
import ctypes
import hashlib
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
W_INSTR_GREEDY = 18.5
SFU_PER_CD = 2.0
NCU_FACTS = os.path.join(ROOT, "profiles", "r02_is_score_ncu.json")    # written by tools/ncu_is_score.py from an ncu capture


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


# ----------------------------------------------------------------------------------------------------------------------
# CPU legs: the reference itself if TensorFlow ever imports; else the restatement -- (2a) numpy, shaped like
# the reference (one group at a time, [K, G] candidates materialised, float32), single process; (2b) the C port of the
# oracle on all host cores.  Plus the greedy sampler's C port.  All on bounded samples.
# ----------------------------------------------------------------------------------------------------------------------
def tensorflow_status():
    try:
        import tensorflow  # noqa: F401
        return "importable (version {}), but tensorflow_probability / the reference's TF-1.x session path were not attempted"\
            .format(getattr(tensorflow, "__version__", "?"))
    except Exception as e:      # ModuleNotFoundError in this image
        return "not importable here ({}): the reference's own TF path cannot be timed".format(type(e).__name__)


def _is_problem(segs):
    from oracle import oracle as O
    O.build()
    (q_loc, q_scale, p_loc, p_scale), mg = segs[1]
    t0 = time.perf_counter()
    t_loc, t_scale, kl, mask = O.is_prologue(q_loc, q_scale, p_loc, p_scale, DIM_KL_BIT_LIMIT)
    starts = O.group(kl, N_BITS_PER_GROUP, mg, True)
    total, K = O.group_totals(kl, starts)
    a, b = O.is_coef(t_loc, t_scale)
    return O, a, b, starts, K, time.perf_counter() - t0


def cpu_reference_sample(segs, threads, target_seconds=15.0):
    """Time the oracle port (oracle/cwq_oracle.c, pthreads over groups) on a bounded sample: the first
    groups of the first image's level-1 latents.  Returns groups/s etc. of the CPU path."""
    O, a, b, starts, K, t_pro = _is_problem(segs)
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


def cpu_numpy_reference_shaped(segs, target_seconds=6.0):
    """Leg (2a): what the reference does per group (CIS:233-245, 48-63) with numpy standing in for the TF ops: the [K, G]
    table of candidates materialised, log q - log p summed over dims, argmax, row `index` kept -- one group at a time, one
    process, float32.  The candidate table comes from the spec's stream (C, libm), as TF's stateless_normal would."""
    O, a, b, starts, K, t_pro = _is_problem(segs)
    t0 = time.perf_counter()
    done, cd, cand = 0, 0.0, 0.0
    for g in range(starts.size - 1):
        s, e = int(starts[g]), int(starts[g + 1])
        Kg = int(K[g])
        if e > s:
            z = O.default_normal_source((SEED + g) & 0xFFFFFFFF, 0, Kg, e - s)           # [K, G] float32
            w = np.sum(z * (a[s:e][None, :] * z + b[s:e][None, :]), axis=1, dtype=np.float32)
            idx = int(np.argmax(w))
            _ = z[idx]
        done += 1
        cd += Kg * (e - s)
        cand += Kg
        if time.perf_counter() - t0 > target_seconds:
            break
    dt = time.perf_counter() - t0 + t_pro * done / (starts.size - 1)
    return {"value": done / dt, "unit": UNIT, "cores": 1, "kind": "port", "candidates_per_sec": cand / dt,
            "candidate_dims_per_sec": cd / dt,
            "sample": "first {} level-1 groups of image 0 ({:.3g} candidate-dims), numpy restatement shaped like the reference "
                      "(one group at a time, [K,G] materialised), 1 process".format(done, cd)}


def cpu_greedy_sample(threads, target_seconds=8.0):
    """The greedy sampler's C port (oracle) on the first groups of one 768x512 image's level-1 latents at the reference
    defaults (30 steps x 2^14 candidates, <= 4095 dims per group): BASELINE configs[0] / configs[4]'s CPU side."""
    from compression_without_quantization_b200.synthetic import pln_level_sizes, synthetic_posterior
    from oracle import oracle as O
    O.build()
    n1, _ = pln_level_sizes(512, 768)
    q_loc, q_scale, p_loc, p_scale = synthetic_posterior(n1, 0)
    t_loc, t_scale, kl = O.greedy_prologue(q_loc, q_scale, p_loc, p_scale)
    starts = O.group(kl, 30 * 14, 12, False)
    s = O.greedy_shard_scale(1.0, 30)
    ng = starts.size - 1
    n_sample = min(ng, max(threads, 1))
    t0 = time.perf_counter()
    O.greedy_encode_groups(t_loc, t_scale, starts, 30, 1 << 14, s, SEED, 0, n_sample, threads)
    dt = time.perf_counter() - t0
    if dt < target_seconds / 3 and n_sample < ng:          # one more, larger, round inside the budget
        n_sample = min(ng, int(n_sample * min(4.0, target_seconds / max(dt, 1e-3) * 0.6)))
        t0 = time.perf_counter()
        O.greedy_encode_groups(t_loc, t_scale, starts, 30, 1 << 14, s, SEED, 0, n_sample, threads)
        dt = time.perf_counter() - t0
    cd = float(np.sum(np.diff(starts)[:n_sample])) * 30 * (1 << 14)
    return {"value": n_sample / dt, "unit": UNIT, "cores": threads, "kind": "port", "candidate_dims_per_sec": cd / dt,
            "sample": "first {} of {} greedy groups of one 768x512 image ({:.3g} candidate-dims), oracle C port, {} threads".format(
                n_sample, ng, cd, threads)}


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
                         "tensorflow": tensorflow_status(),
                         "note": "CPU restatement of the reference (C port of the oracle, all host cores): far faster than the "
                                 "reference's TF session loop (published: 81-102 s per Kodak image ~ 150 groups/s)"},
        "e2e": {"value": gps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def call_digests(results):
    """sha256 per call over what the reference call returns (latents, code, group starts, outlier indices / values);
    None for calls another rank owns."""
    out = []
    for r in results:
        if r is None:
            out.append(None)
            continue
        s, code, starts, (oi, oq) = r
        h = hashlib.sha256()
        for a in (s, starts, oi, oq):
            h.update(np.ascontiguousarray(a).tobytes())
        h.update(code if isinstance(code, (bytes, bytearray)) else str(code).encode())
        out.append(h.hexdigest())
    return out


def secondary_measurements(dev, n1, n2):
    """The other single-GPU BASELINE configs, each with its own rate: configs[1] (one image, importance sampler),
    configs[2] (1M-dim sweep at 8 / 12 / 16 bits), configs[0] (one image, greedy), and the latency of the drop-in call
    sequence of pln.py (level 2, then level 1, numpy in / numpy out)."""
    import torch
    import compression_without_quantization_b200 as cwq
    from compression_without_quantization_b200 import _lib, engine as E
    from compression_without_quantization_b200.synthetic import Normal, synthetic_posterior
    lib = _lib.load()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    out = {}

    def is_case(name, arrays, specs, steps=10):
        t = [torch.from_numpy(np.concatenate(a)).to(dev) for a in zip(*arrays)]
        for _ in range(3):
            job = E.ISJob(t[0], t[1], t[2], t[3], specs).launch(7)
        torch.cuda.synchronize()
        lib.cwq_profile_enable(1)
        ms = float("inf")
        for _ in range(2):      # best of two runs of `steps` back-to-back steps: a one-off cudaMalloc of the caching allocator
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)   # otherwise lands in a 0.2 ms step
            e0.record()
            for _ in range(steps):
                job = E.ISJob(t[0], t[1], t[2], t[3], specs).launch(7)
            e1.record()
            torch.cuda.synchronize()
            ms = min(ms, e0.elapsed_time(e1) / steps)
        score = []
        for i in range(lib.cwq_profile_count()):
            v = ctypes.c_float(0)
            lib.cwq_profile_get_ms(i, ctypes.cast(ctypes.pointer(v), ctypes.c_void_p))
            score.append(v.value)
        lib.cwq_profile_enable(0)
        job.sync_groups()
        sizes = (job.groups.starts[1:] - job.groups.starts[:-1]).double()
        K = job.K.double()
        cd = float((K * sizes).sum())
        # candidate-dims the kernel has to GENERATE: a group of G dims costs the Philox blocks of ceil(G/4) quads (3 per 16
        # dims), so partially filled units count in full here; the ratio says how much of the gap to the 16-dim case is
        # group-size mix rather than kernel overhead
        nq = torch.div(sizes + 3, 4, rounding_mode="floor")
        blocks = 3 * torch.div(nq, 4, rounding_mode="floor") + torch.clamp(nq % 4, max=3)
        cd_generated = float((K * blocks).sum()) * 16.0 / 3.0
        k_ms = float(np.median(score)) if score else None
        peak = sms * 128 * 1.965e9 / W_INSTR_IS
        out[name] = {"dims": int(t[0].numel()), "groups": job.groups.n_groups, "mean_K": float(K.mean()), "candidate_dims": cd,
                     "step_ms": ms, "score_kernel_ms": k_ms, "candidate_dims_per_sec": cd / ms * 1e3,
                     "candidate_dims_per_sec_kernel": cd / k_ms * 1e3 if k_ms else None,
                     "frac_of_issue_roofline_step": cd / ms * 1e3 / peak,
                     "frac_of_issue_roofline_kernel": cd / k_ms * 1e3 / peak if k_ms else None,
                     "generated_over_scored_candidate_dims": cd_generated / cd if cd else None}

    is_case("configs[1] one 768x512 image, IS 20 bits", [synthetic_posterior(n2, 1000), synthetic_posterior(n1, 0)],
            [E.SegmentSpec(n2, SEED, 20, 2, 16, True), E.SegmentSpec(n1, SEED, 20, 4, 16, True)])
    for bits in (8, 12, 16):
        is_case("configs[2] 1M dims, IS {} bits".format(bits), [synthetic_posterior(1 << 20, 77)],
                [E.SegmentSpec(1 << 20, SEED, bits, 4, bits - 2, True)])

    # configs[0]: greedy, one image, reference defaults
    a = synthetic_posterior(n1, 0)
    t = [torch.from_numpy(x).to(dev) for x in a]
    gjob = E.GreedyJob(t[0], t[1], t[2], t[3], [E.SegmentSpec(n1, SEED, 30 * 14, 12, float("inf"), False)], 30, 14).prepare()
    gjob.encode()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        gjob.encode()
    e1.record()
    torch.cuda.synchronize()
    gms = e0.elapsed_time(e1) / 2
    gcd = float(n1) * 30 * (1 << 14)
    out["configs[0] one 768x512 image, greedy 30 x 2^14"] = {
        "dims": n1, "groups": gjob.groups.n_groups, "candidate_dims": gcd, "step_ms": gms, "candidate_dims_per_sec": gcd / gms * 1e3,
        "groups_per_sec": gjob.groups.n_groups / gms * 1e3,
        "frac_of_issue_roofline_step": gcd / gms * 1e3 / (sms * 128 * 1.965e9 / W_INSTR_GREEDY)}

    # the drop-in call sequence of pln.py:350,423 for one image: level 2 then level 1, numpy in, numpy / bytes out
    d2, d1 = synthetic_posterior(n2, 1000), synthetic_posterior(n1, 0)

    def one_image():
        r2 = cwq.code_grouped_importance_sample(None, Normal(d2[0], d2[1]), Normal(d2[2], d2[3]), SEED, 20, 2, 16)
        r1 = cwq.code_grouped_importance_sample(None, Normal(d1[0], d1[1]), Normal(d1[2], d1[3]), SEED, 20, 4, 16)
        return r2, r1
    for _ in range(3):
        one_image()
    torch.cuda.synchronize()
    lat = []
    for _ in range(8):
        t0 = time.perf_counter()
        one_image()
        lat.append(1e3 * (time.perf_counter() - t0))
    dev_ms = out["configs[1] one 768x512 image, IS 20 bits"]["step_ms"]
    # what the copies of that call sequence cost alone: the eight pageable numpy inputs up (the caller's arrays are not
    # pinned), result-sized buffers down through pinned staging -- the part of the overhead no host code can remove
    r2, r1 = one_image()
    down_bytes = sum(x[0].nbytes + len(x[1]) + x[2].nbytes + x[3][0].nbytes + x[3][1].nbytes for x in (r2, r1))
    down_dev = torch.empty(down_bytes, dtype=torch.uint8, device=dev)
    floor = []
    for _ in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for arrs in (d2, d1):
            keep = [E.as_device_f32(x, dev) for x in arrs]
            torch.cuda.synchronize()
        E.to_host([down_dev])
        floor.append(1e3 * (time.perf_counter() - t0))
    out["drop-in latency, one image (level 2 then level 1, numpy in/out)"] = {
        "ms_median": float(np.median(lat)), "ms_min": float(np.min(lat)), "device_ms_both_levels_one_launch": dev_ms,
        "host_overhead_ms": float(np.median(lat)) - dev_ms,
        "copies_alone_ms": float(np.median(floor)),
        "copies_note": "%d B of pageable numpy inputs up in 8 copies, %d B of results down through pinned staging" % (
            sum(x.nbytes for arrs in (d2, d1) for x in arrs), down_bytes)}

    # the decoder's side of the same image (pln.py:744,783): bitstream + group starts + outliers in, latents out

    # the caller owns the group-start lists (pln.py:735-741, 767-770 builds them from the decoded group sizes); the decoder
    # appends N to the list it is given (CIS:294), so every call gets its own copy
    starts2, starts1 = r2[2][:-1].tolist(), r1[2][:-1].tolist()

    def decode_image():
        a = cwq.decode_grouped_importance_sample(None, r2[1], list(starts2), Normal(d2[2], d2[3]), 20, SEED, r2[3][0], r2[3][1])
        b = cwq.decode_grouped_importance_sample(None, r1[1], list(starts1), Normal(d1[2], d1[3]), 20, SEED, r1[3][0], r1[3][1])
        return a, b
    for _ in range(3):
        dec = decode_image()
    assert np.array_equal(dec[1].view(np.uint32), r1[0].view(np.uint32)) and np.array_equal(dec[0].view(np.uint32), r2[0].view(np.uint32))
    dl = []
    for _ in range(8):
        t0 = time.perf_counter()
        decode_image()
        dl.append(1e3 * (time.perf_counter() - t0))
    out["drop-in decode latency, one image (both levels, bitstream in, numpy out)"] = {
        "ms_median": float(np.median(dl)), "ms_min": float(np.min(dl)), "round_trip_exact": True,
        "reference_published_decode_s_per_kodak_image": "41.9-49.0 (Compression.ipynb:306-309, hardware unstated)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
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
    e2e, result_digest, digest_matches_single = None, None, None
    if not args.no_e2e:
        off = np.cumsum([0] + [a[0].size for a, _ in segs])
        targets = [Normal(host[0][off[i]:off[i + 1]], host[1][off[i]:off[i + 1]]) for i in range(len(segs))]
        proposals = [Normal(host[2][off[i]:off[i + 1]], host[3][off[i]:off[i + 1]]) for i in range(len(segs))]
        mgs = [mg for _, mg in segs]
        shard = D.make_is_shard() if world > 1 else None

        def step_e2e(sh):
            return cwq.code_grouped_importance_sample_batch(targets, proposals, SEED, N_BITS_PER_GROUP, mgs,
                                                            DIM_KL_BIT_LIMIT, shard=sh)
        for _ in range(2):
            res = step_e2e(shard)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = step_e2e(shard)
        barrier()
        dt = (time.perf_counter() - t0) / args.steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        d2h = sum(r[0].nbytes + len(r[1]) // 8 + r[2].nbytes + r[3][0].nbytes + r[3][1].nbytes for r in res if r is not None)
        d2h_t = torch.tensor([float(d2h)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(d2h_t, op=dist.ReduceOp.SUM)
        e2e = {"value": n_groups / dt, "unit": UNIT, "h2d_bytes_per_step": 16 * n_total, "d2h_bytes_per_step": int(d2h_t.item()),
               "ms_per_step": 1e3 * dt,
               "api": "code_grouped_importance_sample_batch ({} reference calls: {} images x 2 levels), pinned host tensors in, "
                      "numpy latents / '0'-'1' bitstreams / group starts / outliers out".format(len(segs), args.images),
               "bytes_note": "whole job, all ranks: every input byte is uploaded by exactly one rank (1/N slices, all-gathered "
                             "over NVLink) and every call's results are downloaded by exactly one rank" if world > 1 else
                             "one GPU: all inputs up, all results down"}
        # ---- correctness record (outside every timed region) ----
        mine = call_digests(res)
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            merged = [next((g[i] for g in gathered if g[i] is not None), None) for i in range(len(mine))]
        else:
            merged = mine
        if rank == 0:
            result_digest = hashlib.sha256("".join(d or "-" for d in merged).encode()).hexdigest()
            if world > 1:
                single = call_digests(step_e2e(None))       # the same batch, unsharded, on this GPU alone
                digest_matches_single = bool(all(m is not None for m in merged) and merged == single)
        barrier()

    secondary = None
    if world == 1 and not args.no_secondary:
        secondary = secondary_measurements(dev, n1, n2)

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
    # facts of one `ncu --set full` capture of this kernel on the full 24-image launch with this code (tools/ncu_is_score.py)
    ncu = None
    if os.path.exists(NCU_FACTS):
        try:
            ncu = json.load(open(NCU_FACTS))
        except Exception:
            ncu = None
    instr_per_cd = (ncu or {}).get("executed_thread_instr_per_candidate_dim")
    same_launch = bool(ncu) and world == 1 and int(ncu.get("images", -1)) == args.images
    roofline = {
        "bound": "on-chip issue, XU (MUFU) and FMA-heavy (IMAD.WIDE) pipes co-limited; not HBM, not tensor",
        "kernel": "k_is_score", "achieved": achieved / 1e12, "peak": issue_peak / 1e12,
        "unit": "Tinstr/s", "frac": achieved / issue_peak,
        # the same rate counted in the thread-instructions the kernel actually EXECUTES (ncu smsp__inst_executed x 32 /
        # candidate-dims): this is the issue-slot utilisation; `frac` uses SURVEY.md 8d's algorithmic 17.5 as the unit
        "frac_executed": (cd_per_s_kernel * instr_per_cd / issue_peak) if instr_per_cd else None,
        "executed_instr_per_candidate_dim": instr_per_cd,
        "sfu_frac": cd_per_s_kernel * SFU_PER_CD / sfu_peak,
        "frac_at_observed_clock": (achieved / (sms * 128 * f_obs)) if f_obs else None,
        # ideal-independent-chains microbenchmark of this kernel's instruction mix: 23.6 cycles per warp-candidate-dim
        # per sub-partition (tools/microbench/overlap.cu "v2 mix", profiles/r01_pipe_overlap_v2_mix.log)
        "frac_of_measured_mix_ceiling": cd_per_s_kernel / (sms * 4 * 32 / 23.6 * f_max),
        "mix_ceiling_note": "denominator = what tools/microbench/overlap reached issuing the ROUND-1 instruction mix of this loop as "
                            "ideal independent chains (23.6 cycles per warp-candidate-dim per sub-partition, 1.58e12 cd/s); the "
                            "loop has shed instructions since (folded scale factor, skipped half-quads), so ~1.0 means 'as fast as "
                            "that older, heavier mix could ever run', not 'nothing left'",
        "candidate_dims_per_sec_kernel": cd_per_s_kernel, "kernel_ms": score_ms_max,
        "kernel_share_of_step": score_ms_max / ms_per_step if ms_per_step > 0 else None,
        # dram__bytes_read.sum + dram__bytes_write.sum of the ncu capture, per launch; null when the capture is not of this
        # very launch shape (other image count / sharded range)
        "traffic": (ncu.get("dram_bytes_read", 0) + ncu.get("dram_bytes_write", 0)) if same_launch else None,
        "traffic_source": (os.path.relpath(NCU_FACTS, ROOT) if ncu else None),
        "algorithmic_bytes": 8.0 * n_total / world + 16.0 * n_groups / world,
        "peak_source": "{} SMs x 128 FP32 lanes x {:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json); the path is on-chip "
                       "issue/SFU bound, not HBM or tensor bound (SURVEY.md 8d); pipe widths: profiles/r01_pipes_microbench*.log"
                       .format(sms, f_max / 1e6),
        "hbm_note": "HBM traffic of the kernel is ~8 B/dim + 16 B/group against >=1e3 candidate evaluations per dim",
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:   # the CPU legs are reported at N = 1 only
        threads = os.cpu_count() or 1
        c = cpu_reference_sample(segs, threads, target_seconds=12.0)
        cpu = {"value": c["groups_per_s"], "unit": UNIT, "cores": threads, "kind": "port", "sample": c["sample"],
               "candidates_per_sec": c["candidates_per_s"], "candidate_dims_per_sec": c["candidate_dims_per_s"],
               "tensorflow": tensorflow_status(),
               "numpy_reference_shaped": cpu_numpy_reference_shaped(segs, target_seconds=5.0),
               "greedy": cpu_greedy_sample(threads, target_seconds=8.0),
               "note": "CPU restatement of the reference (TensorFlow/TFP not installed, the reference itself cannot run); every "
                       "GPU/CPU ratio has to be read with `cores`"}

    s = ms_per_step * 1e-3
    line = {
        "metric": METRIC, "value": n_groups / s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "{}x(768x512) PLN latents (N1={}, N2={} per image), importance sampler, n_bits_per_group=20, "
                               "max_group_size_bits=4/2, dim_kl_bit_limit=16, seed=42 (BASELINE configs[3])".format(args.images, n1, n2),
                   "dims": n_total, "groups": n_groups, "candidates": candidates, "candidate_dims": cand_dims,
                   "parallelism": "groups sharded in {} contiguous work-balanced ranges + all-gather of indices".format(world),
                   "l2": "compute-bound path: inputs are 76 MB (< L2) but are read once per step by the O(N) stages only; "
                         "every step rewrites >100 MB of scratch (kl/coef/gid/rows/starts) between reads"},
        "candidates_per_sec": candidates / s, "candidate_dims_per_sec": cand_dims / s,
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "result_digest": result_digest, "digest_matches_single": digest_matches_single, "secondary": secondary,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
