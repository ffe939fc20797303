#!/usr/bin/env python
"""Facts about the dominant kernel (k_is_score) on the bench's own launch -- all 24 images, final code -- from one ncu
capture: executed instructions, DRAM traffic, duration, pipe utilisation.  Writes gpurun_out/r02_is_score_ncu.json (copied
to profiles/ and read by bench.py for roofline.traffic / roofline.frac_executed) and keeps the raw CSV beside it.

    python tools/ncu_is_score.py            # on a GPU box: runs ncu on itself (--child does one step of the workload)
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
METRICS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "dram__bytes_read.sum",
           "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum",
           "sm__inst_executed_pipe_alu.sum", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "lts__t_bytes.sum"]


def child(images):
    import numpy as np
    import torch
    import bench
    from compression_without_quantization_b200 import engine as E
    segs, n1, n2 = bench.workload(images)
    specs = [E.SegmentSpec(a[0].size, bench.SEED, bench.N_BITS_PER_GROUP, mg, bench.DIM_KL_BIT_LIMIT, True) for a, mg in segs]
    dev = E.cuda_device()
    t = [torch.from_numpy(np.concatenate([a[j] for a, _ in segs])).to(dev) for j in range(4)]
    for _ in range(2):
        job = E.ISJob(t[0], t[1], t[2], t[3], specs).launch(7)
    torch.cuda.synchronize()
    job.sync_groups()
    sizes = (job.groups.starts[1:] - job.groups.starts[:-1]).double()
    cd = float((job.K.double() * sizes).sum())
    json.dump({"images": images, "candidate_dims": cd, "groups": job.groups.n_groups, "dims": int(t[0].numel())},
              open(os.path.join(OUT, "r02_is_score_child.json"), "w"))


def main():
    images = 24
    if "--child" in sys.argv:
        return child(images)
    os.makedirs(OUT, exist_ok=True)
    raw = os.path.join(OUT, "r02_is_score_ncu_raw.csv")
    cmd = ["ncu", "--metrics", ",".join(METRICS), "--clock-control", "none", "-k", "regex:k_is_score", "-s", "1", "-c", "1", "--csv",
           "--log-file", raw, sys.executable, os.path.abspath(__file__), "--child"]
    if "--parse" not in sys.argv:     # --parse: only rebuild the JSON from an existing raw CSV
        subprocess.check_call(cmd, cwd=ROOT)
    rows = list(csv.DictReader(l for l in open(raw) if not l.startswith("==")))
    m = {}
    for r in rows:
        try:
            m[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            pass                      # 'n/a': the metric does not exist on this chip
    units = {r["Metric Name"]: r["Metric Unit"] for r in rows}
    c = json.load(open(os.path.join(OUT, "r02_is_score_child.json")))

    def to_bytes(name):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units.get(name, "byte"), 1.0)
        return m.get(name, 0.0) * scale
    dur = m["gpu__time_duration.sum"] * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(units["gpu__time_duration.sum"], 1e-9)
    facts = {
        "what": "ncu capture of k_is_score on the bench launch ({} images, both levels, one GPU), one launch after a warm-up step; "
                "--clock-control none; durations under ncu are cold-cache, use bench.py's event timing for rates".format(c["images"]),
        "images": c["images"], "dims": c["dims"], "groups": c["groups"], "candidate_dims": c["candidate_dims"],
        "duration_ms_under_ncu": dur * 1e3,
        "warp_instructions_executed": m["smsp__inst_executed.sum"],
        "thread_instructions_executed": m["smsp__thread_inst_executed.sum"],
        "executed_thread_instr_per_candidate_dim": m["smsp__inst_executed.sum"] * 32.0 / c["candidate_dims"],
        "active_thread_instr_per_candidate_dim": m["smsp__thread_inst_executed.sum"] / c["candidate_dims"],
        "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
        "l2_bytes": to_bytes("lts__t_bytes.sum"),
        "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "pipe_xu_active_pct": m.get("sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active"),
        "pipe_fmaheavy_active_pct": m.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": m.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "inst_pipe_xu": m.get("sm__inst_executed_pipe_xu.sum"), "inst_pipe_fma": m.get("sm__inst_executed_pipe_fma.sum"),
        "inst_pipe_fmaheavy": m.get("sm__inst_executed_pipe_fmaheavy.sum"), "inst_pipe_alu": m.get("sm__inst_executed_pipe_alu.sum"),
        "registers_per_thread": m.get("launch__registers_per_thread"),
    }
    json.dump(facts, open(os.path.join(OUT, "r02_is_score_ncu.json"), "w"), indent=1)
    print(json.dumps(facts))


if __name__ == "__main__":
    main()
