"""Host timeline of the chunked batch pipeline (coded_importance_sampler._encode_batch_pipelined) on the bench workload:
where the host is when each chunk is queued / collected, to see what is left exposed around the scoring kernels.

    python tools/profile_pipeline.py [n_chunks]                                          # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/profile_pipeline.py [n_chunks]   # sharded, rank 0 prints
"""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from compression_without_quantization_b200 import coded_importance_sampler as CIS, distributed as D, engine as E
from compression_without_quantization_b200.synthetic import Normal

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
segs, n1, n2 = bench.workload(24)
host = [torch.from_numpy(np.concatenate([a[j] for a, _ in segs])).pin_memory() for j in range(4)]
off = np.cumsum([0] + [a[0].size for a, _ in segs])
targets = [Normal(host[0][off[i]:off[i + 1]], host[1][off[i]:off[i + 1]]) for i in range(len(segs))]
proposals = [Normal(host[2][off[i]:off[i + 1]], host[3][off[i]:off[i + 1]]) for i in range(len(segs))]
mgs = [mg for _, mg in segs]
shard = D.make_is_shard(results=os.environ.get("CWQ_RESULTS", "own"), slice_uploads=os.environ.get("CWQ_SLICE", "1") == "1") if world > 1 else None
specs = CIS._batch_specs(targets, proposals, 42, 20, mgs, 16)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for arg in (sys.argv[1:] or ["4"]):
    # "F/4" = full uploads on every rank, "A/4" = every rank returns all results, "4" / "1:6:1" = the sharded defaults
    flags, layout = arg.split("/") if "/" in arg else ("", arg)
    if world > 1:
        shard = D.make_is_shard(results="all" if "A" in flags else "own", slice_uploads="F" not in flags)
    weights = [float(x) for x in layout.split(":")] if ":" in layout else None
    if weights is None:
        chunks = CIS._plan_chunks([sp.n for sp in specs], int(layout))
    else:   # explicit weights, e.g. 1:6:1
        sizes, total = [sp.n for sp in specs], float(sum(sp.n for sp in specs))
        cuts, acc, start, chunks = np.cumsum(weights)[:-1] / sum(weights) * total, 0, 0, []
        for i, nn in enumerate(sizes):
            acc += nn
            if len(chunks) < len(cuts) and acc >= cuts[len(chunks)] and i + 1 < len(sizes):
                chunks.append((start, i + 1)); start = i + 1
        chunks.append((start, len(sizes)))
    times = []
    for rep in range(8):
        barrier()
        trace = []
        t0 = time.perf_counter()
        CIS._encode_batch_pipelined(targets, proposals, specs, chunks, False, shard, trace)
        barrier()
        times.append(time.perf_counter() - t0)
    if world > 1:      # the slowest rank's own timeline matters: gather every rank's last mark
        lasts = [None] * world
        dist.all_gather_object(lasts, [(n_, round(1e3 * t_, 2)) for n_, t_ in trace])
    if rank == 0:
        if world > 1:
            print("   per-rank (first mark .. last mark): " + " | ".join("%.2f..%.2f" % (l[0][1], l[-1][1]) for l in lasts))
        print("layout %s chunks %s world %d: e2e median %.2f ms (min %.2f) over %d reps" % (
            arg, chunks, world, 1e3 * float(np.median(times[2:])), 1e3 * min(times[2:]), len(times) - 2), flush=True)
        for name, t in trace:
            print("   %-24s %8.2f ms" % (name, 1e3 * t))
if world > 1:
    dist.destroy_process_group()
