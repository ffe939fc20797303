"""Sharded batch API, NCCL all-gather vs peer-memory exchange of the indices: per-step e2e times and the number of
cudaMalloc'd segments torch's allocator created during the timed steps (torchrun --nproc-per-node N ...)."""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import compression_without_quantization_b200 as cwq
from compression_without_quantization_b200 import distributed as D
from compression_without_quantization_b200.synthetic import Normal

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
segs, n1, n2 = bench.workload(24)
host = [torch.from_numpy(np.concatenate([a[j] for a, _ in segs])).pin_memory() for j in range(4)]
off = np.cumsum([0] + [a[0].size for a, _ in segs])
targets = [Normal(host[0][off[i]:off[i + 1]], host[1][off[i]:off[i + 1]]) for i in range(len(segs))]
proposals = [Normal(host[2][off[i]:off[i + 1]], host[3][off[i]:off[i + 1]]) for i in range(len(segs))]
mgs = [mg for _, mg in segs]
for mode in ("nccl", "p2p", "nccl", "p2p"):
    shard = D.make_is_shard(p2p=mode == "p2p")
    times = []
    for rep in range(12):
        if rep == 2:
            seg0 = torch.cuda.memory_stats()["segment.all.allocated"]
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        cwq.code_grouped_importance_sample_batch(targets, proposals, 42, 20, mgs, 16, shard=shard)
        dist.barrier(); torch.cuda.synchronize()
        times.append(1e3 * (time.perf_counter() - t0))
    segs_made = torch.cuda.memory_stats()["segment.all.allocated"] - seg0
    t = torch.tensor(times[2:], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(mode, "per-step ms (max over ranks):", [round(x, 1) for x in t.tolist()], " cudaMalloc'd segments in 10 steps (rank 0):", segs_made, flush=True)
dist.destroy_process_group()
