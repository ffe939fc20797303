"""One sharded importance-sampler step with a world of one rank (no NCCL): the kernels a rank of a multi-GPU run
launches around the scoring kernel, for an ncu launch list (ncu cannot wrap a multi-rank command)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from compression_without_quantization_b200 import _lib as L, distributed as D, engine as E

n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
segs, n1, n2 = bench.workload(n_images)
specs = [E.SegmentSpec(a[0].size, 42, 20, mg, 16, True) for a, mg in segs]
dev = E.cuda_device()
devt = [torch.from_numpy(np.concatenate([a[j] for a, _ in segs])).to(dev) for j in range(4)]
shard = D.make_is_shard()
for rep in range(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    job = E.ISJob(*devt, specs).launch(L.STAGE_GROUP)
    shard(job)
    job.launch(L.STAGE_FINISH)
    e1.record()
    torch.cuda.synchronize()
    print("sharded step (world 1): %.3f ms" % e0.elapsed_time(e1), flush=True)
