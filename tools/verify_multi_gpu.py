"""torchrun --nproc-per-node N tools/verify_multi_gpu.py : the sharded encode (contiguous work-balanced group
ranges + all-gather of indices) returns, on every rank, exactly what the single-GPU encode returns."""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import compression_without_quantization_b200 as cwq  # noqa: E402
from compression_without_quantization_b200 import distributed as D  # noqa: E402
from compression_without_quantization_b200.synthetic import Normal, synthetic_posterior  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

sizes = [196608, 2304, 50000, 1]
data = [synthetic_posterior(n, 300 + i, i % 2 == 1) for i, n in enumerate(sizes)]
t = [Normal(d[0], d[1]) for d in data]
p = [Normal(d[2], d[3]) for d in data]
mg = [4, 2, 3, 4]


def digest(res):
    h = hashlib.sha256()
    for s, code, starts, (oi, oq) in res:
        for a in (s, starts, oi, oq):
            h.update(np.ascontiguousarray(a).tobytes())
        h.update(code)
    return h.hexdigest()


def own_matches(res, ref):
    """results="own": every call is returned by exactly one rank, and what it returns is the single-GPU result"""
    mine = [i for i, r in enumerate(res) if r is not None]
    ok = all(digest([res[i]]) == digest([ref[i]]) for i in mine)
    owners = torch.zeros(len(res), dtype=torch.int32, device="cuda")
    owners[mine] = 1
    dist.all_reduce(owners)
    return ok and bool((owners == 1).all().item())


sharded = cwq.code_grouped_importance_sample_batch(t, p, 42, 16, mg, 12, shard=D.make_is_shard(results="all"))
single = cwq.code_grouped_importance_sample_batch(t, p, 42, 16, mg, 12)
ok_is = digest(sharded) == digest(single)
# the chunked pipeline (coded_importance_sampler._encode_batch_pipelined), sharded, at 2 and 3 chunks
from compression_without_quantization_b200 import coded_importance_sampler as CIS  # noqa: E402
CIS.PIPELINE_MIN_DIMS = 0
for n_chunks in (2, 3):
    CIS.PIPELINE_CHUNKS = n_chunks
    ok_is = ok_is and digest(cwq.code_grouped_importance_sample_batch(t, p, 42, 16, mg, 12, shard=D.make_is_shard(results="all"))) == digest(single)
    # every rank uploads 1/world of the inputs and returns its own calls (the default of the sharded batch API)
    ok_is = ok_is and own_matches(cwq.code_grouped_importance_sample_batch(t, p, 42, 16, mg, 12, shard=D.make_is_shard()), single)
    ok_is = ok_is and own_matches(cwq.code_grouped_importance_sample_batch(t, p, 42, 16, mg, 12,
                                                                            shard=D.make_is_shard(slice_uploads=False)), single)
CIS.PIPELINE_MIN_DIMS = 1 << 62
gs = cwq.code_grouped_greedy_sample_batch(t[2:], p[2:], 3, 8, [7, 8], 6, shard=D.make_greedy_shard())
g1 = cwq.code_grouped_greedy_sample_batch(t[2:], p[2:], 3, 8, [7, 8], 6)
ok_gr = all(a[1] == b[1] and a[2] == b[2] and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) for a, b in zip(gs, g1))
flags = torch.tensor([int(ok_is), int(ok_gr)], device="cuda")
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print("world", world, "importance sharded == single :", bool(flags[0]),
          " greedy sharded == single:", bool(flags[1]),
          " groups:", sum(len(r[2]) - 1 for r in single), flush=True)
dist.destroy_process_group()
sys.exit(0 if flags.min().item() == 1 else 1)
