"""Device time of one importance-sampler step: stage-by-stage methods vs the fused cwq_is_encode_pipeline call."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from compression_without_quantization_b200 import _lib as L, engine as E

n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 24
segs, n1, n2 = bench.workload(n_images)
specs = [E.SegmentSpec(a[0].size, 42, 20, mg, 16, True) for a, mg in segs]
dev = E.cuda_device()
devt = [torch.from_numpy(np.concatenate([a[j] for a, _ in segs])).to(dev) for j in range(4)]


def staged():
    return E.ISJob(*devt, specs).prepare(sync=False).encode().decode().pack()


def fused():
    return E.ISJob(*devt, specs).launch(7)


def fused_split():
    j = E.ISJob(*devt, specs).launch(1)
    return j.launch(2).launch(4)


for name, fn in (("staged", staged), ("fused", fused), ("fused_split", fused_split), ("staged", staged), ("fused", fused)):
    for _ in range(3):
        job = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(5):
        job = fn()
    e1.record()
    th = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("%-12s device %.3f ms/step   host enqueue %.3f ms/step" % (name, e0.elapsed_time(e1) / 5, 1e3 * th / 5), flush=True)
