import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from compression_without_quantization_b200 import _lib, engine as E
from compression_without_quantization_b200.synthetic import synthetic_posterior
dev = E.cuda_device()
for bits in (8, 12):
    a = synthetic_posterior(1 << 20, 77)
    t = [torch.from_numpy(x).to(dev) for x in a]
    specs = [E.SegmentSpec(1 << 20, 42, bits, 4, bits - 2, True)]
    for _ in range(3):
        job = E.ISJob(*t, specs).launch(7)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    hs = []
    evs[0].record()
    for i in range(20):
        h0 = time.perf_counter()
        job = E.ISJob(*t, specs).launch(7)
        evs[i + 1].record()
        hs.append(1e3 * (time.perf_counter() - h0))
    torch.cuda.synchronize()
    print(bits, "host ms", " ".join("%.2f" % h for h in hs))
    print(bits, "gpu  ms", " ".join("%.2f" % evs[i].elapsed_time(evs[i + 1]) for i in range(20)), flush=True)
