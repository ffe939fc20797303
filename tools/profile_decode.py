"""Decode-side timing of the batch API (24 images x 2 levels): parse, group tables, regenerate, outliers."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import compression_without_quantization_b200 as cwq
from compression_without_quantization_b200.synthetic import Normal

segs, n1, n2 = bench.workload(24)
t = [Normal(a[0], a[1]) for a, _ in segs]
p = [Normal(a[2], a[3]) for a, _ in segs]
mgs = [mg for _, mg in segs]
res = cwq.code_grouped_importance_sample_batch(t, p, 42, 20, mgs, 16)
codes = [r[1] for r in res]
starts = [list(r[2]) for r in res]
oi, oq = [r[3][0] for r in res], [r[3][1] for r in res]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dec = cwq.decode_grouped_importance_sample_batch(codes, starts, p, 42, oi, oq)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ng = sum(len(s) - 1 for s in starts)
    print("decode batch: %.2f ms  (%d groups, %.3g groups/s, %d bits)" % (1e3 * dt, ng, ng / dt, sum(len(c) for c in codes)))
ok = all(np.array_equal(np.delete(d, o), np.delete(r[0], o)) for d, r, o in zip(dec, res, oi))
print("round trip ok:", ok)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
cwq.decode_grouped_importance_sample_batch(codes, starts, p, 42, oi, oq); torch.cuda.synchronize()
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
