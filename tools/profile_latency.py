"""Latency of the drop-in call sequence of pln.py for one 768x512 image (level 2, then level 1; numpy in, numpy / bytes out) and
a cProfile of the level-2 call, whose cost is host overhead only."""
import cProfile, pstats, io, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import compression_without_quantization_b200 as cwq
from compression_without_quantization_b200.synthetic import Normal, synthetic_posterior, pln_level_sizes
n1, n2 = pln_level_sizes(512, 768)
d2, d1 = synthetic_posterior(n2, 1000), synthetic_posterior(n1, 0)
def one_image():
    r2 = cwq.code_grouped_importance_sample(None, Normal(d2[0], d2[1]), Normal(d2[2], d2[3]), 42, 20, 2, 16)
    r1 = cwq.code_grouped_importance_sample(None, Normal(d1[0], d1[1]), Normal(d1[2], d1[3]), 42, 20, 4, 16)
    return r2, r1
def level2():
    return cwq.code_grouped_importance_sample(None, Normal(d2[0], d2[1]), Normal(d2[2], d2[3]), 42, 20, 2, 16)
for _ in range(5): one_image()
torch.cuda.synchronize()
for name, fn in (("both levels", one_image), ("level 2 only", level2)):
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); fn(); ts.append(1e3 * (time.perf_counter() - t0))
    print(name, "median ms", np.median(ts), "min", min(ts))
pr = cProfile.Profile()
pr.enable()
for _ in range(50): level2()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:5000])
