"""Host-side phase timing of the public batch API (where does e2e - device time go?)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import compression_without_quantization_b200 as cwq
from compression_without_quantization_b200 import coded_importance_sampler as CIS, engine as E
from compression_without_quantization_b200.synthetic import Normal

segs, n1, n2 = bench.workload(24)
host = [torch.from_numpy(np.concatenate([a[j] for a, _ in segs])).pin_memory() for j in range(4)]
off = np.cumsum([0] + [a[0].size for a, _ in segs])
targets = [Normal(host[0][off[i]:off[i + 1]], host[1][off[i]:off[i + 1]]) for i in range(len(segs))]
proposals = [Normal(host[2][off[i]:off[i + 1]], host[3][off[i]:off[i + 1]]) for i in range(len(segs))]
mgs = [mg for _, mg in segs]

def T():
    torch.cuda.synchronize(); return time.perf_counter()

for rep in range(3):
    t0 = T()
    dev = E.cuda_device()
    q_loc = [E.as_device_f32(t.loc, dev) for t in targets]; q_scale = [E.as_device_f32(t.scale, dev) for t in targets]
    p_loc = [E.as_device_f32(p.loc, dev) for p in proposals]; p_scale = [E.as_device_f32(p.scale, dev) for p in proposals]
    t1 = T()
    cat = [torch.cat(x) for x in (q_loc, q_scale, p_loc, p_scale)]
    t2 = T()
    specs = [E.SegmentSpec(int(q_loc[i].numel()), 42, 20, mgs[i], 16, True) for i in range(len(segs))]
    job = E.ISJob(cat[0], cat[1], cat[2], cat[3], specs)
    t3 = T()
    job.prepare(); t4 = T()
    job.encode(); t5 = T()
    job.decode().pack(); t6 = T()
    hostv = CIS._host_views(job, True); t7 = T()
    res = [CIS._segment_result(job, s, hostv, False) for s in range(job.table.count)]; t8 = T()
    print("h2d %.2f cat %.2f table %.2f prepare %.2f encode %.2f decode+pack %.2f d2h %.2f per-seg %.2f total %.2f ms" % tuple(
        1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t8 - t7, t8 - t0)))
