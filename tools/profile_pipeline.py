"""Timeline of the chunked batch pipeline (coded_importance_sampler._encode_batch_pipelined): host time per stage and
GPU completion time per chunk, to see what is left exposed around the scoring kernel."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from compression_without_quantization_b200 import coded_importance_sampler as CIS, engine as E
from compression_without_quantization_b200.synthetic import Normal

segs, n1, n2 = bench.workload(24)
host = [torch.from_numpy(np.concatenate([a[j] for a, _ in segs])).pin_memory() for j in range(4)]
off = np.cumsum([0] + [a[0].size for a, _ in segs])
targets = [Normal(host[0][off[i]:off[i + 1]], host[1][off[i]:off[i + 1]]) for i in range(len(segs))]
proposals = [Normal(host[2][off[i]:off[i + 1]], host[3][off[i]:off[i + 1]]) for i in range(len(segs))]
mgs = [mg for _, mg in segs]
n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else CIS.PIPELINE_CHUNKS

for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    specs = CIS._batch_specs(targets, proposals, 42, 20, mgs, 16)
    chunks = CIS._plan_chunks([sp.n for sp in specs], n_chunks)
    dev = E.cuda_device()
    compute = torch.cuda.current_stream()
    up, down = CIS._copy_streams(dev)
    up.wait_stream(compute)
    marks = [("specs", time.perf_counter() - t0)]
    start_ev = torch.cuda.Event(enable_timing=True); start_ev.record(compute)
    staged = []
    with torch.cuda.stream(up):
        for a, b in chunks:
            tensors = CIS._upload(targets[a:b], proposals[a:b], dev)
            ev = torch.cuda.Event(enable_timing=True); ev.record(up)
            staged.append((tensors, ev))
    marks.append(("uploads enqueued", time.perf_counter() - t0))
    jobs = []
    for (a, b), (tensors, ev) in zip(chunks, staged):
        compute.wait_event(ev)
        for t in tensors:
            t.record_stream(compute)
        job = E.ISJob(*tensors, specs[a:b]).prepare(sync=False)
        job.encode().decode().pack()
        done = torch.cuda.Event(enable_timing=True); done.record(compute)
        jobs.append((job, done))
        marks.append(("chunk %d enqueued" % len(jobs), time.perf_counter() - t0))
    for i, (job, done) in enumerate(jobs):
        down.wait_event(done)
        with torch.cuda.stream(down):
            job.sync_groups()
            marks.append(("chunk %d gpu done (host saw)" % (i + 1), time.perf_counter() - t0))
            hv = CIS._host_views(job, True)
            marks.append(("chunk %d downloaded" % (i + 1), time.perf_counter() - t0))
        res = [CIS._segment_result(job, s, hv, False) for s in range(job.table.count)]
        marks.append(("chunk %d formatted" % (i + 1), time.perf_counter() - t0))
    compute.wait_stream(down)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    if rep >= 2:
        print("rep", rep, "chunks", chunks, "total %.2f ms" % (1e3 * total))
        for name, t in marks:
            print("   %-32s %8.2f ms" % (name, 1e3 * t))
        print("   gpu: uploads done at", ["%.2f" % start_ev.elapsed_time(ev) for _, ev in staged],
              " chunks done at", ["%.2f" % start_ev.elapsed_time(d) for _, d in jobs])
