#!/usr/bin/env python
"""A/B of builds of libcwq.so with the same ABI on ONE box: for every library given on the command line a fresh process
(CWQ_LIB=<path>) times the scoring kernel and the whole step of the headline batch, one image and the 1 M-dim sweep.

    python tools/ab_variants.py build_variants/a.so build_variants/b.so [--rounds 2]

ptxas schedules the 231-instruction candidate loop differently whenever the code around it changes, and that alone moves
the kernel by about +-1 %; comparing variants on different boxes or in different processes' warm-up states cannot resolve
that, interleaved rounds on one box can."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    import ctypes
    import numpy as np
    import torch
    import bench
    from compression_without_quantization_b200 import _lib, engine as E
    from compression_without_quantization_b200.synthetic import pln_level_sizes, synthetic_posterior
    lib = _lib.load()
    dev = E.cuda_device()
    n1, n2 = pln_level_sizes(512, 768)

    def case(arrays, specs, steps):
        t = [torch.from_numpy(np.concatenate(a)).to(dev) for a in zip(*arrays)]
        for _ in range(3):
            E.ISJob(t[0], t[1], t[2], t[3], specs).launch(7)
        torch.cuda.synchronize()
        lib.cwq_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            E.ISJob(t[0], t[1], t[2], t[3], specs).launch(7)
        e1.record()
        torch.cuda.synchronize()
        score = []
        for i in range(lib.cwq_profile_count()):
            v = ctypes.c_float(0)
            lib.cwq_profile_get_ms(i, ctypes.cast(ctypes.pointer(v), ctypes.c_void_p))
            score.append(v.value)
        lib.cwq_profile_enable(0)
        return round(e0.elapsed_time(e1) / steps, 4), round(float(np.median(score)), 4)

    out = {}
    segs, _, _ = bench.workload(24)
    specs = [E.SegmentSpec(a[0].size, 42, 20, mg, 16, True) for a, mg in segs]
    out["batch24"] = case([a for a, _ in segs], specs, 6)
    out["image"] = case([a for a, _ in segs[:2]], specs[:2], 20)
    a = synthetic_posterior(1 << 20, 77)
    for bits in (8, 12, 16):
        out["1M_%d" % bits] = case([a], [E.SegmentSpec(1 << 20, 42, bits, 4, bits - 2, True)], 20)
    print(json.dumps(out), flush=True)


def main():
    if "--child" in sys.argv:
        return child()
    libs = [a for a in sys.argv[1:] if not a.startswith("--") and a.endswith(".so")]
    rounds = int(sys.argv[sys.argv.index("--rounds") + 1]) if "--rounds" in sys.argv else 2
    for r in range(rounds):
        for lib in libs:
            env = dict(os.environ, CWQ_LIB=os.path.abspath(lib))
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True)
            line = [l for l in res.stdout.splitlines() if l.startswith("{")]
            print(r, os.path.basename(lib), line[-1] if line else "FAILED " + res.stderr[-300:], flush=True)


if __name__ == "__main__":
    main()
