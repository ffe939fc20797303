"""The candidate stream with the reference's name (reference code/misc.py:3-17).

`stateless_normal_sample(loc, scale, num_samples, seed)` returns `[num_samples, *loc.shape]` draws
`scale * z + loc`, z ~ N(0,1) a pure function of (seed, row, column).  The reference obtains z from
`tf.random.stateless_normal(seed=[seed, 42])`; the engine replaces that third-party stream by its own
counter-based Philox4x32-10 + Box-Muller keyed (seed, 42) / (row, column // 4), generated in registers
inside the coders and materialised only here (debug view / oracle injection seam).
"""
import numpy as np

from . import engine as E


def stateless_normal_sample(loc, scale, num_samples, seed):
    dev = E.cuda_device()
    shape = tuple(np.shape(loc)) if not hasattr(loc, "shape") else tuple(loc.shape)
    l, s = E.as_device_f32(loc, dev), E.as_device_f32(scale, dev)
    z = E.draw_candidates(int(seed) & 0xFFFFFFFF, 0, int(num_samples), l.numel(), dev)
    out = s[None] * z + l[None]
    return out.reshape((int(num_samples),) + shape)


def candidate_rows(stream_seed, k0, nk, G):
    """float32[nk, G] numpy: rows k0.. of the RAW draws y of stream `stream_seed` -- the `normal_source` the oracle takes
    (the injection seam of the parity harness).  The N(0,1) draw is float32(sqrt(2 ln 2)) * y; the candidate loops score y
    itself with that factor folded into their coefficients."""
    return E.draw_candidates(stream_seed, k0, nk, G, raw=True).cpu().numpy()
