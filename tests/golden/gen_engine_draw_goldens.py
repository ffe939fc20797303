"""Goldens for BASELINE.json's parity harness -- "a reference harness that injects the identical Philox candidate draws
into the reference samplers; chosen indices and bitstreams must be bit-exact":

  step 1, on a B200 (`python tests/golden/gen_engine_draw_goldens.py --dump`): for a few small cases the ENGINE's own
          candidate draws (MUFU Box-Muller, not reproducible on a CPU) are materialised through the injection seam
          (misc.candidate_rows = cwq_draw_candidates) and saved: every table any sampler asks for, keyed by
          (stream seed, columns).  Output: gpurun_out/engine_draws.npz (copied to tests/golden/).
  step 2, in the container that holds /root/reference (`python tests/golden/gen_engine_draw_goldens.py`): the UNMODIFIED
          reference samplers run under oracle/tf_shim with tf.random.stateless_normal answering from those saved tables.
          Output: tests/golden/reference_engine_draws.npz.

tests/test_oracle_cpu.py checks oracle(injected tables) == these goldens; tests/test_gpu_parity.py checks that the engine,
drawing for itself on the GPU, returns exactly the reference's bitstreams, group tables, outliers and latents.
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

IS_CASES = [  # name, n, key, nontrivial prior, seed, n_bits_per_group, max_group_size_bits, dim_kl_bit_limit
    ("xis0", 120, 40, False, 42, 10, 3, 8), ("xis1", 90, 41, True, 1234567, 9, 2, 7), ("xis2", 33, 42, True, 7, 11, 4, 9),
]
GR_CASES = [  # name, n, key, nontrivial prior, seed, n_steps, n_bits_per_step, max_group_size_bits
    ("xgr0", 60, 50, False, 42, 3, 6, 4), ("xgr1", 45, 51, True, 99, 4, 5, 3),
]
PG_CASES = [  # name, G, key, seed, n_steps, n_bits_per_step, rho   (per-group greedy, proposal loc != 0)
    ("xpg0", 20, 60, 42, 3, 6, 1.0), ("xpg1", 7, 61, 5, 4, 5, 0.8),
]


def posterior(n, key, nontrivial):
    rng = np.random.default_rng(9500 + key)
    q_scale = np.exp(-np.abs(rng.normal(0.0, 0.7, n)))
    q_loc = 1.2 * rng.normal(size=n) * (rng.random(n) < 0.35)
    if nontrivial:
        p_loc = rng.normal(size=n)
        p_scale = np.log1p(np.exp(rng.normal(size=n)))
        q_loc = p_loc + p_scale * q_loc
        q_scale = p_scale * q_scale
    else:
        p_loc, p_scale = np.zeros(n), np.ones(n)
    f = np.float32
    q_loc, q_scale = q_loc.astype(f), q_scale.astype(f)
    if n > 50:
        q_loc[n // 3], q_scale[n // 3] = q_loc[n // 3] + 6.0 * (p_scale[n // 3] if nontrivial else 1.0), q_scale[n // 3] * 0.05  # an outlier
    return q_loc, q_scale, p_loc.astype(f), p_scale.astype(f)


def pg_inputs(G, key):
    rng = np.random.default_rng(9500 + key)
    f = np.float32
    p_loc = rng.normal(size=G).astype(f)
    p_scale = np.log1p(np.exp(rng.normal(size=G))).astype(f)
    t_loc = (p_loc + p_scale * rng.normal(size=G) * 0.8).astype(f)
    t_scale = (p_scale * np.exp(-np.abs(rng.normal(0, .7, G)))).astype(f)
    return t_loc, t_scale, p_loc, p_scale


class Tables:
    """normal_source that records (dump) or replays (golden run) whole tables keyed by (stream seed, columns)."""

    def __init__(self, backing=None, stored=None):
        self.backing, self.stored = backing, {} if stored is None else stored

    def __call__(self, seed, k0, nk, G):
        key = "{}_{}".format(int(seed) & 0xFFFFFFFF, int(G))
        have = self.stored.get(key)
        if have is None or have.shape[0] < k0 + nk:
            if self.backing is None:
                raise KeyError("no saved engine draws for stream {} rows {}..{}".format(key, k0, k0 + nk))
            rows = max(k0 + nk, 0 if have is None else have.shape[0])
            have = np.asarray(self.backing(seed, 0, rows, G), dtype=np.float32).reshape(rows, G)
            self.stored[key] = have
        return have[k0:k0 + nk]


def dump():
    from compression_without_quantization_b200.misc import candidate_rows
    from oracle import oracle as O
    O.build()
    t = Tables(backing=candidate_rows)
    for name, n, key, nontrivial, seed, bits, m, limit in IS_CASES:
        q_loc, q_scale, p_loc, p_scale = posterior(n, key, nontrivial)
        t((seed + 0x9E3779B9) & 0xFFFFFFFF, 0, n, 1)      # the reference's target.sample() draws EVERY dim (CIS:150), then gathers
        s, code, starts, (oi, oq) = O.code_grouped_importance_sample(None, O.Normal(q_loc, q_scale), O.Normal(p_loc, p_scale), seed, bits,
                                                                     m, limit, normal_source=t)
        O.decode_grouped_importance_sample(None, code, [int(x) for x in starts[:-1]], O.Normal(p_loc, p_scale), bits, seed, oi, oq,
                                           normal_source=t)
    for name, n, key, nontrivial, seed, steps, b, m in GR_CASES:
        q_loc, q_scale, p_loc, p_scale = posterior(n, key, nontrivial)
        O.code_grouped_greedy_sample(None, O.Normal(q_loc, q_scale), O.Normal(p_loc, p_scale), steps, b, seed, m, normal_source=t)
    for name, G, key, seed, steps, b, rho in PG_CASES:
        t_loc, t_scale, p_loc, p_scale = pg_inputs(G, key)
        O.code_greedy_sample(t_loc, t_scale, p_loc, p_scale, b, steps, seed, rho=rho, normal_source=t)
    out = os.path.join(ROOT, "gpurun_out", "engine_draws.npz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, **t.stored)
    print("wrote", out, len(t.stored), "tables", sum(v.nbytes for v in t.stored.values()), "bytes")


def golden():
    from oracle import oracle as O
    from oracle import tf_shim
    draws = np.load(os.path.join(HERE, "engine_draws.npz"))
    t = Tables(stored={k: draws[k] for k in draws.files})
    tf = tf_shim.install()
    tf_shim.set_normal_source(O.as_standard_normal(t))     # the saved tables are the engine's RAW draws
    sys.path.insert(0, "/root/reference/code")
    import tqdm as _tqdm
    _tqdm.tqdm = lambda it, *a, **k: it
    import coded_greedy_sampler as CGS  # reference, unmodified
    import coded_importance_sampler as CIS  # reference, unmodified
    import tensorflow_probability as tfp  # the shim
    tfd = tfp.distributions
    sess = tf.Session()
    sink = io.StringIO()
    out = {}
    for name, n, key, nontrivial, seed, bits, m, limit in IS_CASES:
        q_loc, q_scale, p_loc, p_scale = posterior(n, key, nontrivial)
        tf_shim.set_outlier_seed(seed)
        target, proposal = tfd.Normal(loc=q_loc, scale=q_scale), tfd.Normal(loc=p_loc, scale=p_scale)
        with redirect_stdout(sink):
            sample, bitcode, starts, (o_idx, o_q) = CIS.code_grouped_importance_sample(sess, target, proposal, seed, bits,
                                                                                       max_group_size_bits=m, dim_kl_bit_limit=limit)
            decoded = CIS.decode_grouped_importance_sample(sess, bitcode, [int(x) for x in starts[:-1]], proposal, bits, seed, o_idx, o_q)
        out.update({name + "_in": np.stack([q_loc, q_scale, p_loc, p_scale]), name + "_par": np.array([seed, bits, m, limit]),
                    name + "_sample": np.asarray(sample, np.float32), name + "_bitcode": np.frombuffer(bitcode, np.uint8),
                    name + "_starts": np.asarray(starts, np.int64), name + "_oidx": np.asarray(o_idx, np.int64),
                    name + "_oq": np.asarray(o_q, np.uint16), name + "_decoded": np.asarray(decoded, np.float32)})
        print(name, "groups", len(starts) - 1, "bits", len(bitcode), "outliers", len(o_idx))
    for name, n, key, nontrivial, seed, steps, b, m in GR_CASES:
        q_loc, q_scale, p_loc, p_scale = posterior(n, key, nontrivial)
        target, proposal = tfd.Normal(loc=q_loc, scale=q_scale), tfd.Normal(loc=p_loc, scale=p_scale)
        with redirect_stdout(sink):
            sample, bitcode, starts = CGS.code_grouped_greedy_sample(sess, target, proposal, steps, b, seed, max_group_size_bits=m)
            decoded = CGS.decode_grouped_greedy_sample(sess, bitcode, list(starts[:-1]), proposal, b, steps, seed)
        out.update({name + "_in": np.stack([q_loc, q_scale, p_loc, p_scale]), name + "_par": np.array([seed, steps, b, m]),
                    name + "_sample": np.asarray(sample, np.float32), name + "_bitcode": np.frombuffer(bitcode.encode(), np.uint8),
                    name + "_starts": np.asarray(starts, np.int64), name + "_decoded": np.asarray(decoded, np.float32)})
        print(name, "groups", len(starts) - 1, "bits", len(bitcode))
    for name, G, key, seed, steps, b, rho in PG_CASES:
        t_loc, t_scale, p_loc, p_scale = pg_inputs(G, key)
        with redirect_stdout(sink):
            best, code = sess.run(list(CGS.code_greedy_sample(t_loc, t_scale, p_loc, p_scale, b, steps, seed, rho=rho)))
            decoded = sess.run(CGS.decode_greedy_sample(code, p_loc, p_scale, b, steps, seed, rho=rho))
        code = code.decode("utf-8") if isinstance(code, bytes) else str(code)
        out.update({name + "_in": np.stack([t_loc, t_scale, p_loc, p_scale]), name + "_par": np.array([seed, steps, b, rho]),
                    name + "_sample": np.asarray(best, np.float32), name + "_bitcode": np.frombuffer(code.encode(), np.uint8),
                    name + "_decoded": np.asarray(decoded, np.float32)})
        print(name, "dims", G, "bits", len(code))
    np.savez_compressed(os.path.join(HERE, "reference_engine_draws.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_engine_draws.npz"))


if __name__ == "__main__":
    dump() if "--dump" in sys.argv else golden()
