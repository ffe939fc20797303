"""Generates tests/golden/reference_samplers.npz by executing the UNMODIFIED reference sources
/root/reference/code/{coded_importance_sampler,coded_greedy_sampler,misc,binary_io}.py.

TensorFlow / TFP are not installed here, so the three modules are imported under oracle/tf_shim (a
lazy numpy stand-in for exactly the TF-1.x graph surface they touch; its stated assumptions are in its
docstring).  The one third-party op that is replaced ON PURPOSE is tf.random.stateless_normal -- the
injection seam of BASELINE.json's north_star -- which returns the spec's Philox4x32-10 + Box-Muller
stream (oracle.default_normal_source, libm floats).  Everything else -- the grouping loop and its
inequalities, seeds per group and per greedy step, K = ceil(exp(KL)), argmax, Elias-delta / LSB-first
coding and concatenation, un-rescaling, outlier handling, return types, list mutation -- is the
reference's own Python, run as shipped.

    python tests/golden/gen_reference_goldens.py        # needs /root/reference; writes the .npz next to it
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from oracle import tf_shim  # noqa: E402

tf = tf_shim.install()
tf_shim.set_normal_source(O.as_standard_normal(O.default_normal_source))   # tf.random.stateless_normal: N(0,1) = sqrt(2 ln 2) * raw
sys.path.insert(0, "/root/reference/code")
import tqdm as _tqdm  # noqa: E402

_tqdm.tqdm = lambda it, *a, **k: it  # silence progress bars (side effect only)
import coded_greedy_sampler as CGS  # noqa: E402  (reference, unmodified)
import coded_importance_sampler as CIS  # noqa: E402  (reference, unmodified)
import tensorflow_probability as tfp  # noqa: E402  (the shim)

tfd = tfp.distributions
sess = tf.Session()


def posterior(n, key, nontrivial):
    rng = np.random.default_rng(9000 + key)
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
    return q_loc.astype(f), q_scale.astype(f), p_loc.astype(f), p_scale.astype(f)


out = {}
IS_CASES = [  # name, n, key, nontrivial, seed, n_bits_per_group, max_group_size_bits, dim_kl_bit_limit
    ("is0", 400, 0, False, 42, 12, 4, 10), ("is1", 300, 1, True, 7, 10, 2, 8), ("is2", 257, 2, True, 123456, 14, 3, 12),
    ("is3", 1, 3, False, 42, 20, 4, 16), ("is4", 64, 4, False, 5, 8, 5, 7),
]
sink = io.StringIO()
for name, n, key, nontrivial, seed, bits, m, limit in IS_CASES:
    q_loc, q_scale, p_loc, p_scale = posterior(n, key, nontrivial)
    if name == "is4":
        q_loc[0], q_scale[0] = 0.0, 0.0055        # ~6.8 bits: over the 8-bit group budget (6.56 bits), under the dim limit
        q_loc[10], q_scale[10] = 5.0, 0.05        # outliers (> 7 bits)
        q_loc[33], q_scale[33] = -4.0, 0.01
    tf_shim.set_outlier_seed(seed)
    target, proposal = tfd.Normal(loc=q_loc, scale=q_scale), tfd.Normal(loc=p_loc, scale=p_scale)
    with redirect_stdout(sink):
        sample, bitcode, starts, (o_idx, o_q) = CIS.code_grouped_importance_sample(
            sess, target, proposal, seed, bits, max_group_size_bits=m, dim_kl_bit_limit=limit)
        starts_only, group_kls = CIS.code_grouped_importance_sample(
            sess, target, proposal, seed, bits, max_group_size_bits=m, dim_kl_bit_limit=limit, return_group_indices_only=True)
        lst = [int(x) for x in starts[:-1]]
        if n > 1:
            decoded = CIS.decode_grouped_importance_sample(sess, bitcode, lst, proposal, bits, seed, o_idx, o_q)
        else:
            # N = 1: the reference squeezes the sample to a scalar and tf.scatter_nd(shape=[]) cannot run (CIS:348-359)
            decoded, lst = np.zeros(0, np.float32), lst + [n]
    assert isinstance(bitcode, bytes) and lst[-1] == n and np.array_equal(starts, starts_only)
    out.update({name + "_in": np.stack([q_loc, q_scale, p_loc, p_scale]), name + "_par": np.array([seed, bits, m, limit]),
                name + "_sample": np.asarray(sample, np.float32), name + "_bitcode": np.frombuffer(bitcode, np.uint8),
                name + "_starts": np.asarray(starts, np.int64), name + "_oidx": np.asarray(o_idx, np.int64),
                name + "_oq": np.asarray(o_q, np.uint16), name + "_decoded": np.asarray(decoded, np.float32),
                name + "_group_kls": np.asarray(group_kls, np.float64)})
    print(name, "groups", len(starts) - 1, "bits", len(bitcode), "outliers", len(o_idx))

GR_CASES = [  # name, n, key, nontrivial, seed, n_steps, n_bits_per_step, max_group_size_bits
    ("gr0", 200, 10, False, 42, 3, 6, 5), ("gr1", 130, 11, True, 9, 4, 5, 4), ("gr2", 33, 12, True, 77, 2, 7, 12),
]
for name, n, key, nontrivial, seed, steps, b, m in GR_CASES:
    q_loc, q_scale, p_loc, p_scale = posterior(n, key, nontrivial)
    target, proposal = tfd.Normal(loc=q_loc, scale=q_scale), tfd.Normal(loc=p_loc, scale=p_scale)
    with redirect_stdout(sink):
        sample, bitcode, starts = CGS.code_grouped_greedy_sample(sess, target, proposal, steps, b, seed, max_group_size_bits=m)
        lst = list(starts[:-1])
        decoded = CGS.decode_grouped_greedy_sample(sess, bitcode, lst, proposal, b, steps, seed)
    assert isinstance(bitcode, str) and isinstance(starts, list) and lst[-1] == n
    out.update({name + "_in": np.stack([q_loc, q_scale, p_loc, p_scale]), name + "_par": np.array([seed, steps, b, m]),
                name + "_sample": np.asarray(sample, np.float32), name + "_bitcode": np.frombuffer(bitcode.encode(), np.uint8),
                name + "_starts": np.asarray(starts, np.int64), name + "_decoded": np.asarray(decoded, np.float32)})
    print(name, "groups", len(starts) - 1, "bits", len(bitcode))

# ---- the reference's DEFAULT sizes (miracle_arguments.py:147-190): a 20-bit importance-sampler budget with 16-dim groups
# (K up to 385 751 candidates, CIS:48-51) and a greedy group of ~500 dims at 30 steps x 2^14 candidates (CGS:51-78) ----
def default_size_cases():
    f = np.float32
    rng = np.random.default_rng(9100)
    n = 40                                            # 16 + 16 + 7 + the last dim on its own
    q_scale = np.full(n, 0.3) * np.exp(0.05 * rng.normal(size=n))
    q_loc = 0.15 * rng.normal(size=n)                 # ~0.75 nats per dim: a 16-dim group holds ~12 nats, K ~ 1.6e5
    isd = (q_loc.astype(f), q_scale.astype(f), np.zeros(n, f), np.ones(n, f))
    grd = posterior(600, 30, False)                   # first group: ~520 dims inside 30 * 14 bits
    return isd, grd


isd_in, grd_in = default_size_cases()
tf_shim.set_outlier_seed(42)
with redirect_stdout(sink):
    target, proposal = tfd.Normal(loc=isd_in[0], scale=isd_in[1]), tfd.Normal(loc=isd_in[2], scale=isd_in[3])
    sample, bitcode, starts, (o_idx, o_q) = CIS.code_grouped_importance_sample(sess, target, proposal, 42, 20, max_group_size_bits=4,
                                                                               dim_kl_bit_limit=16)
    decoded = CIS.decode_grouped_importance_sample(sess, bitcode, [int(x) for x in starts[:-1]], proposal, 20, 42, o_idx, o_q)
    _, isd_group_kls = CIS.code_grouped_importance_sample(sess, target, proposal, 42, 20, max_group_size_bits=4, dim_kl_bit_limit=16,
                                                          return_group_indices_only=True)
t_loc, t_scale, kl = O.greedy_prologue(*[isd_in[k] for k in (0, 1, 2, 3)])
tot, Ks = O.group_totals(kl, np.asarray(starts, np.int64))
assert Ks.max() >= 100000 and len(o_idx) == 0, (Ks, o_idx)
out.update({"isd_in": np.stack(isd_in), "isd_par": np.array([42, 20, 4, 16]), "isd_sample": np.asarray(sample, np.float32),
            "isd_bitcode": np.frombuffer(bitcode, np.uint8), "isd_starts": np.asarray(starts, np.int64),
            "isd_oidx": np.asarray(o_idx, np.int64), "isd_oq": np.asarray(o_q, np.uint16),
            "isd_decoded": np.asarray(decoded, np.float32), "isd_K": Ks.astype(np.int64),
            "isd_group_kls": np.asarray(isd_group_kls, np.float64)})
print("isd groups", len(starts) - 1, "K", list(Ks), "bits", len(bitcode))
with redirect_stdout(sink):
    target, proposal = tfd.Normal(loc=grd_in[0], scale=grd_in[1]), tfd.Normal(loc=grd_in[2], scale=grd_in[3])
    sample, bitcode, starts = CGS.code_grouped_greedy_sample(sess, target, proposal, 30, 14, 42, max_group_size_bits=12)
    decoded = CGS.decode_grouped_greedy_sample(sess, bitcode, list(starts[:-1]), proposal, 14, 30, 42)
assert max(np.diff(starts)) >= 400, starts
out.update({"grd_in": np.stack(grd_in), "grd_par": np.array([42, 30, 14, 12]), "grd_sample": np.asarray(sample, np.float32),
            "grd_bitcode": np.frombuffer(bitcode.encode(), np.uint8), "grd_starts": np.asarray(starts, np.int64),
            "grd_decoded": np.asarray(decoded, np.float32)})
print("grd groups", len(starts) - 1, "sizes", list(np.diff(starts)), "bits", len(bitcode))

# per-group greedy entry points with a non-zero proposal loc (CGS:29-89, 93-167: shard prior N(p_loc / n_steps, ...))
PG_CASES = [  # name, G, key, seed, n_steps, n_bits_per_step, rho
    ("pg0", 27, 20, 42, 5, 6, 1.0), ("pg1", 5, 21, 7, 2, 7, 0.8), ("pg2", 36, 22, 123, 4, 6, 1.0),
]
for name, G, key, seed, steps, b, rho in PG_CASES:
    rng = np.random.default_rng(9000 + key)
    f = np.float32
    p_loc = rng.normal(size=G).astype(f)
    p_scale = np.log1p(np.exp(rng.normal(size=G))).astype(f)
    t_loc = (p_loc + p_scale * rng.normal(size=G) * 0.8).astype(f)
    t_scale = (p_scale * np.exp(-np.abs(rng.normal(0, .7, G)))).astype(f)
    with redirect_stdout(sink):
        best, code = sess.run(list(CGS.code_greedy_sample(t_loc, t_scale, p_loc, p_scale, b, steps, seed, rho=rho)))
        decoded = sess.run(CGS.decode_greedy_sample(code, p_loc, p_scale, b, steps, seed, rho=rho))
    code = code.decode("utf-8") if isinstance(code, bytes) else str(code)
    out.update({name + "_in": np.stack([t_loc, t_scale, p_loc, p_scale]), name + "_par": np.array([seed, steps, b, rho]),
                name + "_sample": np.asarray(best, np.float32), name + "_bitcode": np.frombuffer(code.encode(), np.uint8),
                name + "_decoded": np.asarray(decoded, np.float32)})
    print(name, "dims", G, "bits", len(code))

np.savez_compressed(os.path.join(HERE, "reference_samplers.npz"), **out)
print("wrote", os.path.join(HERE, "reference_samplers.npz"))
