"""Live cross-check against the UNMODIFIED reference sources (only where /root/reference exists, i.e. in the build container;
skipped elsewhere): random small problems through the reference's own coded_importance_sampler.py / coded_greedy_sampler.py
(executed under oracle/tf_shim, tf.random.stateless_normal answering from the spec stream) and through the oracle -- group
starts, bitstreams, outlier side information bit for bit, latents to 1e-5.  The committed goldens (tests/golden/) are fixed
cases of the same comparison; this sweeps parameters the goldens do not."""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest

REF = "/root/reference/code"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "coded_importance_sampler.py")),
                                reason="the reference sources are only present in the build container")


@pytest.fixture(scope="module")
def ref():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from oracle import oracle as O
    from oracle import tf_shim
    O.build()
    tf = tf_shim.install()
    tf_shim.set_normal_source(O.as_standard_normal(O.default_normal_source))
    sys.path.insert(0, REF)
    import tqdm as _tqdm
    _tqdm.tqdm = lambda it, *a, **k: it
    import coded_greedy_sampler as CGS
    import coded_importance_sampler as CIS
    import tensorflow_probability as tfp
    return O, tf_shim, CIS, CGS, tfp.distributions, tf.Session()


def _posterior(rng, n, nontrivial):
    q_scale = np.exp(-np.abs(rng.normal(0.0, 0.7, n)))
    q_loc = 1.2 * rng.normal(size=n) * (rng.random(n) < 0.35)
    if nontrivial:
        p_loc, p_scale = rng.normal(size=n), np.log1p(np.exp(rng.normal(size=n)))
        q_loc, q_scale = p_loc + p_scale * q_loc, p_scale * q_scale
    else:
        p_loc, p_scale = np.zeros(n), np.ones(n)
    f = np.float32
    return q_loc.astype(f), q_scale.astype(f), p_loc.astype(f), p_scale.astype(f)


def test_importance_sampler_random_sweep_against_the_reference(ref):
    O, shim, CIS, CGS, tfd, sess = ref
    rng = np.random.default_rng(20261018)
    sink = io.StringIO()
    for case in range(24):
        n = int(rng.integers(2, 260))
        bits, m = int(rng.integers(6, 13)), int(rng.integers(0, 6))
        limit = int(rng.integers(4, 12))
        seed = int(rng.integers(0, 2 ** 31 - 1))
        q_loc, q_scale, p_loc, p_scale = _posterior(rng, n, case % 2 == 1)
        if case % 5 == 0:
            q_loc[0], q_scale[0] = p_loc[0], p_scale[0] * 1e-3            # dim 0 alone over the budget / an outlier
        shim.set_outlier_seed(seed)
        target, proposal = tfd.Normal(loc=q_loc, scale=q_scale), tfd.Normal(loc=p_loc, scale=p_scale)
        with redirect_stdout(sink):
            r_sample, r_code, r_starts, (r_oi, r_oq) = CIS.code_grouped_importance_sample(
                sess, target, proposal, seed, bits, max_group_size_bits=m, dim_kl_bit_limit=limit)
            r_dec = CIS.decode_grouped_importance_sample(sess, r_code, [int(x) for x in r_starts[:-1]], proposal, bits, seed, r_oi, r_oq)
        ctx = (case, n, bits, m, limit, seed)
        s, code, starts, (oi, oq) = O.code_grouped_importance_sample(None, O.Normal(q_loc, q_scale), O.Normal(p_loc, p_scale), seed,
                                                                     bits, m, limit, outlier_mode="reference")
        assert np.array_equal(starts, r_starts), ctx
        assert code == r_code, ctx
        assert np.array_equal(oi, r_oi) and np.array_equal(oq, np.asarray(r_oq, np.uint16)), ctx
        assert np.allclose(s, r_sample, rtol=1e-5, atol=1e-6), ctx
        d = O.decode_grouped_importance_sample(None, code, [int(x) for x in starts[:-1]], O.Normal(p_loc, p_scale), bits, seed, oi, oq,
                                               outlier_mode="reference")
        assert np.allclose(d, r_dec, rtol=1e-5, atol=1e-6), ctx


def test_greedy_sampler_random_sweep_against_the_reference(ref):
    O, shim, CIS, CGS, tfd, sess = ref
    rng = np.random.default_rng(42)
    sink = io.StringIO()
    for case in range(16):
        n = int(rng.integers(2, 160))
        steps, b, m = int(rng.integers(1, 6)), int(rng.integers(2, 8)), int(rng.integers(1, 8))
        seed = int(rng.integers(0, 2 ** 20))
        rho = float(rng.choice([1.0, 0.9, 1.2]))
        q_loc, q_scale, p_loc, p_scale = _posterior(rng, n, case % 2 == 0)
        target, proposal = tfd.Normal(loc=q_loc, scale=q_scale), tfd.Normal(loc=p_loc, scale=p_scale)
        with redirect_stdout(sink):
            r_sample, r_code, r_starts = CGS.code_grouped_greedy_sample(sess, target, proposal, steps, b, seed, max_group_size_bits=m, rho=rho)
            r_dec = CGS.decode_grouped_greedy_sample(sess, r_code, list(r_starts[:-1]), proposal, b, steps, seed, rho=rho)
        ctx = (case, n, steps, b, m, seed, rho)
        s, code, starts = O.code_grouped_greedy_sample(None, O.Normal(q_loc, q_scale), O.Normal(p_loc, p_scale), steps, b, seed, m, rho=rho)
        assert starts == [int(x) for x in r_starts], ctx
        assert code == r_code, ctx
        assert np.allclose(s, r_sample, rtol=1e-5, atol=1e-6), ctx
        d = O.decode_grouped_greedy_sample(None, code, starts[:-1], O.Normal(p_loc, p_scale), b, steps, seed, rho=rho)
        assert np.allclose(d, r_dec, rtol=1e-5, atol=1e-6), ctx
        # the per-group entry points with this case's first group and a non-zero proposal loc
        G = starts[1]
        if G > 0:
            with redirect_stdout(sink):
                rb, rc = sess.run(list(CGS.code_greedy_sample(q_loc[:G], q_scale[:G], p_loc[:G] + 0.5, p_scale[:G], b, steps, seed, rho=rho)))
            rc = rc.decode() if isinstance(rc, bytes) else str(rc)
            ob, oc = O.code_greedy_sample(q_loc[:G], q_scale[:G], p_loc[:G] + 0.5, p_scale[:G], b, steps, seed, rho=rho)
            assert oc == rc and np.allclose(ob, rb, rtol=1e-5, atol=1e-6), ctx
