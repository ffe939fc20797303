"""CPU tests of the oracle (test infrastructure): Philox known answers, the reference's own bit helpers,
the grouping examples verified in SURVEY.md appendix A, and the goldens produced by running the
UNMODIFIED reference sampler sources under oracle/tf_shim (tests/golden/gen_reference_goldens.py)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def test_philox4x32_10_known_answers(oracle):
    # Random123 kat_vectors (philox4x32 10), also SURVEY.md appendix A
    kats = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
            ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
            ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
             [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kats:
        assert list(oracle.philox4x32_10(ctr, key)) == want


def test_stream_blocks_are_philox_on_counter_s_k_0_j(oracle):
    w = oracle.stream_words(42, 7, 3, 5)
    for i in range(3):
        for j in range(5):
            assert list(w[i, j]) == list(oracle.philox4x32_10([42, 7 + i, 0, j], [0x6A09E667, 42]))


def test_stream_units_three_blocks_plus_harvested_quad(oracle):
    """unit u of a row owns blocks 3u..3u+2; quads 0-2 are those blocks, quad 3 is made of the
    top bytes of the three blocks (block 3u in bits 0-7, 3u+1 in bits 8-15, 3u+2 in bits 16-23)."""
    seed, k0, nk = 1000 * 45 + 3, 16380, 6
    blocks = oracle.stream_words(seed, k0, nk, 9)
    quads = oracle.quad_words(seed, k0, nk, 11)          # 2 full units + 3 quads
    for u in range(3):
        for b in range(3):
            if 4 * u + b < 11:
                assert np.array_equal(quads[:, 4 * u + b], blocks[:, 3 * u + b])
    for u in range(2):
        top = blocks[:, 3 * u:3 * u + 3] >> 24
        assert np.array_equal(quads[:, 4 * u + 3], top[:, 0] | (top[:, 1] << 8) | (top[:, 2] << 16))
    # the normals are Box-Muller of the low 23 bits of those words (float64 here, libm float32 in the oracle)
    z = oracle.normal_rows(seed, k0, nk, 44)
    m = (quads & 0x7FFFFF).astype(np.float64) / 2.0 ** 23
    u1, th = 1.0 - m[..., 0::2], 2 * np.pi * m[..., 1::2] - np.pi
    r = np.sqrt(-2 * np.log(u1))
    want = np.stack([r * np.cos(th), r * np.sin(th)], -1).reshape(nk, 44)
    assert np.abs(z - want).max() < 2e-5


def test_normal_stream_statistics_and_addressing(oracle):
    z = oracle.normal_rows(42, 0, 40000, 16)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3 and abs(np.mean(z ** 4) - 3) < 0.05
    assert np.abs(z).max() < 5.7   # 23-bit uniforms: r <= sqrt(-2 ln 2^-23)
    assert np.array_equal(oracle.normal_rows(42, 100, 5, 7), z[100:105, :7])
    assert not np.array_equal(oracle.normal_rows(43, 0, 4, 16), z[:4])


def test_harvested_quad_is_standard_normal_and_uncorrelated_with_the_direct_quads(oracle):
    """Dims 12..15 of a unit come from the top bytes of the words whose low 23 bits give dims 0..11:
    same distribution, no linear correlation between any two dims of a unit (nor between their squares)."""
    z = oracle.normal_rows(7, 0, 60000, 32).astype(np.float64)
    for d in range(32):
        col = z[:, d]
        assert abs(col.mean()) < 0.02 and abs(col.std() - 1) < 0.02 and abs(np.mean(col ** 4) - 3) < 0.15
        assert abs(np.mean(np.abs(col) < 0.6744897501960817) - 0.5) < 0.01      # quartiles of N(0,1)
    c = np.corrcoef(z.T)
    c2 = np.corrcoef((z ** 2).T)
    off = ~np.eye(32, dtype=bool)
    assert np.abs(c[off]).max() < 0.02 and np.abs(c2[off]).max() < 0.02


def test_harvested_quad_ks_chi2_and_joint_independence(oracle):
    """Stream v2 takes the fourth quad of every 16 dims from the TOP BYTES of the three Philox blocks whose low 23 bits
    give the other twelve dims.  Goodness of fit of the harvested dims (Kolmogorov-Smirnov against N(0,1),
    chi-square of their 23-bit uniforms and of each harvested byte) and independence of a harvested dim from the three
    direct dims that share its Philox words (chi-square on 16 x 16 contingency tables of the probability transforms).
    Fixed seeds: the p-values are constants, the thresholds only have to sit below them."""
    from scipy import stats
    n = 200000
    z = oracle.normal_rows(11, 0, n, 16).astype(np.float64)
    words = oracle.quad_words(11, 0, n, 4)                       # [n, 4 quads, 4 words]
    pvals = {}
    for d in range(12, 16):
        pvals["ks%d" % d] = stats.kstest(z[:, d], "norm").pvalue
    for d in range(0, 12, 5):                                     # a few direct dims as the yardstick
        pvals["ks%d" % d] = stats.kstest(z[:, d], "norm").pvalue
    # uniforms of the harvested quad: 23-bit mantissas in 256 equal bins, and each harvested byte on its own
    h = words[:, 3, :] & 0x7FFFFF
    for i in range(4):
        pvals["chi_u%d" % i] = stats.chisquare(np.bincount(h[:, i] >> 15, minlength=256)).pvalue
        for byte in range(2):
            pvals["chi_b%d_%d" % (i, byte)] = stats.chisquare(np.bincount((h[:, i] >> (8 * byte)) & 0xFF, minlength=256)).pvalue
        pvals["chi_b%d_2" % i] = stats.chisquare(np.bincount((h[:, i] >> 16) & 0x7F, minlength=128)).pvalue
    # joint: harvested dim 12 + j shares words with dims j, 4 + j, 8 + j (same word position in blocks 0, 1, 2)
    u = stats.norm.cdf(z)
    bins = np.minimum((u * 16).astype(np.int64), 15)
    for j in range(4):
        for other in (j, 4 + j, 8 + j):
            table = np.zeros((16, 16))
            np.add.at(table, (bins[:, 12 + j], bins[:, other]), 1)
            pvals["ind%d_%d" % (12 + j, other)] = stats.chi2_contingency(table)[1]
    # the word-level version of the same question: top byte against the low 23 bits of the SAME word
    blocks = oracle.stream_words(11, 0, n, 3)                     # [n, 3 blocks, 4 words]
    for b in range(3):
        table = np.zeros((16, 16))
        np.add.at(table, ((blocks[:, b, 0] >> 28).astype(np.int64), ((blocks[:, b, 0] >> 19) & 0xF).astype(np.int64)), 1)
        pvals["word%d" % b] = stats.chi2_contingency(table)[1]
    low = {k: v for k, v in pvals.items() if v < 1e-3}
    assert not low, low
    # 41 tests: uniformly distributed p-values put about two below 0.05; a defect would put most of them near 0
    assert sum(v < 0.05 for v in pvals.values()) <= 6, sorted(pvals.items(), key=lambda kv: kv[1])[:8]


def test_deterministic_log_exp_accuracy(oracle):
    rng = np.random.default_rng(0)
    xs = np.exp(rng.uniform(-30, 30, 4000)).astype(np.float32)
    got = np.array([oracle.detlogf(float(x)) for x in xs])
    assert np.max(np.abs(got - np.log(xs.astype(np.float64))) / np.maximum(np.abs(np.log(xs.astype(np.float64))), 1e-2)) < 3e-7
    xe = rng.uniform(-10, 21.4, 4000).astype(np.float32)
    ge = np.array([oracle.detexpf(float(x)) for x in xe])
    assert np.max(np.abs(ge / np.exp(xe.astype(np.float64)) - 1)) < 3e-7
    assert oracle.detexpf(0.0) == 1.0 and oracle.detlogf(1.0) == 0.0


def test_binary_io_against_reference_module_goldens(oracle):
    g = json.load(open(os.path.join(GOLD, "binary_io.json")))
    for x, code in g["elias_delta"]:
        assert oracle.elias_delta_code(x) == code
        assert oracle.elias_delta_decode((code + "0110").encode()) == (x, len(code))
    for v, bits, s in g["to_bit_string"]:
        assert oracle.to_bit_string(v, bits) == s and oracle.from_bit_string(s) == v
    with pytest.raises(Exception, match="bigger than what we can encode"):
        oracle.to_bit_string(8, 3)


def test_grouping_examples_from_survey(oracle):
    kl = np.full(40, 1e-6, np.float32)
    assert list(oracle.group(kl, 20, 2, True)) == list(range(0, 40, 4)) + [39, 40]
    assert list(oracle.group(kl, 20, 4, True)) == [0, 16, 32, 39, 40]
    assert list(oracle.group(kl, 20, 3, False)) == [0, 7, 14, 21, 28, 35, 39, 40]
    assert list(oracle.group(np.array([100, 1e-3, 1e-3, 1e-3], np.float32), 20, 4, True)) == [0, 0, 1, 3, 4]
    assert list(oracle.group(np.array([0.1], np.float32), 20, 4, True)) == [0, 0, 1]
    assert oracle.size_limit(4, True) == 16 and oracle.size_limit(12, False) == 4095 and oracle.size_limit(2, True) == 4


def test_grouping_matches_verbatim_python_loop(oracle):
    """CIS:178-204 / CGS:223-252 transcribed literally (numpy float32 accumulator, float64 budget)."""
    rng = np.random.default_rng(3)
    for strict, m, bits in [(True, 4, 20), (True, 2, 8), (False, 5, 30), (False, 12, 420)]:
        kl_divs = (rng.gamma(0.6, 0.9, 3000)).astype(np.float32)
        starts = [0]
        size, acc = 0, 0
        budget = bits * np.log(2) - 1
        n = kl_divs.size
        for idx in range(n):
            gb = np.log(size + 1) / np.log(2)
            hit = (gb > m or acc + kl_divs[idx] > budget) if strict else (gb >= m or acc + kl_divs[idx] >= budget)
            if hit or idx == n - 1:
                starts.append(idx)
                size, acc = 1, kl_divs[idx]
            else:
                acc += kl_divs[idx]
                size += 1
        starts += [n]
        assert list(oracle.group(kl_divs, bits, m, strict)) == starts


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "reference_samplers.npz"))


@pytest.mark.parametrize("name", ["is0", "is1", "is2", "is3", "is4", "isd"])
def test_importance_sampler_matches_unmodified_reference_run(oracle, gold, name):
    q_loc, q_scale, p_loc, p_scale = gold[name + "_in"]
    seed, bits, m, limit = [int(x) for x in gold[name + "_par"]]
    s, code, starts, (oi, oq) = oracle.code_grouped_importance_sample(None, oracle.Normal(q_loc, q_scale),
                                                                      oracle.Normal(p_loc, p_scale), seed, bits, m, limit,
                                                                      outlier_mode="reference")
    assert np.array_equal(starts, gold[name + "_starts"])
    assert code == gold[name + "_bitcode"].tobytes()
    assert np.array_equal(oi, gold[name + "_oidx"]) and np.array_equal(oq, gold[name + "_oq"])
    assert np.allclose(s, gold[name + "_sample"], rtol=1e-5, atol=1e-6)
    st2, kls = oracle.code_grouped_importance_sample(None, oracle.Normal(q_loc, q_scale), oracle.Normal(p_loc, p_scale), seed,
                                                     bits, m, limit, return_group_indices_only=True)
    assert np.allclose(kls, gold[name + "_group_kls"], rtol=2e-5, atol=1e-6)   # TFP KL formula vs spec kl01: same value
    if gold[name + "_decoded"].size:
        lst = [int(x) for x in starts[:-1]]
        d = oracle.decode_grouped_importance_sample(None, code, lst, oracle.Normal(p_loc, p_scale), bits, seed, oi, oq,
                                                    outlier_mode="reference")
        assert lst[-1] == q_loc.size
        assert np.allclose(d, gold[name + "_decoded"], rtol=1e-5, atol=1e-6)
        # the hardened outlier path: the encoder's sample IS the decoder's, outliers included
        sh = oracle.code_grouped_importance_sample(None, oracle.Normal(q_loc, q_scale), oracle.Normal(p_loc, p_scale), seed, bits,
                                                   m, limit)[0]
        dh = oracle.decode_grouped_importance_sample(None, code, [int(x) for x in starts[:-1]], oracle.Normal(p_loc, p_scale), bits,
                                                     seed, oi, oq)
        assert np.array_equal(sh, dh)
        inl = np.ones(q_loc.size, bool)
        inl[oi] = False
        assert np.array_equal(sh[inl], s[inl]) and np.all(np.abs(sh[oi] - s[oi]) <= 30 / 65535 + 1e-6)


def test_is4_covers_outliers_and_empty_first_group(gold):
    assert list(gold["is4_oidx"]) == [10, 33, 43] and list(gold["is4_starts"][:3]) == [0, 0, 1]


def test_default_size_goldens_are_default_sized(gold):
    """isd: 20-bit budget, 16-dim groups with K > 1e5 candidates (CIS:48-51); grd: a 558-dim greedy group at 30 x 2^14."""
    assert list(gold["isd_par"]) == [42, 20, 4, 16] and gold["isd_K"].max() >= 100000
    assert list(np.diff(gold["isd_starts"])[:2]) == [16, 16]
    assert list(gold["grd_par"]) == [42, 30, 14, 12] and np.diff(gold["grd_starts"]).max() >= 500


@pytest.mark.parametrize("name", ["gr0", "gr1", "gr2", "grd"])
def test_greedy_sampler_matches_unmodified_reference_run(oracle, gold, name):
    q_loc, q_scale, p_loc, p_scale = gold[name + "_in"]
    seed, steps, b, m = [int(x) for x in gold[name + "_par"]]
    s, code, starts = oracle.code_grouped_greedy_sample(None, oracle.Normal(q_loc, q_scale), oracle.Normal(p_loc, p_scale), steps,
                                                        b, seed, m)
    assert starts == [int(x) for x in gold[name + "_starts"]]
    ref = gold[name + "_bitcode"].tobytes().decode()
    d = oracle.decode_grouped_greedy_sample(None, code, starts[:-1], oracle.Normal(p_loc, p_scale), b, steps, seed)
    assert np.array_equal(d, s)
    if name != "grd":
        assert code == ref
        assert np.allclose(s, gold[name + "_sample"], rtol=1e-5, atol=1e-6)
        assert np.allclose(d, gold[name + "_decoded"], rtol=1e-5, atol=1e-6)
        return
    # default sizes: the 558- and the 41-dim group are identical to the reference's.  The last dim is a group of ONE dim
    # searched with 2^14 candidates per step: TFP's log_prob (-u^2/2 - log sigma - log sqrt(2 pi), float32, ~6e-8 apart near
    # its maximum) cannot tell candidates within |u| < 3.4e-4 of the best apart and takes the first of them; the spec's
    # score u^2 can, and takes the closest.  Both land within that resolution of the same point.
    per = steps * b
    agree = [code[i * per:(i + 1) * per] == ref[i * per:(i + 1) * per] for i in range(len(starts) - 1)]
    assert agree[:2] == [True, True] and np.diff(starts)[2] == 1
    cut = starts[2]
    assert np.allclose(s[:cut], gold[name + "_sample"][:cut], rtol=1e-5, atol=1e-6)
    assert abs(float(s[cut]) - float(gold[name + "_sample"][cut])) <= 2e-3 * float(q_scale[cut])
    dref = oracle.decode_grouped_greedy_sample(None, ref, starts[:-1], oracle.Normal(p_loc, p_scale), b, steps, seed)
    assert np.allclose(dref, gold[name + "_decoded"], rtol=1e-5, atol=1e-6)      # the oracle decodes the reference's stream


@pytest.mark.parametrize("name", ["pg0", "pg1", "pg2"])
def test_per_group_greedy_with_proposal_loc_matches_unmodified_reference_run(oracle, gold, name):
    """CGS:29-89 / 93-167 with p_loc != 0: shard prior N(p_loc / n_steps, rho p_scale / sqrt(n_steps))."""
    t_loc, t_scale, p_loc, p_scale = gold[name + "_in"]
    seed, steps, b = [int(x) for x in gold[name + "_par"][:3]]
    rho = float(gold[name + "_par"][3])
    assert np.any(p_loc != 0)
    best, code = oracle.code_greedy_sample(t_loc, t_scale, p_loc, p_scale, b, steps, seed, rho=rho)
    assert code.encode() == gold[name + "_bitcode"].tobytes()
    assert np.allclose(best, gold[name + "_sample"], rtol=1e-5, atol=1e-6)
    d = oracle.decode_greedy_sample(code, p_loc, p_scale, b, steps, seed, rho=rho)
    assert np.allclose(d, gold[name + "_decoded"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(d, best)


# ------------------------------------------------------------------------------------------------
# BASELINE.json's parity harness: the unmodified reference fed the ENGINE's own draws (saved from a B200 through the
# injection seam, tests/golden/gen_engine_draw_goldens.py); the GPU tests compare the engine with the same files.
# ------------------------------------------------------------------------------------------------
def _engine_draw_source():
    import sys
    sys.path.insert(0, GOLD)
    import gen_engine_draw_goldens as G
    draws = np.load(os.path.join(GOLD, "engine_draws.npz"))
    return G, G.Tables(stored={k: draws[k] for k in draws.files}), np.load(os.path.join(GOLD, "reference_engine_draws.npz"))


def test_oracle_equals_reference_run_on_engine_draws(oracle):
    G, src, gold = _engine_draw_source()
    for name, n, key, nontrivial, seed, bits, m, limit in G.IS_CASES:
        q_loc, q_scale, p_loc, p_scale = gold[name + "_in"]
        s, code, starts, (oi, oq) = oracle.code_grouped_importance_sample(
            None, oracle.Normal(q_loc, q_scale), oracle.Normal(p_loc, p_scale), seed, bits, m, limit, normal_source=src,
            outlier_mode="reference")
        assert np.array_equal(starts, gold[name + "_starts"]) and code == gold[name + "_bitcode"].tobytes(), name
        assert np.array_equal(oi, gold[name + "_oidx"]) and np.array_equal(oq, gold[name + "_oq"]), name
        assert np.allclose(s, gold[name + "_sample"], rtol=1e-5, atol=1e-6), name
    assert any(gold[c[0] + "_oidx"].size for c in G.IS_CASES)          # the outlier path is covered
    for name, n, key, nontrivial, seed, steps, b, m in G.GR_CASES:
        q_loc, q_scale, p_loc, p_scale = gold[name + "_in"]
        s, code, starts = oracle.code_grouped_greedy_sample(None, oracle.Normal(q_loc, q_scale), oracle.Normal(p_loc, p_scale), steps, b,
                                                            seed, m, normal_source=src)
        assert starts == [int(x) for x in gold[name + "_starts"]] and code.encode() == gold[name + "_bitcode"].tobytes(), name
        assert np.allclose(s, gold[name + "_sample"], rtol=1e-5, atol=1e-6), name
    for name, Gd, key, seed, steps, b, rho in G.PG_CASES:
        t_loc, t_scale, p_loc, p_scale = gold[name + "_in"]
        best, code = oracle.code_greedy_sample(t_loc, t_scale, p_loc, p_scale, b, steps, seed, rho=rho, normal_source=src)
        assert code.encode() == gold[name + "_bitcode"].tobytes(), name
        assert np.allclose(best, gold[name + "_sample"], rtol=1e-5, atol=1e-6), name


def test_oracle_round_trips_and_injection_seam(oracle):
    rng = np.random.default_rng(1)
    n = 500
    q_loc = (rng.normal(size=n) * (rng.random(n) < .4)).astype(np.float32)
    q_scale = np.exp(-np.abs(rng.normal(0, .7, n))).astype(np.float32)
    p = oracle.Normal(np.zeros(n, np.float32), np.ones(n, np.float32))
    s, code, starts, (oi, oq) = oracle.code_grouped_importance_sample(None, oracle.Normal(q_loc, q_scale), p, 42, 10, 3, 6)
    d = oracle.decode_grouped_importance_sample(None, code, [int(x) for x in starts[:-1]], p, 10, 42, oi, oq)
    assert np.array_equal(d, s)
    # an injected source that returns the default stream must change nothing; a different one must
    calls = []

    def src(seed, k0, nk, G):
        calls.append((seed, k0, nk, G))
        return oracle.default_normal_source(seed, k0, nk, G)
    s2, code2, _, _ = oracle.code_grouped_importance_sample(None, oracle.Normal(q_loc, q_scale), p, 42, 10, 3, 6, normal_source=src)
    assert code2 == code and np.array_equal(s2, s) and calls[0][0] in (42, (42 + 0x9E3779B9) & 0xFFFFFFFF)
    s3, code3, _, _ = oracle.code_grouped_importance_sample(None, oracle.Normal(q_loc, q_scale), p, 42, 10, 3, 6,
                                                            normal_source=lambda sd, k0, nk, G: oracle.default_normal_source(sd + 1, k0, nk, G))
    assert code3 != code


def test_quantizer_round_trip(oracle):
    x = np.array([-40, -30, -1.234, 0.0, 5.5, 29.9999, 30, 99], np.float32)
    q = oracle.quantize_u16(x)
    assert q[0] == 0 and q[1] == 0 and q[-1] == 65535 and q[-2] == 65535
    assert np.all(np.abs(oracle.dequantize_u16(q)[1:-1] - np.clip(x, -30, 30)[1:-1]) <= 30 / 65535 + 1e-6)
