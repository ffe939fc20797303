"""Image-level codec orchestration (pln.py:213-627, 638-817 minus the networks) on the GPU: file written by
the engine == file assembled from the CPU oracle's outputs (fed the engine's draws), and decode(encode)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPE1, SHAPE2 = [1, 8, 12, 16], [1, 2, 3, 6]


def _setup(syn, key):
    n1, n2 = int(np.prod(SHAPE1)), int(np.prod(SHAPE2))
    q1 = syn.synthetic_posterior(n1, key)
    q2 = syn.synthetic_posterior(n2, key + 1)
    counts1 = np.arange(1, 18) * 3 + 5          # 1 + 2**4 group-size symbols
    counts2 = np.array([3, 10, 20, 30, 400])    # 1 + 2**2
    return n1, n2, q1, q2, counts1, counts2


def _synthesis_2(latents2):
    """Stand-in for synthesis_transform_2: any deterministic map level-2 latents -> level-1 prior."""
    import torch
    m = latents2.float().mean()
    n1 = int(np.prod(SHAPE1))
    i = torch.arange(n1, device=latents2.device, dtype=torch.float32)
    loc = 0.3 * torch.sin(i * 0.37 + m)
    scale = 0.9 + 0.2 * torch.cos(i * 0.11 + 2 * m) ** 2
    return loc, scale


@pytest.mark.parametrize("use_is", [True, False])
def test_codec_file_matches_oracle_assembly_and_round_trips(tmp_path, oracle, use_is):
    import torch
    import compression_without_quantization_b200 as cwq
    from compression_without_quantization_b200 import synthetic as syn
    from compression_without_quantization_b200.misc import candidate_rows
    n1, n2, q1, q2, counts1, counts2 = _setup(syn, 50)
    path = str(tmp_path / "img.miracle")
    seed, steps, bps, gm = 42, 3, 6, 4
    p2 = syn.Normal(np.zeros(n2, np.float32), np.ones(n2, np.float32))
    (s2, s1), summ = cwq.code_image_latents(syn.Normal(q1[0], q1[1]), syn.Normal(q2[0], q2[1]), p2, _synthesis_2, SHAPE1, SHAPE2, seed,
                                            steps, bps, gm, path, use_importance_sampling=use_is,
                                            second_level_n_bits_per_group=12, second_level_max_group_size_bits=2,
                                            second_level_dim_kl_bit_limit=10, first_level_n_bits_per_group=12,
                                            first_level_max_group_size_bits=4, first_level_dim_kl_bit_limit=10,
                                            second_level_group_dist_counts=counts2, first_level_group_dist_counts=counts1,
                                            image_shape=[1, 128, 192, 3])
    assert summ["actual_byte_size"] == os.path.getsize(path) and summ["bpp"] > 0

    # ---- the same file assembled from the oracle (reference order of operations, pln.py:302-537) ----
    perm_1, perm_2 = cwq.latent_permutations(seed, n1, n2)
    rs = np.random.RandomState(seed)
    assert np.array_equal(perm_1, rs.permutation(n1)) and np.array_equal(perm_2, rs.permutation(n2))
    O = oracle
    o2 = O.code_grouped_importance_sample(None, O.Normal(q2[0][perm_2], q2[1][perm_2]), O.Normal(p2.loc[perm_2], p2.scale[perm_2]),
                                          seed, 12, 2, 10, normal_source=candidate_rows)
    assert np.array_equal(o2[0].view(np.uint32), s2.view(np.uint32))
    inv2 = np.empty(n2, np.float32)
    inv2[perm_2] = o2[0]
    p1_loc, p1_scale = (t.cpu().numpy() for t in _synthesis_2(torch.from_numpy(inv2.reshape(SHAPE2)).cuda()))
    if use_is:
        o1 = O.code_grouped_importance_sample(None, O.Normal(q1[0][perm_1], q1[1][perm_1]), O.Normal(p1_loc[perm_1], p1_scale[perm_1]),
                                              seed, 12, 4, 10, normal_source=candidate_rows)
        code1, starts1 = o1[1].decode(), np.asarray(o1[2])
    else:
        o1 = O.code_grouped_greedy_sample(None, O.Normal(q1[0][perm_1], q1[1][perm_1]), O.Normal(p1_loc[perm_1], p1_scale[perm_1]),
                                          steps, bps, seed, gm, normal_source=candidate_rows)
        code1, starts1 = o1[1], np.asarray(o1[2])
    assert np.array_equal(o1[0].view(np.uint32), s1.view(np.uint32))
    extras = [seed, steps, bps, 12, 12, len(code1), len(o2[1])] + SHAPE1[1:3] + SHAPE2[1:3]
    vle = [[int(v) for v in o2[3][0]], [int(v) for v in o2[3][1]]]
    vlb = [24, 16]
    if use_is:
        vle += [[int(v) for v in o1[3][0]], [int(v) for v in o1[3][1]]]
        vlb += [24, 16]
    gi1 = cwq.ArithmeticCoder(counts1).encode(np.concatenate((np.diff(starts1), [0])))
    gi2 = cwq.ArithmeticCoder(counts2).encode(np.concatenate((np.diff(o2[2]), [0])))
    want = str(tmp_path / "want.miracle")
    cwq.write_bin_code(code1 + o2[1].decode(), want, extras=extras, extra_var_bits=[gi1, gi2], var_length_extras=vle, var_length_bits=vlb)
    assert open(path, "rb").read() == open(want, "rb").read()

    # ---- decode the file ----
    d1, d2 = cwq.decode_image_latents(path, _synthesis_2, SHAPE1[3], SHAPE2[3], use_importance_sampling=use_is,
                                      second_level_group_dist_counts=counts2, first_level_group_dist_counts=counts1)
    assert list(d1.shape) == SHAPE1 and list(d2.shape) == SHAPE2
    enc1 = np.empty(n1, np.float32)
    enc1[perm_1] = s1
    enc2 = np.empty(n2, np.float32)
    enc2[perm_2] = s2
    # decode(encode) is exact, outlier dims included (the encoder returns the dequantised outlier values)
    assert np.array_equal(d2.reshape(-1).view(np.uint32), enc2.view(np.uint32))
    assert np.array_equal(d1.reshape(-1).view(np.uint32), enc1.view(np.uint32))


def test_codec_round_trip_with_level2_outliers_and_wide_length_fields(tmp_path):
    """ADVICE r1 / SURVEY 8f N2+N4: (a) level-2 outliers: the level-1 prior the encoder builds from its level-2 sample equals the
    one the decoder builds from the 16-bit outlier values, so level 1 round-trips exactly; (b) more than 65 535 level-1
    outliers: the container switches to 32-bit length fields and decodes; the reference format raises."""
    import compression_without_quantization_b200 as cwq
    from compression_without_quantization_b200 import synthetic as syn
    shape1, shape2 = [1, 40, 60, 32], [1, 5, 6, 8]
    n1, n2 = int(np.prod(shape1)), int(np.prod(shape2))

    def synthesis(latents2):
        import torch
        m = latents2.float().abs().mean()      # depends on every level-2 value, outliers included
        i = torch.arange(n1, device=latents2.device, dtype=torch.float32)
        return 0.3 * torch.sin(i * 0.37 + m), 0.9 + 0.2 * torch.cos(i * 0.11 + 2 * m) ** 2
    q1, q2 = [np.array(a) for a in syn.synthetic_posterior(n1, 60)], [np.array(a) for a in syn.synthetic_posterior(n2, 61)]
    q2[0][::17], q2[1][::17] = 6.0, 0.01       # level-2 outliers (KL >> 10 bits)
    q1[0][:70000], q1[1][:70000] = 7.0, 0.01   # 70 000 level-1 outliers: over the 16-bit count
    counts1, counts2 = np.arange(1, 18) * 3 + 5, np.array([3, 10, 20, 30, 400])
    p2 = syn.Normal(np.zeros(n2, np.float32), np.ones(n2, np.float32))
    kw = dict(second_level_n_bits_per_group=12, second_level_max_group_size_bits=2, second_level_dim_kl_bit_limit=10,
              first_level_n_bits_per_group=12, first_level_max_group_size_bits=4, first_level_dim_kl_bit_limit=10,
              second_level_group_dist_counts=counts2, first_level_group_dist_counts=counts1)
    path = str(tmp_path / "big.miracle")
    args = (syn.Normal(q1[0], q1[1]), syn.Normal(q2[0], q2[1]), p2, synthesis, shape1, shape2, 42, 3, 6, 4)
    (s2, s1), _ = cwq.code_image_latents(*args, path, use_importance_sampling=True, **kw)
    extras = cwq.read_bin_code(path, num_extras=11)[1]
    assert extras[1] == (1 << 31) | 3           # the wide-length marker on n_steps
    d1, d2 = cwq.decode_image_latents(path, synthesis, shape1[3], shape2[3], use_importance_sampling=True,
                                      second_level_group_dist_counts=counts2, first_level_group_dist_counts=counts1)
    perm_1, perm_2 = cwq.latent_permutations(42, n1, n2)
    enc1, enc2 = np.empty(n1, np.float32), np.empty(n2, np.float32)
    enc1[perm_1], enc2[perm_2] = s1, s2
    assert np.array_equal(d2.reshape(-1).view(np.uint32), enc2.view(np.uint32))
    assert np.array_equal(d1.reshape(-1).view(np.uint32), enc1.view(np.uint32))
    with pytest.raises(ValueError):             # the reference's 16-bit fields cannot hold this file (binary_io.py:110)
        cwq.code_image_latents(*args, str(tmp_path / "ref.miracle"), use_importance_sampling=True, wide_length_fields=False, **kw)


def test_empirical_dists_feed_the_codec(tmp_path):
    """miracle.py:521-667 then compress/decompress: histograms from a few 'images', smoothed, used as the
    arithmetic coder's model for a round trip."""
    import compression_without_quantization_b200 as cwq
    from compression_without_quantization_b200 import synthetic as syn
    n1, n2 = int(np.prod(SHAPE1)), int(np.prod(SHAPE2))
    p2 = syn.Normal(np.zeros(n2, np.float32), np.ones(n2, np.float32))

    def images():
        for k in range(3):
            q1, q2 = syn.synthetic_posterior(n1, 80 + k), syn.synthetic_posterior(n2, 90 + k)
            yield syn.Normal(q1[0], q1[1]), syn.Normal(q2[0], q2[1]), p2, _synthesis_2, SHAPE2
    kw = dict(second_level_n_bits_per_group=10, second_level_max_group_size_bits=2, second_level_dim_kl_bit_limit=8,
              first_level_n_bits_per_group=10, first_level_max_group_size_bits=4, first_level_dim_kl_bit_limit=8)
    gs1, gs2, si1, si2 = cwq.build_empirical_dists(images(), 42, **kw)
    assert gs1.size == 17 and gs2.size == 5 and si1.size == 1 + 2 ** 10
    assert gs1[0] == 3 and gs2[0] == 3 and si1[0] == 3
    assert gs1[1:].sum() == si1[1:].sum() and gs2[1:].sum() == si2[1:].sum()     # one index per group
    assert (np.arange(17) * gs1).sum() == 3 * n1 and (np.arange(5) * gs2).sum() == 3 * n2
    q1, q2 = syn.synthetic_posterior(n1, 80), syn.synthetic_posterior(n2, 90)
    path = str(tmp_path / "e.miracle")
    (s2, s1), _ = cwq.code_image_latents(syn.Normal(q1[0], q1[1]), syn.Normal(q2[0], q2[1]), p2, _synthesis_2, SHAPE1, SHAPE2, 42, 3, 6,
                                         4, path, use_importance_sampling=True, second_level_group_dist_counts=gs2 + 1,
                                         first_level_group_dist_counts=gs1 + 1, **kw)
    d1, d2 = cwq.decode_image_latents(path, _synthesis_2, SHAPE1[3], SHAPE2[3], second_level_group_dist_counts=gs2 + 1,
                                      first_level_group_dist_counts=gs1 + 1)
    perm_1, perm_2 = cwq.latent_permutations(42, n1, n2)
    enc1 = np.empty(n1, np.float32)
    enc1[perm_1] = s1
    assert np.mean(d1.reshape(-1) == enc1) > 0.99
