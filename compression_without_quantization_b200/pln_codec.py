"""Image-level codec orchestration around the coders: the part of the reference's
`ProbabilisticLadderNetwork.code_image_greedy` / `decode_image_greedy` (code/pln.py:213-627, 638-817)
that is NOT the conv nets -- seeded permutation of the latents, level 2 -> level 1 sequencing, group-size
side information through the arithmetic coder, outlier lists, the `.miracle` container and the size
summaries.  The networks stay with the caller: it passes the flat posterior / prior parameters and a
`synthesis_2` callback that maps the decoded level-2 latents to the level-1 prior
(`self.synthesis_transform_2`, pln.py:397-412, 757-767).

File layout (pln.py:491-537, 660-671): 11 big-endian u32 extras
`[seed, n_steps, n_bits_per_step, n_bits_group_1, n_bits_group_2, len(code1), len(code2), H1, W1, H2, W2]`,
two 16-bit-length-prefixed AC blobs (group sizes of level 1, level 2, each terminated by symbol 0),
outlier lists of level 2 (and of level 1 with importance sampling) as 24-bit indices / 16-bit values,
then code1 + code2 zero padded to a byte.

Extension for large images: the reference's 16-bit length fields (binary_io.py:97,110) overflow once a
blob holds more than 65 535 coded bits or an outlier list more than 65 535 entries (a 3840x2160 image).  With
`wide_length_fields="auto"` such a file is written with 32-bit length fields and bit 31 of the second extra (n_steps, always
small) set as the marker; `decode_image_latents` recognises the marker.  Files that fit the reference format are written
byte-identically to it; `wide_length_fields=False` keeps the reference's behaviour (an exception on overflow).

Known reference behaviours kept: the decoder assumes a standard-normal level-2
prior (pln.py:729-730); an empty first group has size 0, which is also the AC end symbol, so such a file
cannot be decoded (the encoder warns).  The greedy level-1 path, which cannot run in the reference
(`list - list`, pln.py:477), works here.
"""
import os

import numpy as np
import torch

from . import engine as E
from .binary_io import read_bin_code, write_bin_code
from .coded_greedy_sampler import code_grouped_greedy_sample, decode_grouped_greedy_sample
from .coded_importance_sampler import code_grouped_importance_sample, decode_grouped_importance_sample
from .coding import ArithmeticCoder


WIDE_LENGTH_FLAG = 1 << 31     # bit 31 of extras[1] (n_steps): the file has 32-bit length fields


class _Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale, self.dtype = loc, scale, np.float32


def latent_permutations(seed, num_dim_1, num_dim_2, use_permutation=True):
    """pln.py:302-310 / 697-705: `np.random.seed(seed)`, then the level-1 and the level-2 permutation from
    numpy's legacy MT19937 stream, in that order."""
    if not use_permutation:
        return np.arange(num_dim_1, dtype=np.int32), np.arange(num_dim_2, dtype=np.int32)
    rs = np.random.RandomState(int(seed))
    return rs.permutation(int(num_dim_1)).astype("int32"), rs.permutation(int(num_dim_2)).astype("int32")


def _counts(c):
    return np.load(c) if isinstance(c, (str, bytes, os.PathLike)) else np.asarray(c)


def _flat32(x, dev):
    return E.as_device_f32(x, dev)


def _forward(x, perm_d):      # tfp Permute.forward: y[i] = x[perm[i]]
    return x.index_select(0, perm_d)


def _inverse(y, perm_d):      # x[perm[i]] = y[i]
    x = torch.empty_like(y)
    x.index_copy_(0, perm_d, y)
    return x


def _kl_nats(q_loc, q_scale, p_loc, p_scale):
    r = (q_scale.double() / p_scale.double()) ** 2
    return float((0.5 * ((q_loc.double() - p_loc.double()) / p_scale.double()) ** 2 + 0.5 * (r - 1 - torch.log(r))).sum().item())


def code_image_latents(q1, q2, p2, synthesis_2, first_level_shape, second_level_shape, seed, n_steps, n_bits_per_step,
                       greedy_max_group_size_bits, comp_file_path, rho=1., use_importance_sampling=False, use_permutation=True,
                       second_level_n_bits_per_group=20, second_level_max_group_size_bits=4, second_level_dim_kl_bit_limit=12,
                       first_level_n_bits_per_group=20, first_level_max_group_size_bits=3, first_level_dim_kl_bit_limit=12,
                       outlier_index_bytes=3, outlier_sample_bytes=2, second_level_group_dist_counts=None,
                       first_level_group_dist_counts=None, image_shape=None, verbose=False, wide_length_fields="auto"):
    """pln.py:213-627 without the networks.  q1, q2, p2: objects with flat .loc/.scale (level-1 posterior,
    level-2 posterior and prior, UNPERMUTED); synthesis_2(latents2 with shape second_level_shape) ->
    (p1_loc, p1_scale) flat, unpermuted.  Returns ((sample2, sample1), summaries) like the reference
    (samples in permuted order, as pln.py:627 returns them) and writes comp_file_path."""
    dev = E.cuda_device()
    q1_loc, q1_scale = _flat32(q1.loc, dev), _flat32(q1.scale, dev)
    q2_loc, q2_scale = _flat32(q2.loc, dev), _flat32(q2.scale, dev)
    p2_loc, p2_scale = _flat32(p2.loc, dev), _flat32(p2.scale, dev)
    n1, n2 = q1_loc.numel(), q2_loc.numel()
    perm_1, perm_2 = latent_permutations(seed, n1, n2, use_permutation)
    perm_1d, perm_2d = torch.from_numpy(perm_1.astype(np.int64)).to(dev), torch.from_numpy(perm_2.astype(np.int64)).to(dev)

    # ---- level 2 (pln.py:350-359) ----
    q2p = _Normal(_forward(q2_loc, perm_2d), _forward(q2_scale, perm_2d))
    p2p = _Normal(_forward(p2_loc, perm_2d), _forward(p2_scale, perm_2d))
    sample2, code2, group_indices2, outlier_extras2 = code_grouped_importance_sample(
        None, q2p, p2p, seed, second_level_n_bits_per_group, second_level_max_group_size_bits, second_level_dim_kl_bit_limit)
    group_differences2 = group_indices2[1:] - group_indices2[:-1]

    # ---- level-1 prior from the decoded level-2 sample (pln.py:388-412) ----
    sample2_d = torch.from_numpy(np.ascontiguousarray(sample2)).to(dev)
    latents2 = _inverse(sample2_d, perm_2d).reshape(tuple(int(v) for v in second_level_shape))
    p1_loc_u, p1_scale_u = synthesis_2(latents2)
    p1_loc_u, p1_scale_u = _flat32(p1_loc_u, dev), _flat32(p1_scale_u, dev)
    q1p = _Normal(_forward(q1_loc, perm_1d), _forward(q1_scale, perm_1d))
    p1p = _Normal(_forward(p1_loc_u, perm_1d), _forward(p1_scale_u, perm_1d))

    # ---- level 1 (pln.py:420-472) ----
    if use_importance_sampling:
        sample1, code1, group_indices1, outlier_extras1 = code_grouped_importance_sample(
            None, q1p, p1p, seed, first_level_n_bits_per_group, first_level_max_group_size_bits, first_level_dim_kl_bit_limit)
        code1 = code1.decode("utf-8")
    else:
        sample1, code1, group_indices1 = code_grouped_greedy_sample(None, q1p, p1p, n_steps, n_bits_per_step, seed,
                                                                    greedy_max_group_size_bits, rho=rho)
        group_indices1 = np.asarray(group_indices1, dtype=np.int64)
        outlier_extras1 = None
    group_differences1 = group_indices1[1:] - group_indices1[:-1]
    bitcode = code1 + code2.decode("utf-8")

    # ---- file (pln.py:491-537) ----
    extras = [int(seed), int(n_steps), int(n_bits_per_step), int(first_level_n_bits_per_group), int(second_level_n_bits_per_group),
              len(code1), len(code2)] + [int(v) for v in first_level_shape[1:3]] + [int(v) for v in second_level_shape[1:3]]
    var_length_extras = [[int(v) for v in outlier_extras2[0]], [int(v) for v in outlier_extras2[1]]]
    var_length_bits = [outlier_index_bytes * 8, outlier_sample_bytes * 8]
    if use_importance_sampling:
        var_length_extras += [[int(v) for v in outlier_extras1[0]], [int(v) for v in outlier_extras1[1]]]
        var_length_bits += [outlier_index_bytes * 8, outlier_sample_bytes * 8]
    coder2 = ArithmeticCoder(_counts(second_level_group_dist_counts), precision=32)
    coder1 = ArithmeticCoder(_counts(first_level_group_dist_counts), precision=32)
    if verbose and (group_differences1.min() == 0 or group_differences2.min() == 0):
        print("warning: an empty group has size 0, the arithmetic coder's end symbol; the reference format cannot decode this file")
    gi1_code = coder1.encode(np.concatenate((group_differences1, [0])))
    gi2_code = coder2.encode(np.concatenate((group_differences2, [0])))
    overflow = max([len(gi1_code), len(gi2_code)] + [len(v) for v in var_length_extras]) > 65535
    wide = wide_length_fields is True or (wide_length_fields == "auto" and overflow)
    if wide:
        extras[1] |= WIDE_LENGTH_FLAG
    write_bin_code(bitcode, comp_file_path, extras=extras, extra_var_bits=[gi1_code, gi2_code],
                   var_length_extras=var_length_extras, var_length_bits=var_length_bits, extra_bytes=4 if wide else 2)

    # ---- summaries (pln.py:543-577) ----
    total_kls = [_kl_nats(q1_loc, q1_scale, p1_loc_u, p1_scale_u), _kl_nats(q2_loc, q2_scale, p2_loc, p2_scale)]
    total_kl = sum(total_kls)
    theoretical_byte_size = (total_kl + 2 * np.log(total_kl + 1)) / np.log(2) / 8
    extra_byte_size = len(gi1_code) + len(gi2_code) + 9 * 2 // 8       # pln.py:548, as written
    actual_byte_size = os.path.getsize(comp_file_path)
    summaries = {
        "image_shape": list(image_shape) if image_shape is not None else None,
        "theoretical_byte_size": float(theoretical_byte_size),
        "actual_byte_size": actual_byte_size,
        "extra_byte_size": extra_byte_size,
        "actual_no_extra": actual_byte_size - extra_byte_size,
    }
    if image_shape is not None:
        pixels = image_shape[1] * image_shape[2]
        second_level_extra = len(gi2_code) // 8 + 1
        summaries["second_bpp"] = (len(code2) / 8 + second_level_extra) * 8 / pixels
        summaries["bpp"] = 8 * actual_byte_size / pixels
    summaries["first_level_groups"] = int(len(group_indices1) - 1)
    summaries["second_level_groups"] = int(len(group_indices2) - 1)
    if verbose:
        print("Theoretical size: {:.2f} bytes".format(theoretical_byte_size))
        print("Actual size: {:.2f} bytes".format(actual_byte_size))
        print("Efficiency: {:.3f}".format(actual_byte_size / theoretical_byte_size))
    return (sample2, sample1), summaries


def decode_image_latents(comp_file_path, synthesis_2, first_level_latent_channels, second_level_latent_channels,
                         use_importance_sampling=True, rho=1., use_permutation=True, second_level_group_dist_counts=None,
                         first_level_group_dist_counts=None, verbose=False):
    """pln.py:638-812 without the networks -> (decoded_first_level with shape [1, H1, W1, C1],
    decoded_second_level with shape [1, H2, W2, C2]) as float32 numpy arrays; the caller feeds the first
    to its synthesis_transform_1 (pln.py:815)."""
    dev = E.cuda_device()
    num_var_length_extras = 4 if use_importance_sampling else 2
    _, extras, _, _ = read_bin_code(comp_file_path, num_extras=11)
    wide = bool(extras[1] & WIDE_LENGTH_FLAG)
    code, extras, extra_var_bits, var_length_extras = read_bin_code(comp_file_path, num_extras=11, num_extra_var_bits=2,
                                                                    num_var_length_extras=num_var_length_extras,
                                                                    extra_bytes=4 if wide else 2)
    extras[1] &= ~WIDE_LENGTH_FLAG
    coder2 = ArithmeticCoder(_counts(second_level_group_dist_counts), precision=32)
    coder1 = ArithmeticCoder(_counts(first_level_group_dist_counts), precision=32)
    seed, n_steps, n_bits_per_step = extras[0], extras[1], extras[2]
    first_level_n_bits_per_group, second_level_n_bits_per_group = extras[3], extras[4]
    first_code_length, second_code_length = extras[5], extras[6]
    first_level_shape = [1] + extras[7:9] + [first_level_latent_channels]
    second_level_shape = [1] + extras[9:] + [second_level_latent_channels]
    n1, n2 = int(np.prod(first_level_shape)), int(np.prod(second_level_shape))
    perm_1, perm_2 = latent_permutations(seed, n1, n2, use_permutation)
    perm_1d, perm_2d = torch.from_numpy(perm_1.astype(np.int64)).to(dev), torch.from_numpy(perm_2.astype(np.int64)).to(dev)
    group_differences2 = coder2.decode_fast(extra_var_bits[1])[:-1]     # chop the terminating 0 (pln.py:708)
    group_differences1 = coder1.decode_fast(extra_var_bits[0])[:-1]
    code1 = code[:first_code_length]
    code2 = code[first_code_length:first_code_length + second_code_length]

    # ---- level 2: the proposal is the standard normal (pln.py:729-730) ----
    group_indices2 = [0]
    for d in group_differences2:
        group_indices2.append(group_indices2[-1] + d)
    proposal2 = _Normal(torch.zeros(n2, device=dev), torch.ones(n2, device=dev))
    decoded2 = decode_grouped_importance_sample(None, code2, group_indices2[:-1], proposal2, second_level_n_bits_per_group, seed,
                                                var_length_extras[0], var_length_extras[1])
    d2 = torch.from_numpy(np.ascontiguousarray(decoded2)).to(dev)
    latents2 = _inverse(d2, perm_2d).reshape(tuple(second_level_shape))
    p1_loc_u, p1_scale_u = synthesis_2(latents2)
    prior_1 = _Normal(_forward(_flat32(p1_loc_u, dev), perm_1d), _forward(_flat32(p1_scale_u, dev), perm_1d))

    # ---- level 1 ----
    group_indices1 = [0]
    for d in group_differences1:
        group_indices1.append(group_indices1[-1] + d)
    if use_importance_sampling:
        decoded1 = decode_grouped_importance_sample(None, code1, group_indices1[:-1], prior_1, first_level_n_bits_per_group, seed,
                                                    var_length_extras[2], var_length_extras[3])
    else:
        decoded1 = decode_grouped_greedy_sample(None, code1, group_indices1[:-1], prior_1, n_bits_per_step, n_steps, seed, rho=rho)
    d1 = _inverse(torch.from_numpy(np.ascontiguousarray(decoded1)).to(dev), perm_1d)
    return (d1.reshape(tuple(first_level_shape)).cpu().numpy(), latents2.cpu().numpy())


def build_empirical_dists(images, seed, second_level_n_bits_per_group=20, second_level_max_group_size_bits=4,
                          second_level_dim_kl_bit_limit=12, first_level_n_bits_per_group=20, first_level_max_group_size_bits=3,
                          first_level_dim_kl_bit_limit=12, use_permutation=True, with_sample_indices=True):
    """The histogram part of the reference's `build_empirical_dists` (miracle.py:521-667) without the
    networks: `images` yields (q1, q2, p2, synthesis_2, second_level_shape) per image; the result is
    (group_sizes1, group_sizes2, sample_indices1, sample_indices2) int64 count arrays with symbol 0 (the
    arithmetic coder's end symbol) counted once per image, exactly as miracle.py:621-662 accumulates them.
    The arithmetic coder needs every count positive (coding.pyx builds a strict search tree), so symbols that
    never occurred must be smoothed by the caller before use, as in the reference."""
    dev = E.cuda_device()
    gs2 = np.zeros(1 + 2 ** second_level_max_group_size_bits, dtype=np.int64)
    gs1 = np.zeros(1 + 2 ** first_level_max_group_size_bits, dtype=np.int64)
    si2 = np.zeros(1 + 2 ** second_level_n_bits_per_group, dtype=np.int64) if with_sample_indices else None
    si1 = np.zeros(1 + 2 ** first_level_n_bits_per_group, dtype=np.int64) if with_sample_indices else None
    for q1, q2, p2, synthesis_2, second_level_shape in images:
        q1_loc, q1_scale = _flat32(q1.loc, dev), _flat32(q1.scale, dev)
        q2_loc, q2_scale = _flat32(q2.loc, dev), _flat32(q2.scale, dev)
        p2_loc, p2_scale = _flat32(p2.loc, dev), _flat32(p2.scale, dev)
        n1, n2 = q1_loc.numel(), q2_loc.numel()
        perm_1, perm_2 = latent_permutations(seed, n1, n2, use_permutation)
        perm_1d = torch.from_numpy(perm_1.astype(np.int64)).to(dev)
        perm_2d = torch.from_numpy(perm_2.astype(np.int64)).to(dev)
        q2p = _Normal(_forward(q2_loc, perm_2d), _forward(q2_scale, perm_2d))
        p2p = _Normal(_forward(p2_loc, perm_2d), _forward(p2_scale, perm_2d))
        sample2, idx2, starts2, _ = code_grouped_importance_sample(None, q2p, p2p, seed, second_level_n_bits_per_group,
                                                                   second_level_max_group_size_bits, second_level_dim_kl_bit_limit,
                                                                   return_indices=True)
        u, c = np.unique(np.diff(starts2), return_counts=True)      # miracle.py:621-625
        gs2[u] += c
        gs2[0] += 1
        latents2 = _inverse(torch.from_numpy(np.ascontiguousarray(sample2)).to(dev), perm_2d).reshape(
            tuple(int(v) for v in second_level_shape))
        p1_loc_u, p1_scale_u = synthesis_2(latents2)
        q1p = _Normal(_forward(q1_loc, perm_1d), _forward(q1_scale, perm_1d))
        p1p = _Normal(_forward(_flat32(p1_loc_u, dev), perm_1d), _forward(_flat32(p1_scale_u, dev), perm_1d))
        _, idx1, starts1, _ = code_grouped_importance_sample(None, q1p, p1p, seed, first_level_n_bits_per_group,
                                                             first_level_max_group_size_bits, first_level_dim_kl_bit_limit,
                                                             return_indices=True)
        u, c = np.unique(np.diff(starts1), return_counts=True)      # miracle.py:630-636
        gs1[u] += c
        gs1[0] += 1
        if with_sample_indices:
            for idx, hist in ((idx2, si2), (idx1, si1)):               # miracle.py:645-662 (indices are 1-based, shifted by one)
                u, c = np.unique(np.asarray(idx, dtype=np.int64), return_counts=True)
                hist[u + 1] += c
                hist[0] += 1
    return gs1, gs2, si1, si2
