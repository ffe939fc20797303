"""CPU ORACLE for the MIRACLE sampler hot path -- test infrastructure, NOT product code.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module.  It restates, on numpy + a small C library (oracle/cwq_oracle.c), the
reference functions below (paths under /root/reference/code), following the canonical arithmetic
spec of DESIGN.md section 3:

    coded_importance_sampler.py:29-79    code_importance_sample
    coded_importance_sampler.py:82-109   decode_importance_sample
    coded_importance_sampler.py:112-274  code_grouped_importance_sample
    coded_importance_sampler.py:277-363  decode_grouped_importance_sample
    coded_greedy_sampler.py:29-89        code_greedy_sample
    coded_greedy_sampler.py:93-167       decode_greedy_sample
    coded_greedy_sampler.py:170-296      code_grouped_greedy_sample
    coded_greedy_sampler.py:299-364      decode_grouped_greedy_sample
    misc.py:3-17                         stateless_normal_sample   (the injection seam)
    binary_io.py:7-67                    elias_delta_code/decode, to_bit_string/from_bit_string

The ``normal_source`` argument of every sampler is the seam: a callable
``normal_source(stream_seed, k0, nk, G) -> float32[nk, G]`` returning rows k0..k0+nk-1 of the
N(0,1) candidate table the reference would have drawn with
``tf.random.stateless_normal([K, G], seed=[stream_seed, 42])``.  The default evaluates the spec's
Philox4x32-10 + Box-Muller stream with libm; the GPU parity tests pass the engine's own draws
(``cwq_draw_candidates``) so that indices and bitstreams can be compared bit for bit.

Parity status: pinned against (i) Random123 Philox KATs, (ii) the reference's binary_io.py,
(iii) goldens produced by running the unmodified reference sampler sources under oracle/tf_shim
(tests/golden/).  TensorFlow's own kernels are NOT available here: that part stays unpinned.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcwq_oracle.so")

_c_f32p = ctypes.POINTER(ctypes.c_float)
_c_u32p = ctypes.POINTER(ctypes.c_uint32)
_c_i32p = ctypes.POINTER(ctypes.c_int32)
_c_i64p = ctypes.POINTER(ctypes.c_int64)
_c_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile oracle/cwq_oracle.c with the committed Makefile (no-op when up to date)."""
    src = os.path.join(_HERE, "cwq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.cwqo_detlogf.restype = ctypes.c_float
        L.cwqo_detlogf.argtypes = [ctypes.c_float]
        L.cwqo_detexpf.restype = ctypes.c_float
        L.cwqo_detexpf.argtypes = [ctypes.c_float]
        L.cwqo_kl01.restype = ctypes.c_float
        L.cwqo_kl01.argtypes = [ctypes.c_float, ctypes.c_float]
        L.cwqo_group.restype = ctypes.c_int64
        L.cwqo_group.argtypes = [_c_f32p, ctypes.c_int64, ctypes.c_double, ctypes.c_int64, ctypes.c_int, _c_i64p]
        L.cwqo_is_encode_group.restype = ctypes.c_uint32
        L.cwqo_is_encode_group.argtypes = [_c_f32p, _c_f32p, ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint32,
                                           _c_f32p, _c_f32p]
        L.cwqo_greedy_shard_scale.restype = ctypes.c_float
        L.cwqo_greedy_shard_scale.argtypes = [ctypes.c_float, ctypes.c_int32]
        L.cwqo_greedy_step_scale.restype = ctypes.c_float
        L.cwqo_greedy_step_scale.argtypes = [ctypes.c_float]
        L.cwqo_normal_scale.restype = ctypes.c_float
        L.cwqo_quantize_u16.restype = ctypes.c_uint16
        L.cwqo_quantize_u16.argtypes = [ctypes.c_float]
        L.cwqo_dequantize_u16.restype = ctypes.c_float
        L.cwqo_dequantize_u16.argtypes = [ctypes.c_uint16]
        L.cwqo_outlier_normal.restype = ctypes.c_float
        L.cwqo_outlier_normal.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        _lib = L
    return _lib


def _f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1))


def _p(a, t):
    return a.ctypes.data_as(t)


def _u32(x):
    return int(x) & 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------
# Low-level pieces
# ----------------------------------------------------------------------------------------------
LN2 = np.log(2)  # float64, as the reference uses it (CIS:142,181)
OUTLIER_STREAM_OFFSET = 0x9E3779B9


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32).copy()
    k = np.asarray(key, dtype=np.uint32).copy()
    out = np.zeros(4, dtype=np.uint32)
    lib().cwqo_philox4x32_10(_p(c, _c_u32p), _p(k, _c_u32p), _p(out, _c_u32p))
    return out


STREAM_KEY = (0x6A09E667, 42)   # block(s, k, j) = philox(ctr = (s, k, 0, j), key = STREAM_KEY)


def stream_words(stream_seed, k0, nk, nb):
    """Raw Philox words of blocks 0..nb-1 of rows k0..k0+nk-1: uint32[nk, nb, 4]."""
    out = np.zeros((nk, nb, 4), dtype=np.uint32)
    lib().cwqo_stream_words(ctypes.c_uint32(_u32(stream_seed)), ctypes.c_uint32(_u32(k0)), ctypes.c_uint32(nk),
                            ctypes.c_uint32(nb), _p(out, _c_u32p))
    return out


def quad_words(stream_seed, k0, nk, nq):
    """The four uniform words (before the 23-bit mask) of quads 0..nq-1 of rows k0..k0+nk-1: uint32[nk, nq, 4].
    Quads 0-2 of a 16-dim unit are the unit's three blocks, quad 3 is harvested from their top bytes."""
    out = np.zeros((nk, nq, 4), dtype=np.uint32)
    lib().cwqo_quad_words(ctypes.c_uint32(_u32(stream_seed)), ctypes.c_uint32(_u32(k0)), ctypes.c_uint32(nk),
                          ctypes.c_uint32(nq), _p(out, _c_u32p))
    return out


NORMAL_SCALE = np.float32(1.1774100)     # float32(sqrt(2 ln 2)): the N(0,1) draw is NORMAL_SCALE * (raw draw), DESIGN.md 3.1


def _rows(stream_seed, k0, nk, G, raw):
    out = np.zeros((nk, G), dtype=np.float32)
    if nk > 0 and G > 0:
        lib().cwqo_normals(ctypes.c_uint32(_u32(stream_seed)), ctypes.c_uint32(_u32(k0)), ctypes.c_uint32(nk),
                           ctypes.c_uint32(G), ctypes.c_int(raw), _p(out, _c_f32p))
    return out


def default_normal_source(stream_seed, k0, nk, G):
    """The injection seam: rows k0..k0+nk-1 of the [K, G] table of RAW draws y of the spec stream (y ~ N(0, 1/(2 ln 2));
    the candidate loops score y directly).  A `normal_source` passed to the functions below must return raw draws too --
    the engine's are compression_without_quantization_b200.misc.candidate_rows."""
    return _rows(stream_seed, k0, nk, G, 1)


def normal_rows(stream_seed, k0, nk, G):
    """misc.py:10-11 restated on the spec stream: rows of the [K, G] N(0,1) table, z = NORMAL_SCALE * y."""
    return _rows(stream_seed, k0, nk, G, 0)


def as_standard_normal(normal_source):
    """N(0,1) view of a raw-draw source: what stands in for tf.random.stateless_normal in the reference harness."""
    return lambda seed, k0, nk, G: (NORMAL_SCALE * np.asarray(normal_source(seed, k0, nk, G), dtype=np.float32)).astype(np.float32)


def stateless_normal_sample(loc, scale, num_samples, seed, normal_source=None):
    """misc.py:3-17: [num_samples, *loc.shape] normals, then scale * z + loc."""
    loc = np.asarray(loc, dtype=np.float32)
    scale = np.asarray(scale, dtype=np.float32)
    z = _normal(normal_source, seed, 0, int(num_samples), int(loc.size)).reshape((int(num_samples),) + loc.shape)
    return (scale[None] * z + loc[None]).astype(np.float32)


def detlogf(x):
    return float(lib().cwqo_detlogf(ctypes.c_float(x)))


def detexpf(x):
    return float(lib().cwqo_detexpf(ctypes.c_float(x)))


def kl01(mu, sigma):
    mu = _f32(mu)
    sigma = _f32(sigma)
    return np.array([lib().cwqo_kl01(ctypes.c_float(m), ctypes.c_float(s)) for m, s in zip(mu, sigma)],
                    dtype=np.float32)


def size_limit(max_group_size_bits, strict):
    """Smallest running group size at which the reference's size test fires.

    CIS:185-187 tests ``np.log(size + 1) / np.log(2) > max_group_size_bits`` (strict) and CGS:230-232
    the same with ``>=``; both are evaluated here verbatim around 2**bits."""
    m = max_group_size_bits
    guess = int(np.floor(2.0 ** float(m)))
    for size in range(max(0, guess - 3), guess + 4):
        bits = np.log(size + 1) / np.log(2)
        if (bits > m) if strict else (bits >= m):
            return size
    raise ValueError("size limit not bracketed")


def is_prologue(q_loc, q_scale, p_loc, p_scale, dim_kl_bit_limit):
    q_loc, q_scale, p_loc, p_scale = map(_f32, (q_loc, q_scale, p_loc, p_scale))
    n = q_loc.size
    t_loc = np.empty(n, np.float32)
    t_scale = np.empty(n, np.float32)
    kl = np.empty(n, np.float32)
    mask = np.empty(n, np.uint8)
    lib().cwqo_is_prologue(_p(q_loc, _c_f32p), _p(q_scale, _c_f32p), _p(p_loc, _c_f32p), _p(p_scale, _c_f32p),
                           ctypes.c_int64(n), ctypes.c_float(dim_kl_bit_limit), _p(t_loc, _c_f32p),
                           _p(t_scale, _c_f32p), _p(kl, _c_f32p), _p(mask, _c_u8p))
    return t_loc, t_scale, kl, mask


def greedy_prologue(q_loc, q_scale, p_loc, p_scale):
    q_loc, q_scale, p_loc, p_scale = map(_f32, (q_loc, q_scale, p_loc, p_scale))
    n = q_loc.size
    t_loc = np.empty(n, np.float32)
    t_scale = np.empty(n, np.float32)
    kl = np.empty(n, np.float32)
    lib().cwqo_greedy_prologue(_p(q_loc, _c_f32p), _p(q_scale, _c_f32p), _p(p_loc, _c_f32p), _p(p_scale, _c_f32p),
                               ctypes.c_int64(n), _p(t_loc, _c_f32p), _p(t_scale, _c_f32p), _p(kl, _c_f32p))
    return t_loc, t_scale, kl


def group(kl, n_bits_per_group, max_group_size_bits, strict):
    """G1 (CIS:164-204 strict / CGS:207-252 non-strict) -> int64 starts of length n_groups + 1."""
    kl = _f32(kl)
    n = kl.size
    n_nats_per_group = n_bits_per_group * np.log(2) - 1  # CIS:181, CGS:226 (float64)
    starts = np.empty(n + 2, np.int64)
    ng = lib().cwqo_group(_p(kl, _c_f32p), ctypes.c_int64(n), ctypes.c_double(float(n_nats_per_group)),
                          ctypes.c_int64(size_limit(max_group_size_bits, strict)), ctypes.c_int(1 if strict else 0),
                          _p(starts, _c_i64p))
    return starts[:ng + 1].copy()


def group_totals(kl, starts):
    kl = _f32(kl)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    ng = starts.size - 1
    total = np.empty(ng, np.float32)
    K = np.empty(ng, np.int32)
    lib().cwqo_group_totals(_p(kl, _c_f32p), _p(starts, _c_i64p), ctypes.c_int64(ng), _p(total, _c_f32p),
                            _p(K, _c_i32p))
    return total, K


def is_coef(t_loc, t_scale):
    t_loc, t_scale = _f32(t_loc), _f32(t_scale)
    a = np.empty_like(t_loc)
    b = np.empty_like(t_loc)
    lib().cwqo_is_coef(_p(t_loc, _c_f32p), _p(t_scale, _c_f32p), ctypes.c_int64(t_loc.size), _p(a, _c_f32p),
                       _p(b, _c_f32p))
    return a, b


def is_encode_group(a, b, K, stream_seed, draws=None):
    """CIS:55-63 -> (index, best score).  draws: optional injected float32[K, G]."""
    a, b = _f32(a), _f32(b)
    G = a.size
    score = ctypes.c_float(0)
    inj = None
    if draws is not None:
        draws = np.ascontiguousarray(draws, dtype=np.float32).reshape(int(K), G)
        inj = _p(draws, _c_f32p)
    idx = lib().cwqo_is_encode_group(_p(a, _c_f32p), _p(b, _c_f32p), ctypes.c_int32(G), ctypes.c_uint32(int(K)),
                                     ctypes.c_uint32(_u32(stream_seed)), inj, ctypes.byref(score))
    return int(idx), float(score.value)


def is_encode_groups(a, b, starts, K, seed, g0=0, g1=None, n_threads=1):
    """CIS:233-245 over groups [g0, g1) with the oracle's own draws, on n_threads host threads."""
    a, b = _f32(a), _f32(b)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    K = np.ascontiguousarray(K, dtype=np.int32)
    ng = starts.size - 1
    g1 = ng if g1 is None else g1
    idx = np.zeros(ng, np.uint32)
    z = np.zeros(a.size, np.float32)
    lib().cwqo_is_encode_groups(_p(a, _c_f32p), _p(b, _c_f32p), _p(starts, _c_i64p), _p(K, _c_i32p),
                                ctypes.c_int64(g0), ctypes.c_int64(g1), ctypes.c_uint32(_u32(seed)),
                                ctypes.c_int(n_threads), _p(idx, _c_u32p), _p(z, _c_f32p))
    return idx, z


def greedy_shard_scale(rho, n_steps):
    return float(lib().cwqo_greedy_shard_scale(ctypes.c_float(rho), ctypes.c_int32(n_steps)))


def greedy_shard_loc(p_loc, p_scale, n_steps):
    """CGS:42 in rescaled coordinates: (p_loc / n_steps) / p_scale, float32."""
    p_loc, p_scale = _f32(p_loc), _f32(p_scale)
    out = np.zeros(max(p_loc.size, 1), np.float32)
    lib().cwqo_greedy_shard_loc(_p(p_loc, _c_f32p), _p(p_scale, _c_f32p), ctypes.c_int64(p_loc.size), ctypes.c_int32(n_steps),
                                _p(out, _c_f32p))
    return out[:p_loc.size]


def greedy_encode_group(t_loc, t_scale, n_steps, n_samples, shard_scale, group_seed, draws=None, shard_loc=None):
    t_loc, t_scale = _f32(t_loc), _f32(t_scale)
    G = t_loc.size
    idx = np.zeros(n_steps, np.uint32)
    best = np.zeros(max(G, 1), np.float32)
    inj = None
    if draws is not None:
        draws = np.ascontiguousarray(draws, dtype=np.float32).reshape(n_steps, n_samples, G)
        inj = _p(draws, _c_f32p)
    sl = None if shard_loc is None else _p(_f32(shard_loc), _c_f32p)
    lib().cwqo_greedy_encode_group(_p(t_loc, _c_f32p), _p(t_scale, _c_f32p), sl, ctypes.c_int32(G),
                                   ctypes.c_int32(n_steps), ctypes.c_uint32(n_samples), ctypes.c_float(shard_scale),
                                   ctypes.c_uint32(_u32(group_seed)), inj, _p(idx, _c_u32p), _p(best, _c_f32p))
    return idx, best[:G]


def greedy_encode_groups(t_loc, t_scale, starts, n_steps, n_samples, shard_scale, seed, g0=0, g1=None, n_threads=1):
    t_loc, t_scale = _f32(t_loc), _f32(t_scale)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    ng = starts.size - 1
    g1 = ng if g1 is None else g1
    idx = np.zeros((ng, n_steps), np.uint32)
    best = np.zeros(t_loc.size, np.float32)
    lib().cwqo_greedy_encode_groups(_p(t_loc, _c_f32p), _p(t_scale, _c_f32p), _p(starts, _c_i64p),
                                    ctypes.c_int64(g0), ctypes.c_int64(g1), ctypes.c_int32(n_steps),
                                    ctypes.c_uint32(n_samples), ctypes.c_float(shard_scale),
                                    ctypes.c_uint32(_u32(seed)), ctypes.c_int(n_threads), _p(idx, _c_u32p),
                                    _p(best, _c_f32p))
    return idx, best


def quantize_u16(x):
    return np.array([lib().cwqo_quantize_u16(ctypes.c_float(v)) for v in _f32(x)], dtype=np.uint16)


def dequantize_u16(q):
    return np.array([lib().cwqo_dequantize_u16(ctypes.c_uint16(int(v))) for v in np.asarray(q).reshape(-1)],
                    dtype=np.float32)


# ----------------------------------------------------------------------------------------------
# binary_io.py:7-67 restated (integer arithmetic instead of float logs; equal for all x <= 2**21,
# SURVEY appendix A; checked against the reference module in tests/golden/binary_io.json)
# ----------------------------------------------------------------------------------------------
def to_bit_string(num, num_bits):
    num = int(num)
    if num >= 2 ** int(num_bits):
        raise Exception("The number {} (>= {}) is bigger than what we can encode!".format(num, 2 ** int(num_bits)))
    return ''.join('1' if (num >> i) & 1 else '0' for i in range(int(num_bits)))


def from_bit_string(bitcode):
    if isinstance(bitcode, bytes):
        bitcode = bitcode.decode("utf-8")
    return sum(1 << i for i, ch in enumerate(bitcode) if ch == "1")


def elias_delta_code(x):
    x = int(x)
    n = x.bit_length() - 1
    l = (n + 1).bit_length() - 1
    return "0" * l + to_bit_string(n + 1, l + 1)[::-1] + to_bit_string(x, n + 1)[::-1][1:]


def elias_delta_decode(x):
    if isinstance(x, str):
        x = x.encode("utf-8")
    l = 0
    while x[l] == 48:
        l += 1
    x = x[l:]
    n_plus_one = from_bit_string(x[:l + 1][::-1])
    x = x[l + 1:]
    num = from_bit_string(x[:n_plus_one - 1][::-1] + b"1")
    return num, 2 * l + n_plus_one


# ----------------------------------------------------------------------------------------------
# The reference's eight entry points, restated
# ----------------------------------------------------------------------------------------------
class Normal:
    """Stand-in for tfd.Normal: only .loc, .scale and .dtype are read by the reference (CIS:124-138)."""

    def __init__(self, loc, scale):
        self.loc = np.asarray(loc)
        self.scale = np.asarray(scale)
        self.dtype = np.result_type(self.loc.dtype, self.scale.dtype)


def _check_f32(dist, what):
    if np.dtype(dist.dtype) != np.float32:
        raise Exception("{} datatype must be float32!".format(what))


def _draw(normal_source, stream_seed, k0, nk, G):
    """raw draws y (what the encode loops are fed)"""
    src = normal_source or default_normal_source
    return np.ascontiguousarray(src(_u32(stream_seed), k0, nk, G), dtype=np.float32).reshape(nk, G)


def _normal(normal_source, stream_seed, k0, nk, G):
    """N(0,1) draws z = NORMAL_SCALE * y (what a latent is made of), one float32 multiply"""
    return (NORMAL_SCALE * _draw(normal_source, stream_seed, k0, nk, G)).astype(np.float32)


def code_importance_sample(t_loc, t_scale, p_loc, p_scale, n_coding_bits, seed, return_index_only=False,
                           normal_source=None):
    """CIS:29-79 for one group -> (best_sample f32[1, G], bitcode str | index + 1)."""
    t_loc, t_scale, p_loc, p_scale = map(_f32, (t_loc, t_scale, p_loc, p_scale))
    G = t_loc.size
    r_loc, r_scale, kl = greedy_prologue(t_loc, t_scale, p_loc, p_scale)  # rescale + KL, no masking
    starts = np.array([0, G], np.int64)
    _, K = group_totals(kl, starts)
    K = int(K[0])
    a, b = is_coef(r_loc, r_scale)
    draws = _draw(normal_source, seed, 0, K, G) if normal_source is not None else None
    index, _ = is_encode_group(a, b, K, seed, draws)
    z = _normal(normal_source, seed, index, 1, G)
    best_sample = (p_scale[None] * z + p_loc[None]).astype(np.float32)
    if return_index_only:
        return best_sample, index + 1
    return best_sample, elias_delta_code(index + 1)


def decode_importance_sample(sample_index, p_loc, p_scale, seed, use_index=False, normal_source=None):
    """CIS:82-109 -> row `index` of the candidate table (and code length / index when parsing a string)."""
    p_loc, p_scale = _f32(p_loc), _f32(p_scale)
    G = p_loc.size
    if use_index:
        index = int(sample_index) - 1
        z = _normal(normal_source, seed, index, 1, G)
        return (p_scale[None] * z + p_loc[None]).astype(np.float32)
    num, code_length = elias_delta_decode(sample_index)
    index = num - 1
    z = _normal(normal_source, seed, index, 1, G)
    return (p_scale[None] * z + p_loc[None]).astype(np.float32), code_length, index


def code_grouped_importance_sample(sess, target, proposal, seed, n_bits_per_group, max_group_size_bits=4,
                                   dim_kl_bit_limit=12, return_group_indices_only=False, return_indices=False,
                                   return_indices_only=False, normal_source=None, outlier_mode="hardened"):
    """CIS:112-274.  outlier_mode "reference": the returned sample carries the unquantised outlier draws (CIS:267);
    "hardened" (what the engine does): it carries dequantise(quantise(draw)), i.e. exactly what the decoder
    reconstructs, so that encoder and decoder agree on the level-2 sample the level-1 prior is built from."""
    _check_f32(target, "Target")
    _check_f32(proposal, "Proposal")
    q_loc, q_scale, p_loc, p_scale = map(_f32, (target.loc, target.scale, proposal.loc, proposal.scale))
    num_dimensions = q_loc.size
    # CIS:133-163
    t_loc, t_scale, kl_divs, mask = is_prologue(q_loc, q_scale, p_loc, p_scale, dim_kl_bit_limit)
    outlier_indices = np.nonzero(mask & 2)[0].astype(np.int64)
    # CIS:150 target.sample(), seeded by the spec (the reference's draw is unseeded)
    z_out = np.array([_normal(normal_source, _u32(seed) + OUTLIER_STREAM_OFFSET, int(i) & 0xFFFFFFFF, 1, 1)[0, 0]
                      for i in outlier_indices], dtype=np.float32)
    target_samples_at_outliers = (q_scale[outlier_indices] * z_out + q_loc[outlier_indices]).astype(np.float32)
    outlier_samples = quantize_u16(target_samples_at_outliers)
    outlier_extras = (outlier_indices, outlier_samples)
    # CIS:164-204
    group_start_indices = group(kl_divs, n_bits_per_group, max_group_size_bits, strict=True)
    group_total, K = group_totals(kl_divs, group_start_indices)
    if return_group_indices_only:
        # CIS:192 appends the KL of the group that just closed, starting with the empty pre-group
        group_kls = [float(x) / LN2 for x in group_total[:-1]]
        return group_start_indices, group_kls
    # CIS:233-245
    a, b = is_coef(t_loc, t_scale)
    n_groups = group_start_indices.size - 1
    indices = []
    sample = np.zeros(num_dimensions, np.float32)
    for i in range(n_groups):
        s, e = int(group_start_indices[i]), int(group_start_indices[i + 1])
        G = e - s
        ss = _u32(seed + i)
        draws = _draw(normal_source, ss, 0, int(K[i]), G) if normal_source is not None else None
        index, _ = is_encode_group(a[s:e], b[s:e], int(K[i]), ss, draws)
        indices.append(index + 1)
        if G > 0:
            sample[s:e] = _normal(normal_source, ss, index, 1, G)[0]
    if return_indices_only:
        return tuple(indices)
    # CIS:262-267
    sample = (p_scale * sample + p_loc).astype(np.float32)
    if outlier_mode == "reference":
        sample[outlier_indices] = target_samples_at_outliers
    else:
        sample[outlier_indices] = dequantize_u16(outlier_samples)
    if return_indices:
        return sample, indices, group_start_indices, outlier_extras
    bitcode = ''.join(elias_delta_code(ix) for ix in indices).encode("utf-8")
    return sample, bitcode, group_start_indices, outlier_extras


def decode_grouped_importance_sample(sess, bitcode, group_start_indices, proposal, n_bits_per_group, seed,
                                     outlier_indices, outlier_samples, use_indices=False, normal_source=None,
                                     outlier_mode="hardened"):
    """CIS:277-363 (use_indices follows the evident intent of CIS:88-96; the shipped branch cannot run).
    outlier_mode "reference": dense scatter + `updates == 0` sentinel (CIS:351-361); "hardened": the listed dims are
    overwritten, whatever their dequantised value."""
    _check_f32(proposal, "Proposal")
    p_loc, p_scale = _f32(proposal.loc), _f32(proposal.scale)
    num_dimensions = p_loc.size
    group_start_indices.append(num_dimensions)  # CIS:294 mutates the caller's list
    if isinstance(bitcode, str):
        bitcode = bitcode.encode("utf-8")
    sample = np.zeros(num_dimensions, np.float32)
    pos = 0
    for i in range(len(group_start_indices) - 1):
        s, e = int(group_start_indices[i]), int(group_start_indices[i + 1])
        if use_indices:
            index = int(bitcode[i]) - 1
        else:
            num, codelength = elias_delta_decode(bitcode[pos:])
            pos += codelength
            index = num - 1
        if e > s:
            sample[s:e] = _normal(normal_source, _u32(seed + i), index, 1, e - s)[0]
    sample = (p_scale * sample + p_loc).astype(np.float32)
    # CIS:351-361
    outlier_indices = np.asarray(outlier_indices, dtype=np.int64).reshape(-1)
    if outlier_mode == "reference":
        updates = np.zeros(num_dimensions, np.float32)
        updates[outlier_indices] = dequantize_u16(outlier_samples)
        return np.where(updates == 0, sample, updates).astype(np.float32)
    sample[outlier_indices] = dequantize_u16(outlier_samples)
    return sample


def code_greedy_sample(t_loc, t_scale, p_loc, p_scale, n_bits_per_step, n_steps, seed, rho=1., normal_source=None):
    """CGS:29-89 for one group -> (best_sample f32[G], bitcode str)."""
    t_loc, t_scale, p_loc, p_scale = map(_f32, (t_loc, t_scale, p_loc, p_scale))
    G = t_loc.size
    n_samples = int(2 ** n_bits_per_step)
    # The shard prior is N(p_loc / n_steps, rho p_scale / sqrt(n_steps)) (CGS:42-45): every step moves the running sample by
    # p_loc / n_steps before the target scores it (CGS:57).  Spec: everything is divided by p_scale only (no loc shift), the
    # shard loc becomes (p_loc / n_steps) / p_scale and the final sample is p_scale * best.
    zeros = np.zeros(G, np.float32)
    r_loc, r_scale, _ = greedy_prologue(t_loc, t_scale, zeros, p_scale)
    s = greedy_shard_scale(rho, n_steps)
    sl = greedy_shard_loc(p_loc, p_scale, n_steps)
    draws = None
    if normal_source is not None:
        draws = np.stack([_draw(normal_source, 1000 * _u32(seed) + i, 0, n_samples, G) for i in range(n_steps)])
    idx, best = greedy_encode_group(r_loc, r_scale, n_steps, n_samples, s, seed, draws, shard_loc=sl)
    best_sample = (p_scale * best + zeros).astype(np.float32)
    return best_sample, ''.join(to_bit_string(int(i), n_bits_per_step) for i in idx)


def _greedy_replay(G, indices, n_steps, shard_scale, group_seed, normal_source, shard_loc=None):
    out = np.zeros(G, np.float32)
    step_scale = float(lib().cwqo_greedy_step_scale(ctypes.c_float(shard_scale)))      # shard_scale * NORMAL_SCALE: the rows are raw
    for i in range(n_steps):
        z = np.ascontiguousarray(_draw(normal_source, 1000 * _u32(group_seed) + i, int(indices[i]), 1, G)[0])
        if shard_loc is not None:
            out = (out + shard_loc).astype(np.float32)
        if G > 0:
            lib().cwqo_axpy_fma(ctypes.c_float(step_scale), _p(z, _c_f32p), _p(out, _c_f32p), ctypes.c_int64(G))
    return out


def decode_greedy_sample(sample_index, p_loc, p_scale, n_bits_per_step, n_steps, seed, rho=1., normal_source=None):
    """CGS:93-167."""
    p_loc, p_scale = _f32(p_loc), _f32(p_scale)
    G = p_loc.size
    if isinstance(sample_index, bytes):
        sample_index = sample_index.decode("utf-8")
    indices = [from_bit_string(sample_index[i:i + n_bits_per_step])
               for i in range(0, n_bits_per_step * n_steps, n_bits_per_step)]
    s = greedy_shard_scale(rho, n_steps)
    sl = greedy_shard_loc(p_loc, p_scale, n_steps)
    if normal_source is None:
        out = np.zeros(max(G, 1), np.float32)
        idx = np.asarray(indices, dtype=np.uint32)
        lib().cwqo_greedy_decode_group(ctypes.c_int32(G), ctypes.c_int32(n_steps), ctypes.c_float(s), _p(sl, _c_f32p) if G else None,
                                       ctypes.c_uint32(_u32(seed)), _p(idx, _c_u32p), _p(out, _c_f32p))
        best = out[:G]
    else:
        best = _greedy_replay(G, indices, n_steps, s, seed, normal_source, sl)
    return (p_scale * best + np.zeros(G, np.float32)).astype(np.float32)


def code_grouped_greedy_sample(sess, target, proposal, n_steps, n_bits_per_step, seed, max_group_size_bits=12,
                               adaptive=True, backfitting_steps=0, use_log_prob=False, rho=1., normal_source=None):
    """CGS:170-296 -> (sample f32[N], bitcode str, group_start_indices list incl. N)."""
    _check_f32(target, "Target")
    _check_f32(proposal, "Proposal")
    n_bits_per_group = n_bits_per_step * n_steps
    q_loc, q_scale, p_loc, p_scale = map(_f32, (target.loc, target.scale, proposal.loc, proposal.scale))
    num_dimensions = q_loc.size
    t_loc, t_scale, kl_divs = greedy_prologue(q_loc, q_scale, p_loc, p_scale)
    starts = group(kl_divs, n_bits_per_group, max_group_size_bits, strict=False)
    n_samples = int(2 ** n_bits_per_step)
    s = greedy_shard_scale(rho, n_steps)
    codes = []
    sample = np.zeros(num_dimensions, np.float32)
    for i in range(starts.size - 1):
        a, e = int(starts[i]), int(starts[i + 1])
        G = e - a
        gs = _u32(seed + i)
        draws = None
        if normal_source is not None:
            draws = np.stack([_draw(normal_source, 1000 * gs + j, 0, n_samples, G) for j in range(n_steps)])
        idx, best = greedy_encode_group(t_loc[a:e], t_scale[a:e], n_steps, n_samples, s, gs, draws)
        sample[a:e] = best
        codes.append(''.join(to_bit_string(int(k), n_bits_per_step) for k in idx))
    sample = (p_scale * sample + p_loc).astype(np.float32)
    return sample, ''.join(codes), [int(x) for x in starts]


def decode_grouped_greedy_sample(sess, bitcode, group_start_indices, proposal, n_bits_per_step, n_steps, seed,
                                 adaptive=True, rho=1., normal_source=None):
    """CGS:299-364."""
    _check_f32(proposal, "Proposal")
    n_bits_per_group = n_bits_per_step * n_steps
    p_loc, p_scale = _f32(proposal.loc), _f32(proposal.scale)
    num_dimensions = p_loc.size
    group_start_indices += [num_dimensions]  # CGS:323 mutates the caller's list
    if isinstance(bitcode, bytes):
        bitcode = bitcode.decode("utf-8")
    samples = []
    zero = np.zeros(1, np.float32)
    for i in range(len(group_start_indices) - 1):
        code = bitcode[n_bits_per_group * i: n_bits_per_group * (i + 1)]
        if code == '':
            break
        G = int(group_start_indices[i + 1]) - int(group_start_indices[i])
        samples.append(decode_greedy_sample(code, np.zeros(G, np.float32), np.ones(G, np.float32), n_bits_per_step,
                                            n_steps, _u32(seed + i), rho, normal_source))
    sample = np.concatenate(samples + [zero[:0]]).astype(np.float32)
    return (p_scale[:sample.size] * sample + p_loc[:sample.size]).astype(np.float32)
