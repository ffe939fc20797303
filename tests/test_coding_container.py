"""Arithmetic coder (C++ in libcwq.so) and the .miracle container against goldens produced by the
reference's own coding.pyx / binary_io.py (tests/golden/gen_coding_goldens.py).  Host-only: no GPU needed."""
import json
import os

import numpy as np
import pytest

from compression_without_quantization_b200 import binary_io as bio
from compression_without_quantization_b200.coding import ArithmeticCoder

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "coding.npz"))


@pytest.mark.parametrize("case", range(6))
def test_arithmetic_coder_matches_reference_bit_for_bit(gold, case):
    P, msg = gold["ac%d_P" % case], gold["ac%d_msg" % case]
    want = gold["ac%d_code" % case].tobytes().decode()
    coder = ArithmeticCoder(P, precision=32)
    code = coder.encode(msg)
    assert isinstance(code, list) and ''.join(code) == want
    assert coder.decode_fast(want) == [int(x) for x in msg]
    assert coder.decode_fast(want + "0110") == [int(x) for x in msg]      # trailing payload bits are ignored
    if msg.size <= 4000:
        assert coder.decode(want) == [int(x) for x in msg]
    assert np.array_equal(coder.C, np.concatenate([[0], np.cumsum(P)[:-1]])) and coder.R == int(P.sum())


def test_arithmetic_coder_round_trip_group_sizes_alphabet():
    rng = np.random.default_rng(1)
    P = rng.integers(1, 5000, 17)          # 1 + 2**first_level_max_group_size_bits symbols (miracle.py:565-566)
    coder = ArithmeticCoder(P)
    for n in (1, 2, 50, 20000):
        msg = np.concatenate([rng.integers(1, 17, n - 1), [0]])
        assert coder.decode_fast(coder.encode(msg)) == [int(x) for x in msg]


def test_arithmetic_coder_skewed_histogram_below_one_bit_per_symbol():
    """A message with far more symbols than code bits (2000 x symbol 3 + end symbol -> 21 bits) decodes;
    the reference's coding.pyx has no limit on the symbols-per-bit ratio either."""
    coder = ArithmeticCoder(np.array([1, 1, 1, 1000000]))
    msg = [3] * 2000 + [0]
    code = coder.encode(msg)
    assert len(code) < 64
    assert coder.decode_fast(''.join(code)) == msg
    assert coder.decode(''.join(code)) == msg
    long_msg = [3] * 300000 + [1, 2, 0]
    assert coder.decode_fast(coder.encode(long_msg)) == long_msg


def test_arithmetic_coder_rejects_bad_input():
    with pytest.raises(Exception, match="positive count"):
        ArithmeticCoder(np.array([3, 0, 2])).encode([1, 0])
    with pytest.raises(Exception, match="outside the alphabet"):
        ArithmeticCoder(np.array([3, 1, 2])).encode([5, 0])


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_container_is_byte_identical_and_reads_back(tmp_path, name):
    meta = json.load(open(os.path.join(GOLD, "containers.json")))[name]
    path = str(tmp_path / "x.miracle")
    evb = None if meta["extra_var_bits"] is None else [list(x) for x in meta["extra_var_bits"]]
    bio.write_bin_code(meta["code"], path, extras=meta["extras"], extra_var_bits=evb,
                       var_length_extras=meta["var_length_extras"], var_length_bits=meta["var_length_bits"])
    want = open(os.path.join(GOLD, "container_%s.miracle" % name), "rb").read()
    assert open(path, "rb").read() == want
    back = bio.read_bin_code(path, num_extras=len(meta["extras"] or []), num_extra_var_bits=len(meta["extra_var_bits"] or []),
                             num_var_length_extras=len(meta["var_length_extras"] or []))
    ref = meta["read_back"]
    assert back[0] == ref[0] and back[1] == ref[1] and back[2] == ref[2] and back[3] == ref[3]
    assert back[0].startswith(meta["code"])


def test_container_wide_length_fields_extension(tmp_path):
    """more than 65 535 outliers / coded bits per blob (a 3840x2160 image).  The file written with
    extra_bytes=4 was read back by the REFERENCE's reader when the golden was made (containers.json 'wide'); here the
    engine rebuilds the same bytes, reads them back, and the default 16-bit format still raises like the reference."""
    import hashlib
    meta = json.load(open(os.path.join(GOLD, "containers.json")))["wide"]
    n = 70000
    blob = ''.join('1' if (i * i) % 7 < 3 else '0' for i in range(n + 1))
    idx = [(i * 37) % (1 << 24) for i in range(n)]
    val = [(i * 7919) % 65536 for i in range(n)]
    code = ''.join('1' if (i * 5) % 11 < 4 else '0' for i in range(1234))
    extras = [42, (1 << 31) | 30, 14]
    path = str(tmp_path / "wide.miracle")
    bio.write_bin_code(code, path, extras=extras, extra_var_bits=[list(blob), list(blob[:77])],
                       var_length_extras=[idx, val], var_length_bits=[24, 16], extra_bytes=4)
    raw = open(path, "rb").read()
    assert len(raw) == meta["file_bytes"] and hashlib.sha256(raw).hexdigest() == meta["file_sha256"] and meta["reference_reader_ok"]
    back = bio.read_bin_code(path, num_extras=3, num_extra_var_bits=2, num_var_length_extras=2, extra_bytes=4)
    assert back[1] == extras and back[2] == [blob, blob[:77]] and back[3] == [idx, val]
    assert back[0][:len(code)] == code and len(back[0]) == meta["payload_bits"]
    with pytest.raises(ValueError):        # the reference format: bytes([n // 256, ...]) overflows (binary_io.py:97,110)
        bio.write_bin_code(code, path, extras=extras, extra_var_bits=[list(blob)])
    with pytest.raises(ValueError):
        bio.write_bin_code(code, path, var_length_extras=[idx], var_length_bits=[24])


def test_container_errors_like_the_reference(tmp_path):
    with pytest.raises(Exception, match="Each var length extra needs to have a bitlength associated!"):
        bio.write_bin_code("01", str(tmp_path / "y"), var_length_extras=[[1]], var_length_bits=None)
    with pytest.raises(ValueError):   # 16-bit length field overflow (BIO:97)
        bio.write_bin_code("01", str(tmp_path / "z"), extra_var_bits=["0" * 70000])
