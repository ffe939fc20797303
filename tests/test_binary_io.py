"""The package's bit helpers (reference names) against goldens made by the reference's binary_io.py."""
import json
import os

import pytest

from compression_without_quantization_b200 import binary_io as bio

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "binary_io.json")


def test_elias_delta_matches_reference():
    g = json.load(open(GOLD))
    for x, code in g["elias_delta"]:
        assert bio.elias_delta_code(x) == code
        assert bio.elias_delta_decode((code + "0110").encode("utf-8")) == (x, len(code))
        assert bio.elias_delta_decode(code + "1") == (x, len(code))


def test_fixed_width_matches_reference():
    g = json.load(open(GOLD))
    for v, bits, s in g["to_bit_string"]:
        assert bio.to_bit_string(v, bits) == s
        assert bio.from_bit_string(s) == v and bio.from_bit_string(s.encode("utf-8")) == v
    with pytest.raises(Exception) as e:
        bio.to_bit_string(8, 3)
    assert str(e.value) == g["to_bit_string_overflow"]
