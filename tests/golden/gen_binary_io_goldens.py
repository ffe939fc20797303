"""Generates tests/golden/binary_io.json by importing the reference's own code/binary_io.py
(pure numpy, importable here; it cannot travel to the GPU box, the JSON does).

    python tests/golden/gen_binary_io_goldens.py      # needs /root/reference
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/code")
import binary_io as ref  # noqa: E402  (the reference module, unmodified)

rng = np.random.default_rng(20261018)
xs = sorted(set([1, 2, 3, 4, 7, 8, 15, 16, 17, 31, 32, 255, 256, 1000, 65535, 65536, 385750, 385751, 2 ** 20 - 1, 2 ** 20,
                 2 ** 21] + [int(v) for v in rng.integers(1, 2 ** 21, 400)]))
elias = []
for x in xs:
    code = ref.elias_delta_code(np.int64(x))
    num, length = ref.elias_delta_decode((code + "0110").encode("utf-8"))
    assert num == x and length == len(code)
    elias.append([x, code])
fixed = []
for bits in (1, 3, 8, 14, 20):
    for v in [0, 1, 2 ** bits - 1] + [int(t) for t in rng.integers(0, 2 ** bits, 20)]:
        s = ref.to_bit_string(v, bits)
        assert ref.from_bit_string(s) == v and ref.from_bit_string(s.encode("utf-8")) == v
        fixed.append([v, bits, s])
out = {"source": "/root/reference/code/binary_io.py (elias_delta_code, elias_delta_decode, to_bit_string, from_bit_string)",
       "elias_delta": elias, "to_bit_string": fixed,
       "to_bit_string_overflow": "The number 8 (>= 8) is bigger than what we can encode!"}
try:
    ref.to_bit_string(8, 3)
    raise SystemExit("expected an exception")
except Exception as e:  # the reference raises a plain Exception (BIO:43-44)
    assert str(e) == out["to_bit_string_overflow"]
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "binary_io.json"), "w") as f:
    json.dump(out, f)
print("wrote", len(elias), "elias and", len(fixed), "fixed-width vectors")
