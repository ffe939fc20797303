"""Generates tests/golden/coding.npz and tests/golden/container_*.miracle from the reference's own
code/coding.pyx, code/data_structures.py and code/binary_io.py.

binary_io.py and data_structures.py are imported as they are.  coding.pyx is Cython and does not build
under the Cython 3.3 of this image; its ArithmeticCoder is executed here as
plain Python after a MECHANICAL removal of the Cython type annotations (the regexes below; no logic is
touched) -- the integer arithmetic is the same in Python ints as in C longs for every value reached
(width <= 2^32, counts < 2^31).

    python tests/golden/gen_coding_goldens.py      # needs /root/reference
"""
import io
import os
import re
import sys
from contextlib import redirect_stderr, redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/code")
import binary_io as ref_bio  # noqa: E402
import tqdm as _tqdm  # noqa: E402

_tqdm.tqdm = lambda it, *a, **k: it

src = open("/root/reference/code/coding.pyx").read()
src = src.replace("cimport numpy as np", "")
src = re.sub(r"^ctypedef .*$", "", src, flags=re.M)
src = src.replace("np.ndarray P", "P")
src = re.sub(r"^(\s*)cdef\s+[\w\.]+\s+(\w+\s*=.*)$", r"\1\2", src, flags=re.M)     # cdef T x = v  ->  x = v
src = re.sub(r"^\s*cdef\s+[\w\.]+\s+[\w, ]+$", "", src, flags=re.M)                # cdef T a, b    ->  (dropped)
ns = {}
with redirect_stdout(io.StringIO()), redirect_stderr(io.StringIO()):
    exec(compile(src, "coding.pyx (annotations stripped)", "exec"), ns)
ArithmeticCoder = ns["ArithmeticCoder"]

rng = np.random.default_rng(20261018)
out = {}
cases = []
for ci, (n_sym, n_msg, scale) in enumerate([(17, 3000, 1000), (5, 500, 50), (2, 200, 3), (257, 4000, 20000), (17, 1, 10), (65, 12000, 100000)]):
    P = rng.integers(1, scale + 1, n_sym).astype(np.int64)
    p = P[1:] / P[1:].sum() if n_sym > 1 else None
    msg = rng.choice(np.arange(1, n_sym), size=n_msg - 1, p=p) if n_msg > 1 else np.zeros(0, np.int64)
    msg = np.concatenate([msg, [0]]).astype(np.int64)      # symbol 0 terminates (pln.py:521-525)
    with redirect_stdout(io.StringIO()):
        coder = ArithmeticCoder(P, precision=32)
        code = coder.encode(msg)
        dec_fast = coder.decode_fast(''.join(code))
        dec_slow = coder.decode(''.join(code)) if n_msg <= 3000 else dec_fast
    assert list(dec_fast) == list(msg) and list(dec_slow) == list(msg)
    out["ac%d_P" % ci] = P
    out["ac%d_msg" % ci] = msg
    out["ac%d_code" % ci] = np.frombuffer(''.join(code).encode(), np.uint8)
    cases.append(ci)
    print("ac case", ci, "symbols", n_sym, "message", n_msg, "code bits", len(code))
out["ac_cases"] = np.array(cases)
np.savez_compressed(os.path.join(HERE, "coding.npz"), **out)

# ---- .miracle container (binary_io.py:69-197), written by the reference itself ----
def bits(n):
    return ''.join(rng.choice(["0", "1"], n))

containers = {
    "a": dict(code=bits(4321), extras=[42, 30, 14, 20, 20, 3000, 1321, 32, 48, 8, 12], extra_var_bits=[list(bits(777)), list(bits(64))],
              var_length_extras=[[5, 70000, 16777215], [1, 65535, 32768], [], []], var_length_bits=[24, 16, 24, 16]),
    "b": dict(code=bits(8), extras=[1, 2, 3], extra_var_bits=None, var_length_extras=None, var_length_bits=None),
    "c": dict(code=bits(1000), extras=None, extra_var_bits=[list(bits(5))], var_length_extras=[[3, 1, 2]], var_length_bits=[2]),
}
meta = {}
for name, c in containers.items():
    path = os.path.join(HERE, "container_%s.miracle" % name)
    evb = None if c["extra_var_bits"] is None else [list(x) for x in c["extra_var_bits"]]
    ref_bio.write_bin_code(c["code"], path, extras=c["extras"], extra_var_bits=evb,
                           var_length_extras=c["var_length_extras"], var_length_bits=c["var_length_bits"])
    back = ref_bio.read_bin_code(path, num_extras=len(c["extras"] or []), num_extra_var_bits=len(c["extra_var_bits"] or []),
                                 num_var_length_extras=len(c["var_length_extras"] or []))
    meta[name] = {"code": c["code"], "extras": c["extras"],
                  "extra_var_bits": None if c["extra_var_bits"] is None else [''.join(x) for x in c["extra_var_bits"]],
                  "var_length_extras": c["var_length_extras"], "var_length_bits": c["var_length_bits"],
                  "read_back": [back[0], [int(x) for x in back[1]], list(back[2]), [[int(v) for v in x] for x in back[3]]]}
    print("container", name, os.path.getsize(path), "bytes")

# ---- the 32-bit length-field extension: written by the ENGINE's writer, read by the REFERENCE's reader,
# whose `extra_bytes` parameter (binary_io.py:136) already allows it.  Content is formula-generated (tests rebuild it).
import hashlib  # noqa: E402
import json  # noqa: E402
import tempfile  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from compression_without_quantization_b200 import binary_io as our_bio  # noqa: E402


def wide_case():
    n = 70000
    blob = ''.join('1' if (i * i) % 7 < 3 else '0' for i in range(n + 1))
    idx = [(i * 37) % (1 << 24) for i in range(n)]
    val = [(i * 7919) % 65536 for i in range(n)]
    code = ''.join('1' if (i * 5) % 11 < 4 else '0' for i in range(1234))
    return dict(code=code, extras=[42, (1 << 31) | 30, 14], extra_var_bits=[list(blob), list(blob[:77])],
                var_length_extras=[idx, val], var_length_bits=[24, 16])


w = wide_case()
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "wide.miracle")
    our_bio.write_bin_code(w["code"], path, extras=w["extras"], extra_var_bits=[list(x) for x in w["extra_var_bits"]],
                           var_length_extras=w["var_length_extras"], var_length_bits=w["var_length_bits"], extra_bytes=4)
    raw = open(path, "rb").read()
    back = ref_bio.read_bin_code(path, num_extras=3, num_extra_var_bits=2, num_var_length_extras=2, extra_bytes=4)
assert back[1] == w["extras"] and back[2] == [''.join(x) for x in w["extra_var_bits"]] and [list(x) for x in back[3]] == w["var_length_extras"]
assert back[0][:len(w["code"])] == w["code"]
meta["wide"] = {"file_sha256": hashlib.sha256(raw).hexdigest(), "file_bytes": len(raw),
                "reference_reader_ok": True, "payload_bits": len(back[0])}
print("container wide", len(raw), "bytes; the reference's reader (extra_bytes=4) returns what was written")
json.dump(meta, open(os.path.join(HERE, "containers.json"), "w"))
