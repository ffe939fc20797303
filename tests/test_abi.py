"""The C-ABI boundary: libcwq.so loads without a GPU, exports every symbol include/cwq.h declares, the
ctypes table mirrors the header, host-only entry points work, and the product refuses to run without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cwq.h")


def header_symbols():
    text = open(HEADER).read()
    return re.findall(r"CWQ_API\s+[\w\s\*]+?\b(cwq_\w+)\s*\(", text)


def test_header_declares_the_expected_entry_points():
    names = header_symbols()
    assert len(names) == len(set(names)) >= 30
    for must in ("cwq_is_prologue", "cwq_group", "cwq_is_encode", "cwq_is_decode", "cwq_greedy_encode", "cwq_greedy_decode",
                 "cwq_elias_delta_pack", "cwq_fixed_pack", "cwq_draw_candidates", "cwq_philox4x32_10"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from compression_without_quantization_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.cwq_abi_version() == 3


def test_ctypes_table_mirrors_the_header():
    from compression_without_quantization_b200 import _lib
    assert sorted(_lib.SIGNATURES) == sorted(header_symbols())
    text = open(HEADER).read()
    for name, (_, args) in _lib.SIGNATURES.items():
        decl = re.search(r"CWQ_API[^;]*?\b" + name + r"\s*\(([^;]*?)\)\s*;", text, re.S).group(1)
        n_args = 0 if decl.strip() == "void" else decl.count(",") + 1
        assert n_args == len(args), (name, n_args, len(args))
    assert ctypes.sizeof(_lib.Segment) == 48
    _lib.load()


def test_host_elias_delta_parser_needs_no_gpu():
    from compression_without_quantization_b200 import binary_io as bio, engine as E
    values = [1, 2, 3, 17, 1000, 385751, 2 ** 20, 5]
    code = "".join(bio.elias_delta_code(v) for v in values)
    idx, used = E.elias_delta_parse(code + "010", len(values))
    assert list(idx) == [v - 1 for v in values] and used == len(code)
    with pytest.raises(Exception, match="malformed|truncated"):
        E.elias_delta_parse("0000000", 1)
    assert E.elias_delta_parse("", 0)[1] == 0


def test_bit_string_conversions_are_inverse():
    from compression_without_quantization_b200 import engine as E
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 2, 1003).astype(np.uint8)
    code = (bits + 48).tobytes()
    packed, n = E.ascii_to_bits(code)
    assert n == 1003 and E.bits_to_ascii(packed, n) == code
    assert E.bits_to_ascii(packed, 100, 37) == code[37:137]
    with pytest.raises(Exception):
        E.ascii_to_bits("0120")


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import compression_without_quantization_b200 as cwq
    from compression_without_quantization_b200.synthetic import Normal
    z, o = np.zeros(8, np.float32), np.ones(8, np.float32)
    with pytest.raises(cwq.CwqError, match="no CPU fallback"):
        cwq.code_grouped_importance_sample(None, Normal(z, o), Normal(z, o), 42, 20)
    with pytest.raises(cwq.CwqError, match="no CPU fallback"):
        cwq.code_grouped_greedy_sample(None, Normal(z, o), Normal(z, o), 2, 3, 42)
    with pytest.raises(cwq.CwqError):
        cwq.stateless_normal_sample(z, o, 4, 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "compression_without_quantization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for bad in ("import oracle", "from oracle", "libcwq_oracle", '#include "../../oracle', "cwqo_philox", "cwqo_is_"):
                    assert bad not in text, (f, bad)


def test_segment_table_and_size_limits():
    from compression_without_quantization_b200 import engine as E
    assert E.size_limit(4, True) == 16 and E.size_limit(2, True) == 4 and E.size_limit(12, False) == 4095
    assert E.size_limit(0, True) == 1 and E.size_limit(0, False) == 0
    assert E.is_float32(np.float32) and E.is_float32(np.dtype("float32")) and not E.is_float32(np.float64)
    assert abs(E.greedy_shard_scale(1.0, 30) - 1 / np.sqrt(30)) < 1e-7
