"""ctypes binding of libcwq.so (include/cwq.h).  There is no fallback: if the library or a CUDA
device is missing, every compute call raises CwqError."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CWQ_LIB") or os.path.join(_HERE, "libcwq.so")   # CWQ_LIB: A/B builds of the same ABI (tools/variants)

ABI_VERSION = 3


class CwqError(RuntimeError):
    pass


class Segment(ctypes.Structure):
    """struct cwq_segment (include/cwq.h)."""
    _fields_ = [
        ("start", ctypes.c_int64),
        ("end", ctypes.c_int64),
        ("n_nats_per_group", ctypes.c_double),
        ("size_limit", ctypes.c_int64),
        ("dim_kl_bit_limit", ctypes.c_float),
        ("seed", ctypes.c_uint32),
        ("strict", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class ISBuffers(ctypes.Structure):
    """struct cwq_is_buffers (include/cwq.h): every output and the scratch of cwq_is_encode_pipeline."""
    _fields_ = [(name, ctypes.c_void_p) for name in
                ("kl", "coef", "mask", "gid", "sample", "outlier_idx", "outlier_q", "starts", "group_seed", "total_kl", "K",
                 "index", "bit_off", "seg_group_off", "seg_outlier_off", "bits")] + [
        ("bits_capacity_bytes", ctypes.c_size_t), ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("score_range_dev", ctypes.c_void_p), ("tile_segment", ctypes.c_void_p)]


STAGE_GROUP, STAGE_SCORE, STAGE_FINISH = 1, 2, 4

_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I32 = ctypes.c_int32
_U32 = ctypes.c_uint32
_SZ = ctypes.c_size_t
_F32 = ctypes.c_float

# name -> (restype, argtypes); mirrors include/cwq.h one to one (tests/test_abi.py checks it)
SIGNATURES = {
    "cwq_abi_version": (ctypes.c_int, []),
    "cwq_last_error": (ctypes.c_char_p, []),
    "cwq_device_info": (ctypes.c_int, [ctypes.c_int, _P, _P, _P, _P]),
    "cwq_launch_count": (ctypes.c_longlong, []),
    "cwq_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "cwq_profile_count": (ctypes.c_int, []),
    "cwq_profile_kind": (ctypes.c_int, [ctypes.c_int]),
    "cwq_profile_get_ms": (ctypes.c_int, [ctypes.c_int, _P]),
    "cwq_is_prologue": (ctypes.c_int, [_P, _P, _P, _P, _I64, _P, _I32, _P, _P, _P, _P]),
    "cwq_greedy_prologue": (ctypes.c_int, [_P, _P, _P, _P, _I64, _P, _P, _P]),
    "cwq_group_workspace_bytes": (_SZ, [_I64, _I32]),
    "cwq_group": (ctypes.c_int, [_P, _I64, _P, _P, _I32, _P, _P, _P, _P, _P, _SZ, _P]),
    "cwq_expand_groups": (ctypes.c_int, [_P, _I64, _I64, _P, _I32, _P, _P, _P, _P]),
    "cwq_group_totals": (ctypes.c_int, [_P, _P, _I64, _P, _P, _P, _P]),
    "cwq_is_encode_workspace_bytes": (_SZ, [_I64]),
    "cwq_is_encode": (ctypes.c_int, [_P, _P, _P, _P, _I64, _P, _I64, _I64, _I64, _P, _P, _SZ, _P]),
    "cwq_is_encode_device_range": (ctypes.c_int, [_P, _P, _P, _P, _I64, _P, _I64, _P, _P, _SZ, _P]),
    "cwq_shard_ranges_workspace_bytes": (_SZ, [_I64]),
    "cwq_is_shard_ranges": (ctypes.c_int, [_P, _P, _I64, _P, _I32, _P, _P, _SZ, _P]),
    "cwq_is_decode": (ctypes.c_int, [_P, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "cwq_is_pipeline_workspace_bytes": (_SZ, [_I64, _I32, _I64]),
    "cwq_is_pipeline_shard_ranges": (ctypes.c_int, [_I64, _I32, _I64, _I32, _P, _P, _P]),
    "cwq_set_fused_front": (ctypes.c_int, [ctypes.c_int]),
    "cwq_is_encode_pipeline": (ctypes.c_int, [_P, _P, _P, _P, _I64, _P, _P, _I32, _I64, _I32, _P, _P]),
    "cwq_outlier_workspace_bytes": (_SZ, [_I64]),
    "cwq_outlier_encode": (ctypes.c_int, [_P, _P, _P, _I64, _P, _I32, _P, _P, _P, _P, _P, _SZ, _P]),
    "cwq_outlier_decode": (ctypes.c_int, [_P, _P, _P, _I64, _P]),
    "cwq_greedy_workspace_bytes": (_SZ, [_I64, _I64, _I32]),
    "cwq_greedy_encode": (ctypes.c_int, [_P, _P, _P, _I64, _I64, _I64, _I32, _I32, _F32, _P, _I64, _I64, _I32, _P, _P, _P, _SZ, _P]),
    "cwq_greedy_decode": (ctypes.c_int, [_P, _P, _P, _P, _I64, _I32, _F32, _P, _P, _P, _P, _P]),
    "cwq_greedy_shard_loc": (ctypes.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "cwq_unrescale": (ctypes.c_int, [_P, _P, _P, _I64, _P, _P]),
    "cwq_pack_workspace_bytes": (_SZ, [_I64]),
    "cwq_elias_delta_pack": (ctypes.c_int, [_P, _I64, _P, _P, _SZ, _P, _P, _SZ, _P]),
    "cwq_fixed_pack": (ctypes.c_int, [_P, _I64, _I32, _P, _SZ, _P]),
    "cwq_fixed_unpack": (ctypes.c_int, [_P, _I64, _I32, _P, _P]),
    "cwq_bits_to_ascii": (ctypes.c_int, [_P, _I64, _P, _P, _P]),
    "cwq_elias_delta_parse_host": (ctypes.c_int, [_P, _I64, _I64, _P, _P]),
    "cwq_ac_encode_host": (ctypes.c_int, [_P, _I32, _P, _I64, _I32, _P, _I64, _P]),
    "cwq_ac_decode_host": (ctypes.c_int, [_P, _I32, _P, _I64, _I32, _I32, _P, _I64, _P, _P]),
    "cwq_draw_candidates": (ctypes.c_int, [_U32, _U32, _U32, _U32, _I32, _P, _P]),
    "cwq_stream_words": (ctypes.c_int, [_U32, _U32, _U32, _U32, _P, _P]),
    "cwq_philox4x32_10": (ctypes.c_int, [_P, _P, _P, _I64, _P]),
}

_lib = None


def load():
    """Load libcwq.so and bind every symbol of the ABI; raises CwqError when that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CwqError("libcwq.so not found at {} -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)".format(LIB_PATH))
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise CwqError("cannot load {}: {}".format(LIB_PATH, e))
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise CwqError("libcwq.so does not export {}".format(name))
        fn.restype = res
        fn.argtypes = args
    if lib.cwq_abi_version() != ABI_VERSION:
        raise CwqError("libcwq.so ABI {} != expected {}".format(lib.cwq_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().cwq_last_error()
        raise CwqError("{} failed ({}): {}".format(what, rc, msg.decode("utf-8", "replace") if msg else ""))


def call(name, *args):
    """Call an int-returning entry point and raise on a non-zero status."""
    check(getattr(load(), name)(*args), name)
