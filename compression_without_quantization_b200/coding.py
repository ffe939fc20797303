"""Arithmetic coder with the reference's interface (reference code/coding.pyx:27-310), backed by the
C++ implementation in libcwq.so (csrc/cwq_coding.cpp).

    coder = ArithmeticCoder(counts, precision=32)
    code = coder.encode(message)            # list of '0'/'1' characters, message must end with symbol 0
    message = coder.decode_fast(''.join(code))

pln.py uses it for the group sizes of both levels (pln.py:517-529, 708-709, 736-740)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import CwqError


class ArithmeticCoder(object):
    def __init__(self, P, precision=32):
        self._P = np.ascontiguousarray(np.asarray(P), dtype=np.int64)
        if self._P.ndim != 1 or self._P.size == 0:
            raise CwqError("P must be a non-empty 1-D array of counts")
        if np.any(np.asarray(P) != self._P):
            raise CwqError("counts must be integers")
        self._precision = int(precision)
        self.C = np.concatenate([[0], np.cumsum(self._P)[:-1]]).astype(np.int64)   # coding.pyx:42-48
        self.D = np.cumsum(self._P).astype(np.int64)
        self.R = int(self.D[-1])

    def _ptr(self, a):
        return ctypes.c_void_p(a.ctypes.data)

    def encode(self, message):
        msg = np.ascontiguousarray(np.asarray(message), dtype=np.int64).reshape(-1)
        cap = (msg.size + 2) * (self._precision + 2)
        out = np.empty(cap, dtype=np.uint8)
        n_bits = ctypes.c_int64(0)
        _lib.call("cwq_ac_encode_host", self._ptr(self._P), int(self._P.size), self._ptr(msg) if msg.size else ctypes.c_void_p(0),
                  int(msg.size), self._precision, self._ptr(out), int(cap), ctypes.cast(ctypes.pointer(n_bits), ctypes.c_void_p))
        return list(out[:n_bits.value].tobytes().decode("ascii"))

    def _decode(self, code, fast):
        if isinstance(code, (list, tuple)):
            code = ''.join(code)
        if isinstance(code, str):
            code = code.encode("ascii")
        buf = np.frombuffer(code, dtype=np.uint8)
        cap = max(16, buf.size + 16)
        while True:
            out = np.empty(cap, dtype=np.int64)
            n = ctypes.c_int64(0)
            used = ctypes.c_int64(0)
            rc = _lib.load().cwq_ac_decode_host(self._ptr(self._P), int(self._P.size),
                                                self._ptr(buf) if buf.size else ctypes.c_void_p(0), int(buf.size),
                                                self._precision, 1 if fast else 0, self._ptr(out), int(cap),
                                                ctypes.cast(ctypes.pointer(n), ctypes.c_void_p),
                                                ctypes.cast(ctypes.pointer(used), ctypes.c_void_p))
            if rc == -3 and cap < (1 << 31):   # CWQ_ERR_WORKSPACE: message longer than guessed; n holds its true length
                cap = max(int(n.value), 2 * cap)
                continue
            _lib.check(rc, "cwq_ac_decode_host")
            self.bits_used = int(used.value)
            return [int(v) for v in out[:n.value]]

    def decode(self, code):
        return self._decode(code, fast=False)

    def decode_fast(self, code, verbose=False):
        msg = self._decode(code, fast=True)
        if verbose:
            print("Remaining bits at end of code: {}".format(len(code) - self.bits_used))
        return msg
