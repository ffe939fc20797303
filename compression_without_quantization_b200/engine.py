"""Host-side engine: sequences the libcwq.so stages for one or many reference calls ("segments").

PyTorch is used only for device memory, streams and (in distributed.py) torch.distributed; all
arithmetic happens in the CUDA kernels behind the C ABI (include/cwq.h).  The stage order follows
the reference's grouped functions:

    importance sampler (CIS:112-274)   prologue -> group -> totals -> encode -> decode -> outliers -> pack
    greedy sampler     (CGS:170-296)   prologue -> group -> encode -> unrescale -> pack
"""
import collections
import ctypes
import functools
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import CwqError, Segment


# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------
def cuda_device():
    if not torch.cuda.is_available():
        raise CwqError("no CUDA device visible: the engine has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def current_stream_ptr():
    """cudaStream_t of torch's current stream as an int.  torch.cuda.current_stream() costs ~15 us in Python (it builds a
    Stream object); the raw query is ~1 us, and a small encode makes a dozen of these calls."""
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:
        return torch.cuda.current_stream().cuda_stream


def _stream():
    return ctypes.c_void_p(current_stream_ptr())


def is_float32(dtype):
    """True for numpy / torch / TF-style float32 dtype objects (the reference tests `is tf.float32`)."""
    if dtype is None:
        return False
    if dtype is torch.float32:
        return True
    try:
        return np.dtype(dtype) == np.float32
    except TypeError:
        pass
    name = getattr(dtype, "name", None) or str(dtype)
    return name in ("float32", "<dtype: 'float32'>", "torch.float32")


def as_device_f32(x, device=None, what="array"):
    """Flat contiguous float32 CUDA tensor from numpy / torch / any DLPack producer (zero-copy when the
    data already lives on the device)."""
    device = device or cuda_device()
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        if x.dtype != np.float32:
            raise Exception("{} datatype must be float32!".format(what))
        t = torch.from_numpy(np.ascontiguousarray(x))
    elif hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
    elif hasattr(x, "numpy"):
        t = torch.from_numpy(np.ascontiguousarray(x.numpy()))
    else:
        a = np.asarray(x)
        if a.dtype != np.float32:
            raise Exception("{} datatype must be float32!".format(what))
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype != torch.float32:
        raise Exception("{} datatype must be float32!".format(what))
    return t.to(device, non_blocking=True).reshape(-1).contiguous()


def _adjacent_span(ts):
    """If the contiguous CPU tensors `ts` are consecutive pieces of ONE tensor (views of a common base, back to back), a flat
    view over all of them; else None.  Uses only cheap attribute reads: this runs per chunk in front of the first upload."""
    first = ts[0]
    base = first._base if first._base is not None else first
    pos = first.storage_offset()
    for t in ts:
        if (t._base if t._base is not None else t) is not base or t.storage_offset() != pos or not t.is_contiguous():
            return None
        pos += t.numel()
    return torch.as_strided(first.reshape(-1), (pos - first.storage_offset(),), (1,))


def host_span(arrays):
    """One flat contiguous float32 CPU tensor holding the concatenation of `arrays` (CPU tensors / numpy): a view when they
    are consecutive pieces of one buffer (a batch sliced into per-call pieces), else a copy."""
    ts = [a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a)) for a in arrays]
    if len(ts) == 1:
        return ts[0].reshape(-1).contiguous()
    span = _adjacent_span(ts) if all(t.dim() == 1 for t in ts) else None
    return span if span is not None else torch.cat([t.reshape(-1) for t in ts])


def concat_to_device(arrays, device=None, what="array"):
    """Concatenation of flat float32 arrays on the device.  When the inputs are consecutive views of one host
    buffer (a batch sliced into per-call pieces, as bench.py and pln-style callers produce) the whole span
    goes over in ONE copy instead of one per piece."""
    device = device or cuda_device()
    if len(arrays) > 1 and all(isinstance(a, torch.Tensor) and a.device.type == "cpu" and a.dtype == torch.float32
                               and a.dim() == 1 for a in arrays):
        span = _adjacent_span(arrays)
        if span is not None:
            return span.to(device, non_blocking=True)
    parts = [as_device_f32(a, device, what) for a in arrays]
    return parts[0] if len(parts) == 1 else torch.cat(parts)


@functools.lru_cache(maxsize=256)
def size_limit(max_group_size_bits, strict):
    """Running group size at which the reference's size test fires: CIS:185-187 evaluates
    ``np.log(size + 1) / np.log(2) > max_group_size_bits`` (CGS:230-232: ``>=``); the same expression is
    evaluated here around 2**bits so the kernels can compare integers."""
    m = max_group_size_bits
    guess = int(np.floor(2.0 ** float(m)))
    for size in range(max(0, guess - 3), guess + 4):
        bits = np.log(size + 1) / np.log(2)
        if (bits > m) if strict else (bits >= m):
            return size
    raise CwqError("cannot bracket the group size limit for max_group_size_bits={}".format(m))


@dataclass(frozen=True)
class SegmentSpec:
    """One reference call: a flat latent vector of n dims with its own parameters."""
    n: int
    seed: int
    n_bits_per_group: float
    max_group_size_bits: float
    dim_kl_bit_limit: float = float("inf")
    strict: bool = True  # True: importance sampler ('>'), False: greedy ('>=')


_table_cache = collections.OrderedDict()   # (device, table bytes) -> device tensor; tables are read-only on the device


def _segment_table_on_device(raw, device):
    """Device copy of a segment table.  Repeated calls with the same parameters (a codec encoding image after image, a
    benchmark loop) reuse one upload; a new table goes through pinned memory with an asynchronous copy -- a pageable
    upload would block the host until everything already queued on the stream has run, which serialises the chunks of
    the batch pipeline."""
    key = (device.type, device.index, current_stream_ptr() if device.type == "cuda" else 0, raw)
    hit = _table_cache.get(key)
    if hit is not None:
        _table_cache.move_to_end(key)
        return hit
    staging = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    if device.type == "cuda":
        staging = staging.pin_memory()
    dev = staging.to(device, non_blocking=True)
    _table_cache[key] = dev
    while len(_table_cache) > 64:
        _table_cache.popitem(last=False)
    return dev


_buffer_pool = {}      # (n, n_segments, workspace bytes, device, stream) -> released (tensors, cwq_is_buffers) of ISJob.launch
_segment_tables = collections.OrderedDict()   # (specs, device, stream) -> SegmentTable: a codec / benchmark repeats its batches


def segment_table(specs, device):
    """SegmentTable of a tuple of specs, built once per (specs, device, stream): constructing one costs ~10 us per call of
    the batch in Python, which is what the GPU waits for at the start of a small encode."""
    dev = torch.device(device)
    key = (tuple(specs), dev.type, dev.index, current_stream_ptr() if dev.type == "cuda" else 0)
    hit = _segment_tables.get(key)
    if hit is not None:
        _segment_tables.move_to_end(key)
        return hit
    table = SegmentTable(key[0], dev)
    _segment_tables[key] = table
    while len(_segment_tables) > 64:
        _segment_tables.popitem(last=False)
    return table


class SegmentTable:
    def __init__(self, specs, device):
        self.specs = list(specs)
        self.count = len(self.specs)
        if self.count == 0:
            raise CwqError("at least one segment is required")
        self.host = (Segment * self.count)()
        off = 0
        for i, sp in enumerate(self.specs):
            if sp.n <= 0:
                raise CwqError("segments must be non-empty")
            n_nats = sp.n_bits_per_group * np.log(2) - 1  # CIS:181 / CGS:226, float64
            self.host[i] = Segment(off, off + sp.n, float(n_nats), size_limit(sp.max_group_size_bits, sp.strict),
                                   float(sp.dim_kl_bit_limit), int(sp.seed) & 0xFFFFFFFF, 1 if sp.strict else 0, 0)
            off += sp.n
        self.n = off
        self.starts = [int(s.start) for s in self.host]
        # behind the table, in the same upload: the segment of every dim 1024 k (cwq_is_buffers.tile_segment), which lets
        # the fused kernels find a dim's segment with one load instead of a binary search at kernel start
        seg_bytes = bytes(self.host)
        pad = (-len(seg_bytes)) % 16
        tile_seg = (np.searchsorted(np.asarray(self.starts, dtype=np.int64), np.arange(0, off, 1024, dtype=np.int64),
                                    side="right") - 1).astype(np.int32)
        self.dev = _segment_table_on_device(seg_bytes + b"\0" * pad + tile_seg.tobytes(), torch.device(device))
        self.tile_segment_ptr = self.dev.data_ptr() + len(seg_bytes) + pad
        self.max_size_limit = max(int(s.size_limit) for s in self.host)


class _Groups:
    """Group tables on the device.  After cwq_group the group count is known only on the device
    (`n_groups_dev`); `sync()` brings it and the per-segment offsets to the host and trims the views."""

    def __init__(self, starts, group_seed, gid, seg_group_off, seg_group_off_host=None, capacity=None):
        self.starts = starts
        self.group_seed = group_seed
        self.gid = gid
        self.seg_group_off = seg_group_off
        self.seg_group_off_host = seg_group_off_host
        self.capacity = capacity if capacity is not None else int(group_seed.numel())
        self.n_groups_dev = seg_group_off[-1:]          # int64 scalar on the device
        self._n_groups = None if seg_group_off_host is None else int(seg_group_off_host[-1])

    @property
    def synced(self):
        return self._n_groups is not None

    @property
    def n_groups(self):
        if self._n_groups is None:
            self.sync()
        return self._n_groups

    def sync(self):
        if self._n_groups is None:
            self.set_host(self.seg_group_off.cpu().numpy())
        return self

    def fetch_async(self, side):
        """Start copying the per-segment group offsets (and with them the group count) to the host on stream `side`,
        behind everything queued so far on the current stream; finish_fetch() waits for that copy only."""
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        host = torch.empty(self.seg_group_off.shape, dtype=self.seg_group_off.dtype, device="cpu", pin_memory=True)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            host.copy_(self.seg_group_off, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(side)
        self._fetch = (host, copied)

    def finish_fetch(self):
        host, copied = self._fetch
        copied.synchronize()
        return self.set_host(host.numpy().copy())

    def set_host(self, seg_group_off_host):
        """Install the host copy of the per-segment group offsets (however it was brought over) and trim the views."""
        self.seg_group_off_host = np.asarray(seg_group_off_host, dtype=np.int64)
        self._n_groups = ng = int(self.seg_group_off_host[-1])
        self.starts = self.starts[:ng + 1]
        self.group_seed = self.group_seed[:ng]
        return self


# above this many dims the capacity-sized scratch of the synchronisation-free pipeline is not worth its memory
ASYNC_MAX_DIMS = 1 << 25


def _run_group(kl, table, sync=True):
    dev = kl.device
    n, ns = table.n, table.count
    starts = torch.empty(n + ns + 1, dtype=torch.int64, device=dev)
    group_seed = torch.empty(n + ns, dtype=torch.int32, device=dev)
    gid = torch.empty(n, dtype=torch.int32, device=dev)
    seg_off = torch.empty(ns + 1, dtype=torch.int64, device=dev)
    wsb = _lib.load().cwq_group_workspace_bytes(n, ns)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    _lib.call("cwq_group", _ptr(kl), n, _ptr(table.dev), ctypes.cast(table.host, ctypes.c_void_p), ns, _ptr(starts),
              _ptr(group_seed), _ptr(gid), _ptr(seg_off), _ptr(ws), wsb, _stream())
    g = _Groups(starts, group_seed, gid, seg_off, None, capacity=n + ns)
    return g.sync() if sync else g


def _forced_groups(starts_host_per_segment, table, device):
    """Group tables from caller-supplied starts (decoder side / per-group entry points)."""
    flat = []
    seg_off = [0]
    for s, st in enumerate(starts_host_per_segment):
        if isinstance(st, list):
            st = np.fromiter(st, dtype=np.int64, count=len(st))     # the reference hands Python lists over (pln.py:735-741)
        else:
            st = np.asarray(st, dtype=np.int64).reshape(-1)
        if st.size < 2 or st[0] != 0 or st[-1] != table.specs[s].n or np.any(np.diff(st) < 0):
            raise CwqError("group_start_indices must start at 0, end at the number of dimensions and be sorted")
        flat.append(st[:-1] + table.starts[s])
        seg_off.append(seg_off[-1] + st.size - 1)
    flat.append(np.array([table.n], dtype=np.int64))
    starts = torch.from_numpy(np.concatenate(flat)).to(device)
    seg_off_host = np.asarray(seg_off, dtype=np.int64)
    seg_off_dev = torch.from_numpy(seg_off_host).to(device)
    ng = int(seg_off_host[-1])
    gid = torch.empty(table.n, dtype=torch.int32, device=device)
    group_seed = torch.empty(ng, dtype=torch.int32, device=device)
    _lib.call("cwq_expand_groups", _ptr(starts), ng, table.n, _ptr(table.dev), table.count, _ptr(seg_off_dev), _ptr(gid),
              _ptr(group_seed), _stream())
    return _Groups(starts, group_seed, gid, seg_off_dev, seg_off_host, capacity=ng)


def to_host(tensors):
    """Device tensors -> numpy arrays through pinned staging buffers (torch's caching host allocator
    recycles them), all copies asynchronous on the current stream, ONE synchronisation."""
    staged = []
    for t in tensors:
        h = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
        h.copy_(t, non_blocking=True)
        staged.append(h)
    torch.cuda.current_stream().synchronize()
    return [h.numpy() for h in staged]


def bits_to_ascii(bits_host, n_bits, bit_begin=0):
    """Packed MSB-first bits -> the reference's '0'/'1' byte string (BIO:76-79 in reverse)."""
    if n_bits == 0:
        return b""
    b0 = bit_begin // 8
    b1 = (bit_begin + n_bits + 7) // 8
    u = np.unpackbits(bits_host[b0:b1])
    lo = bit_begin - 8 * b0
    return (u[lo:lo + n_bits] + np.uint8(48)).tobytes()


def device_bits_to_ascii(bits, n_bits, n_bits_dev=None):
    """Packed MSB-first bits on the device -> '0'/'1' characters on the device (uint8[n_bits]); the expansion runs at
    HBM speed instead of costing the host an unpackbits pass per call.  With n_bits_dev (int64 on the device) the
    count is read there and n_bits is the capacity of the result."""
    n_bits = int(n_bits)
    out = torch.empty(8 * ((n_bits + 7) // 8), dtype=torch.uint8, device=bits.device)
    if n_bits:
        _lib.call("cwq_bits_to_ascii", _ptr(bits), n_bits, _ptr(n_bits_dev), _ptr(out), _stream())
    return out[:n_bits]


def ascii_to_bits(code):
    """'0'/'1' str or bytes -> (packed MSB-first uint8 array, number of bits)."""
    if isinstance(code, str):
        code = code.encode("utf-8")
    a = np.frombuffer(code, dtype=np.uint8)
    if a.size and (a.min() < 48 or a.max() > 49):
        raise CwqError("bitcode must consist of '0' and '1' characters")
    return np.packbits(a - np.uint8(48)), a.size


# ----------------------------------------------------------------------------------------------
# importance sampler
# ----------------------------------------------------------------------------------------------
class ISJob:
    """Stages of the grouped importance sampler over a batch of segments (device-resident)."""

    def __init__(self, q_loc, q_scale, p_loc, p_scale, specs, forced_starts=None):
        self.dev = q_loc.device
        self.q_loc, self.q_scale, self.p_loc, self.p_scale = q_loc, q_scale, p_loc, p_scale
        self.table = segment_table(specs, self.dev)
        self.n = self.table.n
        if not (q_loc.numel() == q_scale.numel() == p_loc.numel() == p_scale.numel() == self.n):
            raise CwqError("loc/scale arrays must all have {} elements".format(self.n))
        self.forced_starts = forced_starts
        self.groups = None

    def prepare(self, sync=True):
        """Prologue, grouping, candidate counts.  sync=False keeps the group count on the device so that
        encode()/decode()/pack() are enqueued without waiting for the grouping; per-group arrays then have
        capacity N + n_segments until sync_groups() trims them (used by the batch API for N <= 2^25)."""
        n, t = self.n, self.table
        if n > ASYNC_MAX_DIMS:
            sync = True
        self.kl = torch.empty(n, dtype=torch.float32, device=self.dev)
        self.coef = torch.empty(2 * n, dtype=torch.float32, device=self.dev)
        self.mask = torch.empty(n, dtype=torch.uint8, device=self.dev)
        _lib.call("cwq_is_prologue", _ptr(self.q_loc), _ptr(self.q_scale), _ptr(self.p_loc), _ptr(self.p_scale), n,
                  _ptr(t.dev), t.count, _ptr(self.kl), _ptr(self.coef), _ptr(self.mask), _stream())
        if self.forced_starts is None:
            self.groups = _run_group(self.kl, t, sync=sync)
        else:
            self.groups = _forced_groups(self.forced_starts, t, self.dev)
        cap = self.groups.n_groups if self.groups.synced else self.groups.capacity
        self._cap = cap
        self.total_kl = torch.empty(cap, dtype=torch.float32, device=self.dev)
        self.K = torch.empty(cap, dtype=torch.int32, device=self.dev)
        _lib.call("cwq_group_totals", _ptr(self.kl), _ptr(self.groups.starts), cap, self._ng_dev(), _ptr(self.total_kl),
                  _ptr(self.K), _stream())
        self.index = torch.zeros(cap, dtype=torch.int32, device=self.dev)  # uint32 payload
        return self

    def _ng_dev(self):
        return ctypes.c_void_p(0) if self.groups.synced else _ptr(self.groups.n_groups_dev)

    # ---- the same stages through ONE C call (cwq_is_encode_pipeline) ----
    def _alloc_buffers(self):
        """Every output and the shared scratch of cwq_is_encode_pipeline.  One tensor per array on purpose: a single
        ~80 B/dim arena was measured to make torch's caching allocator split and re-allocate its large blocks every few
        steps (sporadic +10 ms on a 113 ms step); same-sized small blocks recycle cleanly."""
        n, ns = self.n, self.table.count
        cap = n + ns
        bits_cap = ((cap * 42 + 31) // 32) * 4
        wsb = _lib.load().cwq_is_pipeline_workspace_bytes(n, ns, max(self.table.max_size_limit, 1))
        f32, i32, i64, u8, i16 = torch.float32, torch.int32, torch.int64, torch.uint8, torch.int16
        layout = [("kl", f32, n), ("coef", f32, 2 * n), ("mask", u8, n), ("gid", i32, n), ("sample", f32, n),
                  ("outlier_idx", i64, n), ("outlier_q", i16, n), ("starts", i64, cap + 1), ("group_seed", i32, cap),
                  ("total_kl", f32, cap), ("K", i32, cap), ("index", i32, cap), ("bit_off", i64, cap + 1),
                  ("seg_group_off", i64, ns + 1), ("seg_outlier_off", i64, ns + 1), ("bits", u8, bits_cap), ("workspace", u8, wsb)]
        # a codec / benchmark repeats its shapes: the tensors and the filled-in cwq_is_buffers of a released job are reused
        # (17 allocations and as many pointer look-ups are ~0.1 ms of Python in front of the first kernel of every chunk)
        self._pool_key = (n, ns, int(wsb), self.dev.index, current_stream_ptr())
        pooled = _buffer_pool.get(self._pool_key)
        if pooled:
            views, buf = pooled.pop()
        else:
            views = {name: torch.empty(max(count, 1), dtype=dt, device=self.dev) for name, dt, count in layout}
            buf = _lib.ISBuffers()
            for name in views:
                setattr(buf, name, views[name].data_ptr())
            buf.bits_capacity_bytes = bits_cap
            buf.workspace_bytes = wsb
        self._storage = views          # keeps the scratch alive as long as the job
        buf.score_range_dev = None
        buf.tile_segment = self.table.tile_segment_ptr
        self._buffers = buf
        for name in ("kl", "coef", "mask", "total_kl", "K", "index", "sample", "outlier_idx", "outlier_q", "seg_outlier_off",
                     "bits", "bit_off"):
            setattr(self, name, views[name])
        self.groups = _Groups(views["starts"], views["group_seed"], views["gid"], views["seg_group_off"], None, capacity=cap)
        self._cap = cap

    def release(self):
        """Hand the job's device buffers back for reuse by a later job of the same shape ON THE SAME STREAM (stream order
        then protects them).  Call only when nothing on another stream still reads them: the batch API does after its
        downloads have been synchronised.  The job must not be used afterwards."""
        if getattr(self, "_storage", None) is None or not hasattr(self, "_pool_key"):
            return
        slot = _buffer_pool.setdefault(self._pool_key, [])
        if len(slot) < 3 and len(_buffer_pool) <= 16:
            slot.append((self._storage, self._buffers))
        self._storage = None
        for name in ("kl", "coef", "mask", "total_kl", "K", "index", "sample", "outlier_idx", "outlier_q", "seg_outlier_off", "bits",
                     "bit_off", "groups", "ascii"):
            if hasattr(self, name):
                delattr(self, name)

    def launch(self, stages):
        """Enqueue the requested stages (bit set of _lib.STAGE_GROUP / STAGE_SCORE / STAGE_FINISH) with one call into
        the library; the group count stays on the device.  Results are the same tensors the stage-by-stage methods
        produce (prepare / encode / decode / pack), with capacity N + n_segments until sync_groups() trims them."""
        if self.forced_starts is not None:
            raise CwqError("launch() runs the grouping itself; use prepare()/encode()/decode() with forced group starts")
        if self.n > ASYNC_MAX_DIMS:
            raise CwqError("launch() keeps capacity-sized group tables; above {} dims use the stage methods".format(ASYNC_MAX_DIMS))
        if not hasattr(self, "_buffers"):
            self._alloc_buffers()
        t = self.table
        _lib.call("cwq_is_encode_pipeline", _ptr(self.q_loc), _ptr(self.q_scale), _ptr(self.p_loc), _ptr(self.p_scale), self.n,
                  _ptr(t.dev), ctypes.cast(t.host, ctypes.c_void_p), t.count, max(t.max_size_limit, 1), int(stages),
                  ctypes.cast(ctypes.pointer(self._buffers), ctypes.c_void_p), _stream())
        return self

    def sync_groups(self):
        """Make the group count known on the host (one device->host copy) and trim the per-group arrays."""
        if not self.groups.synced:
            self.groups.sync()
        ng = self.groups.n_groups
        for name in ("total_kl", "K", "index"):
            if hasattr(self, name) and getattr(self, name).numel() != ng:
                setattr(self, name, getattr(self, name)[:ng])
        if hasattr(self, "bit_off") and self.bit_off.numel() != ng + 1:
            self.bit_off = self.bit_off[:ng + 1]
        self._cap = ng
        return self

    def encode(self, g0=None, g1=None):
        """Score groups [g0, g1) (default: all).  A partial range needs the host-side count and synchronises."""
        if g0 is not None or g1 is not None:
            self.sync_groups()
        cap = self._cap
        g0 = 0 if g0 is None else g0
        g1 = cap if g1 is None else g1
        wsb = _lib.load().cwq_is_encode_workspace_bytes(cap)
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.dev)
        if self.forced_starts is None:
            max_size = max(self.table.max_size_limit, 1)
        else:
            max_size = max(int(np.diff(np.asarray(st)).max()) for st in self.forced_starts)
        _lib.call("cwq_is_encode", _ptr(self.coef), _ptr(self.groups.starts), _ptr(self.K), _ptr(self.groups.group_seed),
                  cap, self._ng_dev(), g0, g1, max_size, _ptr(self.index), _ptr(ws), wsb, _stream())
        return self

    def expand_code_capacity(self):
        """Queue the '0'/'1' expansion of the packed code right behind the packing WITHOUT knowing the group count: the
        buffer is sized by the capacity (42 bits per possible group), the kernel reads the bit count on the device (the
        packing leaves it in the last slot of bit_off).  Lets the caller fetch every count in one small copy."""
        cap = self._cap
        self.ascii = device_bits_to_ascii(self.bits, 42 * cap, self.bit_off[cap:cap + 1])
        return self

    def expand_code_async(self):
        """Queue the '0'/'1' expansion of the packed code behind the packing, on the current stream, once the group count
        is known on the host (42 bits per group bound the buffer; the bit count itself stays on the device).  Collecting
        the results then needs copies only -- no kernel that would have to wait for SM room behind a scoring kernel."""
        ng = self.groups.n_groups
        self.ascii = device_bits_to_ascii(self.bits, 42 * ng, self.bit_off[ng:ng + 1])
        return self

    def shard_ranges_device(self, world):
        """int64[world + 1] on the device: contiguous work-balanced ranges of groups, one per rank (cwq_is_shard_ranges)."""
        cap = int(self.K.numel())
        bounds = torch.empty(world + 1, dtype=torch.int64, device=self.dev)
        if hasattr(self, "_buffers"):
            # fused pipeline: the GROUP stage left the work prefix in the workspace -> one tiny kernel
            rc = _lib.load().cwq_is_pipeline_shard_ranges(self.n, self.table.count, max(self.table.max_size_limit, 1), int(world),
                                                          _ptr(bounds), ctypes.cast(ctypes.pointer(self._buffers), ctypes.c_void_p),
                                                          _stream())
            if rc == 0:
                return bounds
        wsb = _lib.load().cwq_shard_ranges_workspace_bytes(cap)
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.dev)
        _lib.call("cwq_is_shard_ranges", _ptr(self.K), _ptr(self.groups.starts), cap, _ptr(self.groups.n_groups_dev), int(world),
                  _ptr(bounds), _ptr(ws), wsb, _stream())
        return bounds

    def encode_device_range(self, range_dev):
        """Score groups [range_dev[0], range_dev[1]) -- two int64 on the device -- without the host knowing the range
        (sharded runs: the rank's range is computed behind the grouping and scoring starts at once)."""
        if hasattr(self, "_buffers"):
            self._range_keepalive = range_dev
            self._buffers.score_range_dev = range_dev.data_ptr()
            try:
                return self.launch(_lib.STAGE_SCORE)
            finally:
                self._buffers.score_range_dev = None
        cap = int(self.K.numel())
        wsb = _lib.load().cwq_is_encode_workspace_bytes(cap)
        if hasattr(self, "_buffers") and self._buffers.workspace_bytes >= wsb:
            ws_ptr, wsb = ctypes.c_void_p(self._buffers.workspace), self._buffers.workspace_bytes
        else:
            self._range_ws = torch.empty(wsb, dtype=torch.uint8, device=self.dev)
            ws_ptr = _ptr(self._range_ws)
        _lib.call("cwq_is_encode_device_range", _ptr(self.coef), _ptr(self.groups.starts), _ptr(self.K), _ptr(self.groups.group_seed),
                  cap, _ptr(range_dev), max(self.table.max_size_limit, 1), _ptr(self.index), ws_ptr, wsb, _stream())
        return self

    def decode(self):
        """Regenerate the chosen rows (un-rescaled) and overwrite outlier dims; fills the side information."""
        n, t, g = self.n, self.table, self.groups
        self.sample = torch.empty(n, dtype=torch.float32, device=self.dev)
        _lib.call("cwq_is_decode", _ptr(g.starts), _ptr(g.group_seed), _ptr(g.gid), _ptr(self.index), n, _ptr(self.p_loc),
                  _ptr(self.p_scale), _ptr(self.sample), _stream())
        n_out_cap = n
        self.outlier_idx = torch.empty(n_out_cap, dtype=torch.int64, device=self.dev)
        self.outlier_q = torch.empty(n_out_cap, dtype=torch.int16, device=self.dev)  # uint16 payload
        self.seg_outlier_off = torch.empty(t.count + 1, dtype=torch.int64, device=self.dev)
        wsb = _lib.load().cwq_outlier_workspace_bytes(n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.dev)
        _lib.call("cwq_outlier_encode", _ptr(self.q_loc), _ptr(self.q_scale), _ptr(self.mask), n, _ptr(t.dev), t.count,
                  _ptr(self.sample), _ptr(self.outlier_idx), _ptr(self.outlier_q), _ptr(self.seg_outlier_off), _ptr(ws),
                  wsb, _stream())
        return self

    def pack(self):
        ngc = self._cap
        cap = ((ngc * 42 + 31) // 32) * 4
        self.bits = torch.empty(cap, dtype=torch.uint8, device=self.dev)
        self.bit_off = torch.empty(ngc + 1, dtype=torch.int64, device=self.dev)
        wsb = _lib.load().cwq_pack_workspace_bytes(ngc)
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.dev)
        _lib.call("cwq_elias_delta_pack", _ptr(self.index), ngc, self._ng_dev(), _ptr(self.bits), cap, _ptr(self.bit_off),
                  _ptr(ws), wsb, _stream())
        return self

    # ---- host views (each does a device->host copy) ----
    def segment_group_range(self, s):
        self.sync_groups()
        off = self.groups.seg_group_off_host
        return int(off[s]), int(off[s + 1])

    def host_starts(self, s):
        g0, g1 = self.segment_group_range(s)
        return (self.groups.starts[g0:g1 + 1] - self.table.starts[s]).cpu().numpy()

    def host_indices(self, s):
        g0, g1 = self.segment_group_range(s)
        return self.index[g0:g1].cpu().numpy().view(np.uint32)


# ----------------------------------------------------------------------------------------------
# greedy sampler
# ----------------------------------------------------------------------------------------------
def greedy_shard_loc(p_loc, p_scale, n_steps):
    """CGS:42 in rescaled coordinates: (p_loc / n_steps) / p_scale on the device."""
    out = torch.empty_like(p_loc)
    if p_loc.numel():
        _lib.call("cwq_greedy_shard_loc", _ptr(p_loc), _ptr(p_scale), int(p_loc.numel()), int(n_steps), _ptr(out), _stream())
    return out


def greedy_shard_scale(rho, n_steps):
    """CGS:42-45 with p = N(0,1): f32(rho) / f32(sqrt(n_steps))."""
    return float(np.float32(rho) / np.float32(np.sqrt(n_steps)))


class GreedyJob:
    def __init__(self, q_loc, q_scale, p_loc, p_scale, specs, n_steps, n_bits_per_step, rho=1.0, forced_starts=None,
                 shard_loc=None):
        """shard_loc: per-dim loc of the proposal shard in rescaled coordinates (cwq_greedy_shard_loc), or None = 0 -- what
        the grouped calls of the reference use (CGS:278-281); the per-group entry points pass it for p_loc != 0."""
        self.shard_loc = shard_loc
        self.dev = q_loc.device
        self.q_loc, self.q_scale, self.p_loc, self.p_scale = q_loc, q_scale, p_loc, p_scale
        self.table = SegmentTable(specs, self.dev)
        self.n = self.table.n
        if not (q_loc.numel() == q_scale.numel() == p_loc.numel() == p_scale.numel() == self.n):
            raise CwqError("loc/scale arrays must all have {} elements".format(self.n))
        self.n_steps, self.n_bits_per_step = int(n_steps), int(n_bits_per_step)
        if self.n_steps <= 0 or not (0 <= self.n_bits_per_step <= 30):
            raise CwqError("n_steps must be positive and 0 <= n_bits_per_step <= 30")
        self.shard_scale = greedy_shard_scale(rho, n_steps)
        self.forced_starts = forced_starts

    def prepare(self):
        n = self.n
        self.kl = torch.empty(n, dtype=torch.float32, device=self.dev)
        self.tparam = torch.empty(2 * n, dtype=torch.float32, device=self.dev)
        _lib.call("cwq_greedy_prologue", _ptr(self.q_loc), _ptr(self.q_scale), _ptr(self.p_loc), _ptr(self.p_scale), n,
                  _ptr(self.kl), _ptr(self.tparam), _stream())
        if self.forced_starts is None:
            self.groups = _run_group(self.kl, self.table)
        else:
            self.groups = _forced_groups(self.forced_starts, self.table, self.dev)
        ng = self.groups.n_groups
        self.indices = torch.zeros(ng * self.n_steps, dtype=torch.int32, device=self.dev)
        self.best = torch.zeros(n, dtype=torch.float32, device=self.dev)
        sizes = self.groups.starts[1:] - self.groups.starts[:-1]
        self.max_group_size = int(sizes.max().item())
        return self

    def encode(self, g0=0, g1=None, mode=0):
        """mode 0: automatic; 1: one persistent CTA per group; 2: lockstep units (few groups)."""
        ng = self.groups.n_groups
        g1 = ng if g1 is None else g1
        wsb = _lib.load().cwq_greedy_workspace_bytes(self.n, ng, self.n_steps)
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.dev)
        _lib.call("cwq_greedy_encode", _ptr(self.tparam), _ptr(self.groups.starts), _ptr(self.groups.group_seed), ng, g0, g1,
                  self.n_steps, self.n_bits_per_step, self.shard_scale, _ptr(self.shard_loc), self.max_group_size, self.n, int(mode),
                  _ptr(self.indices), _ptr(self.best), _ptr(ws), wsb, _stream())
        return self

    def decode(self):
        """Replay the chosen rows from the indices (what the decoder runs), un-rescaled."""
        g = self.groups
        self.sample = torch.empty(self.n, dtype=torch.float32, device=self.dev)
        _lib.call("cwq_greedy_decode", _ptr(g.starts), _ptr(g.group_seed), _ptr(g.gid), _ptr(self.indices), self.n,
                  self.n_steps, self.shard_scale, _ptr(self.shard_loc), _ptr(self.p_loc), _ptr(self.p_scale), _ptr(self.sample),
                  _stream())
        return self

    def unrescale(self):
        if getattr(self, "best_from_decode", False):
            return self  # sharded run: the sample was already rebuilt from the gathered indices (decode())
        self.sample = torch.empty(self.n, dtype=torch.float32, device=self.dev)
        _lib.call("cwq_unrescale", _ptr(self.best), _ptr(self.p_loc), _ptr(self.p_scale), self.n, _ptr(self.sample), _stream())
        return self

    def pack(self):
        count = self.groups.n_groups * self.n_steps
        b = self.n_bits_per_step
        if b == 0 or count == 0:
            self.bits = torch.zeros(4, dtype=torch.uint8, device=self.dev)
            return self
        cap = ((count * b + 31) // 32) * 4
        self.bits = torch.empty(cap, dtype=torch.uint8, device=self.dev)
        _lib.call("cwq_fixed_pack", _ptr(self.indices), count, b, _ptr(self.bits), cap, _stream())
        return self

    def segment_group_range(self, s):
        off = self.groups.seg_group_off_host
        return int(off[s]), int(off[s + 1])

    def host_starts(self, s):
        g0, g1 = self.segment_group_range(s)
        return (self.groups.starts[g0:g1 + 1] - self.table.starts[s]).cpu().numpy()


# ----------------------------------------------------------------------------------------------
# the candidate stream (misc.py:3-17)
# ----------------------------------------------------------------------------------------------
def draw_candidates(stream_seed, k0, nk, G, device=None, raw=False):
    """Rows k0..k0+nk-1 of the candidate table of stream `stream_seed` as a float32 [nk, G] CUDA tensor: N(0,1) draws, or
    with raw=True the stream's raw draws y (z = float32(sqrt(2 ln 2)) * y) that the candidate loops score."""
    device = device or cuda_device()
    out = torch.empty((int(nk), int(G)), dtype=torch.float32, device=device)
    if nk and G:
        _lib.call("cwq_draw_candidates", int(stream_seed) & 0xFFFFFFFF, int(k0) & 0xFFFFFFFF, int(nk), int(G), 1 if raw else 0,
                  _ptr(out), _stream())
    return out


def stream_words(stream_seed, k0, nk, nq, device=None):
    device = device or cuda_device()
    out = torch.empty((int(nk), int(nq), 4), dtype=torch.int32, device=device)
    if nk and nq:
        _lib.call("cwq_stream_words", int(stream_seed) & 0xFFFFFFFF, int(k0) & 0xFFFFFFFF, int(nk), int(nq), _ptr(out),
                  _stream())
    return out


def philox4x32_10(ctr, key, device=None):
    """ctr uint32[n,4], key uint32[n,2] (numpy) -> uint32[n,4] (numpy)."""
    device = device or cuda_device()
    c = torch.from_numpy(np.ascontiguousarray(ctr, dtype=np.uint32).view(np.int32)).to(device)
    k = torch.from_numpy(np.ascontiguousarray(key, dtype=np.uint32).view(np.int32)).to(device)
    n = c.numel() // 4
    out = torch.empty_like(c)
    _lib.call("cwq_philox4x32_10", _ptr(c), _ptr(k), _ptr(out), n, _stream())
    return out.cpu().numpy().view(np.uint32).reshape(-1, 4)


def elias_delta_parse(code, n_codes):
    """Sequential parse of n_codes Elias-delta codes from a '0'/'1' str/bytes -> (uint32 indices (x - 1), bits used)."""
    if isinstance(code, str):
        code = code.encode("utf-8")
    if int(n_codes) == 0:
        return np.zeros(0, dtype=np.uint32), 0
    buf = np.frombuffer(code, dtype=np.uint8)
    if buf.size == 0:
        raise CwqError("empty bitcode for {} groups".format(n_codes))
    out = np.empty(max(int(n_codes), 1), dtype=np.uint32)
    used = ctypes.c_int64(0)
    _lib.call("cwq_elias_delta_parse_host", ctypes.c_void_p(buf.ctypes.data if buf.size else 0), int(buf.size), int(n_codes),
              ctypes.c_void_p(out.ctypes.data), ctypes.cast(ctypes.pointer(used), ctypes.c_void_p))
    return out[:int(n_codes)], int(used.value)
