"""The raw ctypes sequence shown in INTEGRATION.md section 2, executed as written (one importance-sampler
encode straight through the C ABI), must give the package's own result."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_integration_md_stub_runs_and_matches_package():
    import torch
    import compression_without_quantization_b200 as cwq
    from compression_without_quantization_b200 import _lib, synthetic as syn
    lib = ctypes.CDLL(_lib.LIB_PATH)
    P, I64, I32, SZ = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_size_t
    ptr = lambda t: P(t.data_ptr())  # noqa: E731
    st = P(torch.cuda.current_stream().cuda_stream)

    class Segment(ctypes.Structure):
        _fields_ = [("start", I64), ("end", I64), ("n_nats_per_group", ctypes.c_double), ("size_limit", I64),
                    ("dim_kl_bit_limit", ctypes.c_float), ("seed", ctypes.c_uint32), ("strict", I32), ("reserved", I32)]

    arrs = syn.synthetic_posterior(5000, 70, True)
    q_loc, q_scale, p_loc, p_scale = (torch.from_numpy(a).cuda() for a in arrs)
    n = q_loc.numel()
    seg = (Segment * 1)(Segment(0, n, 14 * np.log(2) - 1, 16, 12.0, 42, 1, 0))
    seg_d = torch.frombuffer(bytearray(bytes(seg)), dtype=torch.uint8).cuda()
    kl, coef = torch.empty(n, device="cuda"), torch.empty(2 * n, device="cuda")
    mask = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert lib.cwq_is_prologue(ptr(q_loc), ptr(q_scale), ptr(p_loc), ptr(p_scale), I64(n), ptr(seg_d), I32(1), ptr(kl), ptr(coef),
                               ptr(mask), st) == 0
    starts = torch.empty(n + 2, dtype=torch.int64, device="cuda")
    gseed = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    gid = torch.empty(n, dtype=torch.int32, device="cuda")
    seg_off = torch.empty(2, dtype=torch.int64, device="cuda")
    lib.cwq_group_workspace_bytes.restype = SZ
    wsb = lib.cwq_group_workspace_bytes(I64(n), I32(1))
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    assert lib.cwq_group(ptr(kl), I64(n), ptr(seg_d), ctypes.cast(seg, P), I32(1), ptr(starts), ptr(gseed), ptr(gid), ptr(seg_off),
                         ptr(ws), SZ(wsb), st) == 0
    ng = int(seg_off[1])
    total, K = torch.empty(ng, device="cuda"), torch.empty(ng, dtype=torch.int32, device="cuda")
    assert lib.cwq_group_totals(ptr(kl), ptr(starts), I64(ng), None, ptr(total), ptr(K), st) == 0
    index = torch.zeros(ng, dtype=torch.int32, device="cuda")
    lib.cwq_is_encode_workspace_bytes.restype = SZ
    wsb = lib.cwq_is_encode_workspace_bytes(I64(ng))
    ws2 = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    assert lib.cwq_is_encode(ptr(coef), ptr(starts), ptr(K), ptr(gseed), I64(ng), None, I64(0), I64(ng), I64(16), ptr(index), ptr(ws2),
                             SZ(wsb), st) == 0
    sample = torch.empty(n, device="cuda")
    assert lib.cwq_is_decode(ptr(starts), ptr(gseed), ptr(gid), ptr(index), I64(n), ptr(p_loc), ptr(p_scale), ptr(sample), st) == 0
    torch.cuda.synchronize()

    want = cwq.code_grouped_importance_sample(None, syn.Normal(arrs[0], arrs[1]), syn.Normal(arrs[2], arrs[3]), 42, 14, 4, 12,
                                              return_indices=True)
    assert np.array_equal(starts[:ng + 1].cpu().numpy(), want[2])
    assert [int(i) + 1 for i in index.cpu().numpy().view(np.uint32)] == [int(i) for i in want[1]]
    inl = np.ones(n, bool)
    inl[want[3][0]] = False
    assert np.array_equal(sample.cpu().numpy()[inl].view(np.uint32), want[0][inl].view(np.uint32))
    # a wrong argument is reported, not ignored
    assert lib.cwq_is_encode(ptr(coef), ptr(starts), ptr(K), ptr(gseed), I64(ng), None, I64(5), I64(2), I64(16), ptr(index), ptr(ws2),
                             SZ(wsb), st) != 0
    lib.cwq_last_error.restype = ctypes.c_char_p
    assert b"bad group range" in lib.cwq_last_error()
