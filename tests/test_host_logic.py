"""Host-side logic that needs neither a GPU nor the CUDA library: chunk planning of the batch pipeline, the Philox block
count of a row, shard bounds from device cuts, the reference's group-size test restated on integers."""
import numpy as np
import torch

from compression_without_quantization_b200 import coded_importance_sampler as CIS
from compression_without_quantization_b200 import distributed as D
from compression_without_quantization_b200 import engine as E


def test_plan_chunks_covers_every_call_once_in_order():
    rng = np.random.default_rng(3)
    for case in range(200):
        n_calls = int(rng.integers(1, 60))
        sizes = [int(x) for x in rng.integers(1, 300000, n_calls)]
        n_chunks = int(rng.integers(1, 9))
        chunks = CIS._plan_chunks(sizes, n_chunks)
        assert chunks[0][0] == 0 and chunks[-1][1] == n_calls and len(chunks) <= max(n_chunks, 1)
        assert all(a < b for a, b in chunks) and all(chunks[i][1] == chunks[i + 1][0] for i in range(len(chunks) - 1))


def test_plan_chunks_bench_batch_has_small_first_and_last_chunk():
    sizes = [2304, 196608] * 24
    chunks = CIS._plan_chunks(sizes, 4)
    dims = [sum(sizes[a:b]) for a, b in chunks]
    assert len(chunks) == 4 and dims[0] < dims[1] and dims[3] < dims[2] and sum(dims) == sum(sizes)
    assert abs(dims[0] / sum(dims) - 1 / 8) < 0.05 and abs(dims[1] / sum(dims) - 3 / 8) < 0.05


def test_single_gpu_chunk_layout_of_the_bench_batch():
    """The default single-GPU layout keeps the two stages nothing overlaps (first upload, last download + formatting) at
    one image each; an explicit chunk count keeps the 1:3:...:3:1 rule."""
    sizes = [2304, 196608] * 24
    chunks = CIS._plan_chunks(sizes, CIS.PIPELINE_CHUNKS, CIS._chunk_weights(1))
    assert chunks == [(0, 2), (2, 24), (24, 46), (46, 48)]
    assert CIS._chunk_weights(2) == [1.0, 6.0, 1.0] and CIS._chunk_weights(8) == [1.0, 3.0]


def test_row_blocks_three_per_sixteen_dims():
    sizes = torch.arange(0, 200)
    got = D.row_blocks(sizes).tolist()
    for g, b in zip(sizes.tolist(), got):
        nq = (g + 3) // 4
        assert b == 3 * (nq // 4) + min(nq % 4, 3)
    assert got[16] == 3 and got[12] == 3 and got[13] == 3 and got[8] == 2 and got[4] == 1 and got[17] == 4 and got[0] == 0


def test_balanced_ranges_properties():
    rng = np.random.default_rng(11)
    for world in (1, 2, 3, 8):
        work = rng.integers(0, 1000, 5000).astype(np.int64)
        b = D.balanced_ranges(work, world)
        assert b[0] == 0 and b[-1] == len(work) and all(x <= y for x, y in zip(b, b[1:])) and len(b) == world + 1
        shares = [work[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(abs(s - work.sum() / world) for s in shares) <= work.max()
    assert D.balanced_ranges(np.zeros(0, np.int64), 4) == [0, 0, 0, 0, 0]
    assert D._bounds_from_cuts([7, 5, 9], 8, 4) == [0, 7, 7, 8, 8]     # cuts are clamped and made ascending


def test_size_limit_matches_the_reference_expression():
    # CIS:185-187 `np.log(size + 1) / np.log(2) > m` and CGS:230-232 `>=`, evaluated as the reference evaluates them
    for m in (0, 1, 2, 3, 4, 5, 8, 12, 16):
        for strict in (True, False):
            lim = E.size_limit(m, strict)
            fires = lambda size: (np.log(size + 1) / np.log(2) > m) if strict else (np.log(size + 1) / np.log(2) >= m)
            assert fires(lim) and (lim == 0 or not fires(lim - 1))
    assert E.size_limit(4, True) == 16 and E.size_limit(12, False) == 4095
