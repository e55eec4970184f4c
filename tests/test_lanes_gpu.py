"""GPU: the call-level machinery of the C ABI -- the fast (non-counting) kernels, asynchronous calls, two batches in
flight (submit / wait, two streams, two host threads), the in-stream re-run of overflowed queries -- always against
the oracle, bit for bit.  Reference behaviour mirrored: Index::search is safe for concurrent callers
(api/hnsw_index.hpp:172, shared_lock + thread_local scratch); Index::load commits nothing before it has validated
(api/hnsw_index.hpp:305-443)."""
import threading

import numpy as np
import pytest

import common
from common import co

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _same(ids, dists, oid, od):
    gi, gd = common.sorted_rows(ids, dists)
    wi, wd = common.sorted_rows(oid, od)
    return np.array_equal(gi, wi) and np.array_equal(_bits(gd), _bits(wd))


@pytest.fixture(scope="module")
def case(oracle):
    fab = common.fabricate(6000, 96, 4, seed=21, layers=2, counts=(32, 32, 29, 8), gamma=1.2, gamma_max=2.0, gamma_warmup=4)
    ix = common.gpu_index_from(fab)
    ix.set_option("collect_stats", 0)
    view = oracle.index_view(fab)
    rng = np.random.default_rng(3)
    batches = [rng.standard_normal((n, 96)).astype(np.float32) for n in (300, 1, 257, 64, 300, 33)]
    want = [oracle.search_batch(view, q, 10)[:2] for q in batches]
    return fab, ix, view, batches, want


@pytest.mark.parametrize("dim,bits,k", [(96, 4, 10), (128, 2, 1), (128, 1, 32), (64, 4, 33), (256, 2, 100), (960, 2, 10)])
def test_fast_kernels_equal_the_oracle(oracle, dim, bits, k):
    """collect_stats = 0 selects the kernels the bench runs: the result list in registers for k <= 32, in shared memory
    above; the counting build (what most parity tests run, for the counters) must agree with both."""
    fab = common.fabricate(3000, dim, bits, seed=dim + bits + k, layers=1, counts=(32, 31, 16), gamma=1.3, gamma_max=2.5, gamma_warmup=5)
    ix = common.gpu_index_from(fab)
    q = np.random.default_rng(k).standard_normal((150, dim)).astype(np.float32)
    oid, od, _ = oracle.search_batch(oracle.index_view(fab), q, k)
    ids_c, d_c = ix.search_batch(q, k)
    ix.set_option("collect_stats", 0)
    ids_f, d_f = ix.search_batch(q, k)
    assert _same(ids_c, d_c, oid, od)
    assert _same(ids_f, d_f, oid, od)
    assert np.array_equal(ids_f, ids_c) and np.array_equal(_bits(d_f), _bits(d_c))   # same order inside rows, too
    assert ix.last_stats()["kernel_launches"] in (2, 3)


def test_submit_wait_keeps_two_batches_in_flight(case):
    fab, ix, view, batches, want = case
    got = list(ix.search_batches(batches, 10))
    assert len(got) == len(batches)
    for (ids, d), (oid, od) in zip(got, want):
        assert _same(ids, d, oid, od)
    # tickets waited out of order, and a ticket waited after its lane was reused
    t0 = ix.search_batch_submit(batches[0], 10)
    t1 = ix.search_batch_submit(batches[2], 10)
    r1 = ix.search_batch_wait(t1)
    t2 = ix.search_batch_submit(batches[3], 10)     # takes the lane of t0: waits for it first
    r0 = ix.search_batch_wait(t0)
    r2 = ix.search_batch_wait(t2)
    for (ids, d), j in ((r0, 0), (r1, 2), (r2, 3)):
        assert _same(ids, d, *want[j])
    with pytest.raises(ValueError):
        ix.search_batch_wait(10**9)


def test_pinned_buffers_and_out_arguments(case):
    torch = pytest.importorskip("torch")
    fab, ix, view, batches, want = case
    q = torch.from_numpy(batches[0]).pin_memory()
    ids = torch.empty((q.shape[0], 10), dtype=torch.int64).pin_memory()
    d = torch.empty((q.shape[0], 10), dtype=torch.float32).pin_memory()
    t = ix.search_batch_submit(q, 10, out=(ids, d))
    ri, rd = ix.search_batch_wait(t)
    assert ri.ctypes.data == ids.data_ptr()
    assert _same(ids.numpy(), d.numpy(), *want[0])


def test_two_host_threads_share_one_index(case):
    fab, ix, view, batches, want = case
    errors, results = [], {}

    def worker(tid):
        try:
            for rep in range(6):
                j = (tid + rep) % len(batches)
                ids, d = ix.search_batch(batches[j], 10)
                results[(tid, rep)] = (j, ids, d)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors
    assert len(results) == 24
    for j, ids, d in results.values():
        assert _same(ids, d, *want[j])


def test_device_calls_on_two_streams_are_asynchronous_and_correct(case):
    torch = pytest.importorskip("torch")
    fab, ix, view, batches, want = case
    dev = torch.device("cuda", 0)
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    qs = [torch.from_numpy(b).to(dev) for b in batches]
    torch.cuda.synchronize()
    outs = []
    for i, q in enumerate(qs * 2):
        with torch.cuda.stream(streams[i % 2]):
            outs.append(ix.search_batch(q, 10))
    ix.synchronize()
    torch.cuda.synchronize()
    for i, (ids, d) in enumerate(outs):
        assert _same(ids.cpu().numpy(), d.cpu().numpy(), *want[i % len(batches)])


def test_overflowed_queries_are_rerun_in_stream(oracle):
    """A 64-entry frontier arena overflows for most queries; the re-run is enqueued behind the first pass (no host round
    trip), also with two batches in flight."""
    fab = common.fabricate(4000, 64, 1, seed=9, gamma=1e6, gamma_max=1e7, slacks=(3.0,), floor=0.2)
    ix = common.gpu_index_from(fab)
    view = oracle.index_view(fab)
    rng = np.random.default_rng(6)
    batches = [rng.standard_normal((40, 64)).astype(np.float32) for _ in range(4)]
    want = [oracle.search_batch(view, q, 10) for q in batches]
    ix.set_option("beam_capacity", 64)
    for stats in (1, 0):
        ix.set_option("collect_stats", stats)
        got = list(ix.search_batches(batches, 10))
        for (ids, d), (oid, od, ost) in zip(got, want):
            assert _same(ids, d, oid, od)
        st = ix.last_stats()
        assert st["kernel_launches"] == 3
        if want[-1][2]["max_beam"] > 64:
            assert st["overflow_retries"] > 0


def test_a_rejected_load_leaves_the_index_usable(case, tmp_path):
    fab, ix, view, batches, want = case
    other = common.fabricate(500, 32, 2, seed=1)
    bad = common.write_save_file(other, tmp_path / "other.bin")
    with pytest.raises(RuntimeError, match="Parameter mismatch"):
        ix.load(str(bad))
    assert ix.is_finalized and ix.size == 6000
    ids, d = ix.search_batch(batches[3], 10)
    assert _same(ids, d, *want[3])
    trunc = tmp_path / "trunc.bin"
    good = common.write_save_file(fab, tmp_path / "good.bin")
    trunc.write_bytes(open(good, "rb").read()[:5000])
    with pytest.raises(RuntimeError, match="truncated"):
        ix.load(str(trunc))
    assert ix.is_finalized          # validation failed before anything was committed: the old index still answers
    ids, d = ix.search_batch(batches[3], 10)
    assert _same(ids, d, *want[3])
    # a header whose vertex count would wrap the size arithmetic is refused, not read
    huge = bytearray(open(good, "rb").read())
    huge[28:36] = (2**61).to_bytes(8, "little")
    (tmp_path / "huge.bin").write_bytes(bytes(huge))
    with pytest.raises(RuntimeError, match="truncated"):
        ix.load(str(tmp_path / "huge.bin"))
    ix.load(str(good))
    ids, d = ix.search_batch(batches[3], 10)
    assert _same(ids, d, *want[3])


def test_current_device_is_left_alone(case):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import cphnsw_b200

    fab, ix, view, batches, want = case
    torch.cuda.set_device(1)
    try:
        ids, d = ix.search_batch(batches[3], 10)
        assert torch.cuda.current_device() == 1
        assert _same(ids, d, *want[3])
        other = cphnsw_b200.CPIndex(96, 4, device=0)
        assert torch.cuda.current_device() == 1
        del other
    finally:
        torch.cuda.set_device(0)
