"""Single-process multi-GPU context (bsa_create_multi): the same calls, the same bytes.
On a one-GPU box the multi context still runs its two worker threads / child contexts on that GPU
(tiles pulled from the shared counter, results copied tile-wise into the caller's buffers); with
more GPUs visible every device takes part."""
import ctypes as C

import numpy as np
import pytest

import bioshell_b200 as bs
from bioshell_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _devices():
    n = _lib.lib().bsa_device_count()
    return list(range(n))


@pytest.fixture(scope="module")
def mctx():
    c = bs.Context(_devices())
    yield c
    c.close()


def test_multi_context_reports_its_devices(mctx):
    assert mctx.n_devices == len(_devices())
    one = bs.Context("all")
    assert one.n_devices == len(_devices())
    one.close()


def test_all_vs_all_bytes_identical_to_one_device(ctx, mctx):
    res, off = synth.generate(2500, seed=77, dist=1)
    n = len(off) - 1
    for c in (ctx, mctx):
        c.set_scoring("BLOSUM62", -10, -1)
        c.load_sequences(0, res, off)
    s1, n1 = ctx.all_vs_all(0)
    # results straight into pinned memory from bsa_host_alloc_pinned
    L = _lib.lib()
    npairs = n * (n - 1) // 2
    ps, pn = L.bsa_host_alloc_pinned(npairs * 4), L.bsa_host_alloc_pinned(npairs * 4)
    try:
        s2 = np.ctypeslib.as_array(C.cast(ps, C.POINTER(C.c_int32)), (npairs,))
        n2 = np.ctypeslib.as_array(C.cast(pn, C.POINTER(C.c_uint32)), (npairs,))
        s2[:] = -12345
        n2[:] = 0xFFFFFFFF
        mctx.align_all_pairs(0, 0, np.arange(n, dtype=np.uint32), scores=s2, n_identical=n2)
        assert np.array_equal(s1, s2) and np.array_equal(n1, n2)
        st = mctx.stats()
        assert st["pairs"] == npairs and st["cells"] == ctx.stats()["cells"]
    finally:
        L.bsa_host_free_pinned(ps)
        L.bsa_host_free_pinned(pn)
    # a template sub-range and the score-only one-vs-many form
    a = ctx.align_all_pairs(0, 0, np.arange(n, dtype=np.uint32), 700, 1900)
    b = mctx.align_all_pairs(0, 0, np.arange(n, dtype=np.uint32), 700, 1900)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    qres, qoff = synth.generate(40, seed=78, dist=1)
    for c in (ctx, mctx):
        c.load_sequences(1, qres, qoff)
    a = ctx.one_vs_many(1, 0)
    b = mctx.one_vs_many(1, 0)
    assert np.array_equal(a[0], b[0])


def test_pair_lists_and_errors_through_the_multi_context(ctx, mctx):
    res, off = synth.generate(60, seed=79, dist=0, lo=1, hi=900)
    rng = np.random.default_rng(3)
    q, t = rng.integers(0, 60, 200), rng.integers(0, 60, 200)
    for c in (ctx, mctx):
        c.set_scoring("BLOSUM62", -11, -1)
        c.load_sequences(2, res, off)
    a = ctx.align_pairs_paths(2, 2, q, t)
    b = mctx.align_pairs_paths(2, 2, q, t)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    la = ctx.local_align_pairs(2, 2, q, t)
    lb = mctx.local_align_pairs(2, 2, q, t)
    for k in ("score", "end_q", "end_t", "start_q", "start_t"):
        assert np.array_equal(la[k], lb[k])
    assert la["paths"] == lb["paths"]
    with pytest.raises(_lib.BsaError) as e:
        mctx.set_scoring("BLOSUM62", -1, -5)
    assert e.value.rc == -2
    with pytest.raises(_lib.BsaError):
        mctx.align_all_pairs(0, 0, None, scores=1, n_identical=1, device_out=True)   # device outputs: single-device contexts only
