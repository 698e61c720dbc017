"""CPU: the host staging part of the C ABI (b200fe_host_*): packing / float64->float32 conversion of utterance
lists into one staging buffer and the zero fill of padding rows -- no CUDA device needed."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def pool(lasr_b200):
    lib = lasr_b200._lib.load()
    h = C.c_void_p()
    assert lib.b200fe_host_pool_create(3, C.byref(h)) == 0
    assert lib.b200fe_host_pool_threads(h) == 3
    yield lib, h
    lib.b200fe_host_pool_destroy(h)


def _aligned(n, dtype):
    raw = np.empty(n * np.dtype(dtype).itemsize + 64, dtype=np.uint8)
    o = (-raw.ctypes.data) % 16
    return raw[o:o + n * np.dtype(dtype).itemsize].view(dtype)


@pytest.mark.parametrize("src_dtype,code,dst_dtype", [(np.float64, 2, np.float32), (np.float32, 0, np.float32), (np.int16, 1, np.int16)])
def test_pack_list_of_utterances(pool, src_dtype, code, dst_dtype):
    lib, h = pool
    rng = np.random.default_rng(7)
    lens = np.array([1, 401, 70001, 0, 16000, 123457, 7], dtype=np.int64)           # ragged, one empty, one longer than a task chunk
    if src_dtype == np.int16:
        wavs = [rng.integers(-32768, 32767, n).astype(np.int16) for n in lens]
    else:
        wavs = [rng.normal(0, 0.3, n).astype(src_dtype) for n in lens]
    al = 16 // np.dtype(dst_dtype).itemsize
    offs = np.zeros(len(lens), dtype=np.int64)
    np.cumsum((lens[:-1] + al - 1) // al * al, out=offs[1:])
    total = int(offs[-1] + (lens[-1] + al - 1) // al * al)
    dst = _aligned(total, dst_dtype)
    dst[:] = 99
    ptrs = (C.c_void_p * len(wavs))(*[max(w.ctypes.data, 1) for w in wavs])
    tk = lib.b200fe_host_pack_begin(h, ptrs, lens.ctypes.data, len(wavs), code, dst.ctypes.data, offs.ctypes.data, total)
    assert tk > 0, lib.b200fe_last_error()
    assert lib.b200fe_host_wait(h, tk) == 0
    for w, o, n in zip(wavs, offs, lens):
        assert np.array_equal(dst[o:o + n], w.astype(dst_dtype))                      # float64 -> float32: round to nearest, like .astype
        assert np.all(dst[o + n:o + (n + al - 1) // al * al] == 0)                     # alignment gap cleared
    assert lib.b200fe_host_wait(h, tk) != 0                                            # a ticket can be waited for once


def test_pack_unaligned_sources_and_special_values(pool):
    """The streaming loops (host_simd.cpp, SSE2 or AVX-512F picked at load time): sources at odd element offsets, lengths around the
    vector widths, values that round, overflow or are not numbers -- element for element what numpy's astype(float32) gives."""
    lib, h = pool
    assert lib.b200fe_host_isa() in (0, 1)
    rng = np.random.default_rng(11)
    base = rng.normal(0, 1, 300000)
    base[::97] *= 1e-42                       # float32 denormals
    base[5::101] *= 1e39                      # beyond float32: +-inf
    base[7::211] = np.nan
    base[11::223] = np.inf
    base[13::227] = 1.0 + 2.0 ** -24          # ties round to even
    lens = np.array([0, 1, 15, 16, 17, 31, 33, 63, 64, 65, 127, 129, 1000, 65536, 65537, 70001], dtype=np.int64)
    starts = np.arange(len(lens)) * 3 + 1     # odd element offsets: 8-byte aligned, not 16 / 64
    wavs = [base[st:st + n] for st, n in zip(starts, lens)]
    offs = np.zeros(len(lens), dtype=np.int64)
    np.cumsum((lens[:-1] + 3) // 4 * 4, out=offs[1:])
    total = int(offs[-1] + (lens[-1] + 3) // 4 * 4)
    dst = _aligned(total, np.float32)
    dst[:] = 99
    ptrs = (C.c_void_p * len(wavs))(*[max(w.ctypes.data, 1) for w in wavs])
    tk = lib.b200fe_host_pack_begin(h, ptrs, lens.ctypes.data, len(wavs), 2, dst.ctypes.data, offs.ctypes.data, total)
    assert tk > 0 and lib.b200fe_host_wait(h, tk) == 0
    with np.errstate(over="ignore"):
        for w, o, n in zip(wavs, offs, lens):
            assert np.array_equal(dst[o:o + n], w.astype(np.float32), equal_nan=True)
    # byte copies (int16 / float32 lists) from odd byte offsets
    raw = rng.integers(-32768, 32767, 200000).astype(np.int16)
    lens16 = np.array([1, 7, 8, 9, 31, 32, 33, 255, 70003], dtype=np.int64)
    w16 = [raw[1 + 5 * i:1 + 5 * i + n] for i, n in enumerate(lens16)]
    offs16 = np.zeros(len(lens16), dtype=np.int64)
    np.cumsum((lens16[:-1] + 7) // 8 * 8, out=offs16[1:])
    tot16 = int(offs16[-1] + (lens16[-1] + 7) // 8 * 8)
    d16 = _aligned(tot16, np.int16)
    d16[:] = 99
    p16 = (C.c_void_p * len(w16))(*[w.ctypes.data for w in w16])
    tk = lib.b200fe_host_pack_begin(h, p16, lens16.ctypes.data, len(w16), 1, d16.ctypes.data, offs16.ctypes.data, tot16)
    assert tk > 0 and lib.b200fe_host_wait(h, tk) == 0
    for w, o, n in zip(w16, offs16, lens16):
        assert np.array_equal(d16[o:o + n], w)
    # zero fill of ranges that start and end anywhere
    buf = np.full(5000, 7, dtype=np.uint8)
    zo = np.array([1, 100, 1000, 4093], dtype=np.int64)
    zn = np.array([63, 129, 2049, 7], dtype=np.int64)
    tk = lib.b200fe_host_zero_ranges_begin(h, buf.ctypes.data, zo.ctypes.data, zn.ctypes.data, len(zo))
    assert tk > 0 and lib.b200fe_host_wait(h, tk) == 0
    want = np.full(5000, 7, dtype=np.uint8)
    for o, n in zip(zo, zn):
        want[o:o + n] = 0
    assert np.array_equal(buf, want)


def test_ndarray_data_pointers(pool):
    """b200fe_host_ndarray_data: the data pointers of ndarrays from the addresses of their Python objects (what HostPipeline
    uses instead of 256 __array_interface__ look-ups), on whole arrays, offset views and rows of a 2-D array."""
    lib, h = pool
    base = np.arange(1000, dtype=np.float64)
    m = np.zeros((7, 33), dtype=np.int16)
    arrs = [base, base[3:], base[10:500:1], m[4], np.zeros(0), np.zeros(5, dtype=np.float32)]
    ids = np.fromiter(map(id, arrs), dtype=np.int64, count=len(arrs))
    out = np.zeros(len(arrs), dtype=np.uint64)
    assert lib.b200fe_host_ndarray_data(ids.ctypes.data, len(arrs), out.ctypes.data, 16) == 0
    assert [int(v) for v in out] == [a.__array_interface__["data"][0] for a in arrs]
    ids[2] = 0
    assert lib.b200fe_host_ndarray_data(ids.ctypes.data, len(arrs), out.ctypes.data, 16) != 0
    assert lib.b200fe_host_ndarray_data(ids.ctypes.data, len(arrs), out.ctypes.data, 12) != 0


def test_pack_float64_that_holds_pcm16_values(pool):
    """src_dtype 3: float64 samples k / 32768 (soundfile.read of PCM_16 files) are staged as the int16 k; one sample that is not such
    a value raises the job's flag; the probe looks at windows spread over every utterance."""
    lib, h = pool
    rng = np.random.default_rng(3)
    lens = np.array([1, 31, 32, 33, 401, 70001, 0, 12345], dtype=np.int64)
    ks = [rng.integers(-32768, 32768, n).astype(np.int16) for n in lens]
    ks[4][:4] = [-32768, 32767, 0, -1]
    wavs = [k.astype(np.float64) / 32768.0 for k in ks]
    wavs[3][5], ks[3][5] = -0.0, 0                       # minus zero is the PCM value 0
    offs = np.zeros(len(lens), dtype=np.int64)
    np.cumsum((lens[:-1] + 7) // 8 * 8, out=offs[1:])
    total = int(offs[-1] + (lens[-1] + 7) // 8 * 8)
    dst = _aligned(total, np.int16)

    def run(ws):
        dst[:] = 99
        ptrs = (C.c_void_p * len(ws))(*[max(w.ctypes.data, 1) for w in ws])
        assert lib.b200fe_host_pcm16_probe(ptrs, lens.ctypes.data, len(ws), 8, 64) in (0, 1)
        tk = lib.b200fe_host_pack_begin(h, ptrs, lens.ctypes.data, len(ws), 3, dst.ctypes.data, offs.ctypes.data, total)
        assert tk > 0, lib.b200fe_last_error()
        flag = C.c_int(-1)
        assert lib.b200fe_host_wait_flag(h, tk, C.byref(flag)) == 0
        return flag.value, lib.b200fe_host_pcm16_probe(ptrs, lens.ctypes.data, len(ws), 8, 64)

    flag, probe = run(wavs)
    assert flag == 0 and probe == 1
    for k, o, n in zip(ks, offs, lens):
        assert np.array_equal(dst[o:o + n], k)
        assert np.all(dst[o + n:o + (n + 7) // 8 * 8] == 0)
    # values that are not PCM16: off the grid, out of range (+1.0 = 32768 / 32768), NaN, infinity -- in the middle of a long utterance,
    # where the probe's windows do not look, and at its start, where they do
    for bad in (0.1, 1.0, -1.0000305, np.nan, np.inf, 2.0 ** -16):
        w2 = [w.copy() for w in wavs]
        w2[5][33333] = bad
        flag, probe = run(w2)
        assert flag == 1 and probe == 1, bad
        w2[5][3] = bad
        flag, probe = run(w2)
        assert flag == 1 and probe == 0, bad
    assert lib.b200fe_host_pcm16_probe(None, lens.ctypes.data, 2, 8, 64) < 0


def test_pack_argument_errors(pool):
    lib, h = pool
    w = np.zeros(10)
    ptrs = (C.c_void_p * 1)(w.ctypes.data)
    lens = np.array([10], dtype=np.int64)
    dst = _aligned(16, np.float32)
    off_bad = np.array([2], dtype=np.int64)                                             # not 16-byte aligned
    assert lib.b200fe_host_pack_begin(h, ptrs, lens.ctypes.data, 1, 2, dst.ctypes.data, off_bad.ctypes.data, 16) < 0
    off = np.array([8], dtype=np.int64)
    assert lib.b200fe_host_pack_begin(h, ptrs, lens.ctypes.data, 1, 2, dst.ctypes.data, off.ctypes.data, 16) < 0      # does not fit
    assert lib.b200fe_host_pack_begin(h, ptrs, lens.ctypes.data, 1, 5, dst.ctypes.data, off.ctypes.data, 64) < 0      # bad dtype code
    assert b"host_pack" in lib.b200fe_last_error()


def test_zero_padding_rows(pool):
    lib, h = pool
    B, T, D = 5, 300, 80
    feats = np.full((B, T, D), 3.0, dtype=np.float32)
    valid = np.array([300, 0, 17, 299, 1000], dtype=np.int64)                           # full, empty, partial, counts past T are clipped
    tk = lib.b200fe_host_zero_rows_begin(h, feats.ctypes.data, B, T, D, valid.ctypes.data, 4)
    assert tk > 0 and lib.b200fe_host_wait(h, tk) == 0
    for b, v in enumerate(np.minimum(valid, T)):
        assert np.all(feats[b, :v] == 3.0) and np.all(feats[b, v:] == 0.0)


def test_jobs_complete_in_submission_order(pool):
    lib, h = pool
    big = np.random.default_rng(0).normal(size=2_000_000)
    dst = _aligned(2_000_000, np.float32)
    lens = np.array([big.size], dtype=np.int64)
    off = np.zeros(1, dtype=np.int64)
    ptrs = (C.c_void_p * 1)(big.ctypes.data)
    tickets = [lib.b200fe_host_pack_begin(h, ptrs, lens.ctypes.data, 1, 2, dst.ctypes.data, off.ctypes.data, dst.size) for _ in range(4)]
    assert tickets == sorted(tickets)
    for tk in reversed(tickets):                                                        # waiting out of order is allowed
        assert lib.b200fe_host_wait(h, tk) == 0
    assert np.array_equal(dst, big.astype(np.float32))


def test_stale_ranges_of_a_recycled_batch_buffer():
    """HostPipeline clears only what an earlier batch left inside this batch's padding: brute-force check of the interval logic."""
    import importlib
    hp = importlib.import_module("lighting-asr_b200.host_pipeline")
    rng = np.random.default_rng(0)
    for _ in range(500):
        size = 260

        def rand_intervals():
            pts = np.sort(rng.choice(size, size=2 * int(rng.integers(0, 8)), replace=False))
            return [(int(pts[2 * i]), int(pts[2 * i + 1])) for i in range(len(pts) // 2)]

        dirty, valid = rand_intervals(), rand_intervals()
        view = int(rng.integers(1, size))
        valid = [(a, min(b, view)) for a, b in valid if a < view]
        zero, after = hp.stale_ranges(dirty, valid, view)
        buf = np.zeros(size, dtype=int)
        for a, b in dirty:
            buf[a:b] = 1
        want = buf.copy()
        for a, b in valid:
            want[a:b] = 0
        want[view:] = 0
        got = np.zeros(size, dtype=int)
        for a, b in zero:
            assert b > a
            got[a:b] += 1
        assert np.array_equal(got, want)
        state = buf.copy()
        for a, b in zero:
            state[a:b] = 0
        for a, b in valid:
            state[a:b] = 1
        cover = np.zeros(size, dtype=int)
        for a, b in after:
            cover[a:b] = 1
        assert np.all(cover >= state)


def test_zero_ranges(pool):
    lib, h = pool
    buf = np.full(5000, 7, dtype=np.uint8)
    off = np.array([3, 100, 4000], dtype=np.int64)
    nb = np.array([50, 0, 999], dtype=np.int64)
    tk = lib.b200fe_host_zero_ranges_begin(h, buf.ctypes.data, off.ctypes.data, nb.ctypes.data, 3)
    assert tk > 0 and lib.b200fe_host_wait(h, tk) == 0
    want = np.full(5000, 7, dtype=np.uint8)
    want[3:53] = 0
    want[4000:4999] = 0
    assert np.array_equal(buf, want)
