"""GPU: the streaming handle of the C ABI (b200fe_stream_create / _push / _reset / _destroy) -- S INDEPENDENT streams, ragged
chunk sizes in one push, streams that skip pushes, CUDA-graph replay.  Contract: the concatenation of a
stream's push outputs equals the offline fbank of everything pushed into it, frame for frame (here: bit for bit)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _offline(lasr_b200, audio, **kw):
    fe = lasr_b200.GpuFbankFrontend(**kw)
    n = len(audio)
    pad = (-n) % 4
    w = torch.from_numpy(np.pad(audio.astype(np.float32), (0, pad))).to(DEV).unsqueeze(0)
    win = int(kw.get("sample_frequency", 16000.0) * 0.025)
    if n < win:
        return np.zeros((0, 80), dtype=np.float32)
    f, fl = fe(w, np.array([n], dtype=np.int64))
    return f[0, : int(fl[0])].cpu().numpy()


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("sf", [16000.0, 8000.0])
def test_independent_streams_equal_offline(lasr_b200, graph, sf):
    rng = np.random.default_rng(31 + int(graph) + int(sf))
    S, max_chunk = 7, 1500
    st = lasr_b200.IndependentStreams(S, device=DEV, max_chunk=max_chunk, graph=graph, sample_frequency=sf)
    assert st.max_frames == 1 + (int(sf * 0.025) - 1 + max_chunk - int(sf * 0.025)) // int(sf * 0.010)
    total = {s: np.zeros(0, dtype=np.float32) for s in range(S)}
    got = {s: [] for s in range(S)}
    for push in range(40):
        k = int(rng.integers(1, S + 1))
        ids = rng.choice(S, size=k, replace=False).tolist()           # a different subset every push, in arbitrary order
        chunks = []
        for s in ids:
            ln = int(rng.choice([0, 1, 79, 160, 320, 640, 641, 1500, int(rng.integers(1, max_chunk + 1))]))
            c = rng.uniform(-0.5, 0.5, ln).astype(np.float32)
            chunks.append(c)
            total[s] = np.concatenate([total[s], c])
        outs = st.push(ids, chunks)
        assert len(outs) == k
        for s, o in zip(ids, outs):
            assert o.dim() == 2 and o.shape[1] == 80
            got[s].append(o.cpu().numpy().copy())
    assert st.flags() == 0
    for s in range(S):
        want = _offline(lasr_b200, total[s], sample_frequency=sf)
        have = np.concatenate(got[s]) if got[s] else np.zeros((0, 80), dtype=np.float32)
        assert have.shape == want.shape, (s, have.shape, want.shape)
        if sf == 16000.0:
            assert np.array_equal(have, want), s                       # a frame's arithmetic does not depend on the launch that computes it
        else:
            # 256-point family: two consecutive frames share one complex FFT, so a frame's last bits depend on which frame it is
            # paired with -- and the pairing follows the chunk boundaries
            assert int((np.abs(have - want) > 1e-5 + 1e-4 * np.abs(want)).sum()) == 0, (s, float(np.abs(have - want).max()))    # north_star tolerance


def test_stream_reset_and_global_cmvn(lasr_b200):
    rng = np.random.default_rng(33)
    stats = np.zeros((2, 81))
    stats[0, :80] = rng.normal(5.0, 1.0, 80) * 1000
    stats[1, :80] = (stats[0, :80] / 1000) ** 2 * 1000 + rng.uniform(1.0, 4.0, 80) * 1000
    stats[0, 80] = 1000
    st = lasr_b200.IndependentStreams(3, device=DEV, max_chunk=2000, cmvn="global", cmvn_stats=stats)
    a = rng.uniform(-0.5, 0.5, 1000).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, 1777).astype(np.float32)
    st.push([0, 2], [a, a])
    st.reset([2])                                                       # stream 2 starts over, stream 0 keeps its carry
    o0, o2 = st.push([0, 2], [b, b])
    want0 = _offline(lasr_b200, np.concatenate([a, b]), cmvn="global", cmvn_stats=stats)
    want2 = _offline(lasr_b200, b, cmvn="global", cmvn_stats=stats)
    n_first = 1 + (1000 - 400) // 160
    assert np.array_equal(o0.cpu().numpy(), want0[n_first:]) and np.array_equal(o2.cpu().numpy(), want2)
    with pytest.raises(ValueError):
        st.push([0, 0], [a, a])                                         # a stream at most once per push
    with pytest.raises(ValueError):
        st.push([1], [np.zeros(2001, dtype=np.float32)])
    with pytest.raises(ValueError):
        lasr_b200.IndependentStreams(2, device=DEV, cmvn="utt_mean")


def test_stream_c_abi_flags_and_arguments(lasr_b200):
    import ctypes as C
    st = lasr_b200.IndependentStreams(4, device=DEV, max_chunk=640, graph=False)
    lib = st.lib
    ids = torch.tensor([0, 9], dtype=torch.int32, device=DEV)           # 9 is not a stream of this handle
    lens = torch.tensor([640, 640], dtype=torch.int32, device=DEV)
    chunks = torch.zeros((2, 640), device=DEV)
    out = torch.empty((2, 1, 80), device=DEV)                            # too few output rows for a 640-sample push (4 frames)
    frames = torch.empty((2,), dtype=torch.int64, device=DEV)
    st.push_device(2, ids.data_ptr(), chunks.data_ptr(), 640, lens.data_ptr(), out.data_ptr(), 1, frames.data_ptr())
    assert st.flags() == (2 | 4)
    assert frames.cpu().tolist() == [1, 0]
    assert lib.b200fe_stream_push(st.handle, None, 0, None, 0, None, None, None, None, 1, None, None) == -1
    assert b"stream_push" in lib.b200fe_last_error()
    h = C.c_void_p()
    assert lib.b200fe_stream_create(st._plan.handle, 0, 640, C.byref(h)) == -1
