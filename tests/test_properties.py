"""CPU property tests (hypothesis / exhaustive small cases) for the pieces whose correctness does not need a GPU:
the Viterbi oracle against brute-force path enumeration, the RIFF reader against hand-built files of every
supported sample format, run-length encoding, the CSV exporter's float formatting, the VBx window plan."""
import io
import itertools
import struct

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from inaspeechsegmenter_b200 import export_funcs
from inaspeechsegmenter_b200 import io as iss_io
from inaspeechsegmenter_b200.segmenter import _rle
from oracle import viterbi_oracle as vo


# ------------------------------------------------------------------ Viterbi oracle vs brute force
def _brute_force(emission, transition):
    """argmax over all K^T paths of  log(1/K) + e[0,s0] + sum_t (trans[s_{t-1}, s_t] + e[t, s_t])
    (uniform initial distribution, pyannote_viterbi.py:166-167)."""
    T, K = emission.shape
    best, arg = -np.inf, None
    for path in itertools.product(range(K), repeat=T):
        s = np.log(1.0 / K) + emission[0, path[0]]
        for t in range(1, T):
            s += transition[path[t - 1], path[t]] + emission[t, path[t]]
        if s > best:
            best, arg = s, path
    return np.array(arg)


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 7), st.sampled_from([2, 3]), st.integers(0, 2 ** 31 - 1))
def test_viterbi_oracle_is_the_map_path(T, K, seed):
    rng = np.random.default_rng(seed)
    emission = np.log(rng.dirichlet(np.ones(K), size=T))
    transition = np.log(rng.dirichlet(np.ones(K), size=K))
    want = _brute_force(emission, transition)
    assert np.array_equal(vo.viterbi_numpy(emission, transition), want)
    assert np.array_equal(vo.viterbi_c(emission, transition), want)


def test_viterbi_oracle_with_the_segmenter_transitions():
    """The two transition families the reference uses (viterbi_utils.py:29-49) on hard 0/1 evidence:
    short flips are smoothed away, long runs survive."""
    raw = np.array([0] * 40 + [1] * 3 + [0] * 40 + [1] * 400 + [0] * 5 + [1] * 400)
    states = vo.viterbi_c(vo.pred2logemission(raw), vo.log_trans_exp(150, cost0=-5))
    assert states[:83].sum() == 0 and states[83:].all()
    probs = np.full((300, 3), 0.05)
    probs[:100, 0] = probs[100:104, 1] = probs[104:, 0] = 0.9
    st3 = vo.viterbi_c(np.log(probs / probs.sum(1, keepdims=True)), vo.diag_trans_exp(80, 3))
    assert (st3 == 0).all()


# ------------------------------------------------------------------ RIFF reader
def _wav(tag, bits, nch, sr, payload, extensible=False, odd_chunk=False):
    block = nch * bits // 8
    if extensible:
        fmt = struct.pack('<HHIIHH', 0xFFFE, nch, sr, sr * block, block, bits) + struct.pack('<HHI', 22, bits, 0) + \
            struct.pack('<H', tag) + b'\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71'
    else:
        fmt = struct.pack('<HHIIHH', tag, nch, sr, sr * block, block, bits)
    chunks = b'fmt ' + struct.pack('<I', len(fmt)) + fmt
    if odd_chunk:
        chunks += b'LIST' + struct.pack('<I', 3) + b'abc' + b'\x00'           # odd-sized chunk + pad byte
    chunks += b'data' + struct.pack('<I', len(payload)) + payload
    return b'RIFF' + struct.pack('<I', 4 + len(chunks)) + b'WAVE' + chunks


@pytest.mark.parametrize('fmt', ['u8', 's16', 's24', 's32', 'f32', 'f64'])
@pytest.mark.parametrize('nch', [1, 2])
def test_wav_reader_formats(fmt, nch):
    rng = np.random.default_rng(5)
    n = 257
    if fmt == 'u8':
        ints = rng.integers(0, 256, n * nch); payload = ints.astype(np.uint8).tobytes(); want = (ints - 128) / 128.0; tag, bits = 1, 8
    elif fmt == 's16':
        ints = rng.integers(-32768, 32768, n * nch); payload = ints.astype('<i2').tobytes(); want = ints / 32768.0; tag, bits = 1, 16
    elif fmt == 's24':
        ints = rng.integers(-(1 << 23), 1 << 23, n * nch)
        payload = b''.join(int(v & 0xFFFFFF).to_bytes(3, 'little') for v in ints); want = ints / float(1 << 23); tag, bits = 1, 24
    elif fmt == 's32':
        ints = rng.integers(-(1 << 31), 1 << 31, n * nch); payload = ints.astype('<i4').tobytes(); want = ints / float(1 << 31); tag, bits = 1, 32
    elif fmt == 'f32':
        want = rng.uniform(-1, 1, n * nch).astype(np.float32); payload = want.astype('<f4').tobytes(); tag, bits = 3, 32
    else:
        want = rng.uniform(-1, 1, n * nch); payload = want.astype('<f8').tobytes(); tag, bits = 3, 64
    for ext, odd in ((False, False), (True, True)):
        sig, sr = iss_io.read_wav(io.BytesIO(_wav(tag, bits, nch, 16000, payload, extensible=ext, odd_chunk=odd)))
        assert sr == 16000 and sig.dtype == np.float64
        assert sig.shape == ((n, nch) if nch > 1 else (n,))
        assert np.array_equal(sig.reshape(-1), np.asarray(want, dtype=np.float64))


def test_wav_reader_streamed_size_and_errors():
    pcm = np.arange(-5, 5, dtype='<i2').tobytes()
    raw = bytearray(_wav(1, 16, 1, 16000, pcm))
    raw[-len(pcm) - 4:-len(pcm)] = struct.pack('<I', 0xFFFFFFFF)      # ffmpeg writes -1 when piping
    sig, _ = iss_io.read_wav(io.BytesIO(bytes(raw)), dtype='float32')
    assert np.array_equal(sig, np.arange(-5, 5, dtype=np.float32) / np.float32(32768))
    with pytest.raises(ValueError):
        iss_io.read_wav(io.BytesIO(b'RIFX' + bytes(40)))
    with pytest.raises(NotImplementedError):
        iss_io.read_wav(io.BytesIO(_wav(6, 8, 1, 8000, b'\x00' * 8)))   # A-law


# ------------------------------------------------------------------ RLE + exporter
@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 4), st.integers(1, 9)), min_size=1, max_size=30))
def test_rle_roundtrip(runs):
    merged = []
    for lab, n in runs:                                   # adjacent equal labels merge
        if merged and merged[-1][0] == lab:
            merged[-1][1] += n
        else:
            merged.append([lab, n])
    track = np.concatenate([np.full(n, lab) for lab, n in merged])
    out = _rle(track)
    assert [(lab, b - a) for lab, a, b in out] == [(lab, n) for lab, n in merged]
    assert out[0][1] == 0 and out[-1][2] == len(track) and all(out[i][2] == out[i + 1][1] for i in range(len(out) - 1))


@settings(max_examples=80, deadline=None)
@given(st.floats(0, 1e5, allow_nan=False), st.integers(0, 10 ** 6), st.integers(1, 10 ** 4))
def test_csv_floats_roundtrip_and_match_repr(start_sec, i, n):
    a, b = start_sec + i * .02, start_sec + (i + n) * .02
    line = export_funcs.seg2csv([('speech', a, b)]).split('\n')[1].split('\t')
    assert float(line[1]) == a and float(line[2]) == b      # shortest repr that round-trips (what pandas writes)
    assert line[1] == repr(a) and line[2] == repr(b)


# ------------------------------------------------------------------ VBx window plan
@pytest.mark.parametrize('M', [0, 1, 143, 144, 145, 153, 154, 155, 167, 168, 169, 177, 178, 1000, 2342])
def test_window_plan_matches_reference_loop(M):
    """vbx_segmenter.py:217-246: windows of 144 every 24 frames while start < M - 144, then a tail
    [start + 24, M) if at least 10 frames remain."""
    from inaspeechsegmenter_b200.vbx_segmenter import window_plan
    from oracle import vbx_oracle as vx
    want, start = [], 0
    for start in range(0, M - 144, 24):
        want.append((start, 144, False))
    if M - start - 24 >= 10:
        want.append((start + 24, M - (start + 24), True))
    got = [tuple(w) for w in window_plan(M)]
    assert [(s, n, bool(t)) for s, n, t in got] == want
    assert [(s, n, bool(t)) for s, n, t in vx.window_plan(M)] == want
