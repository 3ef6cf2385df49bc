"""Oracle (test infrastructure): the numeric glue of reference ``segmenter.py``.

``segmenter.py`` cannot be imported in the build container (it imports
TensorFlow at module level, segmenter.py:31), so its numeric glue is restated
here on the CPU, each function citing the lines it follows.  The CNN forward
(``keras.Model.predict``, segmenter.py:163) is delegated to a callable --
``oracle.cnn_oracle.KerasLikeModel`` in tests.

Weights-free pins: the ``noEnergy`` rows of the reference golden CSVs and the
silence fixture (tests/test_oracle_golden.py).
"""
import struct
import warnings

import numpy as np

from . import sidekit_oracle as sk
from .viterbi_oracle import (viterbi_decoding, pred2logemission,
                             diag_trans_exp, log_trans_exp)

PATCH_W = 68       # frames per CNN patch          segmenter.py:149
PATCH_STEP = 2     # patch hop in frames            segmenter.py:149


def read_wav_16k_mono(path, dtype='float32'):
    """What ``soundfile.read(path, dtype)`` yields for the two encodings the
    reference fixtures use (io.py:51-55): PCM16 -> value / 32768, IEEE float32
    verbatim.  Asserts 16 kHz like io.py:52-54."""
    with open(path, 'rb') as f:
        raw = f.read()
    assert raw[:4] == b'RIFF' and raw[8:12] == b'WAVE'
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(raw):
        cid, size = raw[pos:pos + 4], struct.unpack('<I', raw[pos + 4:pos + 8])[0]
        body = raw[pos + 8:pos + 8 + size]
        if cid == b'fmt ':
            fmt = struct.unpack('<HHIIHH', body[:16])
        elif cid == b'data':
            data = body
            break
        pos += 8 + size + (size & 1)
    tag, nch, sr, _, _, bits = fmt
    assert sr == 16000, sr
    if tag == 1 and bits == 16:
        sig = np.frombuffer(data[:len(data) // 2 * 2], dtype='<i2').astype(dtype) / np.dtype(dtype).type(32768)
    elif tag == 3 and bits == 32:
        sig = np.frombuffer(data[:len(data) // 4 * 4], dtype='<f4').astype(dtype)
    else:
        raise NotImplementedError((tag, bits))
    if nch > 1:
        sig = sig.reshape(-1, nch)
    return sig


def media2feats(sig):
    """segmenter.py:53-67 minus the decode: (mspec, loge, difflen); short
    signals (< 68 frames) are padded with rows of ``min(mspec)`` (:60-65)."""
    mspec, loge = sk.logmel_loge(np.asarray(sig).astype(np.float32))
    difflen = 0
    if len(loge) < PATCH_W:
        difflen = PATCH_W - len(loge)
        mspec = np.concatenate((mspec, np.ones((difflen, 24)) * np.min(mspec)))
    return mspec, loge, difflen


def energy_activity(loge, ratio):
    """segmenter.py:69-73 -- global threshold then 2-state Viterbi."""
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)   # mean of empty slice on all-silent input
        thr = np.mean(loge[np.isfinite(loge)]) + np.log(ratio)
    raw = (loge > thr)
    return viterbi_decoding(pred2logemission(raw), log_trans_exp(150, cost0=-5))


def get_patches(mspec, w=PATCH_W, step=PATCH_STEP):
    """segmenter.py:76-88: sliding (w x h) windows hop ``step``, each
    z-normalised over its w*h values (population std), replicated
    w//(2*step) times on the left and w//(2*step)-1+len%2 times on the right,
    plus the all-finite mask.  ``sliding_window_view(...)[::step, 0]`` is the
    skimage ``view_as_windows(mspec, (w, h), step)`` stand-in."""
    h = mspec.shape[1]
    data = np.lib.stride_tricks.sliding_window_view(mspec, (w, h))[::step, 0]
    data = data.reshape(len(data), w * h)
    with np.errstate(invalid='ignore', divide='ignore'):
        data = (data - np.mean(data, axis=1).reshape(-1, 1)) / np.std(data, axis=1).reshape(-1, 1)
    nl = w // (2 * step)
    nr = w // (2 * step) - 1 + len(mspec) % 2
    data = np.vstack([data[:1]] * nl + [data] + [data[-1:]] * nr)
    finite = np.all(np.isfinite(data), axis=1)
    return data.reshape(len(data), w, h), finite


def binidx2seglist(binidx):
    """segmenter.py:91-108 run-length encoding."""
    out, cur, beg, i = [], None, -1, -1
    for i, e in enumerate(binidx):
        if e != cur:
            if cur is not None:
                out.append((cur, beg, i))
            cur, beg = e, i
    out.append((cur, beg, i + 1))
    return out


class DnnSegmenterOracle:
    """segmenter.py:135-179 with ``predict`` injected (a callable
    ``f32[n,68,nmel,1] -> f32[n,K]``)."""

    def __init__(self, predict, nmel, viterbi_arg, inlabel, outlabels):
        self.predict, self.nmel, self.viterbi_arg = predict, nmel, viterbi_arg
        self.inlabel, self.outlabels = inlabel, outlabels
        self.last_probs = None      # kept so tests can compare per-frame softmax

    def __call__(self, mspec, lseg, difflen=0):
        if self.nmel < 24:
            mspec = mspec[:, :self.nmel].copy()
        patches, finite = get_patches(mspec, PATCH_W, PATCH_STEP)
        if difflen > 0 and int(difflen / 2) > 0:
            # the reference slices [:-int(difflen/2)]; for difflen == 1 that is
            # [:-0] == empty (latent bug, no reference test covers it): guarded.
            patches = patches[:-int(difflen / 2)]
            finite = finite[:-int(difflen / 2)]
        batch = [patches[a:b] for lab, a, b in lseg if lab == self.inlabel]
        rawpred = None
        if batch:
            batch = np.expand_dims(np.concatenate(batch), 3)
            rawpred = np.array(self.predict(batch.astype(np.float32)), dtype=np.float32)
            self.last_probs = rawpred          # the 0.5 override below edits it in place (views)
        ret = []
        for lab, a, b in lseg:
            if lab != self.inlabel:
                ret.append((lab, a, b))
                continue
            n = b - a
            r, rawpred = rawpred[:n], rawpred[n:]
            r[finite[a:b] == False, :] = 0.5           # noqa: E712  (segmenter.py:175)
            with np.errstate(divide='ignore'):
                pred = viterbi_decoding(np.log(r), diag_trans_exp(self.viterbi_arg, len(self.outlabels)))
            for lab2, a2, b2 in binidx2seglist(pred):
                ret.append((self.outlabels[int(lab2)], a2 + a, b2 + a))
        return ret


# class attributes of segmenter.py:182-204
VAD_SM = dict(nmel=21, viterbi_arg=150, inlabel='energy', outlabels=('speech', 'music'))
VAD_SMN = dict(nmel=21, viterbi_arg=80, inlabel='energy', outlabels=('speech', 'music', 'noise'))
GENDER = dict(nmel=24, viterbi_arg=80, inlabel='speech', outlabels=('female', 'male'))


def energy_segments(loge, energy_ratio=0.03):
    """segmenter.py:262-267: energy Viterbi, every 2nd frame, RLE."""
    lseg = []
    for lab, a, b in binidx2seglist(energy_activity(loge, energy_ratio)[::2]):
        lseg.append(('noEnergy' if lab == 0 else 'energy', a, b))
    return lseg


def segment_feats(mspec, loge, difflen, start_sec, vad, gender=None, energy_ratio=0.03):
    """segmenter.py:250-276."""
    lseg = energy_segments(loge, energy_ratio)
    lseg = vad(mspec, lseg, difflen)
    if gender is not None:
        lseg = gender(mspec, lseg, difflen)
    return [(lab, start_sec + a * .02, start_sec + b * .02) for lab, a, b in lseg]
