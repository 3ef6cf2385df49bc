"""Oracle (test infrastructure): SIDEKIT-style log-mel + log-energy front-end.

CPU restatement of the arithmetic of reference
``inaSpeechSegmenter/sidekit_mfcc.py`` as called by ``segmenter.py:58``
(``mfcc(sig.astype(float32), get_mspec=True)`` -> only ``loge`` and ``mspec``
are kept; the DCT at ``sidekit_mfcc.py:337`` is dead work for the segmenter
and is not restated).

Pinned bit-for-bit against the real module by ``tests/golden/make_golden.py``
(run in the build container) and ``tests/test_oracle_golden.py``.
"""
import numpy as np

FS = 16000
WIN = 400          # round(0.025 * 16000)          sidekit_mfcc.py:214
HOP = 160          # int(0.01 * 16000)              sidekit_mfcc.py:215-216
NFFT = 512         # 2 ** ceil(log2(400))           sidekit_mfcc.py:220
NBIN = NFFT // 2 + 1
NMEL = 24
PREFAC = 0.97


def num_frames(n_samples: int) -> int:
    """L = int((N - win) / shift) + 1   (sidekit_mfcc.py:254)."""
    if n_samples < WIN:
        return 0
    return int((n_samples - WIN) / HOP) + 1


def mel_filterbank(fs=FS, nfft=NFFT, lowfreq=100, maxfreq=8000, nfilt=NMEL):
    """24 area-normalised HTK-mel triangles, float32 [nfilt, nfft/2+1].

    Follows ``trfbank(fs, nfft, lowfreq, maxfreq, nlinfilt=0, nlogfilt=24)``
    (sidekit_mfcc.py:118-197), the ``nlinfilt == 0`` branch (:146-154):
    edges are equally spaced on the HTK mel scale (:54-63, :86-95), heights
    are ``2 / (f[i+2] - f[i])`` (:177), left slope covers bins
    ``floor(low*nfft/fs)+1 .. floor(cen*nfft/fs)``, the right slope drops its
    last bin (``rid[:-1]``, :189-195).  All edge maths in float64, the
    table itself is float32 (PARAM_TYPE, :51).
    """
    lo_mel = 2595 * np.log10(1 + lowfreq / 700.)
    hi_mel = 2595 * np.log10(1 + maxfreq / 700.)
    step = (hi_mel - lo_mel) / (nfilt + 1)
    mels = np.zeros(nfilt + 2)
    mels[:] = lo_mel + np.arange(nfilt + 2) * step
    edges = 700. * (10 ** (mels / 2595.) - 1)
    heights = 2. / (edges[2:] - edges[:-2])
    bank = np.zeros((nfilt, nfft // 2 + 1), dtype=np.float32)
    bin_hz = np.arange(nfft) / (1. * nfft) * fs
    for i in range(nfilt):
        low, cen, hi = edges[i], edges[i + 1], edges[i + 2]
        b0 = np.floor(low * nfft / fs) + 1
        b1 = np.floor(cen * nfft / fs) + 1
        b2 = min(np.floor(hi * nfft / fs) + 1, nfft)
        left = np.arange(b0, b1, dtype=np.int32)
        right = np.arange(b1, b2, dtype=np.int32)[:-1]
        bank[i][left] = (heights[i] / (cen - low)) * (bin_hz[left] - low)
        bank[i][right] = (heights[i] / (hi - cen)) * (hi - bin_hz[right])
    return bank, edges


def hann_window(n=WIN):
    """``numpy.hanning(400)`` (symmetric, float64)  sidekit_mfcc.py:223."""
    return np.hanning(n)


def frame_signal(sig, win=WIN, hop=HOP):
    """[L, win] copy of overlapping frames (sidekit_mfcc.py:240-263, :216)."""
    sig = np.ascontiguousarray(sig)
    L = num_frames(len(sig))
    if L == 0:
        return np.zeros((0, win), dtype=sig.dtype)
    view = np.lib.stride_tricks.as_strided(
        sig, shape=(L, win), strides=(hop * sig.itemsize, sig.itemsize))
    return view.copy()


def pre_emphasise(frames, pre=PREFAC):
    """Per-frame first difference; the frame's first sample is its own
    predecessor (sidekit_mfcc.py:266-275): y[0] = x[0] - pre*x[0]."""
    prev = np.concatenate([frames[:, :1], frames[:, :-1]], axis=1)
    return frames - prev * pre      # float32 array * python float -> float32


def logmel_loge(sig, chunk=500000):
    """(mspec[L,24] f32, loge[L] f32) for a float32 16 kHz mono signal.

    power_spectrum (sidekit_mfcc.py:200-237): loge = log(sum(y^2)) in float32
    BEFORE windowing (:226); y*hanning promotes to float64 (:231), rfft 512 in
    float64 (:232), re^2+im^2 stored as float32 (:225,:233).  Then
    mspec = log(spec @ fbank.T) in float32, no floor (:334) -- silent frames
    give -inf and that is expected (segmenter.py:55-58).
    """
    sig = np.asarray(sig, dtype=np.float32)
    frames = pre_emphasise(frame_signal(sig))
    L = frames.shape[0]
    with np.errstate(divide='ignore'):
        loge = np.log((frames ** 2).sum(axis=1))
    spec = np.ones((L, NBIN), dtype=np.float32)
    win = hann_window()
    for a in range(0, L, chunk):
        b = min(a + chunk, L)
        mag = np.fft.rfft(frames[a:b] * win, NFFT, axis=-1)
        spec[a:b] = mag.real ** 2 + mag.imag ** 2
    bank = mel_filterbank()[0]
    with np.errstate(divide='ignore'):
        mspec = np.log(np.dot(spec, bank.T))
    return mspec, loge


def logmel_loge_f64(sig):
    """Same quantities with every stage in float64 (no float32 rounding of the
    time-domain pre-emphasis / energy either).  Used only to bound the error
    of BOTH the reference and the CUDA kernel against exact arithmetic."""
    x = np.asarray(sig, dtype=np.float64)
    fr = frame_signal(x)
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)
    y = fr - prev * PREFAC
    with np.errstate(divide='ignore'):
        loge = np.log((y ** 2).sum(axis=1))
        mag = np.fft.rfft(y * hann_window(), NFFT, axis=-1)
        p = mag.real ** 2 + mag.imag ** 2
        mspec = np.log(p @ mel_filterbank()[0].astype(np.float64).T)
    return mspec, loge
